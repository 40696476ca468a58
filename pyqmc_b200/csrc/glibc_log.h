// glibc_log.h -- bit-exact restatement of glibc's double-precision log() as the x86-64 FMA build
// (__log_fma, what the ifunc selects on every AVX2+FMA host) evaluates it, for arguments in (0, 1) and in
// general for positive normal numbers.
//
// Why: numpy's legacy Gaussian generator computes  f = sqrt(-2.0 * log(r2) / r2)  with libm's log()
// (numpy/random/src/legacy/legacy-distributions.c, legacy_gauss).  To reproduce the reference's random
// stream bit for bit ON THE DEVICE (pyqmc/method/mc.py:119 np.random.normal), the device must round exactly
// as that function does.  Algorithm: ARM optimized-routines log (glibc sysdeps/ieee754/dbl-64/e_log.c):
//   x = 2^k z, z in [0x1.6p-1, 0x1.6p0);  i = top 7 mantissa bits of z;  r = fma(z, invc_i, -1)
//   log x = k ln2 + logc_i + log1p(r),  log1p(r) by a degree-5 polynomial; a separate degree-11 polynomial
//   with a split square term handles x in [1 - 2^-4, 1 + 0x1.09p-4).
// The ORDER OF OPERATIONS AND THE FMA CONTRACTIONS below are those of the compiled function (read from the
// disassembly of the image's libm.so.6, glibc 2.39), each written as an explicit fma/mul/add so that neither
// gcc nor nvcc may re-associate or contract differently.  tests/test_glibc_log.py checks qmcb_glibc_log
// (host build of this header) against libm's log() bit for bit; the device build uses the same expressions
// with IEEE round-to-nearest intrinsics (tests/test_gpu_device_rng.py).
#pragma once
#include <cstdint>
#include <cstring>
#include "glibc_log_data.h"

#if defined(__CUDA_ARCH__)
#define QLOG_FN __device__ __forceinline__
#define QLOG_FMA(a, b, c) __fma_rn((a), (b), (c))
#define QLOG_MUL(a, b) __dmul_rn((a), (b))
#define QLOG_ADD(a, b) __dadd_rn((a), (b))
#define QLOG_SUB(a, b) __dsub_rn((a), (b))
#define QLOG_CONST __constant__
#else
#include <cmath>
#define QLOG_FN inline
#define QLOG_FMA(a, b, c) __builtin_fma((a), (b), (c))
#define QLOG_MUL(a, b) qlog_mul((a), (b))
#define QLOG_ADD(a, b) qlog_add((a), (b))
#define QLOG_SUB(a, b) qlog_sub((a), (b))
// volatile round trips keep gcc from contracting a*b+c on the host (this header may be compiled with FMA enabled)
inline double qlog_mul(double a, double b) { volatile double r = a * b; return r; }
inline double qlog_add(double a, double b) { volatile double r = a + b; return r; }
inline double qlog_sub(double a, double b) { volatile double r = a - b; return r; }
#endif

#if defined(__CUDACC__)
__device__ __constant__ double qlog_dev_table[256] = QMCB_LOG_TABLE;
__device__ __constant__ double qlog_dev_A[5] = QMCB_LOG_POLY_A;
__device__ __constant__ double qlog_dev_B[11] = QMCB_LOG_POLY_B;
#endif
static const double qlog_host_table[256] = QMCB_LOG_TABLE;
static const double qlog_host_A[5] = QMCB_LOG_POLY_A;
static const double qlog_host_B[11] = QMCB_LOG_POLY_B;

#if defined(__CUDACC__)
__host__ __device__ __forceinline__
#else
inline
#endif
double qmcb_glibc_log(double x) {
#if defined(__CUDA_ARCH__)
  const double* T = qlog_dev_table;
  const double* A = qlog_dev_A;
  const double* B = qlog_dev_B;
  uint64_t ix = (uint64_t)__double_as_longlong(x);
#else
  const double* T = qlog_host_table;
  const double* A = qlog_host_A;
  const double* B = qlog_host_B;
  uint64_t ix;
  std::memcpy(&ix, &x, 8);
#endif
  const uint64_t LO = 0x3fee000000000000ull;           // 1 - 2^-4
  if (ix - LO <= 0x308ffffffffffull) {                 // x in [1 - 2^-4, 1 + 0x1.09p-4)
    if (ix == 0x3ff0000000000000ull) return 0.0;
    const double r = QLOG_SUB(x, 1.0);
    const double r2 = QLOG_MUL(r, r);
    const double r3 = QLOG_MUL(r, r2);
    double q1 = QLOG_FMA(r, B[2], B[1]);
    double q2 = QLOG_FMA(r, B[5], B[4]);
    double q3 = QLOG_FMA(r, B[8], B[7]);
    q1 = QLOG_FMA(r2, B[3], q1);
    q2 = QLOG_FMA(r2, B[6], q2);
    q3 = QLOG_FMA(r2, B[9], q3);
    q3 = QLOG_FMA(r3, B[10], q3);
    double p = QLOG_FMA(q3, r3, q2);
    p = QLOG_FMA(p, r3, q1);
    // rhi = r + w - w with w = r * 2^27, contracted as the compiler did: fma(r, 2^27, r), then - 2^27 r by fnmadd
    const double big = 134217728.0;
    const double t = QLOG_FMA(r, big, r);
    const double rhi = QLOG_FMA(-big, r, t);
    const double rhi2 = QLOG_MUL(rhi, rhi);
    const double rlo = QLOG_SUB(r, rhi);
    const double hi = QLOG_FMA(rhi2, B[0], r);
    const double d = QLOG_SUB(r, hi);
    const double s = QLOG_ADD(r, rhi);
    double lo = QLOG_FMA(rhi2, B[0], d);
    const double c = QLOG_MUL(B[0], rlo);
    lo = QLOG_FMA(c, s, lo);
    const double y = QLOG_FMA(p, r3, lo);
    return QLOG_ADD(hi, y);
  }
  // (subnormal, zero, negative, inf and nan arguments take libm's slow path; the generator never produces
  // them: r2 is a sum of squares of multiples of 2^-52 below 1, so 2^-104 <= r2 < 1)
  const uint64_t tmp = ix - 0x3fe6000000000000ull;
  const int i = (int)((tmp >> 45) & 127);
  const int k = (int)((int64_t)tmp >> 52);
  const uint64_t iz = ix - (tmp & 0xfff0000000000000ull);
  double z;
#if defined(__CUDA_ARCH__)
  z = __longlong_as_double((long long)iz);
#else
  std::memcpy(&z, &iz, 8);
#endif
  const double invc = T[2 * i], logc = T[2 * i + 1];
  const double kd = (double)k;
  const double w = QLOG_FMA(kd, QMCB_LOG_LN2HI, logc);
  const double r = QLOG_FMA(z, invc, -1.0);
  const double p12 = QLOG_FMA(r, A[2], A[1]);
  const double hi = QLOG_ADD(r, w);
  const double r2 = QLOG_MUL(r, r);
  double lo = QLOG_SUB(w, hi);
  lo = QLOG_ADD(lo, r);
  lo = QLOG_FMA(kd, QMCB_LOG_LN2LO, lo);
  const double r3 = QLOG_MUL(r, r2);
  const double p34 = QLOG_FMA(r, A[4], A[3]);
  lo = QLOG_FMA(r2, A[0], lo);
  const double p = QLOG_FMA(p34, r2, p12);
  const double y = QLOG_FMA(r3, p, lo);
  return QLOG_ADD(y, hi);
}
