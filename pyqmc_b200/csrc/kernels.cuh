// kernels.cuh -- sm_100a kernels of libqmcb200.
//
// Reference statements (relative to /root/reference):
//   Sherman-Morrison row update    pyqmc/wf/slater.py:88-94, 262-291
//   determinant-ratio rows         pyqmc/wf/slater.py:301-380; determinant_tools.py:74-88
//   recompute (slogdet + inverse)  pyqmc/wf/slater.py:227-260
//   Jastrow caches / updates       pyqmc/wf/jastrowspin.py:56-137, 221-249
//   product combination            pyqmc/wf/multiplywf.py:71-132
//   VMC move                       pyqmc/method/mc.py:115-137
//   local energy                   pyqmc/observables/accumulators.py:60-75, energy.py:28-65,
//                                  eval_ecp.py:83-146, 203-275
#pragma once
#include "device_common.cuh"
#include "coop.cuh"
#include "pbc.cuh"

#define QMCB_SLATER 1
#define QMCB_JASTROW 2
#define QMCB_JASTROW3 4

// =========================================================================================
// Slater part of a single-electron query at one point.
//   rat[c] = sum_D c_D ratio_D,c det_D / sum_D c_D det_D,  c = value[, d/dx, d/dy, d/dz[, lap]]
// FAST path: one determinant, identity occupation, n_s <= NMOT: everything in registers.
// General path: MO values go through a per-point scratch column (coalesced over points).
// =========================================================================================
template <int DERIV, int NMOT>
__device__ __forceinline__ void slater_point_fast(const Sys& S, const double* __restrict__ sd,
                                                  const int* __restrict__ si, const State& st, int w,
                                                  int e, double px, double py, double pz,
                                                  double (&rat)[NComp<DERIV>::value],
                                                  double* __restrict__ mo_save) {
  constexpr int NC = NComp<DERIV>::value;
  const int s = e >= S.nup ? 1 : 0;
  const int n = s ? S.ndn : S.nup;
  const int eeff = e - s * S.nup;
  double acc[NC][NMOT];
  eval_mo<DERIV, NMOT>(S, sd, si, s, px, py, pz, 0, acc);
  const double* __restrict__ inv = st.inv[s] + (size_t)w * n * n + eeff;
#pragma unroll
  for (int c = 0; c < NC; ++c) rat[c] = 0.0;
#pragma unroll
  for (int k = 0; k < NMOT; ++k) {
    if (k < n) {
      const double a = inv[k * n];
#pragma unroll
      for (int c = 0; c < NC; ++c) rat[c] = fma(acc[c][k], a, rat[c]);
    }
  }
  if (mo_save != nullptr) {
#pragma unroll
    for (int k = 0; k < NMOT; ++k) mo_save[k] = acc[0][k];
  }
}

template <int DERIV>
__device__ __forceinline__ void slater_point_general(const Sys& S, const double* __restrict__ sd,
                                                     const int* __restrict__ si, const State& st,
                                                     int w, int e, double px, double py, double pz,
                                                     double (&rat)[NComp<DERIV>::value],
                                                     double* __restrict__ mo_save,
                                                     double* __restrict__ scr, size_t scr_stride) {
  // scr: this point's column of the scratch [NC*ldc][scr_stride]
  constexpr int NC = NComp<DERIV>::value;
  const int s = e >= S.nup ? 1 : 0;
  const int n = s ? S.ndn : S.nup;
  const int eeff = e - s * S.nup;
  const int ldc = S.ldc[s];
  // periodic systems: the lattice-summed MO rows of this point were written by k_pbc_mo
  for (int mo0 = 0; mo0 < ldc && !S.pbc; mo0 += 8) {
    double acc[NC][8];
    eval_mo<DERIV, 8>(S, sd, si, s, px, py, pz, mo0, acc);
#pragma unroll
    for (int c = 0; c < NC; ++c)
#pragma unroll
      for (int j = 0; j < 8; ++j) scr[(size_t)(c * ldc + mo0 + j) * scr_stride] = acc[c][j];
  }
  if (mo_save != nullptr)
    for (int j = 0; j < ldc; ++j) mo_save[j] = scr[(size_t)j * scr_stride];
  const int nds = S.nds[s];
  const int* __restrict__ occ = si + S.o_occ[s];
  double num[NC], den = 0.0;
#pragma unroll
  for (int c = 0; c < NC; ++c) num[c] = 0.0;
  for (int d = 0; d < nds; ++d) {
    const double* __restrict__ inv = st.inv[s] + ((size_t)w * nds + d) * n * n + eeff;
    double r[NC];
#pragma unroll
    for (int c = 0; c < NC; ++c) r[c] = 0.0;
    for (int k = 0; k < n; ++k) {
      const double a = inv[k * n];
      const int orb = occ[d * n + k];
#pragma unroll
      for (int c = 0; c < NC; ++c) r[c] = fma(scr[(size_t)(c * ldc + orb) * scr_stride], a, r[c]);
    }
    if (S.ndet == 1) {
#pragma unroll
      for (int c = 0; c < NC; ++c) num[c] = r[c];
      den = 1.0;
    } else {
      const double wgt = st.dv[s][(size_t)w * nds + d] * st.W[s][(size_t)w * nds + d];
      den += wgt;
#pragma unroll
      for (int c = 0; c < NC; ++c) num[c] = fma(r[c], wgt, num[c]);
    }
  }
#pragma unroll
  for (int c = 0; c < NC; ++c) rat[c] = num[c] / den;
}

// Everything a single-electron query needs, for the factors selected by `which`.
template <int DERIV, int NMOT>
struct PointEval {
  static constexpr int NC = NComp<DERIV>::value;
  double rat[NC];  // Slater ratios (value, derivatives)
  double du;       // Jastrow log-ratio
  double gj[3];    // Jastrow grad U
  double lapj;     // Jastrow laplacian U
  __device__ __forceinline__ void run(const Sys& S, const double* sd, const int* si, const State& st,
                                      int which, int w, int e, double px, double py, double pz,
                                      double* mo_save, double* scr, size_t scr_stride) {
    rat[0] = 1.0;
#pragma unroll
    for (int c = 1; c < NC; ++c) rat[c] = 0.0;
    du = 0.0;
    gj[0] = gj[1] = gj[2] = 0.0;
    lapj = 0.0;
    if (which & QMCB_SLATER) {
      if constexpr (NMOT > 0)
        slater_point_fast<DERIV, NMOT>(S, sd, si, st, w, e, px, py, pz, rat, mo_save);
      else
        slater_point_general<DERIV>(S, sd, si, st, w, e, px, py, pz, rat, mo_save, scr, scr_stride);
    }
    if (which & QMCB_JASTROW) jastrow_point<DERIV>(S, sd, si, st, w, e, px, py, pz, du, gj, lapj);
    // the three-body factor adds to the same log-ratio / grad U / lap U (a sum of Jastrow
    // exponents combines exactly like the product rule of multiplywf.py:121-129)
    if (which & QMCB_JASTROW3) jastrow3_point<DERIV>(S, sd, si, st, w, e, px, py, pz, du, gj, lapj);
  }
};

// =========================================================================================
// k_point: wf.testvalue / gradient / gradient_value / gradient_laplacian for electron e.
// One thread per (walker, auxiliary point).
// =========================================================================================
enum { PV_VALUE = 0, PV_GRAD = 1, PV_GRADVAL = 2, PV_GRADLAP = 3, PV_MOSAVE = 4 };

struct PointArgs {
  int which, e, npoints, naip, save;
  const int* idx;       // compacted walker list (or nullptr: all walkers)
  const double* pos;    // [N][naip][3]
  const uint8_t* mask;  // PV_MOSAVE only
  double* o_val;        // [npoints]
  double* o_grad;       // [3][N]
  double* o_lap;        // [N]
  double* scr;          // general-path scratch
  size_t scr_stride;
};

template <int MODE, int NMOT>
__global__ void __launch_bounds__(128) k_point(const Sys S, const State st, const PointArgs pa) {
  const double* sd;
  const int* si;
  stage_tables(S, sd, si);
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= pa.npoints) return;
  const int m = p / pa.naip, q = p - m * pa.naip;
  const int w = pa.idx ? pa.idx[m] : m;
  const double* pos = pa.pos + ((size_t)w * pa.naip + q) * 3;
  const double px = pos[0], py = pos[1], pz = pos[2];
  constexpr int DERIV = (MODE == PV_VALUE || MODE == PV_MOSAVE) ? 0 : (MODE == PV_GRADLAP ? 2 : 1);
  if (MODE == PV_MOSAVE && pa.mask && !pa.mask[w]) return;
  double* mo_save = nullptr;
  if (pa.save && (pa.which & QMCB_SLATER)) {
    const int s = pa.e >= S.nup ? 1 : 0;
    mo_save = st.saved_mo + (size_t)w * S.ldc[s];
  }
  PointEval<DERIV, NMOT> ev;
  ev.run(S, sd, si, st, MODE == PV_MOSAVE ? QMCB_SLATER : pa.which, w, pa.e, px, py, pz, mo_save,
         pa.scr + p, pa.scr_stride);
  if (pa.save) {
    st.saved_pos[(size_t)w * 3 + 0] = px;
    st.saved_pos[(size_t)w * 3 + 1] = py;
    st.saved_pos[(size_t)w * 3 + 2] = pz;
  }
  const int N = st.N;
  if (MODE == PV_VALUE) {
    pa.o_val[p] = ev.rat[0] * exp(ev.du);
  } else if (MODE == PV_GRAD) {
#pragma unroll
    for (int i = 0; i < 3; ++i) {
      double g = ev.gj[i];
      if (pa.which & QMCB_SLATER) g = ev.rat[1 + i] / ev.rat[0] + g;
      pa.o_grad[(size_t)i * N + w] = g;
    }
  } else if (MODE == PV_GRADVAL) {
    // Slater.gradient_value maps non-finite derivatives -> 0, values -> 1 (slater.py:415-417)
    double v = ev.rat[0];
#pragma unroll
    for (int i = 0; i < 3; ++i) {
      double g = ev.gj[i];
      if (pa.which & QMCB_SLATER) {
        double gs = ev.rat[1 + i] / ev.rat[0];
        if (!isfinite(gs)) gs = 0.0;
        g = gs + g;
      }
      pa.o_grad[(size_t)i * N + w] = g;
    }
    if (!isfinite(v)) v = 1.0;
    pa.o_val[w] = v * exp(ev.du);
  } else if (MODE == PV_GRADLAP) {
    double gs[3] = {0.0, 0.0, 0.0}, laps = 0.0;
    if (pa.which & QMCB_SLATER) {
#pragma unroll
      for (int i = 0; i < 3; ++i) gs[i] = ev.rat[1 + i] / ev.rat[0];
      laps = ev.rat[4] / ev.rat[0];
    }
    double lapj = 0.0, cross = 0.0;
    if (pa.which & (QMCB_JASTROW | QMCB_JASTROW3)) {
      lapj = ev.lapj + (ev.gj[0] * ev.gj[0] + ev.gj[1] * ev.gj[1] + ev.gj[2] * ev.gj[2]);
      cross = gs[0] * ev.gj[0] + gs[1] * ev.gj[1] + gs[2] * ev.gj[2];
    }
#pragma unroll
    for (int i = 0; i < 3; ++i) pa.o_grad[(size_t)i * N + w] = gs[i] + ev.gj[i];
    pa.o_lap[w] = (laps + lapj) + cross * 2.0;  // multiplywf.py:121-129
  }
}

// =========================================================================================
// Recompute: MO values of every electron, then per (walker, spin determinant) slogdet+inverse.
// =========================================================================================
__global__ void __launch_bounds__(128) k_mo_all(const Sys S, const State st, int write_values) {
  // MO value / gradient / Laplacian rows of every electron at its current position:
  // values -> mo_all (determinant build, slater.py:239-240), all five -> mocache
  const double* sd;
  const int* si;
  stage_tables(S, sd, si);
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  const int N = st.N;
  if (p >= N * S.ne) return;
  const int w = p / S.ne, e = p - w * S.ne;
  const int s = e >= S.nup ? 1 : 0;
  const double px = CONF(st, S, w, e, 0), py = CONF(st, S, w, e, 1), pz = CONF(st, S, w, e, 2);
  const int ldmax = S.ldc[0] > S.ldc[1] ? S.ldc[0] : S.ldc[1];
  double* out = st.mo_all + ((size_t)w * S.ne + e) * ldmax;
  double* mc = st.mocache + ((size_t)w * S.ne + e) * 5 * ldmax;
  if (S.ldc[s] == 4) {
    double acc[5][4];
    eval_mo<2, 4>(S, sd, si, s, px, py, pz, 0, acc);
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      if (write_values) out[j] = acc[0][j];
#pragma unroll
      for (int c = 0; c < 5; ++c) mc[c * ldmax + j] = acc[c][j];
    }
  } else {
    for (int mo0 = 0; mo0 < S.ldc[s]; mo0 += 8) {
      double acc[5][8];
      eval_mo<2, 8>(S, sd, si, s, px, py, pz, mo0, acc);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        if (write_values) out[mo0 + j] = acc[0][j];
#pragma unroll
        for (int c = 0; c < 5; ++c) mc[c * ldmax + mo0 + j] = acc[c][j];
      }
    }
  }
}

// Gauss-Jordan with partial pivoting on a per-thread matrix held in `a` (stride 1 in local or
// global scratch).  Returns sign and log|det|; a becomes the inverse (zeros if singular).
template <int NMAX>
__global__ void __launch_bounds__(64) k_invert(const Sys S, const State st, int s, double* gscratch) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  const int nds = S.nds[s];
  const int N = st.N;
  if (t >= N * nds) return;
  const int w = t / nds, d = t - w * nds;
  const int n = s ? S.ndn : S.nup;
  if (n == 0) {  // empty determinant = 1
    st.dsign[s][t] = 1.0;
    st.dlog[s][t] = 0.0;
    return;
  }
  const int lo = s ? S.nup : 0;
  const int ldmax = S.ldc[0] > S.ldc[1] ? S.ldc[0] : S.ldc[1];
  const int* __restrict__ occ = S.iblob + S.o_occ[s] + d * n;
  double loc[NMAX > 0 ? NMAX * NMAX : 1];
  int piv[NMAX > 0 ? NMAX : 64];
  double* a = NMAX > 0 ? loc : gscratch + (size_t)t * n * n;
  // M[i][k] = mo(electron lo+i, orbital occ[k])   (slater.py:239-240)
  for (int i = 0; i < n; ++i)
    for (int k = 0; k < n; ++k) a[i * n + k] = st.mo_all[((size_t)w * S.ne + lo + i) * ldmax + occ[k]];
  double sign = 1.0, logdet = 0.0;
  bool singular = false;
  for (int c = 0; c < n; ++c) {
    int p = c;
    double best = fabs(a[c * n + c]);
    for (int r = c + 1; r < n; ++r) {
      const double v = fabs(a[r * n + c]);
      if (v > best) {
        best = v;
        p = r;
      }
    }
    piv[c] = p;
    if (p != c) {
      sign = -sign;
      for (int k = 0; k < n; ++k) {
        const double tmp = a[c * n + k];
        a[c * n + k] = a[p * n + k];
        a[p * n + k] = tmp;
      }
    }
    const double pv = a[c * n + c];
    if (pv == 0.0 || !isfinite(pv)) {
      singular = true;
      break;
    }
    if (pv < 0.0) sign = -sign;
    logdet += log(fabs(pv));
    const double ipv = 1.0 / pv;
    a[c * n + c] = 1.0;
    for (int k = 0; k < n; ++k) a[c * n + k] *= ipv;
    for (int r = 0; r < n; ++r) {
      if (r == c) continue;
      const double f = a[r * n + c];
      a[r * n + c] = 0.0;
      for (int k = 0; k < n; ++k) a[r * n + k] = fma(-f, a[c * n + k], a[r * n + k]);
    }
  }
  double* out = st.inv[s] + (size_t)t * n * n;
  if (singular) {
    for (int i = 0; i < n * n; ++i) out[i] = 0.0;
    st.dsign[s][t] = 0.0;
    st.dlog[s][t] = -INFINITY;
    return;
  }
  for (int c = n - 1; c >= 0; --c) {
    const int p = piv[c];
    if (p != c)
      for (int r = 0; r < n; ++r) {
        const double tmp = a[r * n + c];
        a[r * n + c] = a[r * n + p];
        a[r * n + p] = tmp;
      }
  }
  for (int i = 0; i < n * n; ++i) out[i] = a[i];
  st.dsign[s][t] = sign;
  st.dlog[s][t] = logdet;
}

// Same Gauss-Jordan (same pivoting, same operation order per element) with ONE WARP PER MATRIX for
// 8 < n <= 32: the matrix lives in shared memory with row stride n + 1, lane k owns column k, the
// pivot search runs with lanes over rows.  Used by recompute for large determinants (n = 32 for the
// 2x2x2 diamond supercell), where a thread per matrix is two orders of magnitude slower.
__global__ void __launch_bounds__(128) k_invert_warp(const Sys S, const State st, int s) {
  extern __shared__ __align__(128) unsigned char qmcb_smem[];
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  const long long t = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int nds = S.nds[s];
  const int N = st.N;
  if (t >= (long long)N * nds) return;
  const int w = (int)(t / nds), d = (int)(t - (long long)w * nds);
  const int n = s ? S.ndn : S.nup;
  const int lo = s ? S.nup : 0;
  const int ldmax = S.ldc[0] > S.ldc[1] ? S.ldc[0] : S.ldc[1];
  const int ld = n + 1;
  double* __restrict__ a = reinterpret_cast<double*>(qmcb_smem) + (size_t)wib * (32 * 33 + 32);
  int* __restrict__ piv = reinterpret_cast<int*>(a + 32 * 33);
  const int* __restrict__ occ = S.iblob + S.o_occ[s] + d * n;
  const bool act = lane < n;
  // M[i][k] = mo(electron lo + i, orbital occ[k])   (slater.py:239-240)
  const int myorb = act ? occ[lane] : 0;
  for (int i = 0; i < n; ++i)
    if (act) a[i * ld + lane] = st.mo_all[((size_t)w * S.ne + lo + i) * ldmax + myorb];
  __syncwarp();
  double sign = 1.0, logdet = 0.0;
  bool singular = false;
  for (int c = 0; c < n; ++c) {
    // pivot: first row r >= c with the largest |a[r][c]|
    double v = (lane >= c && act) ? fabs(a[lane * ld + c]) : -1.0;
    int p = lane;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const double ov = __shfl_xor_sync(0xffffffffu, v, o);
      const int op = __shfl_xor_sync(0xffffffffu, p, o);
      if (ov > v || (ov == v && op < p)) {
        v = ov;
        p = op;
      }
    }
    if (lane == 0) piv[c] = p;
    if (p != c) {
      sign = -sign;
      if (act) {
        const double tmp = a[c * ld + lane];
        a[c * ld + lane] = a[p * ld + lane];
        a[p * ld + lane] = tmp;
      }
    }
    __syncwarp();
    const double pv = a[c * ld + c];
    if (pv == 0.0 || !isfinite(pv)) {
      singular = true;
      break;
    }
    if (pv < 0.0) sign = -sign;
    logdet += log(fabs(pv));
    const double ipv = 1.0 / pv;
    const double f = act ? a[lane * ld + c] : 0.0;  // lane r holds a[r][c] (before the row scaling)
    __syncwarp();
    if (act) a[c * ld + lane] = (lane == c ? 1.0 : a[c * ld + lane]) * ipv;
    __syncwarp();
    const double rc = act ? a[c * ld + lane] : 0.0;
    for (int r = 0; r < n; ++r) {
      const double fr = __shfl_sync(0xffffffffu, f, r);
      if (r == c || !act) continue;
      const double x = lane == c ? 0.0 : a[r * ld + lane];
      a[r * ld + lane] = fma(-fr, rc, x);
    }
    __syncwarp();
  }
  double* out = st.inv[s] + (size_t)t * n * n;
  if (singular) {
    for (int i = lane; i < n * n; i += 32) out[i] = 0.0;
    if (lane == 0) {
      st.dsign[s][t] = 0.0;
      st.dlog[s][t] = -INFINITY;
    }
    return;
  }
  for (int c = n - 1; c >= 0; --c) {
    const int p = piv[c];
    if (p != c && act) {  // lane = row: swap columns c and p
      const double tmp = a[lane * ld + c];
      a[lane * ld + c] = a[lane * ld + p];
      a[lane * ld + p] = tmp;
    }
    __syncwarp();
  }
  for (int i = 0; i < n; ++i)
    if (act) out[i * n + lane] = a[i * ld + lane];
  if (lane == 0) {
    st.dsign[s][t] = sign;
    st.dlog[s][t] = logdet;
  }
}

// Multi-determinant caches for walker w:  ref_s = max_d log_s[d], dv_s[d] = sign*exp(log-ref),
// W_s[d] = sum_{D: map_s(D)=d} c_D dv_other[map_other(D)]   (determinant_tools.py:74-88 with a
// per-walker instead of a global reference exponent; the reference cancels in every ratio).
__device__ __forceinline__ void det_cache_warp(const Sys& S, const State& st, int w, int lane, int spin) {
  // one warp per walker, lanes over unique spin determinants.  spin >= 0: only the determinants of
  // that spin changed (single-electron move): refresh ref / dv of that spin and W of the other one.
  for (int s = 0; s < 2; ++s) {
    if (spin >= 0 && s != spin) continue;
    const int nds = S.nds[s];
    double ref = -INFINITY;
    for (int d = lane; d < nds; d += 32) ref = fmax(ref, st.dlog[s][(size_t)w * nds + d]);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) ref = fmax(ref, __shfl_xor_sync(0xffffffffu, ref, o));
    if (!isfinite(ref)) ref = 0.0;
    if (lane == 0) st.ref[s][w] = ref;
    for (int d = lane; d < nds; d += 32)
      st.dv[s][(size_t)w * nds + d] = st.dsign[s][(size_t)w * nds + d] * exp(st.dlog[s][(size_t)w * nds + d] - ref);
  }
  __syncwarp();
  for (int s = 0; s < 2; ++s) {
    if (spin >= 0 && s == spin) continue;  // W_s depends on dv of the OTHER spin only
    const int nds = S.nds[s], o = 1 - s, ndo = S.nds[o];
    const double* __restrict__ dvo = st.dv[o] + (size_t)w * ndo;
    if (S.dense[s] != nullptr) {  // dense matrix-vector product, coalesced over d
      const double* __restrict__ Cm = S.dense[s];
      for (int d = lane; d < nds; d += 32) {
        double a0 = 0.0, a1 = 0.0, a2 = 0.0, a3 = 0.0;  // four independent chains: the sum is latency-bound
        int j = 0;
        for (; j + 3 < ndo; j += 4) {
          a0 = fma(Cm[(size_t)j * nds + d], dvo[j], a0);
          a1 = fma(Cm[(size_t)(j + 1) * nds + d], dvo[j + 1], a1);
          a2 = fma(Cm[(size_t)(j + 2) * nds + d], dvo[j + 2], a2);
          a3 = fma(Cm[(size_t)(j + 3) * nds + d], dvo[j + 3], a3);
        }
        for (; j < ndo; ++j) a0 = fma(Cm[(size_t)j * nds + d], dvo[j], a0);
        st.W[s][(size_t)w * nds + d] = (a0 + a1) + (a2 + a3);
      }
      continue;
    }
    for (int d = lane; d < nds; d += 32) {
      double acc = 0.0;
      const int k1 = S.grp_off[s][d + 1];
      for (int k = S.grp_off[s][d]; k < k1; ++k) acc = fma(S.grp_coef[s][k], dvo[S.grp_other[s][k]], acc);
      st.W[s][(size_t)w * nds + d] = acc;
    }
  }
}

__global__ void __launch_bounds__(128) k_det_cache(const Sys S, const State st, const uint8_t* mask, int spin) {
  const int lane = threadIdx.x & 31;
  const int w = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (w >= st.N) return;
  if (mask && !mask[w]) return;
  det_cache_warp(S, st, w, lane, spin);
}

// wf.value(): sign and log of  [sum_D c_D D_up D_dn] * exp(U)
__global__ void __launch_bounds__(128) k_value(const Sys S, const State st, int which, double* o_sign,
                                               double* o_log) {
  const double* sd;
  const int* si;
  stage_tables(S, sd, si);
  const int w = blockIdx.x * blockDim.x + threadIdx.x;
  const int N = st.N;
  if (w >= N) return;
  double sign = 1.0, lg = 0.0;
  if (which & QMCB_SLATER) {
    if (S.ndet == 1) {
      const double c = S.detc[0];
      sign = st.dsign[0][w] * st.dsign[1][w] * (c > 0.0 ? 1.0 : (c < 0.0 ? -1.0 : 0.0));
      lg = st.dlog[0][w] + st.dlog[1][w] + log(fabs(c));
    } else {
      double val = 0.0;
      const int nds = S.nds[0];
      for (int d = 0; d < nds; ++d)
        val = fma(st.dv[0][(size_t)w * nds + d], st.W[0][(size_t)w * nds + d], val);
      sign = val > 0.0 ? 1.0 : (val < 0.0 ? -1.0 : 0.0);
      lg = log(fabs(val)) + st.ref[0][w] + st.ref[1][w];
    }
    if (sign == 0.0 || !isfinite(lg)) {  // np.nan_to_num in compute_value
      if (isnan(lg)) lg = 0.0;
      if (lg == -INFINITY) lg = -1.7976931348623157e308;
      if (lg == INFINITY) lg = 1.7976931348623157e308;
    }
  }
  if (which & QMCB_JASTROW) {
    double u = 0.0;
    for (int l = 0; l < S.nb; ++l)
      for (int t = 0; t < 3; ++t) u = fma(BVAL(st, S, w, l, t), sd[S.o_bcoef + l * 3 + t], u);
    double ua = 0.0;
    for (int I = 0; I < S.natom; ++I)
      for (int k = 0; k < S.na; ++k)
        for (int t = 0; t < 2; ++t)
          ua = fma(AVAL(st, S, w, I, k, t), sd[S.o_acoef + (I * S.na + k) * 2 + t], ua);
    lg += u + ua;
  }
  if (which & QMCB_JASTROW3) lg += st.val3[w];
  o_sign[w] = sign;
  o_log[w] = lg;
}

// =========================================================================================
// Jastrow caches
// =========================================================================================
__global__ void __launch_bounds__(128) k_jastrow_recompute(const Sys S, const State st) {
  const double* sd;
  const int* si;
  stage_tables(S, sd, si);
  const int w = blockIdx.x * blockDim.x + threadIdx.x;
  const int N = st.N;
  if (w >= N) return;
  const int ne = S.ne, na = S.na, nb = S.nb, I_ = S.natom;
  for (int i = 0; i < I_ * na * 2; ++i) st.avalues[(size_t)w * I_ * na * 2 + i] = 0.0;
  for (int i = 0; i < nb * 3; ++i) st.bvalues[(size_t)w * nb * 3 + i] = 0.0;
  for (int i = 0; i < ne * nb * 2; ++i) st.b_partial[(size_t)w * ne * nb * 2 + i] = 0.0;
  for (int e = 0; e < ne; ++e) {
    const int s = e >= S.nup ? 1 : 0;
    const double px = CONF(st, S, w, e, 0), py = CONF(st, S, w, e, 1),
                 pz = CONF(st, S, w, e, 2);
    for (int I = 0; I < I_; ++I) {
      double dx = px - sd[S.o_xyz + 3 * I], dy = py - sd[S.o_xyz + 3 * I + 1], dz = pz - sd[S.o_xyz + 3 * I + 2];
      if (S.pbc) min_image(S, sd, dx, dy, dz);
      const double r = sqrt(dx * dx + dy * dy + dz * dz);
      for (int k = 0; k < na; ++k) {
        double v = 0.0, g, l;
        if (r < S.rcut_a) radial_func<0>(si[S.o_akind + k], sd[S.o_apar + k], S.rcut_a, r, v, g, l);
        APART(st, S, w, e, I, k) = v;
        AVAL(st, S, w, I, k, s) += v;
      }
    }
    for (int j = e + 1; j < ne; ++j) {
      const int sj = j >= S.nup ? 1 : 0;
      double dx = px - CONF(st, S, w, j, 0), dy = py - CONF(st, S, w, j, 1), dz = pz - CONF(st, S, w, j, 2);
      if (S.pbc) min_image(S, sd, dx, dy, dz);
      const double r = sqrt(dx * dx + dy * dy + dz * dz);
      if (r < S.rcut_b) {
        for (int l = 0; l < nb; ++l) {
          double v, g, ll;
          radial_func<0>(si[S.o_bkind + l], sd[S.o_bpar + l], S.rcut_b, r, v, g, ll);
          BVAL(st, S, w, l, s + sj) += v;
          BPART(st, S, w, e, l, sj) += v;
          BPART(st, S, w, j, l, s) += v;
        }
      }
    }
  }
}

// Cooperative form of the same recompute for many-electron systems: one CTA per walker, threads over
// electrons.  Every electron accumulates its own partial sums over its partners in ascending partner
// order (the order the pair loop above produces), the pair sums follow from the partial sums:
//   bvalues[l][uu] = 1/2 sum_{e up} b_partial[e][l][up],  [ud] = sum_{e up} b_partial[e][l][dn],
//   bvalues[l][dd] = 1/2 sum_{e dn} b_partial[e][l][dn]
__global__ void __launch_bounds__(128) k_jastrow_recompute_coop(const Sys S, const State st) {
  const double* sd;
  const int* si;
  stage_tables(S, sd, si);
  const int w = blockIdx.x;
  const int ne = S.ne, na = S.na, nb = S.nb, I_ = S.natom;
  for (int e = threadIdx.x; e < ne; e += blockDim.x) {
    const double px = CONF(st, S, w, e, 0), py = CONF(st, S, w, e, 1), pz = CONF(st, S, w, e, 2);
    for (int I = 0; I < I_; ++I) {
      double dx = px - sd[S.o_xyz + 3 * I], dy = py - sd[S.o_xyz + 3 * I + 1], dz = pz - sd[S.o_xyz + 3 * I + 2];
      if (S.pbc) min_image(S, sd, dx, dy, dz);
      const double r = sqrt(dx * dx + dy * dy + dz * dz);
      for (int k = 0; k < na; ++k) {
        double v = 0.0, g, l;
        if (r < S.rcut_a) radial_ool<0>(si[S.o_akind + k], sd[S.o_apar + k], S.rcut_a, r, v, g, l);
        APART(st, S, w, e, I, k) = v;
      }
    }
    double bs[8][2];
    for (int l = 0; l < 8; ++l) bs[l][0] = bs[l][1] = 0.0;
    for (int j = 0; j < ne; ++j) {
      if (j == e) continue;
      const int sj = j >= S.nup ? 1 : 0;
      double dx = px - CONF(st, S, w, j, 0), dy = py - CONF(st, S, w, j, 1), dz = pz - CONF(st, S, w, j, 2);
      if (S.pbc) min_image(S, sd, dx, dy, dz);
      const double r = sqrt(dx * dx + dy * dy + dz * dz);
      if (r < S.rcut_b) {
#pragma unroll
        for (int l = 0; l < 8; ++l) {
          if (l < nb) {
            double v, g, ll;
            radial_ool<0>(si[S.o_bkind + l], sd[S.o_bpar + l], S.rcut_b, r, v, g, ll);
            bs[l][sj] += v;
          }
        }
      }
    }
#pragma unroll
    for (int l = 0; l < 8; ++l)
      if (l < nb) {
        BPART(st, S, w, e, l, 0) = bs[l][0];
        BPART(st, S, w, e, l, 1) = bs[l][1];
      }
  }
  __syncthreads();
  for (int t = threadIdx.x; t < I_ * na * 2; t += blockDim.x) {
    const int s2 = t & 1, k = (t >> 1) % na, I = (t >> 1) / na;
    double acc = 0.0;
    const int e0 = s2 ? S.nup : 0, e1 = s2 ? ne : S.nup;
    for (int e = e0; e < e1; ++e) acc += APART(st, S, w, e, I, k);
    AVAL(st, S, w, I, k, s2) = acc;
  }
  for (int t = threadIdx.x; t < nb * 3; t += blockDim.x) {
    const int l = t / 3, sp = t - l * 3;
    double acc = 0.0;
    if (sp == 0)
      for (int e = 0; e < S.nup; ++e) acc += BPART(st, S, w, e, l, 0);
    else if (sp == 1)
      for (int e = 0; e < S.nup; ++e) acc += BPART(st, S, w, e, l, 1);
    else
      for (int e = S.nup; e < ne; ++e) acc += BPART(st, S, w, e, l, 1);
    BVAL(st, S, w, l, sp) = sp == 1 ? acc : 0.5 * acc;
  }
}

// Few-electron form of the same recompute: G lanes per walker.  Lanes over electrons evaluate the radial
// functions once per pair into shared memory; lanes over cache entries then add them in exactly the order
// of the one-thread loop above (pairs in (i, j) order, partners ascending), so the caches are the same sums.
template <int G>
__global__ void __launch_bounds__(128) k_jastrow_recompute_group(const Sys S, const State st) {
  const double* sd;
  const int* si;
  stage_tables(S, sd, si);
  extern __shared__ __align__(128) unsigned char qmcb_smem[];
  const int lane32 = threadIdx.x & 31;
  const int lane = lane32 & (G - 1);
  const unsigned gm = group_mask<G>(lane32);
  const int slot = threadIdx.x / G;
  const int w = (blockIdx.x * blockDim.x + threadIdx.x) / G;
  if (w >= st.N) return;
  const int ne = S.ne, na = S.na, nb = S.nb, I_ = S.natom;
  const size_t tab = (16 + (size_t)S.dwords * 8 + (size_t)S.iwords * 4 + 15) & ~(size_t)15;
  const int per = (S.npair * nb + 1) & ~1;
  double* pv = reinterpret_cast<double*>(qmcb_smem + tab) + (size_t)slot * per;
  for (int e = lane; e < ne; e += G) {
    const double px = CONF(st, S, w, e, 0), py = CONF(st, S, w, e, 1), pz = CONF(st, S, w, e, 2);
    for (int I = 0; I < I_; ++I) {
      double dx = px - sd[S.o_xyz + 3 * I], dy = py - sd[S.o_xyz + 3 * I + 1], dz = pz - sd[S.o_xyz + 3 * I + 2];
      if (S.pbc) min_image(S, sd, dx, dy, dz);
      const double r = sqrt(dx * dx + dy * dy + dz * dz);
      for (int k = 0; k < na; ++k) {
        double v = 0.0, g, l;
        if (r < S.rcut_a) radial_ool<0>(si[S.o_akind + k], sd[S.o_apar + k], S.rcut_a, r, v, g, l);
        APART(st, S, w, e, I, k) = v;
      }
    }
    for (int j = e + 1; j < ne; ++j) {
      double dx = px - CONF(st, S, w, j, 0), dy = py - CONF(st, S, w, j, 1), dz = pz - CONF(st, S, w, j, 2);
      if (S.pbc) min_image(S, sd, dx, dy, dz);
      const double r = sqrt(dx * dx + dy * dy + dz * dz);
      double* out = pv + pair_index(ne, e, j) * nb;
      for (int l = 0; l < nb; ++l) {
        double v = 0.0, g, ll;
        if (r < S.rcut_b) radial_ool<0>(si[S.o_bkind + l], sd[S.o_bpar + l], S.rcut_b, r, v, g, ll);
        out[l] = v;
      }
    }
  }
  __syncwarp(gm);
  for (int t = lane; t < I_ * na * 2; t += G) {
    const int s2 = t & 1, k = (t >> 1) % na, I = (t >> 1) / na;
    double acc = 0.0;
    const int e0 = s2 ? S.nup : 0, e1 = s2 ? ne : S.nup;
    for (int e = e0; e < e1; ++e) acc += APART(st, S, w, e, I, k);
    AVAL(st, S, w, I, k, s2) = acc;
  }
  for (int t = lane; t < ne * nb * 2; t += G) {
    const int tt = t & 1, l = (t >> 1) % nb, x = (t >> 1) / nb;
    double acc = 0.0;
    const int p0 = tt ? S.nup : 0, p1 = tt ? ne : S.nup;
    for (int p = p0; p < p1; ++p) {
      if (p == x) continue;
      acc += pv[pair_index(ne, p < x ? p : x, p < x ? x : p) * nb + l];
    }
    BPART(st, S, w, x, l, tt) = acc;
  }
  for (int t = lane; t < nb * 3; t += G) {
    const int l = t / 3, cls = t - l * 3;
    double acc = 0.0;
    for (int e = 0; e < ne; ++e) {
      const int s = e >= S.nup ? 1 : 0;
      for (int j = e + 1; j < ne; ++j)
        if (s + (j >= S.nup ? 1 : 0) == cls) acc += pv[pair_index(ne, e, j) * nb + l];
    }
    BVAL(st, S, w, l, cls) = acc;
  }
}

// updateinternals of the Jastrow caches for accepted walkers (jastrowspin.py:111-137,221-249)
// and move of the walker coordinates (coord.py:54-62).  New position = st.saved_pos[w].
__global__ void __launch_bounds__(128) k_jastrow_update(const Sys S, const State st, int e, int do_jastrow,
                                                        int move_conf, const uint8_t* mask) {
  const double* sd;
  const int* si;
  stage_tables(S, sd, si);
  const int w = blockIdx.x * blockDim.x + threadIdx.x;
  const int N = st.N;
  if (w >= N) return;
  if (mask && !mask[w]) return;
  const int s = e >= S.nup ? 1 : 0;
  const double nx = st.saved_pos[(size_t)w * 3], ny = st.saved_pos[(size_t)w * 3 + 1], nz = st.saved_pos[(size_t)w * 3 + 2];
  if (do_jastrow) {
    const int na = S.na, nb = S.nb, I_ = S.natom;
    for (int I = 0; I < I_; ++I) {
      double dx = nx - sd[S.o_xyz + 3 * I], dy = ny - sd[S.o_xyz + 3 * I + 1], dz = nz - sd[S.o_xyz + 3 * I + 2];
      if (S.pbc) min_image(S, sd, dx, dy, dz);
      const double r = sqrt(dx * dx + dy * dy + dz * dz);
      for (int k = 0; k < na; ++k) {
        double v = 0.0, g, l;
        if (r < S.rcut_a) radial_func<0>(si[S.o_akind + k], sd[S.o_apar + k], S.rcut_a, r, v, g, l);
        AVAL(st, S, w, I, k, s) += v - APART(st, S, w, e, I, k);
        APART(st, S, w, e, I, k) = v;
      }
    }
    const double ox = CONF(st, S, w, e, 0), oy = CONF(st, S, w, e, 1),
                 oz = CONF(st, S, w, e, 2);
    // new partial sums of electron e (accumulated in partner order, as _b_update does)
    for (int l = 0; l < nb; ++l) {
      double bnew[2] = {0.0, 0.0};
      for (int j = 0; j < S.ne; ++j) {
        if (j == e) continue;
        const int sj = j >= S.nup ? 1 : 0;
        const double jx = CONF(st, S, w, j, 0), jy = CONF(st, S, w, j, 1),
                     jz = CONF(st, S, w, j, 2);
        double dx = nx - jx, dy = ny - jy, dz = nz - jz;
        if (S.pbc) min_image(S, sd, dx, dy, dz);
        const double rn = sqrt(dx * dx + dy * dy + dz * dz);
        dx = ox - jx;
        dy = oy - jy;
        dz = oz - jz;
        if (S.pbc) min_image(S, sd, dx, dy, dz);
        const double ro = sqrt(dx * dx + dy * dy + dz * dz);
        double vn = 0.0, vo = 0.0, g, ll;
        if (rn < S.rcut_b) radial_func<0>(si[S.o_bkind + l], sd[S.o_bpar + l], S.rcut_b, rn, vn, g, ll);
        if (ro < S.rcut_b) radial_func<0>(si[S.o_bkind + l], sd[S.o_bpar + l], S.rcut_b, ro, vo, g, ll);
        bnew[sj] += vn;
        BPART(st, S, w, j, l, s) += vn - vo;
      }
      for (int t = 0; t < 2; ++t) {
        BVAL(st, S, w, l, s + t) += bnew[t] - BPART(st, S, w, e, l, t);
        BPART(st, S, w, e, l, t) = bnew[t];
      }
    }
  }
  if (!move_conf) return;
  CONF(st, S, w, e, 0) = nx;
  CONF(st, S, w, e, 1) = ny;
  CONF(st, S, w, e, 2) = nz;
  if (S.pbc)
    for (int i = 0; i < 3; ++i) st.wrap[((size_t)w * S.ne + e) * 3 + i] = st.saved_wrap[(size_t)w * 3 + i];
}

// =========================================================================================
// Three-body Jastrow caches (three_body_jastrow.py:66-104 recompute, 149-189 updateinternals,
// 657-719 pgradient).  One thread per walker.
// =========================================================================================
__global__ void __launch_bounds__(128) k_jastrow3_recompute(const Sys S, const State st) {
  const double* sd;
  const int* si;
  stage_tables(S, sd, si);
  const int w = blockIdx.x * blockDim.x + threadIdx.x;
  if (w >= st.N) return;
  double av[QMCB_J3_MAXA], ag[1], al[1];
  for (int e = 0; e < S.ne; ++e) {
    j3_a_values<0>(S, sd, si, CONF(st, S, w, e, 0), CONF(st, S, w, e, 1), CONF(st, S, w, e, 2), av, ag, al);
    for (int i = 0; i < S.natom * S.na3; ++i) st.a3v[((size_t)w * S.ne + e) * S.natom * S.na3 + i] = av[i];
  }
  double tot = 0.0;
  for (int e = 0; e < S.ne; ++e) {
    const double px = CONF(st, S, w, e, 0), py = CONF(st, S, w, e, 1), pz = CONF(st, S, w, e, 2);
    for (int i = 0; i < S.natom * S.na3; ++i) av[i] = st.a3v[((size_t)w * S.ne + e) * S.natom * S.na3 + i];
    double P = 0.0, g[3] = {0.0, 0.0, 0.0}, lap = 0.0;
    for (int j = 0; j < S.ne; ++j) {
      if (j == e) continue;
      j3_pair<0>(S, sd, si, st, w, e, j, px, py, pz, av, ag, al, CONF(st, S, w, j, 0), CONF(st, S, w, j, 1),
                 CONF(st, S, w, j, 2), P, g, lap);
    }
    st.P3[(size_t)w * S.ne + e] = P;
    tot += P;
  }
  st.val3[w] = 0.5 * tot;
}

// the same with G lanes per walker, lanes over electrons: a-values of every electron first, then -- after the group's
// writes are visible -- the pair sums P_e; lane 0 adds the P_e in electron order (the one-thread order)
template <int G>
__global__ void __launch_bounds__(128) k_jastrow3_recompute_group(const Sys S, const State st) {
  const double* sd;
  const int* si;
  stage_tables(S, sd, si);
  const int lane32 = threadIdx.x & 31;
  const int lane = lane32 & (G - 1);
  const unsigned gm = group_mask<G>(lane32);
  const int w = (blockIdx.x * blockDim.x + threadIdx.x) / G;
  if (w >= st.N) return;
  const int na_tot = S.natom * S.na3;
  double av[QMCB_J3_MAXA], ag[1], al[1];
  for (int e = lane; e < S.ne; e += G) {
    j3_a_values<0>(S, sd, si, CONF(st, S, w, e, 0), CONF(st, S, w, e, 1), CONF(st, S, w, e, 2), av, ag, al);
    for (int i = 0; i < na_tot; ++i) st.a3v[((size_t)w * S.ne + e) * na_tot + i] = av[i];
  }
  __syncwarp(gm);
  for (int e = lane; e < S.ne; e += G) {
    const double px = CONF(st, S, w, e, 0), py = CONF(st, S, w, e, 1), pz = CONF(st, S, w, e, 2);
    for (int i = 0; i < na_tot; ++i) av[i] = st.a3v[((size_t)w * S.ne + e) * na_tot + i];
    double P = 0.0, g[3] = {0.0, 0.0, 0.0}, lap = 0.0;
    for (int j = 0; j < S.ne; ++j) {
      if (j == e) continue;
      j3_pair<0>(S, sd, si, st, w, e, j, px, py, pz, av, ag, al, CONF(st, S, w, j, 0), CONF(st, S, w, j, 1),
                 CONF(st, S, w, j, 2), P, g, lap);
    }
    st.P3[(size_t)w * S.ne + e] = P;
  }
  __syncwarp(gm);
  if (lane == 0) {
    double tot = 0.0;
    for (int e = 0; e < S.ne; ++e) tot += st.P3[(size_t)w * S.ne + e];
    st.val3[w] = 0.5 * tot;
  }
}

// d U / d ccoeff [N][I][na][na][nb][3]: thread per (walker, I, k, l)
__global__ void __launch_bounds__(128) k_jastrow3_pgrad(const Sys S, const State st, double* __restrict__ out) {
  const double* sd;
  const int* si;
  stage_tables(S, sd, si);
  const int na = S.na3, nb = S.nb3;
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= (long long)st.N * S.natom * na * na) return;
  const int l = (int)(t % na), k = (int)((t / na) % na), I = (int)((t / (na * na)) % S.natom);
  const int w = (int)(t / ((long long)na * na * S.natom));
  double acc[QMCB_J3_MAXB][3];
  for (int m = 0; m < nb; ++m) acc[m][0] = acc[m][1] = acc[m][2] = 0.0;
  for (int i = 0; i < S.ne; ++i)
    for (int j = i + 1; j < S.ne; ++j) {
      double dx = CONF(st, S, w, i, 0) - CONF(st, S, w, j, 0), dy = CONF(st, S, w, i, 1) - CONF(st, S, w, j, 1),
             dz = CONF(st, S, w, i, 2) - CONF(st, S, w, j, 2);
      if (S.pbc) min_image(S, sd, dx, dy, dz);
      const double r = sqrt(dx * dx + dy * dy + dz * dz);
      if (!(r < S.rcut_b3)) continue;
      const int sp = (i >= S.nup ? 1 : 0) + (j >= S.nup ? 1 : 0);
      // (a_k(i) a_l(j) + a_l(i) a_k(j)) / 2
      const double aa = 0.5 * (A3V(st, S, w, i, I, k) * A3V(st, S, w, j, I, l) + A3V(st, S, w, i, I, l) * A3V(st, S, w, j, I, k));
      for (int m = 0; m < nb; ++m) {
        double v, gg, ll;
        radial_ool<0>(si[S.o_b3kind + m], sd[S.o_b3par + m], S.rcut_b3, r, v, gg, ll);
        acc[m][sp] = fma(aa, v, acc[m][sp]);
      }
    }
  for (int m = 0; m < nb; ++m)
    for (int sp = 0; sp < 3; ++sp) out[(t * nb + m) * 3 + sp] = acc[m][sp];
}

// =========================================================================================
// Sherman-Morrison row-replacement update  (slater.py:88-94):
//   t_j = sum_k vec_k inv[k][j];  ratio = t_e;  col_k = inv[k][e] / ratio
//   inv'[k][j] = inv[k][j] - col_k t_j;  inv'[k][e] = col_k
// followed by  sign *= sgn(ratio), log += log|ratio|  (slater.py:290-291).
// Matrix m belongs to walker m / nds, determinant m % nds; its new row is gathered from the
// walker's MO row through the occupation list (or read directly when occ == nullptr).
// Algorithmic traffic per updated matrix: read n^2 + n, write n^2 + 1 doubles.
// =========================================================================================
struct SmArgs {
  int n, e, nds, vec_stride;
  long long nmat;
  double* inv;
  const double* vec;
  const int* occ;  // [nds][n] or nullptr
  const uint8_t* mask;
  double* ratio;  // optional [nmat]
  double* dsign;  // optional [nmat]
  double* dlog;   // optional [nmat]
};

__device__ __forceinline__ double sgn(double x) { return x > 0.0 ? 1.0 : (x < 0.0 ? -1.0 : 0.0); }

// Sherman-Morrison row replacement of one NN x NN matrix held in registers (n <= 8): v = the new row (orbital
// values through the occupation list), e = the replaced row's index.  Returns the determinant ratio.
template <int NN>
__device__ __forceinline__ double sm_thread_apply(double* __restrict__ inv, const double (&v)[NN], int e) {
  double A[NN][NN], t[NN];
  if (NN % 2 == 0) {
    const double2* __restrict__ p = reinterpret_cast<const double2*>(inv);
#pragma unroll
    for (int i = 0; i < NN * NN / 2; ++i) {
      const double2 x = p[i];
      A[(2 * i) / NN][(2 * i) % NN] = x.x;
      A[(2 * i + 1) / NN][(2 * i + 1) % NN] = x.y;
    }
  } else {
#pragma unroll
    for (int i = 0; i < NN * NN; ++i) A[i / NN][i % NN] = inv[i];
  }
#pragma unroll
  for (int j = 0; j < NN; ++j) {
    double s = 0.0;
#pragma unroll
    for (int k = 0; k < NN; ++k) s = fma(v[k], A[k][j], s);
    t[j] = s;
  }
  double ratio = 0.0;
#pragma unroll
  for (int j = 0; j < NN; ++j)
    if (j == e) ratio = t[j];
#pragma unroll
  for (int k = 0; k < NN; ++k) {
    double ake = 0.0;
#pragma unroll
    for (int j = 0; j < NN; ++j)
      if (j == e) ake = A[k][j];
    const double col = ake / ratio;
#pragma unroll
    for (int j = 0; j < NN; ++j) A[k][j] = (j == e) ? col : fma(-col, t[j], A[k][j]);
  }
  if (NN % 2 == 0) {
    double2* __restrict__ p = reinterpret_cast<double2*>(inv);
#pragma unroll
    for (int i = 0; i < NN * NN / 2; ++i)
      p[i] = make_double2(A[(2 * i) / NN][(2 * i) % NN], A[(2 * i + 1) / NN][(2 * i + 1) % NN]);
  } else {
#pragma unroll
    for (int i = 0; i < NN * NN; ++i) inv[i] = A[i / NN][i % NN];
  }
  return ratio;
}

// one thread per matrix, matrix in registers (n <= 8)
template <int NN>
__global__ void __launch_bounds__(128) k_sm_thread(const SmArgs a) {
  const long long m = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (m >= a.nmat) return;
  const long long w = m / a.nds;
  const int d = (int)(m - w * a.nds);
  if (a.mask && !a.mask[w]) return;
  double v[NN];
  const double* __restrict__ vb = a.vec + w * a.vec_stride;
#pragma unroll
  for (int k = 0; k < NN; ++k) v[k] = a.occ ? vb[a.occ[d * NN + k]] : vb[d * NN + k];
  const double ratio = sm_thread_apply<NN>(a.inv + m * (NN * NN), v, a.e);
  if (a.ratio) a.ratio[m] = ratio;
  if (a.dsign) a.dsign[m] *= sgn(ratio);
  if (a.dlog) a.dlog[m] += log(fabs(ratio));
}

// Multi-determinant update of one accepted move in ONE launch (slater.py:262-291 + the caches of
// determinant_tools.py:74-88): one warp per walker, lanes over the unique determinants of the moved spin -- each
// lane applies the register Sherman-Morrison update to its determinants -- then the same warp refreshes dv of that
// spin and W of the other one (det_cache_warp).  Replaces k_sm_thread + k_det_cache for n <= 8.
template <int NN>
__global__ void __launch_bounds__(128) k_det_update(const Sys S, const State st, int s, int e, const uint8_t* mask) {
  const int lane = threadIdx.x & 31;
  const int w = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (w >= st.N) return;
  if (mask && !mask[w]) return;
  const int nds = S.nds[s];
  const int* __restrict__ occ = S.iblob + S.o_occ[s];
  const double* __restrict__ vb = st.saved_mo + (size_t)w * S.ldc[s];
  for (int d = lane; d < nds; d += 32) {
    const size_t m = (size_t)w * nds + d;
    double v[NN];
#pragma unroll
    for (int k = 0; k < NN; ++k) v[k] = vb[occ[d * NN + k]];
    const double ratio = sm_thread_apply<NN>(st.inv[s] + m * (NN * NN), v, e);
    st.dsign[s][m] *= sgn(ratio);
    st.dlog[s][m] += log(fabs(ratio));
  }
  __syncwarp();
  det_cache_warp(S, st, w, lane, s);
}

// one warp per matrix, lane j owns column j (8 < n <= 32); rows are read/written coalesced and
// the matrix stays in registers between the two passes, so HBM sees exactly one read and one
// write of the inverse.
template <int NPAD>
__global__ void __launch_bounds__(256) k_sm_warp(const SmArgs a) {
  const int lane = threadIdx.x & 31;
  const long long m = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (m >= a.nmat) return;
  const long long w = m / a.nds;
  const int d = (int)(m - w * a.nds);
  if (a.mask && !a.mask[w]) return;
  const int n = a.n;
  double* inv = a.inv + m * (long long)n * n;
  const double* __restrict__ vb = a.vec + w * a.vec_stride;
  double vk = 0.0;
  if (lane < n) vk = a.occ ? vb[a.occ[d * n + lane]] : vb[(long long)d * n + lane];
  const double ratio = sm_warp_apply<NPAD>(inv, n, a.e, lane, vk);
  if (lane == 0) {
    if (a.ratio) a.ratio[m] = ratio;
    if (a.dsign) a.dsign[m] *= sgn(ratio);
    if (a.dlog) a.dlog[m] += log(fabs(ratio));
  }
}

// n = 32 with the matrices staged by the bulk-copy (TMA) engine: every warp owns two 8 KB shared-memory slots; one
// elected lane issues cp.async.bulk global -> shared for the NEXT matrix (mbarrier complete_tx) while the warp updates
// the current one in place in shared memory (lane j = column j, conflict-free rows), then hands the slot back to the
// engine with one cp.async.bulk shared -> global.  No matrix element passes through a register file on its way in or
// out, so 12 warps per SM keep 24 matrices (192 KB) in flight instead of the register-limited 7 warps of k_sm_warp<32>.
// Same arithmetic and operation order as sm_warp_apply.  Persistent grid: warps stride over the matrices.
#define QMCB_SM_TMA_WARPS 4
__global__ void __launch_bounds__(QMCB_SM_TMA_WARPS * 32) k_sm_tma32(const SmArgs a) {
  constexpr int N = 32, BYTES = N * N * 8;
  extern __shared__ __align__(128) unsigned char qmcb_smem[];
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  double* slot0 = reinterpret_cast<double*>(qmcb_smem) + (size_t)wib * 2 * N * N;
  uint64_t* bars = reinterpret_cast<uint64_t*>(qmcb_smem + (size_t)QMCB_SM_TMA_WARPS * 2 * BYTES) + wib * 2;
  const long long nwarps = (long long)gridDim.x * QMCB_SM_TMA_WARPS;
  const long long first = (long long)blockIdx.x * QMCB_SM_TMA_WARPS + wib;
  if (lane == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bars)), "r"(1));
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bars + 1)), "r"(1));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncwarp();
  auto active = [&](long long m) { return m < a.nmat && !(a.mask && !a.mask[m / a.nds]); };
  auto issue_load = [&](long long m, int s) {  // lane 0 only
    const uint32_t mb = smem_u32(bars + s);
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(mb), "r"(BYTES) : "memory");
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_u32(slot0 + (size_t)s * N * N)),
                 "l"(a.inv + m * (long long)(N * N)), "r"(BYTES), "r"(mb)
                 : "memory");
  };
  // next matrix this warp really updates, at or after m
  auto next_active = [&](long long m) {
    while (m < a.nmat && !active(m)) m += nwarps;
    return m;
  };
  long long cur = next_active(first);
  unsigned phase[2] = {0u, 0u};
  int s = 0;
  if (lane == 0 && cur < a.nmat) issue_load(cur, 0);
  while (cur < a.nmat) {
    const long long nxt = next_active(cur + nwarps);
    if (lane == 0 && nxt < a.nmat) {
      // the other slot was handed to the engine one iteration ago: wait until it has been READ before refilling it
      asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
      issue_load(nxt, s ^ 1);
    }
    const long long w = cur / a.nds;
    const int d = (int)(cur - w * a.nds);
    const double* __restrict__ vb = a.vec + w * a.vec_stride;
    const double vk = a.occ ? vb[a.occ[d * N + lane]] : vb[(long long)d * N + lane];
    {  // wait for this slot's bytes
      const uint32_t mb = smem_u32(bars + s);
      uint32_t done = 0;
      while (!done) {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(done)
            : "r"(mb), "r"(phase[s])
            : "memory");
      }
      phase[s] ^= 1u;
    }
    double* __restrict__ A = slot0 + (size_t)s * N * N;
    double t = 0.0;
#pragma unroll
    for (int k = 0; k < N; ++k) t = fma(__shfl_sync(0xffffffffu, vk, k), A[k * N + lane], t);
    const double ratio = __shfl_sync(0xffffffffu, t, a.e);
    const double col = A[lane * N + a.e] / ratio;  // lane k: inv[k][e] / ratio
    __syncwarp();
#pragma unroll
    for (int k = 0; k < N; ++k) {
      const double ck = __shfl_sync(0xffffffffu, col, k);
      A[k * N + lane] = (lane == a.e) ? ck : fma(-ck, t, A[k * N + lane]);
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // generic-proxy writes -> visible to the copy engine
    __syncwarp();
    if (lane == 0) {
      asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(a.inv + cur * (long long)(N * N)),
                   "r"(smem_u32(A)), "r"(BYTES)
                   : "memory");
      asm volatile("cp.async.bulk.commit_group;" ::: "memory");
      if (a.ratio) a.ratio[cur] = ratio;
      if (a.dsign) a.dsign[cur] *= sgn(ratio);
      if (a.dlog) a.dlog[cur] += log(fabs(ratio));
    }
    cur = nxt;
    s ^= 1;
  }
  if (lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");  // shared memory must outlive the last stores
}

// =========================================================================================
// Device-resident VMC move of electron e (mc.py:115-137): drift at the old position, proposal,
// drift + ratio at the new position, Metropolis test.  Saves MO row / new position for the
// update kernels and writes the accept mask.
// =========================================================================================
struct MoveArgs {
  int e;
  double tstep;
  const double* gauss;  // [N][3] for this (step, electron)
  const double* unif;   // [N]
  uint8_t* accept;      // [N]
  unsigned long long* nacc;
  double* scr;
  size_t scr_stride;
  double* r2prop;  // DMC: [N] sum over electrons of |gauss + drift|^2 (dmc.py:68, 190-191)
  double* r2acc;   // DMC: [N] the same for accepted moves
};

__device__ __forceinline__ void limdrift_dmc(double (&g)[3], double tau);

__device__ __forceinline__ void limdrift3(double (&g)[3]) {
  // mc.py:76-89 with cutoff = 1; np.linalg.norm = sqrt(sum of squares)
  const double tot = sqrt(__dadd_rn(__dadd_rn(__dmul_rn(g[0], g[0]), __dmul_rn(g[1], g[1])), __dmul_rn(g[2], g[2])));
  if (tot > 1.0) {
    g[0] = g[0] / tot;
    g[1] = g[1] / tot;
    g[2] = g[2] / tot;
  }
}

template <int NMOT>
__global__ void __launch_bounds__(128) k_vmc_move(const Sys S, const State st, const MoveArgs ma) {
  const double* sd;
  const int* si;
  stage_tables(S, sd, si);
  const int w = blockIdx.x * blockDim.x + threadIdx.x;
  const int N = st.N;
  const int which = (S.nmo[0] + S.nmo[1] > 0 ? QMCB_SLATER : 0) | ((S.na + S.nb) > 0 ? QMCB_JASTROW : 0) |
                    ((S.na3 + S.nb3) > 0 ? QMCB_JASTROW3 : 0);
  bool acc = false;
  if (w < N) {
    const int e = ma.e;
    const int s = e >= S.nup ? 1 : 0;
    const double ox = CONF(st, S, w, e, 0), oy = CONF(st, S, w, e, 1),
                 oz = CONF(st, S, w, e, 2);
    PointEval<1, NMOT> ev;
    ev.run(S, sd, si, st, which, w, e, ox, oy, oz, nullptr, ma.scr + w, ma.scr_stride);
    double grad[3];
#pragma unroll
    for (int i = 0; i < 3; ++i) {
      double gs = 0.0;
      if (which & QMCB_SLATER) {
        gs = ev.rat[1 + i] / ev.rat[0];
        if (!isfinite(gs)) gs = 0.0;
      }
      grad[i] = gs + ev.gj[i];
    }
    limdrift3(grad);
    double gauss[3], np_[3];
#pragma unroll
    for (int i = 0; i < 3; ++i) gauss[i] = ma.gauss[(size_t)w * 3 + i];
    // newcoorde = configs[:, e] + gauss + grad * tstep   (no FMA contraction: numpy rounds each op)
    np_[0] = __dadd_rn(__dadd_rn(ox, gauss[0]), __dmul_rn(grad[0], ma.tstep));
    np_[1] = __dadd_rn(__dadd_rn(oy, gauss[1]), __dmul_rn(grad[1], ma.tstep));
    np_[2] = __dadd_rn(__dadd_rn(oz, gauss[2]), __dmul_rn(grad[2], ma.tstep));
    double* mo_save = (which & QMCB_SLATER) ? st.saved_mo + (size_t)w * S.ldc[s] : nullptr;
    ev.run(S, sd, si, st, which, w, e, np_[0], np_[1], np_[2], mo_save, ma.scr + w, ma.scr_stride);
    double ngrad[3];
    double val = 1.0;
#pragma unroll
    for (int i = 0; i < 3; ++i) {
      double gs = 0.0;
      if (which & QMCB_SLATER) {
        gs = ev.rat[1 + i] / ev.rat[0];
        if (!isfinite(gs)) gs = 0.0;
      }
      ngrad[i] = gs + ev.gj[i];
    }
    if (which & QMCB_SLATER) {
      val = ev.rat[0];
      if (!isfinite(val)) val = 1.0;
    }
    val = val * exp(ev.du);
    limdrift3(ngrad);
    double fwd = 0.0, bwd = 0.0;
#pragma unroll
    for (int i = 0; i < 3; ++i) {
      fwd = __dadd_rn(fwd, __dmul_rn(gauss[i], gauss[i]));
      const double b = __dadd_rn(gauss[i], __dmul_rn(ma.tstep, __dadd_rn(grad[i], ngrad[i])));
      bwd = __dadd_rn(bwd, __dmul_rn(b, b));
    }
    const double tprob = exp(__dmul_rn(1.0 / (2.0 * ma.tstep), __dadd_rn(fwd, -bwd)));
    const double aval = fabs(val);
    const double ratio = __dmul_rn(__dmul_rn(aval, aval), tprob);
    acc = ratio > ma.unif[w];
    ma.accept[w] = acc ? 1 : 0;
    st.saved_pos[(size_t)w * 3] = np_[0];
    st.saved_pos[(size_t)w * 3 + 1] = np_[1];
    st.saved_pos[(size_t)w * 3 + 2] = np_[2];
  }
  const unsigned b = __ballot_sync(0xffffffffu, acc);
  if ((threadIdx.x & 31) == 0 && b) atomicAdd(ma.nacc, (unsigned long long)__popc(b));
}

// =========================================================================================
// Cooperative forms for multi-determinant and three-body wave functions (G lanes per walker).
// =========================================================================================
// three-body terms of electron e at (px,py,pz): a-values by lanes over (atom, function) into shared
// memory, pair terms by lanes over partners (three_body_jastrow.py:454-655); adds to du / g / lap.
// per-walker scratch of coop_jastrow3 (doubles): a-values / gradient / Laplacian factors of the moved electron
// [3][natom * na3], then per partner its displacement, in-range flag and the b radial values / derivatives
__host__ __device__ inline int j3_scratch_doubles(const Sys& S) {
  return 3 * S.natom * S.na3 + (S.ne > 1 ? S.ne - 1 : 0) * (4 + 3 * S.nb3);
}

// Three-body terms of electron e at (px, py, pz) with G lanes per walker, in three phases through shared memory:
//   1  lanes over (atom, k): a_k(r_eI) and its derivative factors                   (three_body_jastrow.py:454-520)
//   2  lanes over partners j: minimal-image displacement r_ej, b_m(r_ej) and derivative factors
//   3  lanes over (partner, atom, m) tasks: sum_{kl} C[I,k,l,m,sp] a_k(r_eI) a_l(r_jI) b_m(r_ej) and its gradient /
//      Laplacian terms (521-655) -- 84 tasks for H2O, so all lanes work, where one lane per partner left 9 of 16 idle
template <int WANT, int G>
__device__ __forceinline__ void coop_jastrow3(const Sys& S, const double* __restrict__ sd, const int* __restrict__ si,
                                              const State& st, int w, int e, double px, double py, double pz, int lane,
                                              unsigned gm, double* __restrict__ abuf, double& du, double (&g)[3],
                                              double& lap) {
  const int na_tot = S.natom * S.na3;
  const int na = S.na3, nb = S.nb3, pstride = 4 + 3 * S.nb3;
  double* __restrict__ av = abuf;
  double* __restrict__ ag = abuf + na_tot;
  double* __restrict__ al = abuf + 2 * na_tot;
  double* __restrict__ pb = abuf + 3 * na_tot;
  for (int t = lane; t < na_tot; t += G) {
    const int I = t / S.na3, k = t - I * S.na3;
    double dx = px - sd[S.o_xyz + 3 * I], dy = py - sd[S.o_xyz + 3 * I + 1], dz = pz - sd[S.o_xyz + 3 * I + 2];
    if (S.pbc) min_image(S, sd, dx, dy, dz);
    const double r = sqrt(dx * dx + dy * dy + dz * dz);
    double v = 0.0, gg = 0.0, ll = 0.0;
    if (r < S.rcut_a3) radial_ool<WANT>(si[S.o_a3kind + k], sd[S.o_a3par + k], S.rcut_a3, r, v, gg, ll);
    av[t] = v;
    ag[t] = gg;
    al[t] = ll;
  }
  for (int jj = lane; jj < S.ne - 1; jj += G) {
    const int j = jj < e ? jj : jj + 1;
    double dx = px - CONF(st, S, w, j, 0), dy = py - CONF(st, S, w, j, 1), dz = pz - CONF(st, S, w, j, 2);
    if (S.pbc) min_image(S, sd, dx, dy, dz);
    const double r = sqrt(dx * dx + dy * dy + dz * dz);
    double* __restrict__ q = pb + jj * pstride;
    q[0] = dx;
    q[1] = dy;
    q[2] = dz;
    q[3] = (r < S.rcut_b3) ? 1.0 : 0.0;
    if (r < S.rcut_b3)
      for (int m = 0; m < nb; ++m) {
        double v, gg = 0.0, ll = 0.0;
        radial_ool<WANT>(si[S.o_b3kind + m], sd[S.o_b3par + m], S.rcut_b3, r, v, gg, ll);
        q[4 + m] = v;
        q[4 + nb + m] = gg;
        q[4 + 2 * nb + m] = ll;
      }
  }
  __syncwarp(gm);
  double P = 0.0, gl[3] = {0.0, 0.0, 0.0}, lp = 0.0;
  const int per_partner = S.natom * nb;
#pragma unroll 1
  for (int t = lane; t < (S.ne - 1) * per_partner; t += G) {
    const int jj = t / per_partner, rem = t - jj * per_partner;
    const int I = rem / nb, m = rem - I * nb;
    const double* __restrict__ q = pb + jj * pstride;
    if (q[3] == 0.0) continue;
    const int j = jj < e ? jj : jj + 1;
    const int sp = (e >= S.nup ? 1 : 0) + (j >= S.nup ? 1 : 0);
    const double* __restrict__ C = sd + S.o_c3 + (size_t)I * na * na * nb * 3 + sp;
    double s0 = 0.0, s1 = 0.0, s2 = 0.0;
    for (int k = 0; k < na; ++k) {
      double tt = 0.0;  // sum_l C[I,k,l,m,sp] a_l(r_jI)
      for (int l = 0; l < na; ++l) tt = fma(C[((k * na + l) * nb + m) * 3], A3V(st, S, w, j, I, l), tt);
      s0 = fma(av[I * na + k], tt, s0);
      if (WANT >= 1) s1 = fma(ag[I * na + k], tt, s1);
      if (WANT >= 2) s2 = fma(al[I * na + k], tt, s2);
    }
    const double bv = q[4 + m];
    P = fma(s0, bv, P);
    if (WANT >= 1) {
      double ax = px - sd[S.o_xyz + 3 * I], ay = py - sd[S.o_xyz + 3 * I + 1], az = pz - sd[S.o_xyz + 3 * I + 2];
      if (S.pbc) min_image(S, sd, ax, ay, az);
      const double bg = q[4 + nb + m];
      const double ca = s1 * bv, cb = s0 * bg;
      gl[0] += ca * ax + cb * q[0];
      gl[1] += ca * ay + cb * q[1];
      gl[2] += ca * az + cb * q[2];
      if (WANT >= 2) lp += s2 * bv + 2.0 * s1 * bg * (ax * q[0] + ay * q[1] + az * q[2]) + s0 * q[4 + 2 * nb + m];
    }
  }
  P = group_sum<G>(P, gm);
#pragma unroll
  for (int i = 0; i < 3; ++i) g[i] += group_sum<G>(gl[i], gm);
  lap += group_sum<G>(lp, gm);
  if (WANT != 2) du += P - st.P3[(size_t)w * S.ne + e];
  __syncwarp(gm);
}

// Slater ratios (value, d/dx, d/dy, d/dz) of electron e from MO rows rows[c * ld + orbital], lanes over
// the unique spin determinants (slater.py:301-380, determinant_tools.py:74-88)
template <int G>
__device__ __forceinline__ void coop_det_ratio4(const Sys& S, const int* __restrict__ si, const State& st, int w, int s,
                                                int eeff, const double* __restrict__ rows, int ld, int lane,
                                                unsigned gm, double (&rat)[4]) {
  const int n = s ? S.ndn : S.nup;
  const int nds = S.nds[s];
  const int* __restrict__ occ = si + S.o_occ[s];
  double num[4] = {0.0, 0.0, 0.0, 0.0}, den = 0.0;
  for (int d = lane; d < nds; d += G) {
    const double* __restrict__ inv = st.inv[s] + ((size_t)w * nds + d) * n * n + eeff;
    double r[4] = {0.0, 0.0, 0.0, 0.0};
    for (int k = 0; k < n; ++k) {
      const double a = inv[k * n];
      const int orb = occ[d * n + k];
#pragma unroll
      for (int c = 0; c < 4; ++c) r[c] = fma(rows[c * ld + orb], a, r[c]);
    }
    const double wgt = S.ndet == 1 ? 1.0 : st.dv[s][(size_t)w * nds + d] * st.W[s][(size_t)w * nds + d];
    den += wgt;
#pragma unroll
    for (int c = 0; c < 4; ++c) num[c] = fma(r[c], wgt, num[c]);
  }
  den = group_sum<G>(den, gm);
#pragma unroll
  for (int c = 0; c < 4; ++c) rat[c] = group_sum<G>(num[c], gm) / den;
}

// mc.py:115-137 for electron e with G lanes per walker (general wave functions: any number of
// determinants, optional two- and three-body Jastrow factors).  The drift at the current position comes
// from the cached MO rows; accepted walkers refresh their cached rows (value, gradient, Laplacian) from
// the evaluation at the proposed position.  Saves the value row / position for the update kernels.
// DMC = true: the drift-diffusion move of dmc.py:49-70 (Umrigar drift limit, fixed-node rejection, r^2 bookkeeping),
// the same arithmetic as k_vmc_sweep<G, true>
template <int G, bool DMC = false>
__global__ void __launch_bounds__(128) k_vmc_move_coop(const Sys S, const State st, const MoveArgs ma) {
  const double* sd;
  const int* si;
  stage_tables(S, sd, si);
  extern __shared__ __align__(128) unsigned char qmcb_smem[];
  const CoopLayout L = coop_layout(S);
  const int lane32 = threadIdx.x & 31;
  const int lane = lane32 & (G - 1);
  const unsigned gm = group_mask<G>(lane32);
  const int slot = threadIdx.x / G;
  const int w = (blockIdx.x * blockDim.x + threadIdx.x) / G;
  const int N = st.N;
  if (w >= N) return;
  const size_t tab = (16 + (size_t)S.dwords * 8 + (size_t)S.iwords * 4 + 15) & ~(size_t)15;
  const int j3n = j3_scratch_doubles(S);
  double* ws = reinterpret_cast<double*>(qmcb_smem + tab) + (size_t)slot * (L.total + j3n);
  double* abuf = ws + L.total;
  const bool has_s = S.nmo[0] + S.nmo[1] > 0, has_j = (S.na + S.nb) > 0, has_j3 = (S.na3 + S.nb3) > 0;
  const int ldmax = S.ldc[0] > S.ldc[1] ? S.ldc[0] : S.ldc[1];
  const int e = ma.e;
  const int s = e >= S.nup ? 1 : 0;
  const int eeff = e - s * S.nup;
  const double ox = CONF(st, S, w, e, 0), oy = CONF(st, S, w, e, 1), oz = CONF(st, S, w, e, 2);
  double grad[3] = {0.0, 0.0, 0.0};
  if (has_s) {
    double r[4];
    coop_det_ratio4<G>(S, si, st, w, s, eeff, st.mocache + ((size_t)w * S.ne + e) * 5 * ldmax, ldmax, lane, gm, r);
#pragma unroll
    for (int i = 0; i < 3; ++i) {
      double gs = r[1 + i] / r[0];
      if (!isfinite(gs)) gs = 0.0;
      grad[i] = gs;
    }
  }
  {
    double du = 0.0, gj[3] = {0.0, 0.0, 0.0}, lj = 0.0;
    if (has_j) coop_jastrow<1, G>(S, sd, si, st, w, e, ox, oy, oz, lane, gm, du, gj, lj);
    if (has_j3) coop_jastrow3<1, G>(S, sd, si, st, w, e, ox, oy, oz, lane, gm, abuf, du, gj, lj);
#pragma unroll
    for (int i = 0; i < 3; ++i) grad[i] = grad[i] + gj[i];
  }
  double gauss[3], np_[3];
#pragma unroll
  for (int i = 0; i < 3; ++i) gauss[i] = ma.gauss[(size_t)w * 3 + i];
  if (DMC) {
    limdrift_dmc(grad, ma.tstep);
    np_[0] = __dadd_rn(__dadd_rn(ox, gauss[0]), grad[0]);
    np_[1] = __dadd_rn(__dadd_rn(oy, gauss[1]), grad[1]);
    np_[2] = __dadd_rn(__dadd_rn(oz, gauss[2]), grad[2]);
  } else {
    limdrift3(grad);
    np_[0] = __dadd_rn(__dadd_rn(ox, gauss[0]), __dmul_rn(grad[0], ma.tstep));
    np_[1] = __dadd_rn(__dadd_rn(oy, gauss[1]), __dmul_rn(grad[1], ma.tstep));
    np_[2] = __dadd_rn(__dadd_rn(oz, gauss[2]), __dmul_rn(grad[2], ma.tstep));
  }
  double ngrad[3] = {0.0, 0.0, 0.0}, val = 1.0;
  if (has_s) {
    coop_eval_mo<2, G>(S, L, sd, si, s, np_[0], np_[1], np_[2], ws, lane, gm);
    double r[4];
    coop_det_ratio4<G>(S, si, st, w, s, eeff, ws + L.mo, ldmax, lane, gm, r);
#pragma unroll
    for (int i = 0; i < 3; ++i) {
      double gs = r[1 + i] / r[0];
      if (!isfinite(gs)) gs = 0.0;
      ngrad[i] = gs;
    }
    val = isfinite(r[0]) ? r[0] : 1.0;
  }
  {
    double du = 0.0, gj[3] = {0.0, 0.0, 0.0}, lj = 0.0;
    if (has_j) coop_jastrow<1, G>(S, sd, si, st, w, e, np_[0], np_[1], np_[2], lane, gm, du, gj, lj);
    if (has_j3) coop_jastrow3<1, G>(S, sd, si, st, w, e, np_[0], np_[1], np_[2], lane, gm, abuf, du, gj, lj);
#pragma unroll
    for (int i = 0; i < 3; ++i) ngrad[i] = ngrad[i] + gj[i];
    val = val * exp(du);
  }
  if (DMC)
    limdrift_dmc(ngrad, ma.tstep);
  else
    limdrift3(ngrad);
  double fwd = 0.0, bwd = 0.0, r2 = 0.0;
#pragma unroll
  for (int i = 0; i < 3; ++i) {
    fwd = __dadd_rn(fwd, __dmul_rn(gauss[i], gauss[i]));
    double b;
    if (DMC) {
      const double gd = __dadd_rn(gauss[i], grad[i]);
      r2 = __dadd_rn(r2, __dmul_rn(gd, gd));
      b = __dadd_rn(gd, ngrad[i]);
    } else {
      b = __dadd_rn(gauss[i], __dmul_rn(ma.tstep, __dadd_rn(grad[i], ngrad[i])));
    }
    bwd = __dadd_rn(bwd, __dmul_rn(b, b));
  }
  const double tprob = exp(__dmul_rn(1.0 / (2.0 * ma.tstep), __dadd_rn(fwd, -bwd)));
  const double aval = fabs(val);
  double ratio = __dmul_rn(__dmul_rn(aval, aval), tprob);
  if (DMC) ratio = __dmul_rn(ratio, val > 0.0 ? 1.0 : (val < 0.0 ? -1.0 : 0.0));  // fixed node (dmc.py:65-66)
  const bool acc = __shfl_sync(gm, (ratio > ma.unif[w]) ? 1 : 0, 0, G) != 0;
  if (lane == 0) {
    ma.accept[w] = acc ? 1 : 0;
    if (acc) atomicAdd(ma.nacc, 1ULL);
    if (DMC) {
      ma.r2prop[w] = __dadd_rn(ma.r2prop[w], r2);
      if (acc) ma.r2acc[w] = __dadd_rn(ma.r2acc[w], r2);
    }
    st.saved_pos[(size_t)w * 3] = np_[0];
    st.saved_pos[(size_t)w * 3 + 1] = np_[1];
    st.saved_pos[(size_t)w * 3 + 2] = np_[2];
  }
  if (has_s) {
    const double* __restrict__ mo = ws + L.mo;
    double* __restrict__ sv = st.saved_mo + (size_t)w * S.ldc[s];
    for (int j = lane; j < S.ldc[s]; j += G) sv[j] = mo[j];
    if (acc) {
      double* __restrict__ mc = st.mocache + ((size_t)w * S.ne + e) * 5 * ldmax;
      for (int i = lane; i < 5 * ldmax; i += G) mc[i] = mo[i];
    }
  }
}

// Device-resident move of electron e for GENERAL periodic wave functions (multi-determinant and / or three-body
// factors on a supercell; mc.py:115-137 with make_irreducible, coord.py:168-194), G lanes per walker, in two launches
// around the lattice-summed orbital evaluation (k_pbc_mo* -> st.monew):
//   phase 0  drift at the current position (cached MO rows, lanes over the unique determinants; minimal-image
//            Jastrow; three-body factor), proposal wrapped into the simulation cell -> saved_pos / saved_wrap / gold
//   phase 1  ratio and drift at the proposed position from st.monew, Metropolis test, accept mask; the value row goes
//            to saved_mo for the update kernels, accepted walkers refresh their cached MO rows
// The internal updates (Sherman-Morrison of every determinant + dv / W, Jastrow and three-body caches, coordinates
// and wrap vectors) follow through launch_update, as for the open-boundary general path.
// DMC = true: the drift-diffusion move of dmc.py:49-70 (arithmetic of k_vmc_sweep<G, true>), and two more phases for
// the T-move of electron e (propose_tmoves + dmc.py:170-177) after k_tmove_select left the chosen position in saved_pos:
//   phase 3  the position wrapped into the cell TWICE, as the reference does (compute_tmoves wraps the candidates,
//            eval_ecp.py:117 / coord.py:168-184, and propose_tmoves wraps the selected one again, dmc.py:110) -- the
//            electron keeps its wrap vector plus whatever the second pass still finds (normally nothing)
//   phase 2  after the orbital kernel: value row -> saved_mo, accepted walkers refresh their cached MO rows
template <int G, bool DMC = false>
__global__ void __launch_bounds__(128) k_pbc_move_general(const Sys S, const State st, const MoveArgs ma, int phase) {
  const double* sd;
  const int* si;
  stage_tables(S, sd, si);
  extern __shared__ __align__(128) unsigned char qmcb_smem[];
  const int lane32 = threadIdx.x & 31;
  const int lane = lane32 & (G - 1);
  const unsigned gm = group_mask<G>(lane32);
  const int slot = threadIdx.x / G;
  const int w = (blockIdx.x * blockDim.x + threadIdx.x) / G;
  if (w >= st.N) return;
  const size_t tab = (16 + (size_t)S.dwords * 8 + (size_t)S.iwords * 4 + 15) & ~(size_t)15;
  double* abuf = reinterpret_cast<double*>(qmcb_smem + tab) + (size_t)slot * j3_scratch_doubles(S);
  const bool has_s = S.nmo[0] + S.nmo[1] > 0, has_j = (S.na + S.nb) > 0, has_j3 = (S.na3 + S.nb3) > 0;
  const int ldmax = S.ldc[0] > S.ldc[1] ? S.ldc[0] : S.ldc[1];
  const int e = ma.e;
  const int s = e >= S.nup ? 1 : 0;
  const int eeff = e - s * S.nup;
  if (DMC && phase == 3) {
    if (lane == 0) {
      double o1[3], w1[3], o2[3], w2[3];
      wrap_cell(sd + S.o_lat, sd + S.o_latinv, st.saved_pos[(size_t)w * 3], st.saved_pos[(size_t)w * 3 + 1],
                st.saved_pos[(size_t)w * 3 + 2], o1, w1);
      wrap_cell(sd + S.o_lat, sd + S.o_latinv, o1[0], o1[1], o1[2], o2, w2);
#pragma unroll
      for (int i = 0; i < 3; ++i) {
        st.saved_pos[(size_t)w * 3 + i] = o2[i];
        st.saved_wrap[(size_t)w * 3 + i] = st.wrap[((size_t)w * S.ne + e) * 3 + i] + w2[i];
      }
    }
    return;
  }
  if (DMC && phase == 2) {
    if (has_s && ma.accept[w]) {  // the orbital kernel ran for the accepted walkers only
      const double* __restrict__ rows2 = st.monew + (size_t)w * 5 * ldmax;
      double* __restrict__ sv = st.saved_mo + (size_t)w * S.ldc[s];
      for (int j = lane; j < S.ldc[s]; j += G) sv[j] = rows2[j];
      double* __restrict__ mc = st.mocache + ((size_t)w * S.ne + e) * 5 * ldmax;
      for (int i = lane; i < 5 * ldmax; i += G) mc[i] = rows2[i];
    }
    return;
  }
  const double px = phase == 0 ? CONF(st, S, w, e, 0) : st.saved_pos[(size_t)w * 3];
  const double py = phase == 0 ? CONF(st, S, w, e, 1) : st.saved_pos[(size_t)w * 3 + 1];
  const double pz = phase == 0 ? CONF(st, S, w, e, 2) : st.saved_pos[(size_t)w * 3 + 2];
  const double* __restrict__ rows = phase == 0 ? st.mocache + ((size_t)w * S.ne + e) * 5 * ldmax : st.monew + (size_t)w * 5 * ldmax;
  double grad[3] = {0.0, 0.0, 0.0}, val = 1.0;
  if (has_s) {
    double r[4];
    coop_det_ratio4<G>(S, si, st, w, s, eeff, rows, ldmax, lane, gm, r);
#pragma unroll
    for (int i = 0; i < 3; ++i) {
      double gs = r[1 + i] / r[0];
      if (!isfinite(gs)) gs = 0.0;
      grad[i] = gs;
    }
    val = isfinite(r[0]) ? r[0] : 1.0;
  }
  {
    double du = 0.0, gj[3] = {0.0, 0.0, 0.0}, lj = 0.0;
    if (has_j) coop_jastrow_pbc<1, G>(S, sd, si, st, w, e, px, py, pz, lane, gm, du, gj, lj);
    if (has_j3) coop_jastrow3<1, G>(S, sd, si, st, w, e, px, py, pz, lane, gm, abuf, du, gj, lj);
#pragma unroll
    for (int i = 0; i < 3; ++i) grad[i] = grad[i] + gj[i];
    val = val * exp(du);
  }
  if (DMC)
    limdrift_dmc(grad, ma.tstep);  // the limited drift already carries the (effective) time step
  else
    limdrift3(grad);
  const double* __restrict__ gauss = ma.gauss + (size_t)w * 3;
  if (phase == 0) {
    if (lane == 0) {
      const double nx = __dadd_rn(__dadd_rn(px, gauss[0]), DMC ? grad[0] : __dmul_rn(grad[0], ma.tstep));
      const double ny = __dadd_rn(__dadd_rn(py, gauss[1]), DMC ? grad[1] : __dmul_rn(grad[1], ma.tstep));
      const double nz = __dadd_rn(__dadd_rn(pz, gauss[2]), DMC ? grad[2] : __dmul_rn(grad[2], ma.tstep));
      double o[3], ww[3];
      wrap_cell(sd + S.o_lat, sd + S.o_latinv, nx, ny, nz, o, ww);
#pragma unroll
      for (int i = 0; i < 3; ++i) {
        st.saved_pos[(size_t)w * 3 + i] = o[i];
        st.saved_wrap[(size_t)w * 3 + i] = st.wrap[((size_t)w * S.ne + e) * 3 + i] + ww[i];
        st.gold[(size_t)w * 3 + i] = grad[i];
      }
    }
    return;
  }
  double fwd = 0.0, bwd = 0.0, r2 = 0.0;
#pragma unroll
  for (int i = 0; i < 3; ++i) {
    fwd = __dadd_rn(fwd, __dmul_rn(gauss[i], gauss[i]));
    double b;
    if (DMC) {
      const double gd = __dadd_rn(gauss[i], st.gold[(size_t)w * 3 + i]);
      r2 = __dadd_rn(r2, __dmul_rn(gd, gd));
      b = __dadd_rn(gd, grad[i]);
    } else {
      b = __dadd_rn(gauss[i], __dmul_rn(ma.tstep, __dadd_rn(st.gold[(size_t)w * 3 + i], grad[i])));
    }
    bwd = __dadd_rn(bwd, __dmul_rn(b, b));
  }
  const double tprob = exp(__dmul_rn(1.0 / (2.0 * ma.tstep), __dadd_rn(fwd, -bwd)));
  const double aval = fabs(val);
  double ratio = __dmul_rn(__dmul_rn(aval, aval), tprob);
  if (DMC) ratio = __dmul_rn(ratio, val > 0.0 ? 1.0 : (val < 0.0 ? -1.0 : 0.0));  // fixed node (dmc.py:65-66)
  const bool acc = __shfl_sync(gm, (ratio > ma.unif[w]) ? 1 : 0, 0, G) != 0;
  if (lane == 0) {
    ma.accept[w] = acc ? 1 : 0;
    if (acc) atomicAdd(ma.nacc, 1ULL);
    if (DMC) {
      ma.r2prop[w] = __dadd_rn(ma.r2prop[w], r2);
      if (acc) ma.r2acc[w] = __dadd_rn(ma.r2acc[w], r2);
    }
  }
  if (has_s) {
    double* __restrict__ sv = st.saved_mo + (size_t)w * S.ldc[s];
    for (int j = lane; j < S.ldc[s]; j += G) sv[j] = rows[j];
    if (acc) {
      double* __restrict__ mc = st.mocache + ((size_t)w * S.ne + e) * 5 * ldmax;
      for (int i = lane; i < 5 * ldmax; i += G) mc[i] = rows[i];
    }
  }
}

// per-walker scratch of k_jastrow3_update_coop (doubles): old and new a-values of the moved electron, per partner the
// b values at the old and at the new distance, per (partner, atom, function) task the old and new contraction
__host__ __device__ inline int j3_update_scratch_doubles(const Sys& S) {
  const int np = S.ne > 1 ? S.ne - 1 : 0;
  return 2 * S.natom * S.na3 + 2 * np * S.nb3 + 2 * np * S.natom * S.nb3;
}

// three-body cache update with G lanes per walker (three_body_jastrow.py:149-189): the pair terms of the moved
// electron with its cached a-values at the old position leave the partners' sums, the ones at the accepted position
// enter.  Phases through shared memory: (1) lanes over (atom, k): new a-values; (2) lanes over partners: b_m at the
// old and new distance; (3) lanes over (partner, atom, m) tasks: sum_l C a_l(r_jI) once, contracted with the old and
// the new a_k(r_eI); (4) lanes over partners: the sums over (atom, m) in the order of j3_pair.
template <int G>
__global__ void __launch_bounds__(128) k_jastrow3_update_coop(const Sys S, const State st, int e, int move_conf,
                                                              const uint8_t* mask) {
  const double* sd;
  const int* si;
  stage_tables(S, sd, si);
  extern __shared__ __align__(128) unsigned char qmcb_smem[];
  const int lane32 = threadIdx.x & 31;
  const int lane = lane32 & (G - 1);
  const unsigned gm = group_mask<G>(lane32);
  const int slot = threadIdx.x / G;
  const int w = (blockIdx.x * blockDim.x + threadIdx.x) / G;
  if (w >= st.N) return;
  if (mask && !mask[w]) return;
  const size_t tab = (16 + (size_t)S.dwords * 8 + (size_t)S.iwords * 4 + 15) & ~(size_t)15;
  const int na_tot = S.natom * S.na3, np = S.ne - 1, na = S.na3, nb = S.nb3, ntask = S.natom * nb;
  double* avo = reinterpret_cast<double*>(qmcb_smem + tab) + (size_t)slot * j3_update_scratch_doubles(S);
  double* avn = avo + na_tot;
  double* bo = avn + na_tot;        // [np][nb]
  double* bn = bo + np * nb;        // [np][nb]
  double* so = bn + np * nb;        // [np][natom * nb]
  double* sn = so + np * ntask;     // [np][natom * nb]
  const double nx = st.saved_pos[(size_t)w * 3], ny = st.saved_pos[(size_t)w * 3 + 1], nz = st.saved_pos[(size_t)w * 3 + 2];
  const double ox = CONF(st, S, w, e, 0), oy = CONF(st, S, w, e, 1), oz = CONF(st, S, w, e, 2);
  const size_t abase = ((size_t)w * S.ne + e) * na_tot;
  for (int t = lane; t < na_tot; t += G) {
    const int I = t / na, k = t - I * na;
    avo[t] = st.a3v[abase + t];
    double dx = nx - sd[S.o_xyz + 3 * I], dy = ny - sd[S.o_xyz + 3 * I + 1], dz = nz - sd[S.o_xyz + 3 * I + 2];
    if (S.pbc) min_image(S, sd, dx, dy, dz);
    const double r = sqrt(dx * dx + dy * dy + dz * dz);
    double v = 0.0, gg, ll;
    if (r < S.rcut_a3) radial_ool<0>(si[S.o_a3kind + k], sd[S.o_a3par + k], S.rcut_a3, r, v, gg, ll);
    avn[t] = v;
  }
  for (int jj = lane; jj < np; jj += G) {
    const int j = jj < e ? jj : jj + 1;
    const double jx = CONF(st, S, w, j, 0), jy = CONF(st, S, w, j, 1), jz = CONF(st, S, w, j, 2);
#pragma unroll
    for (int side = 0; side < 2; ++side) {
      double dx = (side ? nx : ox) - jx, dy = (side ? ny : oy) - jy, dz = (side ? nz : oz) - jz;
      if (S.pbc) min_image(S, sd, dx, dy, dz);
      const double r = sqrt(dx * dx + dy * dy + dz * dz);
      double* __restrict__ b = (side ? bn : bo) + jj * nb;
      for (int m = 0; m < nb; ++m) {
        double v = 0.0, gg, ll;
        if (r < S.rcut_b3) radial_ool<0>(si[S.o_b3kind + m], sd[S.o_b3par + m], S.rcut_b3, r, v, gg, ll);
        b[m] = v;  // 0 beyond the cutoff: the pair term vanishes, as the early return of j3_pair
      }
    }
  }
  __syncwarp(gm);
#pragma unroll 1
  for (int t = lane; t < np * ntask; t += G) {
    const int jj = t / ntask, rem = t - jj * ntask;
    const int I = rem / nb, m = rem - I * nb;
    const int j = jj < e ? jj : jj + 1;
    const int sp = (e >= S.nup ? 1 : 0) + (j >= S.nup ? 1 : 0);
    const double* __restrict__ C = sd + S.o_c3 + (size_t)I * na * na * nb * 3 + sp;
    double s_old = 0.0, s_new = 0.0;
    for (int k = 0; k < na; ++k) {
      double tt = 0.0;  // sum_l C[I,k,l,m,sp] a_l(r_jI)
      for (int l = 0; l < na; ++l) tt = fma(C[((k * na + l) * nb + m) * 3], A3V(st, S, w, j, I, l), tt);
      s_old = fma(avo[I * na + k], tt, s_old);
      s_new = fma(avn[I * na + k], tt, s_new);
    }
    so[t] = s_old;
    sn[t] = s_new;
  }
  __syncwarp(gm);
  double newval = 0.0;
  for (int jj = lane; jj < np; jj += G) {
    const int j = jj < e ? jj : jj + 1;
    double Po = 0.0, Pn = 0.0;
    for (int u = 0; u < ntask; ++u) {  // (atom, m) in the order of j3_pair
      const int m = u % nb;
      Po = fma(so[jj * ntask + u], bo[jj * nb + m], Po);
      Pn = fma(sn[jj * ntask + u], bn[jj * nb + m], Pn);
    }
    double pj = st.P3[(size_t)w * S.ne + j];
    pj -= Po;
    pj += Pn;
    st.P3[(size_t)w * S.ne + j] = pj;
    newval += Pn;
  }
  newval = group_sum<G>(newval, gm);
  if (lane == 0) {
    st.val3[w] += newval - st.P3[(size_t)w * S.ne + e];
    st.P3[(size_t)w * S.ne + e] = newval;
  }
  for (int i = lane; i < na_tot; i += G) st.a3v[abase + i] = avn[i];
  __syncwarp(gm);
  if (move_conf && lane == 0) {
    CONF(st, S, w, e, 0) = nx;
    CONF(st, S, w, e, 1) = ny;
    CONF(st, S, w, e, 2) = nz;
    if (S.pbc)
      for (int i = 0; i < 3; ++i) st.wrap[((size_t)w * S.ne + e) * 3 + i] = st.saved_wrap[(size_t)w * 3 + i];
  }
}

// =========================================================================================
// Device-resident sweep, warp per walker: ALL electrons of one VMC step in one launch
// (mc.py:115-137).  Single-determinant wave functions (any occupation, n_s <= 32 staged in
// shared memory).  Per electron: drift at the old position from the cached MO rows (no orbital
// evaluation), proposal, cooperative value/gradient/Laplacian evaluation at the new position,
// Metropolis test, and for accepted moves Sherman-Morrison + Jastrow cache update + refresh of
// the cached MO rows -- all inside the warp, state in L2-resident global memory.
// =========================================================================================
// builds the pair caches of the sweep kernel from the current walker positions: one thread per
// (walker, pair) and per (walker, electron)
__global__ void __launch_bounds__(128) k_pair_cache_build(const Sys S, const State st) {
  const double* sd;
  const int* si;
  stage_tables(S, sd, si);
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const int per = S.npair + S.ne;
  if (t >= (long long)st.N * per) return;
  const int w = (int)(t / per), r0 = (int)(t - (long long)w * per);
  if (r0 < S.npair) {
    int i = 0, rem = r0;
    while (rem >= S.ne - 1 - i) {
      rem -= S.ne - 1 - i;
      ++i;
    }
    const int j = i + 1 + rem;
    double dx = CONF(st, S, w, i, 0) - CONF(st, S, w, j, 0), dy = CONF(st, S, w, i, 1) - CONF(st, S, w, j, 1),
           dz = CONF(st, S, w, i, 2) - CONF(st, S, w, j, 2);
    if (S.pbc) min_image(S, sd, dx, dy, dz);  // periodic fused chain (k_pbc_accept<true>): same caches, minimal image
    const double r = sqrt(dx * dx + dy * dy + dz * dz);
    const int sp = (i >= S.nup ? 1 : 0) + (j >= S.nup ? 1 : 0);
    double gs = 0.0, ls = 0.0;
    for (int l = 0; l < S.nb; ++l) {
      double v = 0.0, gg = 0.0, ll = 0.0;
      if (r < S.rcut_b) radial_ool<2>(si[S.o_bkind + l], sd[S.o_bpar + l], S.rcut_b, r, v, gg, ll);
      BPAIR(st, S, w, r0, l) = v;
      gs += sd[S.o_bcoef + l * 3 + sp] * gg;
      ls += sd[S.o_bcoef + l * 3 + sp] * ll;
    }
    LPAIR(st, S, w, r0) = ls;
    GPAIR(st, S, w, r0, 0) = gs * dx;
    GPAIR(st, S, w, r0, 1) = gs * dy;
    GPAIR(st, S, w, r0, 2) = gs * dz;
  } else {
    const int e = r0 - S.npair;
    const int s = e >= S.nup ? 1 : 0;
    double g0 = 0.0, g1 = 0.0, g2 = 0.0, la = 0.0;
    for (int I = 0; I < S.natom; ++I) {
      double dx = CONF(st, S, w, e, 0) - sd[S.o_xyz + 3 * I], dy = CONF(st, S, w, e, 1) - sd[S.o_xyz + 3 * I + 1],
             dz = CONF(st, S, w, e, 2) - sd[S.o_xyz + 3 * I + 2];
      if (S.pbc) min_image(S, sd, dx, dy, dz);
      const double r = sqrt(dx * dx + dy * dy + dz * dz);
      if (!(r < S.rcut_a)) continue;
      for (int k2 = 0; k2 < S.na; ++k2) {
        double v, gg, ll;
        radial_ool<2>(si[S.o_akind + k2], sd[S.o_apar + k2], S.rcut_a, r, v, gg, ll);
        const double cg = sd[S.o_acoef + (I * S.na + k2) * 2 + s] * gg;
        g0 = fma(cg, dx, g0);
        g1 = fma(cg, dy, g1);
        g2 = fma(cg, dz, g2);
        la = fma(sd[S.o_acoef + (I * S.na + k2) * 2 + s], ll, la);
      }
    }
    ALAP(st, S, w, e) = la;
    AGRAD(st, S, w, e, 0) = g0;
    AGRAD(st, S, w, e, 1) = g1;
    AGRAD(st, S, w, e, 2) = g2;
  }
}

struct SweepArgs {
  double tstep;
  const double* gauss;  // [ne][N][3] for this step
  const double* unif;   // [ne][N]
  uint8_t* accept;      // [ne][N] or nullptr
  unsigned long long* nacc;  // [ne]
  double* ke_e;         // optional [ne][N]: kinetic-energy pieces of the final positions (energy.py:57-65) ...
  double* g2_e;         // ... and |grad ln Psi|^2, from the caches, so the energy pipeline skips k_kinetic
  double* r2prop;       // DMC: [N] sum over electrons of |gauss + drift|^2          (dmc.py:68, 190-191)
  double* r2acc;        // DMC: [N] the same for accepted moves
};

__device__ __forceinline__ void limdrift3(double (&g)[3]);

// Umrigar's drift limiter, result already multiplied by the (effective) time step (dmc.py:22-35)
__device__ __forceinline__ void limdrift_dmc(double (&g)[3], double tau) {
  const double acyrus = 0.5;
  const double v2 = __dadd_rn(__dadd_rn(__dmul_rn(g[0], g[0]), __dmul_rn(g[1], g[1])), __dmul_rn(g[2], g[2]));
  double taueff = tau;
  if (v2 > 1e-8) taueff = (sqrt(__dadd_rn(1.0, __dmul_rn(__dmul_rn(__dmul_rn(2.0, tau), acyrus), v2))) - 1.0) / __dmul_rn(acyrus, v2);
  g[0] = __dmul_rn(g[0], taueff);
  g[1] = __dmul_rn(g[1], taueff);
  g[2] = __dmul_rn(g[2], taueff);
}

// DMC = false: VMC move (mc.py:115-137).  DMC = true: drift-diffusion move of dmc.py:49-70 (Umrigar
// drift limit, fixed-node rejection of sign changes, |gauss + drift|^2 bookkeeping for tdamp).
template <int G, bool DMC>
__global__ void __launch_bounds__(128, 4) k_vmc_sweep(const Sys S, const State st, const SweepArgs a) {
  const double* sd;
  const int* si;
  stage_tables(S, sd, si);
  extern __shared__ __align__(128) unsigned char qmcb_smem[];
  const CoopLayout L = coop_layout(S);
  constexpr int GP = 32 / G;  // walkers per warp
  const int lane32 = threadIdx.x & 31, wib = threadIdx.x >> 5;
  const int lane = lane32 & (G - 1);  // lane within the walker's group
  const unsigned gm = group_mask<G>(lane32);
  const int slot = wib * GP + lane32 / G;
  const int w = blockIdx.x * ((blockDim.x >> 5) * GP) + slot;
  const int N = st.N;
  if (w >= N) return;
  const size_t tab = (16 + (size_t)S.dwords * 8 + (size_t)S.iwords * 4 + 15) & ~(size_t)15;
  double* ws = reinterpret_cast<double*>(qmcb_smem + tab) + (size_t)slot * L.total;
  const bool has_s = S.nmo[0] + S.nmo[1] > 0, has_j = (S.na + S.nb) > 0;
  const int ldmax = S.ldc[0] > S.ldc[1] ? S.ldc[0] : S.ldc[1];
#pragma unroll 1
  for (int e = 0; e < S.ne; ++e) {
    const int s = e >= S.nup ? 1 : 0;
    const int n = s ? S.ndn : S.nup;
    const int eeff = e - s * S.nup;
    const int* __restrict__ occ = si + S.o_occ[s];
    const double ox = CONF(st, S, w, e, 0), oy = CONF(st, S, w, e, 1), oz = CONF(st, S, w, e, 2);
    // ---- drift at the current position (cached MO rows . inverse column)
    double grad[3] = {0.0, 0.0, 0.0};
    if (has_s) {
      const double* __restrict__ mc = st.mocache + ((size_t)w * S.ne + e) * 5 * ldmax;
      const double* __restrict__ inv = st.inv[s] + (size_t)w * n * n + eeff;
      double r = 0.0;
      if (lane < 4)
        for (int k = 0; k < n; ++k) r = fma(mc[lane * ldmax + occ[k]], inv[k * n], r);
      const double r0 = __shfl_sync(gm, r, 0, G);
#pragma unroll
      for (int i = 0; i < 3; ++i) {
        double gs = __shfl_sync(gm, r, 1 + i, G) / r0;
        if (!isfinite(gs)) gs = 0.0;
        grad[i] = gs;
      }
    }
    if (has_j) {
      double gj[3];
      coop_jastrow_cached_grad<G>(S, st, w, e, lane, gm, gj);
#pragma unroll
      for (int i = 0; i < 3; ++i) grad[i] = grad[i] + gj[i];
    }
    double gauss[3], np_[3];
#pragma unroll
    for (int i = 0; i < 3; ++i) gauss[i] = a.gauss[((size_t)e * N + w) * 3 + i];
    if (DMC) {
      limdrift_dmc(grad, a.tstep);
      np_[0] = __dadd_rn(__dadd_rn(ox, gauss[0]), grad[0]);
      np_[1] = __dadd_rn(__dadd_rn(oy, gauss[1]), grad[1]);
      np_[2] = __dadd_rn(__dadd_rn(oz, gauss[2]), grad[2]);
    } else {
      limdrift3(grad);
      np_[0] = __dadd_rn(__dadd_rn(ox, gauss[0]), __dmul_rn(grad[0], a.tstep));
      np_[1] = __dadd_rn(__dadd_rn(oy, gauss[1]), __dmul_rn(grad[1], a.tstep));
      np_[2] = __dadd_rn(__dadd_rn(oz, gauss[2]), __dmul_rn(grad[2], a.tstep));
    }
    // ---- value + drift at the proposed position
    double ngrad[3] = {0.0, 0.0, 0.0}, val = 1.0;
    if (has_s) {
      coop_eval_mo<2, G>(S, L, sd, si, s, np_[0], np_[1], np_[2], ws, lane, gm);
      const double* __restrict__ mo = ws + L.mo;
      const double* __restrict__ inv = st.inv[s] + (size_t)w * n * n + eeff;
      double r = 0.0;
      if (lane < 4)
        for (int k = 0; k < n; ++k) r = fma(mo[lane * ldmax + occ[k]], inv[k * n], r);
      const double r0 = __shfl_sync(gm, r, 0, G);
#pragma unroll
      for (int i = 0; i < 3; ++i) {
        double gs = __shfl_sync(gm, r, 1 + i, G) / r0;
        if (!isfinite(gs)) gs = 0.0;
        ngrad[i] = gs;
      }
      val = isfinite(r0) ? r0 : 1.0;
    }
    double ga[3] = {0.0, 0.0, 0.0}, la = 0.0;
    if (has_j) {
      double du, gj[3];
      coop_jastrow_propose<G>(S, sd, si, st, w, e, np_[0], np_[1], np_[2], lane, gm, ws + L.jtmp, du, gj, ga, la);
#pragma unroll
      for (int i = 0; i < 3; ++i) ngrad[i] = ngrad[i] + gj[i];
      val = val * exp(du);
    }
    if (DMC)
      limdrift_dmc(ngrad, a.tstep);
    else
      limdrift3(ngrad);
    double fwd = 0.0, bwd = 0.0, r2 = 0.0;
#pragma unroll
    for (int i = 0; i < 3; ++i) {
      fwd = __dadd_rn(fwd, __dmul_rn(gauss[i], gauss[i]));
      double b;
      if (DMC) {
        const double gd = __dadd_rn(gauss[i], grad[i]);
        r2 = __dadd_rn(r2, __dmul_rn(gd, gd));
        b = __dadd_rn(gd, ngrad[i]);
      } else {
        b = __dadd_rn(gauss[i], __dmul_rn(a.tstep, __dadd_rn(grad[i], ngrad[i])));
      }
      bwd = __dadd_rn(bwd, __dmul_rn(b, b));
    }
    const double tprob = exp(__dmul_rn(1.0 / (2.0 * a.tstep), __dadd_rn(fwd, -bwd)));
    const double aval = fabs(val);
    double ratio = __dmul_rn(__dmul_rn(aval, aval), tprob);
    if (DMC) ratio = __dmul_rn(ratio, val > 0.0 ? 1.0 : (val < 0.0 ? -1.0 : 0.0));  // fixed node (dmc.py:65-66)
    // every lane of the group evaluated the same numbers; take lane 0's decision
    const bool acc = __shfl_sync(gm, (ratio > a.unif[(size_t)e * N + w]) ? 1 : 0, 0, G) != 0;
    if (lane == 0) {
      if (a.accept) a.accept[(size_t)e * N + w] = acc ? 1 : 0;
      if (acc) atomicAdd(a.nacc + e, 1ULL);
      if (DMC) {
        a.r2prop[w] = __dadd_rn(a.r2prop[w], r2);
        if (acc) a.r2acc[w] = __dadd_rn(a.r2acc[w], r2);
      }
    }
    if (acc) {
      if (has_s) {
        coop_sherman_morrison<G>(S, L, si, st, w, s, eeff, ws, lane, gm);
        double* __restrict__ mc = st.mocache + ((size_t)w * S.ne + e) * 5 * ldmax;
        const double* __restrict__ mo = ws + L.mo;
        for (int i = lane; i < 5 * ldmax; i += G) mc[i] = mo[i];
      }
      coop_jastrow_commit<G>(S, sd, si, st, w, e, np_[0], np_[1], np_[2], lane, gm, has_j, ws + L.jtmp, ga, la);
    }
    __syncwarp(gm);
  }
  if (a.ke_e == nullptr) return;
  // ---- kinetic-energy pieces of the final walkers from the caches (energy.py:57-65, multiplywf.py:121-129),
  // lanes over (electron, component) for the Slater ratios (cached MO rows . inverse column), then lanes
  // over electrons for the Jastrow gradient / Laplacian sums over the pair caches
  double* __restrict__ rat = ws + L.comp;  // [ne][5]; the AO scratch (5 nao doubles) is free after the sweep
  const bool rat_fits = S.ne * 5 <= 5 * S.nao;
  if (has_s && rat_fits) {
#pragma unroll 1
    for (int t = lane; t < S.ne * 5; t += G) {
      const int e = t / 5, c = t - e * 5;
      const int s = e >= S.nup ? 1 : 0;
      const int n = s ? S.ndn : S.nup;
      const int* __restrict__ occ = si + S.o_occ[s];
      const double* __restrict__ mc = st.mocache + ((size_t)w * S.ne + e) * 5 * ldmax + c * ldmax;
      const double* __restrict__ inv = st.inv[s] + (size_t)w * n * n + (e - s * S.nup);
      double r = 0.0;
      for (int k = 0; k < n; ++k) r = fma(mc[occ[k]], inv[k * n], r);
      rat[t] = r;
    }
  }
  __syncwarp(gm);
#pragma unroll 1
  for (int e = lane; e < S.ne; e += G) {
    double gs[3] = {0.0, 0.0, 0.0}, laps = 0.0;
    if (has_s) {
      double r[5];
      if (rat_fits) {
#pragma unroll
        for (int c = 0; c < 5; ++c) r[c] = rat[e * 5 + c];
      } else {
        const int s = e >= S.nup ? 1 : 0;
        const int n = s ? S.ndn : S.nup;
        const int* __restrict__ occ = si + S.o_occ[s];
        const double* __restrict__ mc = st.mocache + ((size_t)w * S.ne + e) * 5 * ldmax;
        const double* __restrict__ inv = st.inv[s] + (size_t)w * n * n + (e - s * S.nup);
#pragma unroll
        for (int c = 0; c < 5; ++c) {
          r[c] = 0.0;
          for (int k = 0; k < n; ++k) r[c] = fma(mc[c * ldmax + occ[k]], inv[k * n], r[c]);
        }
      }
#pragma unroll
      for (int i = 0; i < 3; ++i) gs[i] = r[1 + i] / r[0];
      laps = r[4] / r[0];
    }
    double gj[3] = {0.0, 0.0, 0.0}, lapj = 0.0, cross = 0.0;
    if (has_j) {
      double lj = ALAP(st, S, w, e);
#pragma unroll
      for (int i = 0; i < 3; ++i) gj[i] = AGRAD(st, S, w, e, i);
      for (int j = 0; j < S.ne; ++j) {
        if (j == e) continue;
        const int p = e < j ? pair_index(S.ne, e, j) : pair_index(S.ne, j, e);
        const double sg = e < j ? 1.0 : -1.0;
        gj[0] += sg * GPAIR(st, S, w, p, 0);
        gj[1] += sg * GPAIR(st, S, w, p, 1);
        gj[2] += sg * GPAIR(st, S, w, p, 2);
        lj += LPAIR(st, S, w, p);
      }
      lapj = lj + (gj[0] * gj[0] + gj[1] * gj[1] + gj[2] * gj[2]);
      cross = gs[0] * gj[0] + gs[1] * gj[1] + gs[2] * gj[2];
    }
    const double lap = (laps + lapj) + cross * 2.0;
    double g2 = 0.0;
#pragma unroll
    for (int i = 0; i < 3; ++i) {
      const double g = gs[i] + gj[i];
      g2 += g * g;
    }
    a.ke_e[(size_t)e * N + w] = -0.5 * lap;
    a.g2_e[(size_t)e * N + w] = g2;
  }
}

// =========================================================================================
// Local energy.
// =========================================================================================
struct EnergyScratch {
  double* ke_e;      // [ne][N]
  double* g2_e;      // [ne][N]
  double* ecp_loc;   // [ne][necp][N]
  int* item_of;      // [ne][necp][N]  (-1: masked out)
  int* work;         // [items] -> t = (e*necp + a)*N + w
  double* vls;       // [items][maxchan]  v_l / prob for the non-local channels
  double* contrib;   // [items][max_naip]
  double* ratio;     // [items][max_naip]  (T-moves)
  int* count;        // [1]
  int maxchan;
  double* ewald;     // [2][N]  periodic: Ewald ee, ei per walker (k_ewald)
  double* ecp_pos;   // [ne*necp*N*max_naip][3]  periodic: wrapped quadrature points
  double* ecp_wrap;  // same shape: their wrap vectors
};

// kinetic energy pieces: one thread per (walker, electron)  (energy.py:57-65).  The MO value /
// gradient / Laplacian rows at the current positions come from st.mocache (filled by recompute
// and refreshed by every accepted move), so no orbital is re-evaluated here.
template <int G>
__global__ void __launch_bounds__(128) k_kinetic(const Sys S, const State st, const EnergyScratch es) {
  // G lanes per (walker, electron): lanes 0..4 take the five cached MO components, the Jastrow
  // sums run over (partner | atom, function) tasks
  const double* sd;
  const int* si;
  stage_tables(S, sd, si);
  const int lane32 = threadIdx.x & 31;
  const int lane = lane32 & (G - 1);
  const unsigned gm = group_mask<G>(lane32);
  const int p = (blockIdx.x * blockDim.x + threadIdx.x) / G;
  const int N = st.N;
  if (p >= N * S.ne) return;
  const int w = p / S.ne, e = p - w * S.ne;
  const int which = (S.nmo[0] + S.nmo[1] > 0 ? QMCB_SLATER : 0) | ((S.na + S.nb) > 0 ? QMCB_JASTROW : 0) |
                    ((S.na3 + S.nb3) > 0 ? QMCB_JASTROW3 : 0);
  const double px = CONF(st, S, w, e, 0), py = CONF(st, S, w, e, 1), pz = CONF(st, S, w, e, 2);
  double gs[3] = {0.0, 0.0, 0.0}, laps = 0.0;
  if (which & QMCB_SLATER) {
    const int s = e >= S.nup ? 1 : 0;
    const int n = s ? S.ndn : S.nup;
    const int eeff = e - s * S.nup;
    const int ldmax = S.ldc[0] > S.ldc[1] ? S.ldc[0] : S.ldc[1];
    const double* __restrict__ mc = st.mocache + ((size_t)w * S.ne + e) * 5 * ldmax;
    const int nds = S.nds[s];
    const int* __restrict__ occ = si + S.o_occ[s];
    if (nds >= G) {
      // multi-determinant expansions: lanes over the unique spin determinants, all five components per lane
      // (the C3 expansion has 70 per spin; five busy lanes walking all of them made this the slowest energy kernel)
      double num5[5] = {0.0, 0.0, 0.0, 0.0, 0.0}, den = 0.0;
      for (int d = lane; d < nds; d += G) {
        const double* __restrict__ inv = st.inv[s] + ((size_t)w * nds + d) * n * n + eeff;
        double r[5] = {0.0, 0.0, 0.0, 0.0, 0.0};
        for (int k = 0; k < n; ++k) {
          const double a = inv[k * n];
          const int orb = occ[d * n + k];
#pragma unroll
          for (int c = 0; c < 5; ++c) r[c] = fma(mc[c * ldmax + orb], a, r[c]);
        }
        const double wgt = st.dv[s][(size_t)w * nds + d] * st.W[s][(size_t)w * nds + d];
        den += wgt;
#pragma unroll
        for (int c = 0; c < 5; ++c) num5[c] = fma(r[c], wgt, num5[c]);
      }
      den = group_sum<G>(den, gm);
#pragma unroll
      for (int c = 0; c < 5; ++c) num5[c] = group_sum<G>(num5[c], gm) / den;
#pragma unroll
      for (int i = 0; i < 3; ++i) gs[i] = num5[1 + i] / num5[0];
      laps = num5[4] / num5[0];
    } else {
    double num = 0.0, den = 0.0;
    if (lane < 5) {
      for (int d = 0; d < nds; ++d) {
        const double* __restrict__ inv = st.inv[s] + ((size_t)w * nds + d) * n * n + eeff;
        double r = 0.0;
        for (int k = 0; k < n; ++k) r = fma(mc[lane * ldmax + occ[d * n + k]], inv[k * n], r);
        if (S.ndet == 1) {
          num = r;
          den = 1.0;
        } else {
          const double wgt = st.dv[s][(size_t)w * nds + d] * st.W[s][(size_t)w * nds + d];
          den += wgt;
          num = fma(r, wgt, num);
        }
      }
      num = num / den;
    }
    const double r0 = __shfl_sync(gm, num, 0, G);
#pragma unroll
    for (int i = 0; i < 3; ++i) gs[i] = __shfl_sync(gm, num, 1 + i, G) / r0;
    laps = __shfl_sync(gm, num, 4, G) / r0;
    }
  }
  double lapj = 0.0, cross = 0.0, gj[3] = {0.0, 0.0, 0.0};
  if (which & QMCB_JASTROW) {
    double du, lj;
    if (S.pbc)
      coop_jastrow_pbc<2, G>(S, sd, si, st, w, e, px, py, pz, lane, gm, du, gj, lj);
    else
      coop_jastrow<2, G>(S, sd, si, st, w, e, px, py, pz, lane, gm, du, gj, lj);
    lapj = lj;
  }
  if (which & QMCB_JASTROW3) {
    extern __shared__ __align__(128) unsigned char qmcb_smem[];
    const size_t tab3 = (16 + (size_t)S.dwords * 8 + (size_t)S.iwords * 4 + 15) & ~(size_t)15;
    double* abuf = reinterpret_cast<double*>(qmcb_smem + tab3) + (size_t)(threadIdx.x / G) * j3_scratch_doubles(S);
    double du3 = 0.0;
    coop_jastrow3<2, G>(S, sd, si, st, w, e, px, py, pz, lane, gm, abuf, du3, gj, lapj);
  }
  if (which & (QMCB_JASTROW | QMCB_JASTROW3)) {
    lapj = lapj + (gj[0] * gj[0] + gj[1] * gj[1] + gj[2] * gj[2]);
    cross = gs[0] * gj[0] + gs[1] * gj[1] + gs[2] * gj[2];
  }
  if (lane == 0) {
    const double lap = (laps + lapj) + cross * 2.0;
    double g2 = 0.0;
#pragma unroll
    for (int i = 0; i < 3; ++i) {
      const double g = gs[i] + gj[i];
      g2 += g * g;
    }
    es.ke_e[(size_t)e * N + w] = -0.5 * lap;
    es.g2_e[(size_t)e * N + w] = g2;
  }
}

// ECP radial channels, stochastic mask and work list: one thread per (electron, ECP atom, walker)
// (eval_ecp.py:83-100, 135-157).  e_only >= 0 restricts to one electron (T-moves).
__global__ void __launch_bounds__(128) k_ecp_prepare(const Sys S, const State st, const EnergyScratch es,
                                                     const double* u, int e_only) {
  const double* sd;
  const int* si;
  stage_tables(S, sd, si);
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  const int N = st.N;
  const int ne_loop = e_only >= 0 ? 1 : S.ne;
  if (t >= ne_loop * S.necp * N) return;
  const int ea = t / N, w = t - ea * N;
  const int e = e_only >= 0 ? e_only : ea / S.necp;
  const int a = ea % S.necp;
  const int atom = si[S.o_ecpatom + a];
  double dx = CONF(st, S, w, e, 0) - sd[S.o_xyz + 3 * atom], dy = CONF(st, S, w, e, 1) - sd[S.o_xyz + 3 * atom + 1],
         dz = CONF(st, S, w, e, 2) - sd[S.o_xyz + 3 * atom + 2];
  if (S.pbc) min_image(S, sd, dx, dy, dz);  // configs.dist.dist_i (eval_ecp.py:94)
  const double r = sqrt(dx * dx + dy * dy + dz * dz);
  const int c0 = si[S.o_chanoff + a], c1 = si[S.o_chanoff + a + 1];
  const int nl = c1 - c0;
  double v[8];
  double prob = 0.0;
  for (int c = 0; c < nl; ++c) {
    double acc = 0.0;
    for (int k = si[S.o_termoff + c0 + c]; k < si[S.o_termoff + c0 + c + 1]; ++k) {
      const int pw = si[S.o_tpow + k];
      double rp;  // r ** n for n in -2..4 (eval_ecp.py:194-200)
      switch (pw) {
        case -2: rp = 1.0 / (r * r); break;
        case -1: rp = 1.0 / r; break;
        case 0: rp = 1.0; break;
        case 1: rp = r; break;
        case 2: rp = r * r; break;
        case 3: rp = r * r * r; break;
        default: rp = pow(r, (double)pw); break;
      }
      acc += rp * sd[S.o_tcoef + k] * exp(-sd[S.o_talpha + k] * r * r);
    }
    v[c] = acc;
    if (c < nl - 1) prob += fabs(acc) * (S.ecp_threshold * (double)(4 * c + 3));
  }
  if (S.ecp_threshold > 0.0)
    prob = fmin(1.0, prob);
  else
    prob = 1.0;
  es.ecp_loc[t] = v[nl - 1];
  const bool acc = prob > u[t];
  int item = -1;
  if (acc) {
    item = atomicAdd(es.count, 1);
    es.work[item] = t;
    for (int c = 0; c < nl - 1; ++c) es.vls[(size_t)item * es.maxchan + c] = v[c] / prob;
  }
  es.item_of[t] = item;
}

__device__ __forceinline__ double legendre_p(int l, double x) {
  switch (l) {
    case 0: return 1.0;
    case 1: return x;
    case 2: return 0.5 * (3.0 * x * x - 1.0);
    case 3: return 0.5 * (5.0 * x * x * x - 3.0 * x);
    default: return 0.125 * (35.0 * x * x * x * x - 30.0 * x * x + 3.0);
  }
}

// ECP quadrature points: one thread per (work item, auxiliary point)  (eval_ecp.py:101-123,
// 228-275).  rot: [ne][necp][9] row-major rotation matrices (or [necp][9] when e_only >= 0).
// tmove_tau > 0: also emit ratio / T-move weight / position instead of the energy contraction.
struct EcpPointArgs {
  const double* rot;
  const double* quad;  // device table: per ECP atom points [naip][3] then weights [naip]
  int e_only;
  double tmove_tau;
  double* tm_ratio;   // [N][tot_naip]
  double* tm_weight;  // [N][tot_naip]
  double* tm_pos;     // [N][tot_naip][3]
  double* scr;
  size_t scr_stride;
  // periodic systems run the kernel twice around k_pbc_mo: pass 0 (pos_out != nullptr) only writes
  // the quadrature points wrapped into the simulation cell (make_irreducible, eval_ecp.py:114) and
  // their wrap vectors; pass 1 (pos_in != nullptr) evaluates with the MO rows found in scr
  double* pos_out;
  double* wrap_out;
  const double* pos_in;
};

template <int NMOT>
__global__ void __launch_bounds__(128) k_ecp_points(const Sys S, const State st, const EnergyScratch es,
                                                    const EcpPointArgs ea) {
  const double* sd;
  const int* si;
  stage_tables(S, sd, si);
  const int N = st.N;
  const int which = (S.nmo[0] + S.nmo[1] > 0 ? QMCB_SLATER : 0) | ((S.na + S.nb) > 0 ? QMCB_JASTROW : 0) |
                    ((S.na3 + S.nb3) > 0 ? QMCB_JASTROW3 : 0);
  const int nitems = *es.count;
  const long long total = (long long)nitems * S.max_naip;
  for (long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x; p < total;
       p += (long long)gridDim.x * blockDim.x) {
    const int item = (int)(p / S.max_naip), q = (int)(p - (long long)item * S.max_naip);
    const int t = es.work[item];
    const int eai = t / N, w = t - eai * N;
    const int e = ea.e_only >= 0 ? ea.e_only : eai / S.necp;
    const int a = eai % S.necp;
    const int naip = si[S.o_naip + a];
    if (q >= naip) {
      if (ea.pos_out) ea.pos_out[3 * p] = NAN;  // k_pbc_mo skips this slot
      continue;
    }
    const int atom = si[S.o_ecpatom + a];
    const double ex = CONF(st, S, w, e, 0), ey = CONF(st, S, w, e, 1),
                 ez = CONF(st, S, w, e, 2);
    double rx = ex - sd[S.o_xyz + 3 * atom], ry = ey - sd[S.o_xyz + 3 * atom + 1], rz = ez - sd[S.o_xyz + 3 * atom + 2];
    if (S.pbc) min_image(S, sd, rx, ry, rz);
    const double r = sqrt(rx * rx + ry * ry + rz * rz);
    const double* R = ea.rot + (size_t)eai * 9;
    const double* qt = ea.quad + (size_t)si[S.o_aipoff + a] * 4;
    const double qx = qt[q * 3], qy = qt[q * 3 + 1], qz = qt[q * 3 + 2];
    const double wq = qt[naip * 3 + q];
    // rot_vec = rot . point ; r_ea_i = r * rot_vec
    const double ux = R[0] * qx + R[1] * qy + R[2] * qz, uy = R[3] * qx + R[4] * qy + R[5] * qz,
                 uz = R[6] * qx + R[7] * qy + R[8] * qz;
    const double dx = r * ux, dy = r * uy, dz = r * uz;
    const double cosang = (rx * dx + ry * dy + rz * dz) / (r * sqrt(dx * dx + dy * dy + dz * dz));
    double px = (ex - rx) + dx, py = (ey - ry) + dy, pz = (ez - rz) + dz;
    const double upx = px, upy = py, upz = pz;  // before the wrap: what the T-move table reports
    if (S.pbc) {
      if (ea.pos_out) {
        double o[3], ww[3];
        wrap_cell(sd + S.o_lat, sd + S.o_latinv, px, py, pz, o, ww);
        for (int i = 0; i < 3; ++i) {
          ea.pos_out[3 * p + i] = o[i];
          ea.wrap_out[3 * p + i] = st.wrap[((size_t)w * S.ne + e) * 3 + i] + ww[i];
        }
        continue;
      }
      px = ea.pos_in[3 * p];
      py = ea.pos_in[3 * p + 1];
      pz = ea.pos_in[3 * p + 2];
    }
    PointEval<0, NMOT> ev;
    ev.run(S, sd, si, st, which, w, e, px, py, pz, nullptr, ea.scr + p, ea.scr_stride);
    const double ratio = ev.rat[0] * exp(ev.du);
    const int nlm1 = si[S.o_chanoff + a + 1] - si[S.o_chanoff + a] - 1;
    if (ea.tmove_tau <= 0.0) {
      double acc = 0.0;
      for (int l = 0; l < nlm1; ++l)
        acc += es.vls[(size_t)item * es.maxchan + l] * ((double)(2 * l + 1) * legendre_p(l, cosang) * wq);
      es.contrib[(size_t)item * S.max_naip + q] = ratio * acc;
    } else {
      double wt = 0.0;
      for (int l = 0; l < nlm1; ++l)
        wt += (exp(-ea.tmove_tau * es.vls[(size_t)item * es.maxchan + l]) - 1.0) *
              ((double)(2 * l + 1) * legendre_p(l, cosang) * wq);
      // local channel column: P_l = 0 -> contributes (exp(-tau v_loc) - 1) * 0 = 0
      const size_t o = (size_t)w * S.tot_naip + (size_t)(si[S.o_aipoff + a] - 0) + q;
      ea.tm_ratio[o] = ratio;
      ea.tm_weight[o] = wt;
      ea.tm_pos[o * 3] = upx;  // periodic: the caller's make_irreducible wraps it (coord.py:168-184)
      ea.tm_pos[o * 3 + 1] = upy;
      ea.tm_pos[o * 3 + 2] = upz;
    }
  }
}

// Cooperative form of k_ecp_points for the T-move tables of open-boundary systems (few points per
// launch: one electron, the walkers that passed the stochastic channel mask): G lanes per quadrature
// point -- orbital evaluation through coop_eval_mo, determinant ratios with lanes over determinants,
// Jastrow sums with lanes over tasks / partners.  With ~10^3 points a thread per point leaves the launch
// bound by the latency of one thread's ~60 exponentials and its AO -> MO contraction.
template <int G>
__global__ void __launch_bounds__(64) k_ecp_points_coop(const Sys S, const State st, const EnergyScratch es,
                                                        const EcpPointArgs ea) {
  const double* sd;
  const int* si;
  stage_tables(S, sd, si);
  extern __shared__ __align__(128) unsigned char qmcb_smem[];
  const CoopLayout L = coop_layout(S);
  const int lane32 = threadIdx.x & 31;
  const int lane = lane32 & (G - 1);
  const unsigned gm = group_mask<G>(lane32);
  const int slot = threadIdx.x / G, gper = blockDim.x / G;
  const size_t tab = (16 + (size_t)S.dwords * 8 + (size_t)S.iwords * 4 + 15) & ~(size_t)15;
  const int j3n = j3_scratch_doubles(S);
  double* ws = reinterpret_cast<double*>(qmcb_smem + tab) + (size_t)slot * (L.total + j3n);
  double* abuf = ws + L.total;
  const int N = st.N;
  const bool has_s = S.nmo[0] + S.nmo[1] > 0, has_j = (S.na + S.nb) > 0, has_j3 = (S.na3 + S.nb3) > 0;
  const int nitems = *es.count;
  const long long total = (long long)nitems * S.max_naip;
  for (long long p = (long long)blockIdx.x * gper + slot; p < total; p += (long long)gridDim.x * gper) {
    const int item = (int)(p / S.max_naip), q = (int)(p - (long long)item * S.max_naip);
    const int t = es.work[item];
    const int eai = t / N, w = t - eai * N;
    const int e = ea.e_only >= 0 ? ea.e_only : eai / S.necp;
    const int a = eai % S.necp;
    const int naip = si[S.o_naip + a];
    if (q >= naip) continue;
    const int atom = si[S.o_ecpatom + a];
    const double ex = CONF(st, S, w, e, 0), ey = CONF(st, S, w, e, 1), ez = CONF(st, S, w, e, 2);
    const double rx = ex - sd[S.o_xyz + 3 * atom], ry = ey - sd[S.o_xyz + 3 * atom + 1], rz = ez - sd[S.o_xyz + 3 * atom + 2];
    const double r = sqrt(rx * rx + ry * ry + rz * rz);
    const double* R = ea.rot + (size_t)eai * 9;
    const double* qt = ea.quad + (size_t)si[S.o_aipoff + a] * 4;
    const double qx = qt[q * 3], qy = qt[q * 3 + 1], qz = qt[q * 3 + 2];
    const double wq = qt[naip * 3 + q];
    const double ux = R[0] * qx + R[1] * qy + R[2] * qz, uy = R[3] * qx + R[4] * qy + R[5] * qz,
                 uz = R[6] * qx + R[7] * qy + R[8] * qz;
    const double dx = r * ux, dy = r * uy, dz = r * uz;
    const double cosang = (rx * dx + ry * dy + rz * dz) / (r * sqrt(dx * dx + dy * dy + dz * dz));
    const double px = (ex - rx) + dx, py = (ey - ry) + dy, pz = (ez - rz) + dz;
    const int s = e >= S.nup ? 1 : 0;
    double rat = 1.0;
    if (has_s) {
      coop_eval_mo<0, G>(S, L, sd, si, s, px, py, pz, ws, lane, gm);
      const double* __restrict__ mo = ws + L.mo;
      const int n = s ? S.ndn : S.nup, nds = S.nds[s], eeff = e - s * S.nup;
      const int* __restrict__ occ = si + S.o_occ[s];
      double num = 0.0, den = 0.0;
      for (int d = lane; d < nds; d += G) {
        const double* __restrict__ inv = st.inv[s] + ((size_t)w * nds + d) * n * n + eeff;
        double rr = 0.0;
        for (int k = 0; k < n; ++k) rr = fma(mo[occ[d * n + k]], inv[k * n], rr);
        const double wgt = S.ndet == 1 ? 1.0 : st.dv[s][(size_t)w * nds + d] * st.W[s][(size_t)w * nds + d];
        den += wgt;
        num = fma(rr, wgt, num);
      }
      rat = group_sum<G>(num, gm) / group_sum<G>(den, gm);
    }
    double du = 0.0, gj[3] = {0.0, 0.0, 0.0}, lj = 0.0;
    if (has_j) coop_jastrow<0, G>(S, sd, si, st, w, e, px, py, pz, lane, gm, du, gj, lj);
    if (has_j3) coop_jastrow3<0, G>(S, sd, si, st, w, e, px, py, pz, lane, gm, abuf, du, gj, lj);
    if (lane == 0) {
      const double ratio = rat * exp(du);
      const int nlm1 = si[S.o_chanoff + a + 1] - si[S.o_chanoff + a] - 1;
      if (ea.tmove_tau <= 0.0) {
        double acc = 0.0;
        for (int l = 0; l < nlm1; ++l)
          acc += es.vls[(size_t)item * es.maxchan + l] * ((double)(2 * l + 1) * legendre_p(l, cosang) * wq);
        es.contrib[(size_t)item * S.max_naip + q] = ratio * acc;
      } else {
        double wt = 0.0;
        for (int l = 0; l < nlm1; ++l)
          wt += (exp(-ea.tmove_tau * es.vls[(size_t)item * es.maxchan + l]) - 1.0) *
                ((double)(2 * l + 1) * legendre_p(l, cosang) * wq);
        const size_t o = (size_t)w * S.tot_naip + (size_t)si[S.o_aipoff + a] + q;
        ea.tm_ratio[o] = ratio;
        ea.tm_weight[o] = wt;
        ea.tm_pos[o * 3] = px;
        ea.tm_pos[o * 3 + 1] = py;
        ea.tm_pos[o * 3 + 2] = pz;
      }
    }
    __syncwarp(gm);
  }
}

// =========================================================================================
// Parameter gradients of the Slater factor (slater.py:462-542).
// =========================================================================================
// AO values of every electron: ao [N][ne][A]   (the reference keeps them as _aovals, slater.py:233)
__global__ void __launch_bounds__(128) k_ao_all(const Sys S, const State st, double* __restrict__ ao) {
  const double* sd;
  const int* si;
  stage_tables(S, sd, si);
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= st.N * S.ne) return;
  const int w = p / S.ne, e = p - w * S.ne;
  const double px = CONF(st, S, w, e, 0), py = CONF(st, S, w, e, 1), pz = CONF(st, S, w, e, 2);
  double* out = ao + (size_t)p * S.nao;
  const double* __restrict__ prim = sd + S.o_prim;
  for (int a = 0; a < S.nbatom; ++a) {
    const double x = px - sd[S.o_bxyz + 3 * a], y = py - sd[S.o_bxyz + 3 * a + 1], z = pz - sd[S.o_bxyz + 3 * a + 2];
    const double r2 = x * x + y * y + z * z;
    for (int sh = si[S.o_atsh + a]; sh < si[S.o_atsh + a + 1]; ++sh) {
      double R = 0.0;
      for (int q = si[S.o_shprim + sh]; q < si[S.o_shprim + sh + 1]; ++q) R += prim[2 * q + 1] * exp(-prim[2 * q] * r2);
      double tmp[44];
      const int l = si[S.o_shl + sh];
      switch (l) {
        case 0: sph_store<0, false>(x, y, z, tmp); break;
        case 1: sph_store<1, false>(x, y, z, tmp); break;
        case 2: sph_store<2, false>(x, y, z, tmp); break;
        case 3: sph_store<3, false>(x, y, z, tmp); break;
        case 4: sph_store<4, false>(x, y, z, tmp); break;
        default: sph_store<5, false>(x, y, z, tmp); break;
      }
      for (int m = 0; m < 2 * l + 1; ++m) out[si[S.o_shao + sh] + m] = tmp[4 * m] * R;
    }
  }
}

// Orbitals sum_mu chi_mu(r_p) C[mu][j] at arbitrary points (open boundary conditions): the orbital evaluation the
// density-matrix accumulators need (obdm.py:150-153, 231-233 with MoleculeOrbitalEvaluator.aos/mos,
// orbitals.py:85-96).  Thread per point; the AO values of a shell are contracted into NT orbital accumulators at
// a time, so the (P, nao) AO matrix is never stored.
template <int NT>
__global__ void __launch_bounds__(128) k_orbitals_points(const Sys S, const double* __restrict__ pos, long long P,
                                                         const double* __restrict__ C, int norb, double* __restrict__ out) {
  const double* sd;
  const int* si;
  stage_tables(S, sd, si);
  const long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= P) return;
  const double px = pos[3 * p], py = pos[3 * p + 1], pz = pos[3 * p + 2];
  const double* __restrict__ prim = sd + S.o_prim;
  for (int j0 = 0; j0 < norb; j0 += NT) {
    double acc[NT];
#pragma unroll
    for (int j = 0; j < NT; ++j) acc[j] = 0.0;
    for (int a = 0; a < S.nbatom; ++a) {
      const double x = px - sd[S.o_bxyz + 3 * a], y = py - sd[S.o_bxyz + 3 * a + 1], z = pz - sd[S.o_bxyz + 3 * a + 2];
      const double r2 = x * x + y * y + z * z;
      for (int sh = si[S.o_atsh + a]; sh < si[S.o_atsh + a + 1]; ++sh) {
        double R = 0.0;
        for (int q = si[S.o_shprim + sh]; q < si[S.o_shprim + sh + 1]; ++q) R += prim[2 * q + 1] * exp(-prim[2 * q] * r2);
        double tmp[44];
        const int l = si[S.o_shl + sh];
        switch (l) {
          case 0: sph_store<0, false>(x, y, z, tmp); break;
          case 1: sph_store<1, false>(x, y, z, tmp); break;
          case 2: sph_store<2, false>(x, y, z, tmp); break;
          case 3: sph_store<3, false>(x, y, z, tmp); break;
          case 4: sph_store<4, false>(x, y, z, tmp); break;
          default: sph_store<5, false>(x, y, z, tmp); break;
        }
        for (int m = 0; m < 2 * l + 1; ++m) {
          const double chi = tmp[4 * m] * R;
          const double* __restrict__ row = C + (size_t)(si[S.o_shao + sh] + m) * norb + j0;
#pragma unroll
          for (int j = 0; j < NT; ++j)
            if (j0 + j < norb) acc[j] = fma(chi, row[j], acc[j]);
        }
      }
    }
#pragma unroll
    for (int j = 0; j < NT; ++j)
      if (j0 + j < norb) out[(size_t)p * norb + j0 + j] = acc[j];
  }
}

// d ln Psi / d det_coeff [N][ndet] and the per-walker weights G_s[d] = sum_{D: map_s(D)=d} c_D dPsi_D
__global__ void __launch_bounds__(128) k_pgrad_det(const Sys S, const State st, double* __restrict__ out,
                                                   double* __restrict__ G, int gstride) {
  const int w = blockIdx.x * blockDim.x + threadIdx.x;
  const int N = st.N;
  if (w >= N) return;
  double sign, logpsi;
  if (S.ndet == 1) {
    const double c = S.detc[0];
    sign = st.dsign[0][w] * st.dsign[1][w] * (c > 0.0 ? 1.0 : (c < 0.0 ? -1.0 : 0.0));
    logpsi = st.dlog[0][w] + st.dlog[1][w] + log(fabs(c));
  } else {
    double val = 0.0;
    for (int d = 0; d < S.nds[0]; ++d) val = fma(st.dv[0][(size_t)w * S.nds[0] + d], st.W[0][(size_t)w * S.nds[0] + d], val);
    sign = val > 0.0 ? 1.0 : (val < 0.0 ? -1.0 : 0.0);
    logpsi = log(fabs(val)) + st.ref[0][w] + st.ref[1][w];
  }
  for (int s = 0; s < 2; ++s)
    for (int d = 0; d < S.nds[s]; ++d) G[((size_t)s * N + w) * gstride + d] = 0.0;
  for (int D = 0; D < S.ndet; ++D) {
    const int d0 = S.map[0][D], d1 = S.map[1][D];
    double v = 0.0;
    if (sign != 0.0)
      v = st.dsign[0][(size_t)w * S.nds[0] + d0] * st.dsign[1][(size_t)w * S.nds[1] + d1] *
          exp(st.dlog[0][(size_t)w * S.nds[0] + d0] + st.dlog[1][(size_t)w * S.nds[1] + d1] - logpsi) / sign;
    out[(size_t)w * S.ndet + D] = v;
    const double cv = S.detc[D] * v;
    G[((size_t)0 * N + w) * gstride + d0] += cv;
    G[((size_t)1 * N + w) * gstride + d1] += cv;
  }
}

// d ln Psi / d mo_coeff_s [N][A][nmo_s]: sum_d G_s[d] sum_e ao[e][a] inv[d][col_d(i)][e]
// periodic: ao is [N][ne][nk][A] and MO i uses the AO set of its own k-point (orbitals.py:241-255)
__global__ void __launch_bounds__(128) k_pgrad_mo(const Sys S, const State st, int s, const double* __restrict__ ao,
                                                  const double* __restrict__ G, int gstride, double* __restrict__ out) {
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const int N = st.N, nmo = S.nmo[s], A = S.nao;
  if (t >= (long long)N * A * nmo) return;
  const int i = (int)(t % nmo);
  const int a = (int)((t / nmo) % A);
  const int w = (int)(t / ((long long)nmo * A));
  const int n = s ? S.ndn : S.nup, lo = s ? S.nup : 0, nds = S.nds[s];
  const int* __restrict__ occ = S.iblob + S.o_occ[s];
  double acc = 0.0;
  for (int d = 0; d < nds; ++d) {
    int col = -1;
    for (int k = 0; k < n; ++k)
      if (occ[d * n + k] == i) col = k;
    if (col < 0) continue;
    const double* __restrict__ inv = st.inv[s] + (((size_t)w * nds + d) * n + col) * n;
    double v = 0.0;
    if (S.pbc) {
      const int k = S.iblob[S.o_mok[s] + i];
      for (int e = 0; e < n; ++e) v = fma(ao[(((size_t)w * S.ne + lo + e) * S.nk + k) * A + a], inv[e], v);
    } else {
      for (int e = 0; e < n; ++e) v = fma(ao[((size_t)w * S.ne + lo + e) * A + a], inv[e], v);
    }
    acc = fma(G[((size_t)s * N + w) * gstride + d], v, acc);
  }
  out[t] = acc;
}

// T-move tables of walkers whose channel mask rejected: ratio 1, weight 0, position = the
// electron's current position  (eval_ecp.py:63-72, 104-106)
__global__ void k_tmove_init(const Sys S, const State st, int e, double* ratio, double* weight, double* pos) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long n = (long long)st.N * S.tot_naip;
  if (i >= n) return;
  const int w = (int)(i / S.tot_naip);
  ratio[i] = 1.0;
  weight[i] = 0.0;
  pos[3 * i] = CONF(st, S, w, e, 0);
  pos[3 * i + 1] = CONF(st, S, w, e, 1);
  pos[3 * i + 2] = CONF(st, S, w, e, 2);
}

// Sum everything per walker in the reference's order: out [6][N] = ke, ee, ei, ecp, grad2, total
template <int G>
__global__ void __launch_bounds__(128) k_energy_finalize(const Sys S, const State st, const EnergyScratch es,
                                                         double* out, int scratch_doubles) {
  // G lanes per walker compute the terms in parallel into shared memory; lane 0 then adds them in
  // the reference's order (energy.py:28-44, eval_ecp.py:21-40), so the result does not depend on G.
  const double* sd;
  const int* si;
  stage_tables(S, sd, si);
  extern __shared__ __align__(128) unsigned char qmcb_smem[];
  const int lane32 = threadIdx.x & 31;
  const int lane = lane32 & (G - 1);
  const unsigned gm = group_mask<G>(lane32);
  const int slot = threadIdx.x / G;
  const int w = (blockIdx.x * blockDim.x + threadIdx.x) / G;
  const int N = st.N;
  if (w >= N) return;
  const size_t tab = (16 + (size_t)S.dwords * 8 + (size_t)S.iwords * 4 + 15) & ~(size_t)15;
  double* buf = reinterpret_cast<double*>(qmcb_smem + tab) + (size_t)slot * scratch_doubles;
  const int ne = S.ne;
  // ---- ECP: one term per (electron, ECP atom)
  for (int t = lane; t < ne * S.necp; t += G) {
    const int e = t / S.necp, a = t - e * S.necp;
    const size_t gi = ((size_t)e * S.necp + a) * N + w;
    const int item = es.item_of[gi];
    double nl = 0.0;
    if (item >= 0) {
      const int naip = si[S.o_naip + a];
      for (int q = 0; q < naip; ++q) nl += es.contrib[(size_t)item * S.max_naip + q];
    }
    buf[t] = nl + es.ecp_loc[gi];
  }
  __syncwarp(gm);
  double ke = 0.0, g2 = 0.0, ecp = 0.0, ee = 0.0, ei = 0.0;
  if (lane == 0) {
    for (int e = 0; e < ne; ++e) {
      ke += es.ke_e[(size_t)e * N + w];
      g2 += es.g2_e[(size_t)e * N + w];
      double ecp_e = 0.0;
      for (int a = 0; a < S.necp; ++a) ecp_e += buf[e * S.necp + a];
      ecp += ecp_e;
    }
  }
  __syncwarp(gm);
  if (S.pbc) {  // Coulomb terms come from the Ewald kernel (accumulators.py:52-55, ewald.py:330-354)
    if (lane == 0) {
      ee = es.ewald[w];
      ei = es.ewald[(size_t)N + w];
      out[w] = ke;
      out[(size_t)N + w] = ee;
      out[(size_t)2 * N + w] = ei;
      out[(size_t)3 * N + w] = ecp;
      out[(size_t)4 * N + w] = g2;
      out[(size_t)5 * N + w] = (((ke + ee) + ei) + ecp) + S.e_ii;
    }
    return;
  }
  // ---- electron-electron: pairs (i < j) in row-major order
  const int npair = ne * (ne - 1) / 2;
  for (int t = lane; t < npair; t += G) {
    int i = 0, rem = t;
    while (rem >= ne - 1 - i) {
      rem -= ne - 1 - i;
      ++i;
    }
    const int j = i + 1 + rem;
    const double dx = CONF(st, S, w, i, 0) - CONF(st, S, w, j, 0), dy = CONF(st, S, w, i, 1) - CONF(st, S, w, j, 1),
                 dz = CONF(st, S, w, i, 2) - CONF(st, S, w, j, 2);
    buf[t] = 1.0 / sqrt(dx * dx + dy * dy + dz * dz);
  }
  __syncwarp(gm);
  if (lane == 0)
    for (int t = 0; t < npair; ++t) ee += buf[t];
  __syncwarp(gm);
  // ---- electron-ion
  for (int t = lane; t < S.natom * ne; t += G) {
    const int I = t / ne, i = t - I * ne;
    const double dx = CONF(st, S, w, i, 0) - sd[S.o_xyz + 3 * I], dy = CONF(st, S, w, i, 1) - sd[S.o_xyz + 3 * I + 1],
                 dz = CONF(st, S, w, i, 2) - sd[S.o_xyz + 3 * I + 2];
    buf[t] = 1.0 / sqrt(dx * dx + dy * dy + dz * dz);
  }
  __syncwarp(gm);
  if (lane == 0) {
    for (int I = 0; I < S.natom; ++I) {
      double acc = 0.0;
      for (int i = 0; i < ne; ++i) acc += buf[I * ne + i];
      ei += -sd[S.o_chg + I] * acc;
    }
    out[w] = ke;
    out[(size_t)N + w] = ee;
    out[(size_t)2 * N + w] = ei;
    out[(size_t)3 * N + w] = ecp;
    out[(size_t)4 * N + w] = g2;
    out[(size_t)5 * N + w] = (((ke + ee) + ei) + ecp) + S.e_ii;
  }
}

// Deterministic column sums: out[k] = sum_w in[k][w], one block per column (fixed tree order).
__global__ void __launch_bounds__(256) k_colsum(const double* in, int N, double* out) {
  __shared__ double sh[256];
  const double* col = in + (size_t)blockIdx.x * N;
  double acc = 0.0;
  for (int i = threadIdx.x; i < N; i += 256) acc += col[i];
  sh[threadIdx.x] = acc;
  __syncthreads();
  for (int s = 128; s > 0; s >>= 1) {
    if ((int)threadIdx.x < s) sh[threadIdx.x] += sh[threadIdx.x + s];
    __syncthreads();
  }
  if (threadIdx.x == 0) out[blockIdx.x] = sh[0];
}

// walker coordinates between the staging buffer and the state (both walker-major (N, ne, 3), the host layout)
__global__ void k_conf_in(const double* host_layout, double* conf, int N, int ne) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (size_t)N * ne * 3) return;
  conf[i] = host_layout[i];
}
__global__ void k_conf_out(const double* conf, double* host_layout, int N, int ne) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (size_t)N * ne * 3) return;
  host_layout[i] = conf[i];
}

// =========================================================================================
// DMC block pieces (pyqmc/method/dmc.py).
// =========================================================================================
// T-move selection for electron e (propose_tmoves, dmc.py:73-120, and the acceptance test of
// dmc_propagate 170-177): one thread per walker over its M = tot_naip candidate moves.
struct TmoveSelectArgs {
  int e, M;
  const double* ratio;   // [N][M]
  const double* weight;  // [N][M]
  const double* pos;     // [N][M][3]
  const double* sel_u;   // [N]  select_walker's rand()
  const double* acc_u;   // [N]
  uint8_t* accept;       // [N]
  unsigned long long* ntacc;
  const int* item_of;    // k_tmove_apply: [necp][N] work item of (ECP atom, walker) or -1 (not sampled)
  int* count;            // k_tmove_apply: work-item counter, reset for the next electron
};

__global__ void __launch_bounds__(128) k_tmove_select(const Sys S, const State st, const TmoveSelectArgs a) {
  const int w = blockIdx.x * blockDim.x + threadIdx.x;
  bool acc = false;
  if (w < st.N) {
    const int M = a.M;
    const double* __restrict__ ra = a.ratio + (size_t)w * M;
    const double* __restrict__ wt = a.weight + (size_t)w * M;
    double norm = 1.0;
    {
      double sum = 0.0;
      for (int m = 0; m < M; ++m) {
        const double amp = __dmul_rn(ra[m], wt[m]);
        if (amp > 0.0) sum = __dadd_rn(sum, amp);
      }
      norm = __dadd_rn(1.0, sum);  // EQN 34
    }
    // selected = searchsorted(cumsum(forward / norm), r): number of entries with cdf < r
    const double r = a.sel_u[w];
    int sel = 0;
    double cdf = 0.0;
    for (int m = 0; m < M; ++m) {
      const double amp = __dmul_rn(ra[m], wt[m]);
      const double f = amp > 0.0 ? amp : 0.0;
      cdf = __dadd_rn(cdf, f / norm);
      if (cdf < r) ++sel;
    }
    const bool chosen = sel < M;
    double px = CONF(st, S, w, a.e, 0), py = CONF(st, S, w, a.e, 1), pz = CONF(st, S, w, a.e, 2);
    double acceptance = 0.0;
    if (chosen) {
      px = a.pos[((size_t)w * M + sel) * 3];
      py = a.pos[((size_t)w * M + sel) * 3 + 1];
      pz = a.pos[((size_t)w * M + sel) * 3 + 2];
      const double rev = 1.0 / ra[sel];
      double bsum = 0.0;
      for (int m = 0; m < M; ++m) {
        double b = m == sel ? __dmul_rn(rev, wt[m]) : __dmul_rn(__dmul_rn(ra[m], wt[m]), rev);
        if (b < 0.0) b = 0.0;
        bsum = __dadd_rn(bsum, b);
      }
      acceptance = norm / __dadd_rn(1.0, bsum);
    }
    acc = chosen && (acceptance > a.acc_u[w]);
    a.accept[w] = acc ? 1 : 0;
    st.saved_pos[(size_t)w * 3] = px;
    st.saved_pos[(size_t)w * 3 + 1] = py;
    st.saved_pos[(size_t)w * 3 + 2] = pz;
  }
  const unsigned b = __ballot_sync(0xffffffffu, acc);
  if ((threadIdx.x & 31) == 0 && b) atomicAdd(a.ntacc, (unsigned long long)__popc(b));
}

// T-move selection AND application for electron e in one launch, G lanes per walker: the group leader
// runs the selection arithmetic of k_tmove_select over the walker's candidate table (entries of ECP atoms
// that were not sampled count as ratio 1 / weight 0 / current position, which is what compute_tmoves
// returns for them, eval_ecp.py:62-70); an accepted walker then evaluates the orbitals at the selected
// position (no saved values, dmc.py:176), applies the Sherman-Morrison update, patches the Jastrow caches
// and moves its coordinate.  Replaces k_tmove_init + k_tmove_select + k_point<MOSAVE> + k_sm_* +
// k_jastrow_update_coop of the launch-per-stage chain (single-determinant open-boundary wave functions).
template <int G>
__global__ void __launch_bounds__(128) k_tmove_apply(const Sys S, const State st, const TmoveSelectArgs a) {
  const double* sd;
  const int* si;
  stage_tables(S, sd, si);
  extern __shared__ __align__(128) unsigned char qmcb_smem[];
  const CoopLayout L = coop_layout(S);
  const int lane32 = threadIdx.x & 31;
  const int lane = lane32 & (G - 1);
  const unsigned gm = group_mask<G>(lane32);
  const int slot = threadIdx.x / G, gper = blockDim.x / G;
  const size_t tab = (16 + (size_t)S.dwords * 8 + (size_t)S.iwords * 4 + 15) & ~(size_t)15;
  double* ws = reinterpret_cast<double*>(qmcb_smem + tab) + (size_t)slot * L.total;
  const int N = st.N, e = a.e, M = a.M;
  const int w = blockIdx.x * gper + slot;
  if (blockIdx.x == 0 && threadIdx.x == 0 && a.count) *a.count = 0;  // every reader of the counter ran before this launch
  bool acc = false;
  double px = 0.0, py = 0.0, pz = 0.0;
  if (w < N) {
    // the walker's candidate table into shared memory, all lanes loading (entries of unsampled atoms take
    // the values compute_tmoves gives them: ratio 1, weight 0); the leader then runs k_tmove_select's
    // sequential arithmetic on it
    double* tra = ws;
    double* twt = ws + M;
    {
      const double* __restrict__ ra = a.ratio + (size_t)w * M;
      const double* __restrict__ wt = a.weight + (size_t)w * M;
      for (int ia = 0; ia < S.necp; ++ia) {
        const bool on = a.item_of[(size_t)ia * N + w] >= 0;
        const int m0 = si[S.o_aipoff + ia], m1 = m0 + si[S.o_naip + ia];
        for (int m = m0 + lane; m < m1; m += G) {
          tra[m] = on ? ra[m] : 1.0;
          twt[m] = on ? wt[m] : 0.0;
        }
      }
    }
    __syncwarp(gm);
    if (lane == 0) {
      double sum = 0.0;
      for (int m = 0; m < M; ++m) {
        const double amp = __dmul_rn(tra[m], twt[m]);
        if (amp > 0.0) sum = __dadd_rn(sum, amp);
      }
      const double norm = __dadd_rn(1.0, sum);  // EQN 34
      const double r = a.sel_u[w];
      int sel = 0;
      double cdf = 0.0;
      for (int m = 0; m < M; ++m) {
        const double amp = __dmul_rn(tra[m], twt[m]);
        const double f = amp > 0.0 ? amp : 0.0;
        cdf = __dadd_rn(cdf, f / norm);
        if (cdf < r) ++sel;
      }
      const bool chosen = sel < M;
      double acceptance = 0.0;
      px = CONF(st, S, w, e, 0);
      py = CONF(st, S, w, e, 1);
      pz = CONF(st, S, w, e, 2);
      if (chosen) {
        bool sel_on = false;
        for (int ia = 0; ia < S.necp; ++ia)
          if (sel >= si[S.o_aipoff + ia] && sel < si[S.o_aipoff + ia] + si[S.o_naip + ia])
            sel_on = a.item_of[(size_t)ia * N + w] >= 0;
        if (sel_on) {
          px = a.pos[((size_t)w * M + sel) * 3];
          py = a.pos[((size_t)w * M + sel) * 3 + 1];
          pz = a.pos[((size_t)w * M + sel) * 3 + 2];
        }
        const double rev = 1.0 / tra[sel];
        double bsum = 0.0;
        for (int m = 0; m < M; ++m) {
          double b = m == sel ? __dmul_rn(rev, twt[m]) : __dmul_rn(__dmul_rn(tra[m], twt[m]), rev);
          if (b < 0.0) b = 0.0;
          bsum = __dadd_rn(bsum, b);
        }
        acceptance = norm / __dadd_rn(1.0, bsum);
      }
      acc = chosen && (acceptance > a.acc_u[w]);
      a.accept[w] = acc ? 1 : 0;
    }
    __syncwarp(gm);  // the table in ws is dead from here on: coop_eval_mo reuses the scratch
    const int leader = lane32 & ~(G - 1);
    acc = __shfl_sync(gm, acc ? 1 : 0, leader) != 0;
    px = __shfl_sync(gm, px, leader);
    py = __shfl_sync(gm, py, leader);
    pz = __shfl_sync(gm, pz, leader);
    if (acc) {
      const int s = e >= S.nup ? 1 : 0;
      if (S.nmo[0] + S.nmo[1] > 0) {
        coop_eval_mo<0, G>(S, L, sd, si, s, px, py, pz, ws, lane, gm);
        coop_sherman_morrison<G>(S, L, si, st, w, s, e - s * S.nup, ws, lane, gm);
      }
      coop_jastrow_update<G>(S, sd, si, st, w, e, px, py, pz, lane, gm, (S.na + S.nb) > 0, ws + L.jtmp);
    }
  }
  const unsigned b = __ballot_sync(0xffffffffu, acc && lane == 0 && w < N);
  if (lane32 == 0 && b) atomicAdd(a.ntacc, (unsigned long long)__popc(b));
}

// Weight update of one DMC step (dmc.py:183-198, compute_S 224-235) and the weighted observables:
// prod [7][N] = w * (ke, ee, ei, ecp, grad2, total), w.  eold / v2old carry E_L and v^2 to the next step.
struct DmcWeightArgs {
  double tstep, branchcut, e_trial, e_est;
  const double* energy;  // [6][N] of this step
  const double* r2prop;
  const double* r2acc;
  double* eold;     // [N]
  double* v2old;    // [N]
  double* weights;  // [N]
  double* prod;     // [7][N]
  int init;         // 1: only record eold / v2old (the evaluation before the first step)
};

__device__ __forceinline__ double dmc_compute_s(const DmcWeightArgs& a, double v2, double eloc, int nelec) {
  double e_cut = a.e_est - eloc;
  if (fabs(e_cut) > a.branchcut) e_cut = a.branchcut * (e_cut > 0.0 ? 1.0 : (e_cut < 0.0 ? -1.0 : 0.0));
  const double q = v2 * a.tstep / (double)nelec;
  return (a.e_trial - a.e_est) + e_cut / sqrt(1.0 + q * q);
}

__global__ void __launch_bounds__(128) k_dmc_weights(const Sys S, const State st, const DmcWeightArgs a) {
  const int w = blockIdx.x * blockDim.x + threadIdx.x;
  const int N = st.N;
  if (w >= N) return;
  const double eloc = a.energy[(size_t)5 * N + w], v2 = a.energy[(size_t)4 * N + w];
  if (!a.init) {
    const double tdamp = a.r2acc[w] / a.r2prop[w];
    const double snew = dmc_compute_s(a, v2, eloc, S.ne), sold = dmc_compute_s(a, a.v2old[w], a.eold[w], S.ne);
    const double wt = a.weights[w] * exp(a.tstep * tdamp * (0.5 * snew + 0.5 * sold));
    a.weights[w] = wt;
    for (int k = 0; k < 6; ++k) a.prod[(size_t)k * N + w] = wt * a.energy[(size_t)k * N + w];
    a.prod[(size_t)6 * N + w] = wt;
  }
  a.eold[w] = eloc;
  a.v2old[w] = v2;
}

// =========================================================================================
// Stochastic-reconfiguration accumulator (pyqmc/observables/stochastic_reconfiguration.py:74-118):
//   dp[i][j]  = d ln Psi_i / d p_j gathered from the parameter-gradient arrays (LinearTransform.
//               serialize_gradients, accumulators.py:161-172)
//   f_i       = nodal regularisation polynomial of 1/|grad Psi|^2 (20-46)
//   dppsi     = sum_i w_i f_i dp_ij,  dpH = sum_i E_i w_i f_i dp_ij,  dpidpj = sum_i dp_ij w_i f_i dp_ik
// =========================================================================================
struct SrArgs {
  int P;
  const int* src;          // [P] source array of parameter j
  const long long* off;    // [P] flat offset inside one walker's block of that array
  const double* base[6];   // det_coeff, mo_alpha, mo_beta, acoeff (avalues), bcoeff (bvalues), ccoeff
  long long stride[6];     // doubles per walker
  const double* weights;   // [N] normalised
  const double* energy;    // [6][N]
  double cutoff;
  double* dp;              // [N][P]
  double* wdpr;            // [N][P]  w_i f_i dp_ij
};

__global__ void __launch_bounds__(128) k_sr_gather(const State st, const SrArgs a) {
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const int N = st.N;
  if (t >= (long long)N * a.P) return;
  const int i = (int)(t / a.P), j = (int)(t - (long long)i * a.P);
  const int s = a.src[j];
  const double v = a.base[s][(long long)i * a.stride[s] + a.off[j]];
  const double r = 1.0 / a.energy[(size_t)4 * N + i];
  double f = 1.0;
  if (r < a.cutoff * a.cutoff) {
    const double c2 = a.cutoff * a.cutoff, c4 = c2 * c2, c6 = c4 * c2;
    f = 9.0 / c2 * r + -15.0 / c4 * (r * r) + 7.0 / c6 * (r * r * r);
  }
  a.dp[t] = v;
  a.wdpr[t] = a.weights[i] * (v * f);
}

// column reductions: out[0][j] = sum_i wdpr_ij (dppsi), out[1][j] = sum_i E_i wdpr_ij (dpH); the first
// six extra columns (j = P..P+5) are the weighted energy averages
__global__ void __launch_bounds__(256) k_sr_colsum(const State st, const SrArgs a, double* __restrict__ out) {
  __shared__ double sh[2][256];
  const int j = blockIdx.x, N = st.N, P = a.P;
  double s0 = 0.0, s1 = 0.0;
  for (int i = threadIdx.x; i < N; i += 256) {
    if (j < P) {
      const double v = a.wdpr[(size_t)i * P + j];
      s0 += v;
      s1 = fma(a.energy[(size_t)5 * N + i], v, s1);
    } else {
      s0 = fma(a.weights[i], a.energy[(size_t)(j - P) * N + i], s0);
    }
  }
  sh[0][threadIdx.x] = s0;
  sh[1][threadIdx.x] = s1;
  __syncthreads();
  for (int s = 128; s > 0; s >>= 1) {
    if ((int)threadIdx.x < s) {
      sh[0][threadIdx.x] += sh[0][threadIdx.x + s];
      sh[1][threadIdx.x] += sh[1][threadIdx.x + s];
    }
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    out[j] = sh[0][0];
    if (j < P) out[(size_t)(P + 6) + j] = sh[1][0];
  }
}

// C[j][k] = sum_i A[i][j] B[i][k]  (A, B: [N][P] row-major, C: [P][P]): 64 x 64 output tile per CTA, 16
// rows of A and B staged per iteration, 4 x 4 accumulators per thread -- the one genuinely dense
// product of this path (FP64: no tensor-core path on tcgen05).
__global__ void __launch_bounds__(256) k_gemm_tn(const double* __restrict__ A, const double* __restrict__ B, int N,
                                                 int P, double* __restrict__ C) {
  __shared__ double sa[16][64 + 1], sb[16][64 + 1];
  const int tj = blockIdx.y * 64, tk = blockIdx.x * 64;
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
  double acc[4][4];
#pragma unroll
  for (int u = 0; u < 4; ++u)
#pragma unroll
    for (int v = 0; v < 4; ++v) acc[u][v] = 0.0;
  for (int i0 = 0; i0 < N; i0 += 16) {
    for (int t = threadIdx.x; t < 16 * 64; t += 256) {
      const int r = t >> 6, c = t & 63;
      const int i = i0 + r;
      sa[r][c] = (i < N && tj + c < P) ? A[(size_t)i * P + tj + c] : 0.0;
      sb[r][c] = (i < N && tk + c < P) ? B[(size_t)i * P + tk + c] : 0.0;
    }
    __syncthreads();
#pragma unroll
    for (int r = 0; r < 16; ++r) {
      double av[4], bv[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) av[u] = sa[r][ty * 4 + u];
#pragma unroll
      for (int v = 0; v < 4; ++v) bv[v] = sb[r][tx * 4 + v];
#pragma unroll
      for (int u = 0; u < 4; ++u)
#pragma unroll
        for (int v = 0; v < 4; ++v) acc[u][v] = fma(av[u], bv[v], acc[u][v]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int u = 0; u < 4; ++u)
#pragma unroll
    for (int v = 0; v < 4; ++v) {
      const int j = tj + ty * 4 + u, k = tk + tx * 4 + v;
      if (j < P && k < P) C[(size_t)j * P + k] = acc[u][v];
    }
}

// The same product on the FP64 tensor pipe: mma.sync.aligned.m8n8k4 (DMMA; tcgen05 has no FP64 path, so this is the
// only tensor-core route for the path's one dense product).  64 x 64 output tile per CTA, four warps of 32 x 32
// (4 x 4 DMMA tiles, 32 accumulator registers per thread), 16 walkers staged per iteration; gridDim.z splits the
// walker range so that P / 64 squared tiles still fill the machine -- partial tiles go to Cpart[z][P][P] and are
// added in a fixed order by k_gemm_reduce (deterministic, unlike atomics).
// Fragment layout (PTX ISA, m8n8k4 .f64): A[row = lane / 4][k = lane % 4], B[k = lane % 4][col = lane / 4],
// C[row = lane / 4][col = 2 (lane % 4) + {0, 1}].  Here "A" = dp^T, so both operands read tile[i0 + lane % 4][x0 + lane / 4].
#define QMCB_GEMM_LD 72  // 64 + 8: the fragment loads hit every shared-memory bank pair exactly twice
__global__ void __launch_bounds__(128) k_gemm_tn_dmma(const double* __restrict__ A, const double* __restrict__ B, int N,
                                                      int P, int rows_per_split, double* __restrict__ Cpart) {
  __shared__ double sa[16][QMCB_GEMM_LD], sb[16][QMCB_GEMM_LD];
  const int tj = blockIdx.y * 64, tk = blockIdx.x * 64;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int wj = (warp >> 1) * 32, wk = (warp & 1) * 32;
  const int fr = lane & 3, fc = lane >> 2;
  double acc[4][4][2];
#pragma unroll
  for (int u = 0; u < 4; ++u)
#pragma unroll
    for (int v = 0; v < 4; ++v) acc[u][v][0] = acc[u][v][1] = 0.0;
  const int ibeg = blockIdx.z * rows_per_split, iend = min(N, ibeg + rows_per_split);
  for (int i0 = ibeg; i0 < iend; i0 += 16) {
    for (int t = threadIdx.x; t < 16 * 64; t += 128) {
      const int r = t >> 6, c = t & 63;
      const int i = i0 + r;
      sa[r][c] = (i < iend && tj + c < P) ? A[(size_t)i * P + tj + c] : 0.0;
      sb[r][c] = (i < iend && tk + c < P) ? B[(size_t)i * P + tk + c] : 0.0;
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < 16; kk += 4) {
      double af[4], bf[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) af[u] = sa[kk + fr][wj + 8 * u + fc];
#pragma unroll
      for (int v = 0; v < 4; ++v) bf[v] = sb[kk + fr][wk + 8 * v + fc];
#pragma unroll
      for (int u = 0; u < 4; ++u)
#pragma unroll
        for (int v = 0; v < 4; ++v)
          asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0, %1}, {%2}, {%3}, {%0, %1};"
                       : "+d"(acc[u][v][0]), "+d"(acc[u][v][1])
                       : "d"(af[u]), "d"(bf[v]));
    }
    __syncthreads();
  }
  double* __restrict__ C = Cpart + (size_t)blockIdx.z * P * P;
#pragma unroll
  for (int u = 0; u < 4; ++u)
#pragma unroll
    for (int v = 0; v < 4; ++v) {
      const int j = tj + wj + 8 * u + fc, k = tk + wk + 8 * v + 2 * fr;
      if (j < P && k < P) C[(size_t)j * P + k] = acc[u][v][0];
      if (j < P && k + 1 < P) C[(size_t)j * P + k + 1] = acc[u][v][1];
    }
}

__global__ void __launch_bounds__(256) k_gemm_reduce(const double* __restrict__ Cpart, int nsplit, size_t n,
                                                     double* __restrict__ C) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  double acc = 0.0;
  for (int z = 0; z < nsplit; ++z) acc += Cpart[(size_t)z * n + i];
  C[i] = acc;
}

#include "cplx.cuh"
