// device_rng.cuh -- numpy's legacy global generator (MT19937 + 53-bit doubles + polar Gaussians with one cached
// value) and scipy's Rotation.random(), reproduced BIT FOR BIT on the device.
//
// The reference draws every random number of VMC/DMC from the global legacy np.random stream
// (pyqmc/method/mc.py:119,132; pyqmc/observables/eval_ecp.py:145,263).  Round 1 generated that stream on the
// host (legacy_rng.cpp) and shipped ~1.8 MB of variates per step over PCIe; the sequential host walk of the
// stream (0.22-0.35 ms per C2 step) then bounded the end-to-end rate above the 0.26 ms device step.  Here the
// device continues the SAME stream from the state np.random.get_state() hands over (2.5 KB), so a block's
// variates never exist on the host:
//
//   k_mt_generate  one CTA: the MT19937 recurrence, 624-word state blocks ping-ponged in shared memory
//                  (3 dependent phases of <= 227 independent words per block), tempered words to HBM;
//   k_rng_plan     one CTA walks the "draw program" (the ordered list of uniform / normal / rotation draws of
//                  the block): uniform draws consume 2 words each; a normal draw of n values consumes 4 words
//                  per polar attempt until ceil((n - cached) / 2) attempts were accepted -- found with block
//                  scans over the accept flags (x1^2 + x2^2 < 1, exact IEEE arithmetic) -- which fixes the stream
//                  offset of every draw, the per-chunk accepted-pair prefix and the cached value carried out;
//   k_rng_fill     many CTAs, one per (draw, chunk): converts words to doubles; for accepted attempts evaluates
//                  f = sqrt(-2 log(r2) / r2) with glibc's log (glibc_log.h) and scatters f x2, f x1 to their slots
//                  (scale applied as numpy does: 0.0 + scale * g); rotations: 4 normals -> unit quaternion -> matrix;
//   k_rng_finalize recovers the raw state block the stream stopped in (inverse tempering), position and cached
//                  Gaussian -> the state of the NEXT block's k_mt_generate and what set_state() gets at the end.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>
#include "glibc_log.h"

namespace devrng {

constexpr int KIND_UNIFORM = 0, KIND_NORMAL = 1, KIND_ROTATION = 2;
constexpr int PLAN_THREADS = 1024;
constexpr int CHUNK_ATTEMPTS = 4096;  // polar attempts per (draw, chunk) work item: 4 per thread
constexpr int CHUNK_UNIFORM = 4096;   // doubles per uniform work item

struct Op {            // host-built, one per draw, in consumption order
  double* dst;         // device destination (rotation: 9 doubles)
  long long count;     // values (rotation: 4 normals)
  double scale;
  int kind;
  int chunk0;          // first slot of this draw in the chunk-prefix array
  int maxchunks;       // slots reserved (upper bound of chunks the draw may need)
  int small;           // 1: few values (rotation) -- chunks of 32 attempts handled by one warp
};

struct OpPlan {        // written by k_rng_plan
  long long start_word;
  long long pairs;     // accepted pairs this draw consumes
  double cached_in;
  int has_in;
  int nchunks;
};

struct State {
  uint32_t key[624];
  int pos;             // numpy convention: index of the next word in key, 624 = regenerate first
  int has_gauss;
  double cached;
  long long cur_end;   // scratch: word offset (relative to the generated buffer) where the program stopped
  int error;           // 1: generated buffer exhausted (host sized it too small)
  int pad;
};

struct Work {          // one CTA of k_rng_fill
  int op;
  int chunk;
};

__device__ __forceinline__ uint32_t temper(uint32_t y) {
  y ^= (y >> 11);
  y ^= (y << 7) & 0x9d2c5680u;
  y ^= (y << 15) & 0xefc60000u;
  y ^= (y >> 18);
  return y;
}

__device__ __forceinline__ uint32_t untemper(uint32_t y) {
  y ^= y >> 18;
  y ^= (y << 15) & 0xefc60000u;
  uint32_t x = y;
#pragma unroll
  for (int k = 0; k < 4; ++k) x = y ^ ((x << 7) & 0x9d2c5680u);
  y = x;
  x = y ^ (y >> 11);
  x = y ^ (x >> 11);
  return x;
}

__device__ __forceinline__ uint32_t mt_twist(uint32_t cur, uint32_t nxt, uint32_t far) {
  const uint32_t y = (cur & 0x80000000u) | (nxt & 0x7fffffffu);
  return far ^ (y >> 1) ^ ((0u - (y & 1u)) & 0x9908b0dfu);
}

// numpy's random_double: (a >> 5, b >> 6) -> (a * 2^26 + b) / 2^53; the integer is below 2^53, so this is exact
__device__ __forceinline__ double to_double(uint32_t a, uint32_t b) {
  const unsigned long long v = ((unsigned long long)(a >> 5) << 26) | (unsigned long long)(b >> 6);
  return __dmul_rn((double)v, 1.0 / 9007199254740992.0);
}

// one polar attempt from 4 consecutive words; numpy: x = 2.0 * double - 1.0; r2 = x1*x1 + x2*x2 (each rounded)
__device__ __forceinline__ bool polar(const uint32_t* w, double& x1, double& x2, double& r2) {
  x1 = __dsub_rn(__dmul_rn(2.0, to_double(w[0], w[1])), 1.0);
  x2 = __dsub_rn(__dmul_rn(2.0, to_double(w[2], w[3])), 1.0);
  r2 = __dadd_rn(__dmul_rn(x1, x1), __dmul_rn(x2, x2));
  return !(r2 >= 1.0 || r2 == 0.0);
}

__device__ __forceinline__ double polar_factor(double r2) {
  return __dsqrt_rn(__ddiv_rn(__dmul_rn(-2.0, qmcb_glibc_log(r2)), r2));
}

// W holds `nblocks` consecutive tempered state blocks; block 0 is the state handed in.
__global__ void __launch_bounds__(256) k_mt_generate(const State* st, uint32_t* __restrict__ W, int nblocks) {
  __shared__ uint32_t buf[2][624];
  const int t = threadIdx.x;
  for (int i = t; i < 624; i += 256) {
    const uint32_t v = st->key[i];
    buf[0][i] = v;
    W[i] = temper(v);
  }
  __syncthreads();
  int cur = 0;
  for (int b = 1; b < nblocks; ++b) {
    const uint32_t* o = buf[cur];
    uint32_t* n = buf[cur ^ 1];
    uint32_t* out = W + (size_t)b * 624;
    if (t < 227) {
      const uint32_t v = mt_twist(o[t], o[t + 1], o[t + 397]);
      n[t] = v;
      out[t] = temper(v);
    }
    __syncthreads();
    if (t < 227) {
      const int kk = 227 + t;
      const uint32_t v = mt_twist(o[kk], o[kk + 1], n[t]);
      n[kk] = v;
      out[kk] = temper(v);
    }
    __syncthreads();
    if (t < 170) {
      const int kk = 454 + t;
      const uint32_t nxt = kk < 623 ? o[kk + 1] : n[0];
      const uint32_t v = mt_twist(o[kk], nxt, n[kk - 227]);
      n[kk] = v;
      out[kk] = temper(v);
    }
    __syncthreads();
    cur ^= 1;
  }
}

// exclusive scan of one int per thread over the CTA (blockDim = PLAN_THREADS); returns the prefix, total in *total
__device__ __forceinline__ int block_scan(int v, int* warp_sums, int* total) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  int inc = v;
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    const int o = __shfl_up_sync(0xffffffffu, inc, d);
    if (lane >= d) inc += o;
  }
  if (lane == 31) warp_sums[warp] = inc;
  __syncthreads();
  if (warp == 0) {
    int s = warp_sums[lane];
    int sinc = s;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      const int o = __shfl_up_sync(0xffffffffu, sinc, d);
      if (lane >= d) sinc += o;
    }
    warp_sums[lane] = sinc - s;  // exclusive
    if (lane == 31) *total = sinc;
  }
  __syncthreads();
  const int pre = warp_sums[warp] + inc - v;
  __syncthreads();
  return pre;
}

// accept flags of the attempts thread t owns in (draw, chunk): attempts [first, first + A)
__device__ __forceinline__ int attempt_flags(const uint32_t* __restrict__ W, long long a0, long long first, int A,
                                             long long nwords, int* overflow) {
  int flags = 0;
  for (int q = 0; q < A; ++q) {
    const long long w0 = a0 + 4 * (first + q);
    if (w0 + 4 > nwords) {  // beyond the generated words: not an attempt; an error only if the draw needs it
      *overflow = 1;
      break;
    }
    uint32_t w[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) w[i] = __ldg(W + w0 + i);
    double x1, x2, r2;
    if (polar(w, x1, x2, r2)) flags |= 1 << q;
  }
  return flags;
}

__global__ void __launch_bounds__(PLAN_THREADS) k_rng_plan(const Op* __restrict__ ops, int nops, State* st,
                                                            const uint32_t* __restrict__ W, long long nwords,
                                                            OpPlan* __restrict__ plan, long long* __restrict__ chunk_prefix) {
  __shared__ int warp_sums[32];
  __shared__ int s_total, s_overflow, s_soft, s_end_attempt_lo, s_end_attempt_hi;
  __shared__ double s_cached_out;
  const int t = threadIdx.x;
  long long cur = st->pos;  // all threads track the same scalars (uniform control flow)
  int has = st->has_gauss;
  double cached = st->cached;
  if (t == 0) {
    s_overflow = 0;
    s_soft = 0;
  }
  __syncthreads();
  for (int op = 0; op < nops; ++op) {
    const Op o = ops[op];
    const long long n = o.kind == KIND_ROTATION ? 4 : o.count;
    if (t == 0) {
      plan[op].start_word = cur;
      plan[op].has_in = has;
      plan[op].cached_in = cached;
    }
    if (o.kind == KIND_UNIFORM) {
      if (t == 0) {
        plan[op].pairs = 0;
        plan[op].nchunks = (int)((n + CHUNK_UNIFORM - 1) / CHUNK_UNIFORM);
      }
      cur += 2 * n;
      if (cur > nwords) {
        if (t == 0) s_overflow = 1;
        __syncthreads();
        break;
      }
      continue;
    }
    long long need = n;
    if (has && n > 0) {  // the cached value is handed out first
      need -= 1;
      has = 0;
      cached = 0.0;
    }
    const long long m = (need + 1) / 2;
    long long accepted = 0;
    int chunk = 0;
    const int A = o.small ? 1 : CHUNK_ATTEMPTS / PLAN_THREADS;
    const int nact = o.small ? 32 : PLAN_THREADS;
    long long end_attempt = 0;
    while (accepted < m) {
      if (chunk >= o.maxchunks) {  // reserved prefix slots exhausted: treat as overflow (host bound too tight)
        if (t == 0) s_overflow = 1;
        break;
      }
      const long long first = (long long)chunk * nact * A + (long long)t * A;
      int ovf = 0;
      const int flags = t < nact ? attempt_flags(W, cur, first, A, nwords, &ovf) : 0;
      if (ovf) s_soft = 1;
      const int cnt = __popc(flags);
      int total;
      const int pre = block_scan(cnt, warp_sums, &s_total);
      total = s_total;
      if (t == 0) chunk_prefix[o.chunk0 + chunk] = accepted;
      const long long remaining = m - accepted;
      if (total >= remaining && cnt > 0 && pre < remaining && remaining <= pre + cnt) {
        // this thread owns the attempt that completes the draw: the (remaining - pre)-th set flag
        int k = (int)(remaining - pre), q = 0;
        for (; q < A; ++q)
          if ((flags >> q) & 1)
            if (--k == 0) break;
        const long long e = first + q + 1;
        s_end_attempt_lo = (int)(e & 0xffffffffll);
        s_end_attempt_hi = (int)(e >> 32);
        if ((2 * m - need) == 1) {  // odd request: the second value of the last pair stays cached
          uint32_t w[4];
          for (int i = 0; i < 4; ++i) w[i] = __ldg(W + cur + 4 * (first + q) + i);
          double x1, x2, r2;
          polar(w, x1, x2, r2);
          s_cached_out = __dmul_rn(polar_factor(r2), x1);
        }
      }
      __syncthreads();
      accepted += total;
      ++chunk;
      if (accepted >= m) end_attempt = ((long long)s_end_attempt_hi << 32) | (unsigned int)s_end_attempt_lo;
      const bool starved = s_soft && accepted < m;  // ran past the generated words before the draw completed
      __syncthreads();
      if (starved) {
        if (t == 0) s_overflow = 1;
        break;
      }
    }
    __syncthreads();
    if (s_overflow) break;
    if (m > 0) {
      cur += 4 * end_attempt;
      if ((2 * m - need) == 1) {
        has = 1;
        cached = s_cached_out;
      }
    }
    if (t == 0) {
      plan[op].pairs = m;
      plan[op].nchunks = (chunk == 0 && n > 0) ? 1 : chunk;  // a draw served by the cached value alone still writes it
      s_soft = 0;
    }
    __syncthreads();
  }
  __syncthreads();
  if (t == 0) {
    st->cur_end = cur;
    st->has_gauss = has;
    st->cached = cached;
    st->error = s_overflow;
  }
}

__device__ __forceinline__ void write_rotation(const double* q, double* m) {
  const double n2 = __dadd_rn(__dadd_rn(__dadd_rn(__dmul_rn(q[0], q[0]), __dmul_rn(q[1], q[1])), __dmul_rn(q[2], q[2])),
                              __dmul_rn(q[3], q[3]));
  const double norm = __dsqrt_rn(n2);
  const double x = __ddiv_rn(q[0], norm), y = __ddiv_rn(q[1], norm), z = __ddiv_rn(q[2], norm), w = __ddiv_rn(q[3], norm);
  const double x2 = __dmul_rn(x, x), y2 = __dmul_rn(y, y), z2 = __dmul_rn(z, z), w2 = __dmul_rn(w, w);
  const double xy = __dmul_rn(x, y), zw = __dmul_rn(z, w), xz = __dmul_rn(x, z), yw = __dmul_rn(y, w);
  const double yz = __dmul_rn(y, z), xw = __dmul_rn(x, w);
  m[0] = __dadd_rn(__dsub_rn(__dsub_rn(x2, y2), z2), w2);
  m[3] = __dmul_rn(2.0, __dadd_rn(xy, zw));
  m[6] = __dmul_rn(2.0, __dsub_rn(xz, yw));
  m[1] = __dmul_rn(2.0, __dsub_rn(xy, zw));
  m[4] = __dadd_rn(__dsub_rn(__dadd_rn(-x2, y2), z2), w2);
  m[7] = __dmul_rn(2.0, __dadd_rn(yz, xw));
  m[2] = __dmul_rn(2.0, __dadd_rn(xz, yw));
  m[5] = __dmul_rn(2.0, __dsub_rn(yz, xw));
  m[8] = __dadd_rn(__dadd_rn(__dsub_rn(-x2, y2), z2), w2);
}

__global__ void __launch_bounds__(PLAN_THREADS) k_rng_fill(const Op* __restrict__ ops, const OpPlan* __restrict__ plan,
                                                            const long long* __restrict__ chunk_prefix,
                                                            const Work* __restrict__ work, const State* st,
                                                            const uint32_t* __restrict__ W, long long nwords) {
  __shared__ int warp_sums[32];
  __shared__ int s_total;
  __shared__ double s_q[4];
  if (st->error) return;
  const Work wk = work[blockIdx.x];
  const Op o = ops[wk.op];
  const OpPlan p = plan[wk.op];
  if (wk.chunk >= p.nchunks) return;
  const int t = threadIdx.x;
  if (o.kind == KIND_UNIFORM) {
    const long long i0 = (long long)wk.chunk * CHUNK_UNIFORM;
#pragma unroll
    for (int q = 0; q < CHUNK_UNIFORM / PLAN_THREADS; ++q) {
      const long long i = i0 + (long long)q * PLAN_THREADS + t;
      if (i < o.count) {
        const long long w0 = p.start_word + 2 * i;
        o.dst[i] = to_double(__ldg(W + w0), __ldg(W + w0 + 1));
      }
    }
    return;
  }
  const long long n = o.kind == KIND_ROTATION ? 4 : o.count;
  double* dst = o.kind == KIND_ROTATION ? s_q : o.dst;
  const double scale = o.kind == KIND_ROTATION ? 1.0 : o.scale;
  if (wk.chunk == 0 && t == 0 && p.has_in && n > 0) dst[0] = __dadd_rn(0.0, __dmul_rn(scale, p.cached_in));
  if (p.pairs == 0) return;  // (a rotation always needs pairs: n = 4)
  const int A = o.small ? 1 : CHUNK_ATTEMPTS / PLAN_THREADS;
  const int nact = o.small ? 32 : PLAN_THREADS;
  const long long first = (long long)wk.chunk * nact * A + (long long)t * A;
  int flags = 0;
  double X1[CHUNK_ATTEMPTS / PLAN_THREADS], X2[CHUNK_ATTEMPTS / PLAN_THREADS], R2[CHUNK_ATTEMPTS / PLAN_THREADS];
  if (t < nact) {
#pragma unroll
    for (int q = 0; q < CHUNK_ATTEMPTS / PLAN_THREADS; ++q) {
      if (q < A) {
        const long long w0 = p.start_word + 4 * (first + q);
        if (w0 + 4 <= nwords) {  // (attempts past the generated words lie after the end of the draw)
          uint32_t w[4];
#pragma unroll
          for (int i = 0; i < 4; ++i) w[i] = __ldg(W + w0 + i);
          if (polar(w, X1[q], X2[q], R2[q])) flags |= 1 << q;
        }
      }
    }
  }
  const int cnt = __popc(flags);
  const int pre = block_scan(cnt, warp_sums, &s_total);
  long long pair = chunk_prefix[o.chunk0 + wk.chunk] + pre;
  const int shift = p.has_in && n > 0 ? 1 : 0;
#pragma unroll
  for (int q = 0; q < CHUNK_ATTEMPTS / PLAN_THREADS; ++q) {
    if (q < A && ((flags >> q) & 1)) {
      if (pair < p.pairs) {
        const double f = polar_factor(R2[q]);
        const long long slot = shift + 2 * pair;
        dst[slot] = __dadd_rn(0.0, __dmul_rn(scale, __dmul_rn(f, X2[q])));
        if (slot + 1 < n) dst[slot + 1] = __dadd_rn(0.0, __dmul_rn(scale, __dmul_rn(f, X1[q])));
      }
      ++pair;
    }
  }
  if (o.kind == KIND_ROTATION) {
    __syncthreads();
    if (t == 0) write_rotation(s_q, o.dst);
  }
}

// New generator state after the program: the raw state block the stream stopped in, numpy's position
// convention (1..624 after any draw), cached Gaussian already stored by the plan kernel.
__global__ void __launch_bounds__(256) k_rng_finalize(State* st, const uint32_t* __restrict__ W) {
  __shared__ long long s_block;
  __shared__ int s_pos;
  if (st->error) return;
  if (threadIdx.x == 0) {
    const long long cur = st->cur_end;
    long long b = cur / 624;
    int pos = (int)(cur % 624);
    if (pos == 0 && b > 0) {  // numpy leaves pos = 624 at a block boundary and regenerates on the next draw
      b -= 1;
      pos = 624;
    }
    s_block = b;
    s_pos = pos;
  }
  __syncthreads();
  const uint32_t* blk = W + (size_t)s_block * 624;
  uint32_t v[3];
  int k = 0;
  for (int i = threadIdx.x; i < 624; i += 256) v[k++] = untemper(blk[i]);
  __syncthreads();
  k = 0;
  for (int i = threadIdx.x; i < 624; i += 256) st->key[i] = v[k++];
  if (threadIdx.x == 0) st->pos = s_pos;
}

}  // namespace devrng
