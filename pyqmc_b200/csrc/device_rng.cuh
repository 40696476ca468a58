// device_rng.cuh -- numpy's legacy global generator (MT19937 + 53-bit doubles + polar Gaussians with one cached
// value) and scipy's Rotation.random(), reproduced BIT FOR BIT on the device.
//
// The reference draws every random number of VMC/DMC from the global legacy np.random stream
// (pyqmc/method/mc.py:119,132; pyqmc/observables/eval_ecp.py:145,263).  Round 1 generated that stream on the
// host (legacy_rng.cpp) and shipped ~1.8 MB of variates per step over PCIe; the sequential host walk of the
// stream (0.22-0.35 ms per C2 step) then bounded the end-to-end rate above the 0.26 ms device step.  Here the
// device continues the SAME stream from the state np.random.get_state() hands over (2.5 KB), so a block's
// variates never exist on the host.  The work is split by what is inherently sequential:
//
//   k_mt_generate  (one CTA, own stream, runs ahead of everything else) the MT19937 recurrence: 624-word state
//                  blocks ping-ponged in shared memory, ONE barrier per block (the three dependent phases of the
//                  textbook reload are substituted into each other), RAW state words to HBM -- tempering is
//                  left to the parallel consumers;
//   k_rng_flags    (grid) accept flag of every possible polar attempt -- 4 consecutive words, x1^2 + x2^2 < 1 in
//                  numpy's exact arithmetic -- as bitmaps.  Attempts start at the running stream offset, which only
//                  ever moves by multiples of 2 words, so two alignment classes (offset mod 4) cover every case;
//   k_rng_plan     (one CTA) walks the draw program: a uniform draw of n values moves the offset by 2n words, a
//                  normal draw of n values by 4 words per attempt until ceil((n - cached)/2) attempts were accepted
//                  = position of the m-th set bit of the class bitmap after the offset (popcount scan).  Produces the
//                  stream offset of every draw and the cached Gaussian carried from draw to draw;
//   k_rng_fill     (grid, one CTA per (draw, chunk)) tempers and converts words to doubles; for accepted attempts
//                  evaluates f = sqrt(-2 log(r2) / r2) with glibc's log (glibc_log.h) and scatters f x2, f x1 to their
//                  slots (0.0 + scale * g, as numpy); rotations: 4 normals -> unit quaternion -> matrix;
//   k_rng_rebase / k_rng_finalize   bookkeeping: restart the word buffer at the block the stream is in; export
//                  numpy's (key, pos, has_gauss, cached) for np.random.set_state().
#pragma once
#include <cstdint>
#include <cuda_runtime.h>
#include "glibc_log.h"

namespace devrng {

constexpr int KIND_UNIFORM = 0, KIND_NORMAL = 1, KIND_ROTATION = 2;
constexpr int THREADS = 256;
constexpr int CHUNK_GROUPS = 32 * THREADS;  // polar attempts per (draw, chunk) work item: one bitmap word per thread
constexpr int CHUNK_UNIFORM = 8 * THREADS;  // doubles per uniform work item

struct Op {            // host-built, one per draw, in consumption order
  double* dst;         // device destination (rotation: 9 doubles)
  long long count;     // values (rotation: 4 normals)
  double scale;
  int kind;
  int maxchunks;       // work items reserved for this draw
};

struct OpPlan {        // written by k_rng_plan
  long long start_word;
  long long pairs;     // accepted pairs this draw consumes
  double cached_in;
  int has_in;
  int pad;
};

struct State {
  long long cur;       // stream offset (word index into the buffer) of the next draw
  double cached;
  int has_gauss;
  int error;           // 1: the program needed words / flags that were not generated (host bound too small)
  uint32_t key[624];   // filled by k_rng_finalize: numpy's view of the state
  int pos;
  int pad;
};

struct GenState {
  uint32_t key[624];   // raw state block generated last
  long long cycles, nanos, blocks;  // of the last k_mt_generate launch (clock64 / globaltimer): profiles/devrng_microbench.py
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

// the "magic" part of the recurrence: twist of the word pair (a, b) without the far term
__device__ __forceinline__ uint32_t mt_f(uint32_t a, uint32_t b) {
  const uint32_t y = (a & 0x80000000u) | (b & 0x7fffffffu);
  return (y >> 1) ^ ((0u - (y & 1u)) & 0x9908b0dfu);
}

// Word k of the NEXT state block written directly in terms of the CURRENT block o[0..623].  The textbook
// reload has three dependent phases (words 227.. need the new words 0..226, words 454.. need the new words
// 227..); substituting the recurrence into itself removes the dependency -- every new word is an XOR of one
// old word and at most three twists of old pairs -- so a whole block costs ONE barrier instead of three.
__device__ __forceinline__ uint32_t mt_next_word(const uint32_t* o, int k) {
  if (k < 227) return o[k + 397] ^ mt_f(o[k], o[k + 1]);
  if (k < 454) return o[k + 170] ^ mt_f(o[k - 227], o[k - 226]) ^ mt_f(o[k], o[k + 1]);
  const uint32_t nxt = k < 623 ? o[k + 1] : (o[397] ^ mt_f(o[0], o[1]));  // word 623 pairs with the NEW word 0
  return o[k - 57] ^ mt_f(o[k - 454], o[k - 453]) ^ mt_f(o[k - 227], o[k - 226]) ^ mt_f(o[k], nxt);
}

// Generates raw state blocks [first, first + nblocks) of the buffer from the block generated last (g->key).
// MODE 0: every new word straight from the old block (mt_next_word: up to three twists per word, ONE barrier per
// block).  MODE 1: the 624 twists h[j] = f(o[j], o[j+1]) are computed once into shared memory, then every new word
// is an XOR of one old word and up to three h's (two barriers, ~40 % fewer instructions: the single CTA is
// issue-bound, profiles/devrng_microbench.py).
template <int MODE>
__global__ void __launch_bounds__(640) k_mt_generate(GenState* g, uint32_t* __restrict__ W, long long first, int nblocks) {
  __shared__ uint32_t buf[2][624];
  __shared__ uint32_t h[624];
  const int t = threadIdx.x;
  if (t < 624) buf[0][t] = g->key[t];
  __syncthreads();
  long long c0 = 0, n0 = 0;
  if (t == 0) {
    c0 = clock64();
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(n0));
  }
  int cur = 0;
  // one word per thread: 20 warps hide the dependent shared-memory / ALU latencies of the short per-word chain
  for (int b = 0; b < nblocks; ++b) {
    const uint32_t* o = buf[cur];
    uint32_t* n = buf[cur ^ 1];
    uint32_t* out = W + (size_t)(first + b) * 624;
    if (MODE == 0) {
      if (t < 624) {
        const uint32_t v = mt_next_word(o, t);
        n[t] = v;
        out[t] = v;
      }
    } else {
      if (t < 624) h[t] = mt_f(o[t], t < 623 ? o[t + 1] : (o[397] ^ mt_f(o[0], o[1])));  // word 623 pairs with the NEW word 0
      __syncthreads();
      if (t < 624) {
        const uint32_t v = t < 227 ? (o[t + 397] ^ h[t])
                                   : (t < 454 ? (o[t + 170] ^ h[t - 227] ^ h[t]) : (o[t - 57] ^ h[t - 454] ^ h[t - 227] ^ h[t]));
        n[t] = v;
        out[t] = v;
      }
    }
    __syncthreads();
    cur ^= 1;
  }
  if (t < 624) g->key[t] = buf[cur][t];
  if (t == 0) {
    long long n1;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(n1));
    g->cycles = clock64() - c0;
    g->nanos = n1 - n0;
    g->blocks = nblocks;
  }
}

// Jump-ahead: the state block that starts QMCB_MT_SEG_BLOCKS blocks after the block at `src` (raw words src[0..623],
// followed in the buffer by the 32 blocks generated from it).  With g(x) = x^J mod phi(x), J = 624 SEG - 1 and phi the
// characteristic polynomial of the generator (tools/mt_jump_poly.py -> mt_jump_poly.h), every raw word J positions
// ahead is the XOR of the words of a 19937-word window selected by g:   w[t + J] = XOR_{g_i = 1} w[t + i],  t >= 1,
// so dst[k] = w[1 + k + J], k = 0..623, needs src[1 .. 20560] only.  That lets several single-CTA generators produce
// disjoint segments of the SAME stream concurrently: the sequential recurrence bounded the end-to-end rate.
// One CTA per 4 output words, threads over the polynomial's coefficients, XOR tree through shared memory.
constexpr int JUMP_OUT = 4, JUMP_THREADS = 256, JUMP_DEGREE = 19937;
__global__ void __launch_bounds__(JUMP_THREADS) k_mt_jump(const uint32_t* __restrict__ src, const uint32_t* __restrict__ poly,
                                                          uint32_t* __restrict__ dst, GenState* gnext) {
  __shared__ uint32_t red[JUMP_THREADS][JUMP_OUT + 1];
  const int k0 = blockIdx.x * JUMP_OUT;
  uint32_t acc[JUMP_OUT];
#pragma unroll
  for (int o = 0; o < JUMP_OUT; ++o) acc[o] = 0u;
  for (int i = threadIdx.x; i < JUMP_DEGREE; i += JUMP_THREADS) {
    if ((__ldg(poly + (i >> 5)) >> (i & 31)) & 1u) {
      const uint32_t* __restrict__ p = src + 1 + k0 + i;
#pragma unroll
      for (int o = 0; o < JUMP_OUT; ++o) acc[o] ^= __ldg(p + o);
    }
  }
#pragma unroll
  for (int o = 0; o < JUMP_OUT; ++o) red[threadIdx.x][o] = acc[o];
  __syncthreads();
  for (int stride = JUMP_THREADS / 2; stride > 0; stride >>= 1) {
    if ((int)threadIdx.x < stride) {
#pragma unroll
      for (int o = 0; o < JUMP_OUT; ++o) red[threadIdx.x][o] ^= red[threadIdx.x + stride][o];
    }
    __syncthreads();
  }
  if (threadIdx.x < JUMP_OUT && k0 + (int)threadIdx.x < 624) {
    const uint32_t v = red[0][threadIdx.x];
    dst[k0 + threadIdx.x] = v;
    gnext->key[k0 + threadIdx.x] = v;
  }
}

// numpy's random_double: (a >> 5, b >> 6) -> (a * 2^26 + b) / 2^53; the integer is below 2^53, so this is exact
__device__ __forceinline__ double to_double(uint32_t a, uint32_t b) {
  const unsigned long long v = ((unsigned long long)(a >> 5) << 26) | (unsigned long long)(b >> 6);
  return __dmul_rn((double)(long long)v, 1.0 / 9007199254740992.0);
}

// one polar attempt from 4 consecutive RAW words; numpy: x = 2.0 * double - 1.0; r2 = x1*x1 + x2*x2 (each rounded)
__device__ __forceinline__ bool polar(const uint32_t* __restrict__ w, double& x1, double& x2, double& r2) {
  const uint32_t a = temper(__ldg(w)), b = temper(__ldg(w + 1)), c = temper(__ldg(w + 2)), d = temper(__ldg(w + 3));
  x1 = __dsub_rn(__dmul_rn(2.0, to_double(a, b)), 1.0);
  x2 = __dsub_rn(__dmul_rn(2.0, to_double(c, d)), 1.0);
  r2 = __dadd_rn(__dmul_rn(x1, x1), __dmul_rn(x2, x2));
  return !(r2 >= 1.0 || r2 == 0.0);
}

__device__ __forceinline__ double polar_factor(double r2) {
  return __dsqrt_rn(__ddiv_rn(__dmul_rn(-2.0, qmcb_glibc_log(r2)), r2));
}

// word offset of attempt group `grp` of alignment class c (classes: offsets congruent to r0 and r0 + 2 mod 4)
__device__ __forceinline__ long long group_word(int r0, int c, long long grp) { return ((r0 + 2 * c) & 3) + 4 * grp; }

// Accept flags of the attempt groups [g_lo, g_hi) (multiples of 32) of both classes; every word they read exists.
__global__ void __launch_bounds__(THREADS) k_rng_flags(const uint32_t* __restrict__ W, uint32_t* __restrict__ F0,
                                                        uint32_t* __restrict__ F1, int r0, long long g_lo, long long g_hi) {
  const long long grp = g_lo + (long long)blockIdx.x * THREADS + threadIdx.x;
  const bool live = grp < g_hi;  // (g_hi - g_lo is a multiple of 32: whole warps are live or not)
  double x1, x2, r2;
  const bool ok0 = live && polar(W + group_word(r0, 0, grp), x1, x2, r2);
  const bool ok1 = live && polar(W + group_word(r0, 1, grp), x1, x2, r2);
  const unsigned m0 = __ballot_sync(0xffffffffu, ok0), m1 = __ballot_sync(0xffffffffu, ok1);
  if (live && (threadIdx.x & 31) == 0) {
    F0[grp >> 5] = m0;
    F1[grp >> 5] = m1;
  }
}

// exclusive scan of one int per thread over the CTA (blockDim = THREADS); total broadcast through *total
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
    const int s = lane < THREADS / 32 ? warp_sums[lane] : 0;
    int sinc = s;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      const int o = __shfl_up_sync(0xffffffffu, sinc, d);
      if (lane >= d) sinc += o;
    }
    if (lane < THREADS / 32) warp_sums[lane] = sinc - s;  // exclusive
    if (lane == 31) *total = sinc;
  }
  __syncthreads();
  const int pre = warp_sums[warp] + inc - v;
  __syncthreads();
  return pre;
}

// bitmap word `wi` of a draw that starts at group g0: bits before g0 and at / after `valid` are not attempts of it
__device__ __forceinline__ uint32_t draw_word(const uint32_t* __restrict__ F, long long wi, long long g0, long long valid) {
  if (wi * 32 >= valid) return 0u;
  uint32_t w = __ldg(F + wi);
  if (wi == (g0 >> 5)) w &= ~((1u << (g0 & 31)) - 1u);
  if ((wi + 1) * 32 > valid) w &= (1u << (valid & 31)) - 1u;
  return w;
}

// One CTA walks the draw program.  `valid` = attempt groups whose flags exist; `nwords` = words generated.
__global__ void __launch_bounds__(THREADS) k_rng_plan(const Op* __restrict__ ops, int nops, State* st,
                                                       const uint32_t* __restrict__ W, const uint32_t* __restrict__ F0,
                                                       const uint32_t* __restrict__ F1, int r0, long long valid,
                                                       long long nwords, OpPlan* __restrict__ plan) {
  constexpr int TILE = 256;
  __shared__ Op s_ops[TILE];
  __shared__ int warp_sums[32];
  __shared__ int s_total;
  __shared__ long long s_end_group;
  __shared__ double s_cached_out;
  const int t = threadIdx.x;
  // every thread carries the same running scalars (uniform control flow)
  long long cur = st->cur;
  int has = st->has_gauss;
  double cached = st->cached;
  int error = st->error;
  for (int base = 0; base < nops && !error; base += TILE) {
    __syncthreads();
    if (base + t < nops) s_ops[t] = ops[base + t];
    __syncthreads();
    const int tile_n = min(TILE, nops - base);
    for (int i = 0; i < tile_n && !error; ++i) {
      const Op o = s_ops[i];
      const long long n = o.kind == KIND_ROTATION ? 4 : o.count;
      if (t == 0) {
        OpPlan p;
        p.start_word = cur;
        p.has_in = has;
        p.cached_in = cached;
        p.pairs = 0;
        p.pad = 0;
        if (o.kind != KIND_UNIFORM) {
          const long long need0 = n - ((has && n > 0) ? 1 : 0);
          p.pairs = (need0 + 1) / 2;
        }
        plan[base + i] = p;
      }
      if (o.kind == KIND_UNIFORM) {
        cur += 2 * n;
        if (cur > nwords) error = 1;
        continue;
      }
      long long need = n;
      if (has && n > 0) {  // the cached value is handed out first
        need -= 1;
        has = 0;
        cached = 0.0;
      }
      const long long m = (need + 1) / 2;
      if (m == 0) continue;
      const bool odd = (2 * m - need) == 1;  // the second value of the last pair stays cached
      const int c = ((cur - r0) & 3) ? 1 : 0;
      const uint32_t* F = c ? F1 : F0;
      const long long g0 = (cur - ((r0 + 2 * c) & 3)) >> 2;
      long long accepted = 0;
      long long wi0 = g0 >> 5;
      while (true) {
        const long long wi = wi0 + t;
        const uint32_t w = draw_word(F, wi, g0, valid);
        const int cnt = __popc(w);
        const int pre = block_scan(cnt, warp_sums, &s_total);
        const int total = s_total;
        const long long rem = m - accepted;
        if (total >= rem && cnt > 0 && pre < rem && rem <= pre + cnt) {
          const int bit = __fns(w, 0, (int)(rem - pre));  // position of the (rem - pre)-th set bit
          const long long ge = wi * 32 + bit;
          s_end_group = ge;
          if (odd) {
            double x1, x2, r2;
            polar(W + group_word(r0, c, ge), x1, x2, r2);
            s_cached_out = __dmul_rn(polar_factor(r2), x1);
          }
        }
        __syncthreads();
        if (total >= rem) break;
        accepted += total;
        wi0 += THREADS;
        if (wi0 * 32 >= valid) {  // ran out of generated flags before the draw completed
          error = 1;
          break;
        }
      }
      if (error) break;
      cur = group_word(r0, c, s_end_group + 1);
      if (odd) {
        has = 1;
        cached = s_cached_out;
      }
      __syncthreads();  // s_end_group / s_cached_out are rewritten by the next normal draw
    }
  }
  if (t == 0) {
    st->cur = cur;
    st->has_gauss = has;
    st->cached = cached;
    st->error = error;
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

__global__ void __launch_bounds__(THREADS) k_rng_fill(const Op* __restrict__ ops, const OpPlan* __restrict__ plan,
                                                       const Work* __restrict__ work, const State* st,
                                                       const uint32_t* __restrict__ W, const uint32_t* __restrict__ F0,
                                                       const uint32_t* __restrict__ F1, int r0, long long valid) {
  __shared__ int warp_sums[32];
  __shared__ int s_total;
  __shared__ double s_q[4];
  if (st->error) return;
  const Work wk = work[blockIdx.x];
  const Op o = ops[wk.op];
  const OpPlan p = plan[wk.op];
  const int t = threadIdx.x;
  if (o.kind == KIND_UNIFORM) {
    const long long i0 = (long long)wk.chunk * CHUNK_UNIFORM;
    if (i0 >= o.count) return;
#pragma unroll
    for (int q = 0; q < CHUNK_UNIFORM / THREADS; ++q) {
      const long long i = i0 + (long long)q * THREADS + t;
      if (i < o.count) {
        const long long w0 = p.start_word + 2 * i;
        o.dst[i] = to_double(temper(__ldg(W + w0)), temper(__ldg(W + w0 + 1)));
      }
    }
    return;
  }
  const long long n = o.kind == KIND_ROTATION ? 4 : o.count;
  double* dst = o.kind == KIND_ROTATION ? s_q : o.dst;
  const double scale = o.kind == KIND_ROTATION ? 1.0 : o.scale;
  const int shift = p.has_in && n > 0 ? 1 : 0;
  if (wk.chunk == 0 && t == 0 && shift) dst[0] = __dadd_rn(0.0, __dmul_rn(scale, p.cached_in));
  if (p.pairs == 0) return;  // (a rotation always needs pairs: n = 4)
  const long long cur = p.start_word;
  const int c = ((cur - r0) & 3) ? 1 : 0;
  const uint32_t* F = c ? F1 : F0;
  const long long g0 = (cur - ((r0 + 2 * c) & 3)) >> 2;
  // accepted pairs of this draw before this chunk
  long long before = 0;
  for (int cc = 0; cc < wk.chunk; ++cc) {
    const int cnt = __popc(draw_word(F, (g0 >> 5) + (long long)cc * THREADS + t, g0, valid));
    block_scan(cnt, warp_sums, &s_total);
    before += s_total;
    __syncthreads();
  }
  if (before >= p.pairs) return;  // the draw ended in an earlier chunk (uniform across the CTA)
  const long long wi = (g0 >> 5) + (long long)wk.chunk * THREADS + t;
  uint32_t w = draw_word(F, wi, g0, valid);
  const int pre = block_scan(__popc(w), warp_sums, &s_total);
  long long pair = before + pre;
  while (w && pair < p.pairs) {
    const int bit = __ffs(w) - 1;
    w &= w - 1;
    double x1, x2, r2;
    polar(W + group_word(r0, c, wi * 32 + bit), x1, x2, r2);
    const double f = polar_factor(r2);
    const long long slot = shift + 2 * pair;
    dst[slot] = __dadd_rn(0.0, __dmul_rn(scale, __dmul_rn(f, x2)));
    if (slot + 1 < n) dst[slot + 1] = __dadd_rn(0.0, __dmul_rn(scale, __dmul_rn(f, x1)));
    ++pair;
  }
  if (o.kind == KIND_ROTATION) {
    __syncthreads();
    if (t == 0) write_rotation(s_q, o.dst);
  }
}

// Restart the buffer at the state block the stream is in: block bi moves to the front, the offset follows
// (624 is a multiple of 4, so the alignment classes keep their meaning).
__global__ void __launch_bounds__(320) k_rng_rebase(State* st, GenState* g, uint32_t* __restrict__ W) {
  __shared__ long long s_bi;
  // (cur - 1) / 624: an offset exactly on a block boundary stays "position 624 of the previous block", which is
  // what numpy reports after consuming a whole block
  if (threadIdx.x == 0) s_bi = st->cur > 0 ? (st->cur - 1) / 624 : 0;
  __syncthreads();
  const long long bi = s_bi;
  uint32_t v[2];
  int k = 0;
  for (int i = threadIdx.x; i < 624; i += 320) v[k++] = W[bi * 624 + i];
  __syncthreads();
  k = 0;
  for (int i = threadIdx.x; i < 624; i += 320) {
    W[i] = v[k];
    g->key[i] = v[k++];
  }
  if (threadIdx.x == 0) st->cur -= bi * 624;
}

// numpy's view of the generator after the programs run so far: the raw state block the stream stopped in and
// the position in it (1..624 after any draw: numpy regenerates lazily, on the next draw).
__global__ void __launch_bounds__(320) k_rng_finalize(State* st, const uint32_t* __restrict__ W) {
  __shared__ long long s_block;
  if (threadIdx.x == 0) {
    const long long cur = st->cur;
    long long b = cur / 624;
    int pos = (int)(cur % 624);
    if (pos == 0 && b > 0) {
      b -= 1;
      pos = 624;
    }
    s_block = b;
    st->pos = pos;
  }
  __syncthreads();
  for (int i = threadIdx.x; i < 624; i += 320) st->key[i] = W[s_block * 624 + i];
}

}  // namespace devrng
