// legacy_rng.cpp -- host-side generator of the random variates one VMC block consumes, in the
// exact order and with the exact arithmetic of the reference's use of the global legacy
// numpy RandomState (MT19937):
//   per step, per electron:  np.random.normal(scale=sqrt(tstep), size=(N,3)); np.random.rand(N)
//                                                                   (pyqmc/method/mc.py:119,132)
//   then per electron, per ECP atom: np.random.random(N); scipy Rotation.random().as_matrix()
//                                                        (pyqmc/observables/eval_ecp.py:145,263)
// The caller passes the state obtained from np.random.get_state() and writes the advanced state
// back with np.random.set_state(), so seeded runs stay bit-identical to the reference's stream
// while the draws run ~5x faster than through numpy (no per-call overhead, fused loops).
//
// Algorithms (public, restated): MT19937 (Matsumoto & Nishimura 1998) with numpy's 53-bit double
// construction (a>>5, b>>6); numpy's legacy Gaussian = Marsaglia polar method with one cached
// value; scipy's random rotation = normalised 4-vector of standard normals read as a quaternion
// (x, y, z, w) converted to a matrix.
// numpy rounds every operation separately: no FMA contraction anywhere in this file
#pragma GCC optimize("fp-contract=off")
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstdlib>
#include <atomic>
#if defined(__x86_64__)
#include <immintrin.h>
#endif
#include <chrono>
#include <cstdio>
#include <cstring>
#include <memory>
#include <thread>
#include <vector>

namespace {

// The MT19937 recurrence is inherently sequential, so it runs on its own producer thread that
// stays a few blocks ahead: each ring slot holds the tempered image (the outputs) of one 624-word
// state block.  The consumer (phase A below) walks the tempered blocks in order; the raw state handed
// back to numpy at the end is recovered by inverting the tempering of the block the consumer stopped in
// (copying the raw block into the ring as well cost more than the recurrence and the tempering together).
// With one thread the blocks are generated inline.
struct Ring {
  static constexpr int K = 64;
  uint32_t tb[K][624];
  std::atomic<long> produced{0};  // blocks 0..produced-1 are ready
  std::atomic<long> consumed{0};  // the consumer is working on block `consumed`
  std::atomic<bool> stop{false};
  bool threaded = false;
  uint32_t key[624];  // producer-private running state
};

struct MT {
  const uint32_t* tb;
  int pos;
  long blk;
  Ring* ring;
};

void mt_reload(uint32_t* __restrict__ mt);
void temper_raw(const uint32_t* key, uint32_t* out);

inline void produce_block(Ring& r, long index) {
  mt_reload(r.key);
  const int slot = (int)(index % Ring::K);
  temper_raw(r.key, r.tb[slot]);
}

void producer_loop(Ring* r) {
  long next = 1;  // block 0 is the state handed in by the caller
  while (!r->stop.load(std::memory_order_acquire)) {
    if (next - r->consumed.load(std::memory_order_acquire) < Ring::K - 1) {
      produce_block(*r, next);
      ++next;
      r->produced.store(next, std::memory_order_release);
    } else {
      std::this_thread::yield();
    }
  }
}

inline void refill(MT& s) {
  Ring& r = *s.ring;
  const long want = s.blk + 1;
  if (r.threaded) {
    while (r.produced.load(std::memory_order_acquire) <= want) {
    }
  } else {
    produce_block(r, want);
    r.produced.store(want + 1, std::memory_order_relaxed);
  }
  s.blk = want;
  s.tb = r.tb[want % Ring::K];
  s.pos = 0;
  r.consumed.store(want, std::memory_order_release);
}

// One MT19937 state transition (624 new words).  The recurrence reads mt[kk+1] (not yet rewritten)
// and mt[kk+397 mod 624] (rewritten >= 227 iterations earlier), so blocks of up to 227 iterations are
// independent: the loops are marked ivdep and split so the compiler vectorises them.
__attribute__((target_clones("avx512f", "avx2", "default"))) void mt_reload(uint32_t* __restrict__ mt) {
  constexpr int N = 624, M = 397;
  constexpr uint32_t UPPER = 0x80000000u, LOWER = 0x7fffffffu, MATRIX_A = 0x9908b0dfu;
  uint32_t nxt[N + 1];
  std::memcpy(nxt, mt + 1, (N - 1) * sizeof(uint32_t));  // old mt[kk+1] for kk < N-1
#pragma GCC ivdep
  for (int kk = 0; kk < N - M; kk++) {
    const uint32_t y = (mt[kk] & UPPER) | (nxt[kk] & LOWER);
    mt[kk] = mt[kk + M] ^ (y >> 1) ^ ((0u - (y & 1u)) & MATRIX_A);
  }
  // kk in [227, 454): reads mt[kk-227] (new values written by the loop above)
#pragma GCC ivdep
  for (int kk = N - M; kk < 2 * (N - M); kk++) {
    const uint32_t y = (mt[kk] & UPPER) | (nxt[kk] & LOWER);
    mt[kk] = mt[kk - (N - M)] ^ (y >> 1) ^ ((0u - (y & 1u)) & MATRIX_A);
  }
  // kk in [454, 623): reads mt[kk-227] written by the second loop
#pragma GCC ivdep
  for (int kk = 2 * (N - M); kk < N - 1; kk++) {
    const uint32_t y = (mt[kk] & UPPER) | (nxt[kk] & LOWER);
    mt[kk] = mt[kk - (N - M)] ^ (y >> 1) ^ ((0u - (y & 1u)) & MATRIX_A);
  }
  const uint32_t y = (mt[N - 1] & UPPER) | (mt[0] & LOWER);
  mt[N - 1] = mt[M - 1] ^ (y >> 1) ^ ((0u - (y & 1u)) & MATRIX_A);
}

__attribute__((target_clones("avx512f", "avx2", "default"))) void temper_raw(const uint32_t* key, uint32_t* out) {
  for (int i = 0; i < 624; ++i) {
    uint32_t y = key[i];
    y ^= (y >> 11);
    y ^= (y << 7) & 0x9d2c5680u;
    y ^= (y << 15) & 0xefc60000u;
    y ^= (y >> 18);
    out[i] = y;
  }
}

// inverse of temper_raw (each of the four xor-shift steps is a bijection on 32-bit words)
inline void untemper_block(const uint32_t* t, uint32_t* key) {
  for (int i = 0; i < 624; ++i) {
    uint32_t y = t[i];
    y ^= y >> 18;
    y ^= (y << 15) & 0xefc60000u;
    uint32_t x = y;
    for (int k = 0; k < 4; ++k) x = y ^ ((x << 7) & 0x9d2c5680u);
    y = x;
    x = y ^ (y >> 11);
    x = y ^ (x >> 11);
    key[i] = x;
  }
}

inline uint32_t next32(MT& s) {
  if (s.pos == 624) refill(s);
  return s.tb[s.pos++];
}

inline double next_double(MT& s) {
  const int32_t a = next32(s) >> 5, b = next32(s) >> 6;
  return (a * 67108864.0 + b) / 9007199254740992.0;
}

// 53-bit double from two consecutive outputs: ((a>>5) * 2^26 + (b>>6)) / 2^53.  The integer
// (a>>5)<<26 | (b>>6) is below 2^53, so converting it and scaling by 2^-53 gives the same bits.
void convert_pairs_scalar(const uint32_t* t, double* out, int64_t m) {
  for (int64_t k = 0; k < m; ++k) {
    const int32_t a = t[2 * k] >> 5, b = t[2 * k + 1] >> 6;
    out[k] = (a * 67108864.0 + b) / 9007199254740992.0;
  }
}

#if defined(__x86_64__)
__attribute__((target("avx512f,avx512dq"))) void convert_pairs_avx512(const uint32_t* t, double* out, int64_t m) {
  int64_t k = 0;
  const __m512d scale = _mm512_set1_pd(1.0 / 9007199254740992.0);
  for (; k + 8 <= m; k += 8) {
    const __m512i v = _mm512_loadu_si512(t + 2 * k);                                  // 8 x (lo = a, hi = b)
    const __m512i a = _mm512_srli_epi64(_mm512_and_si512(v, _mm512_set1_epi64(0xffffffffLL)), 5);
    const __m512i b = _mm512_srli_epi64(v, 32 + 6);
    const __m512i x = _mm512_or_si512(_mm512_slli_epi64(a, 26), b);
    _mm512_storeu_pd(out + k, _mm512_mul_pd(_mm512_cvtepi64_pd(x), scale));
  }
  convert_pairs_scalar(t + 2 * k, out + k, m - k);
}

// Polar attempts on 4-output groups: x1 = 2 d1 - 1, x2 = 2 d2 - 1, r2 = x1^2 + x2^2; accepted
// (r2 < 1 and r2 != 0) triples are compress-stored.  Returns the number of attempts consumed; never
// consumes past the attempt that completes `need` accepted pairs.
__attribute__((target("avx512f,avx512dq"))) int polar_block_avx512(const uint32_t* t, int avail, int64_t need,
                                                                    double* X1, double* X2, double* R2,
                                                                    int64_t* naccepted) {
  const __m512d scale = _mm512_set1_pd(1.0 / 9007199254740992.0);
  const __m512d two = _mm512_set1_pd(2.0), one = _mm512_set1_pd(1.0), zero = _mm512_setzero_pd();
  const __m512i lomask = _mm512_set1_epi64(0xffffffffLL);
  // gather indices: attempt j uses outputs 4j..4j+3 -> as 64-bit lanes: d1 = lane 2j, d2 = lane 2j+1
  const __m512i idx1 = _mm512_setr_epi64(0, 2, 4, 6, 8, 10, 12, 14);
  const __m512i idx2 = _mm512_setr_epi64(1, 3, 5, 7, 9, 11, 13, 15);
  int used = 0;
  int64_t np_ = 0;
  while (used + 8 <= avail && np_ + 8 <= need) {
    const __m512i v0 = _mm512_loadu_si512(t + 4 * used);       // attempts 0..3
    const __m512i v1 = _mm512_loadu_si512(t + 4 * used + 16);  // attempts 4..7
    const __m512i p1 = _mm512_permutex2var_epi64(v0, idx1, v1);  // the 8 (a,b) pairs of d1
    const __m512i p2 = _mm512_permutex2var_epi64(v0, idx2, v1);  // the 8 (a,b) pairs of d2
    const __m512i xa = _mm512_or_si512(_mm512_slli_epi64(_mm512_srli_epi64(_mm512_and_si512(p1, lomask), 5), 26),
                                       _mm512_srli_epi64(p1, 38));
    const __m512i xb = _mm512_or_si512(_mm512_slli_epi64(_mm512_srli_epi64(_mm512_and_si512(p2, lomask), 5), 26),
                                       _mm512_srli_epi64(p2, 38));
    const __m512d d1 = _mm512_mul_pd(_mm512_cvtepi64_pd(xa), scale);
    const __m512d d2 = _mm512_mul_pd(_mm512_cvtepi64_pd(xb), scale);
    // 2.0*d - 1.0 and x1*x1 + x2*x2 with separate roundings (no FMA), as the scalar code
    const __m512d x1 = _mm512_sub_pd(_mm512_mul_pd(two, d1), one);
    const __m512d x2 = _mm512_sub_pd(_mm512_mul_pd(two, d2), one);
    const __m512d r2 = _mm512_add_pd(_mm512_mul_pd(x1, x1), _mm512_mul_pd(x2, x2));
    const __mmask8 ok = _mm512_cmp_pd_mask(r2, one, _CMP_LT_OQ) & _mm512_cmp_pd_mask(r2, zero, _CMP_NEQ_OQ);
    _mm512_mask_compressstoreu_pd(X1 + np_, ok, x1);
    _mm512_mask_compressstoreu_pd(X2 + np_, ok, x2);
    _mm512_mask_compressstoreu_pd(R2 + np_, ok, r2);
    np_ += __builtin_popcount((unsigned)ok);
    used += 8;
  }
  *naccepted = np_;
  return used;
}
#endif

static bool have_avx512() {
#if defined(__x86_64__)
  static const bool ok = __builtin_cpu_supports("avx512f") && __builtin_cpu_supports("avx512dq") &&
                         std::getenv("QMCB_RNG_NOAVX512") == nullptr;
  return ok;
#else
  return false;
#endif
}

inline void convert_pairs(const uint32_t* t, double* out, int64_t m) {
#if defined(__x86_64__)
  if (have_avx512()) return convert_pairs_avx512(t, out, m);
#endif
  convert_pairs_scalar(t, out, m);
}

inline void fill_uniform(MT& s, double* out, int64_t n) {
  int64_t i = 0;
  while (i < n) {
    if (s.pos >= 623) {  // block boundary: scalar path
      out[i++] = next_double(s);
      continue;
    }
    const int64_t m = std::min<int64_t>(n - i, (624 - s.pos) / 2);
    convert_pairs(s.tb + s.pos, out + i, m);
    s.pos += (int)(2 * m);
    i += m;
  }
}

struct Gauss {
  int has;
  double cached;
};

// one Marsaglia-polar attempt: consumes exactly two doubles, accepted or not
inline bool polar_attempt(MT& s, double& x1, double& x2, double& r2) {
  x1 = 2.0 * next_double(s) - 1.0;
  x2 = 2.0 * next_double(s) - 1.0;
  r2 = x1 * x1 + x2 * x2;
  return !(r2 >= 1.0 || r2 == 0.0);
}

inline double legacy_gauss(MT& s, Gauss& g) {
  if (g.has) {
    const double t = g.cached;
    g.has = 0;
    g.cached = 0.0;
    return t;
  }
  double x1, x2, r2;
  while (!polar_attempt(s, x1, x2, r2)) {
  }
  const double f = std::sqrt(-2.0 * std::log(r2) / r2);
  g.cached = f * x1;
  g.has = 1;
  return f * x2;
}

// Gaussian "slots": the legacy generator hands out, in order, [cached value if any], then for
// each accepted polar pair p the two values f_p*x2_p, f_p*x1_p.  Phase A (sequential, cheap) runs
// the MT stream and records the accepted pairs plus which slot range each destination consumes;
// phase B (parallel) evaluates f_p = sqrt(-2 log(r2)/r2) once per pair and fills destinations.
struct Segment {
  double* dst;
  int64_t n, first_slot;
  double scale;
  int is_rotation;
};

// uninitialised growable buffer (std::vector::resize would zero-fill ~16 MB per block)
struct RawBuf {
  double* p = nullptr;
  size_t cap = 0;
  double* data() { return p; }
  size_t size() const { return cap; }
  void resize(size_t n) {
    if (n <= cap) return;
    double* q = static_cast<double*>(std::malloc(n * sizeof(double)));
    if (p) {
      std::memcpy(q, p, cap * sizeof(double));
      std::free(p);
    }
    p = q;
    cap = n;
  }
  double& operator[](size_t i) { return p[i]; }
  const double& operator[](size_t i) const { return p[i]; }
};

struct PairBuffers {
  RawBuf x1, x2, r2, f;
};

struct GaussPlan {
  RawBuf &x1, &x2, &r2, &f;
  explicit GaussPlan(PairBuffers& b) : x1(b.x1), x2(b.x2), r2(b.r2), f(b.f) {}
  std::vector<Segment> segs;
  int64_t npairs = 0;      // accepted pairs generated so far
  int64_t nslots = 0;      // slots handed out so far
  int64_t slot_shift = 0;  // 1 if slot 0 is the cached value carried in from the caller
  double carried = 0.0;

  // reserve n consecutive slots, generating pairs as needed
  int64_t take(MT& s, int64_t n) {
    const int64_t first = nslots;
    nslots += n;
    const int64_t need_pairs = (nslots - slot_shift + 1) / 2;
    if ((int64_t)r2.size() < need_pairs + 16) {  // slack: the branch-free loops store before they test
      x1.resize(need_pairs + 16);
      x2.resize(need_pairs + 16);
      r2.resize(need_pairs + 16);
    }
    double* X1 = x1.data();
    double* X2 = x2.data();
    double* R2 = r2.data();
    while (npairs < need_pairs) {
      // bulk path: all complete 4-output attempts left in the tempered block, branch-free
      // compaction (every attempt consumes exactly four outputs, accepted or not)
      const int avail = (624 - s.pos) / 4;
      const int64_t want = need_pairs - npairs;
      if (avail > 0) {
        const uint32_t* t = s.tb + s.pos;
#if defined(__x86_64__)
        if (have_avx512() && avail >= 8 && want >= 8) {
          int64_t got = 0;
          const int u8 = polar_block_avx512(t, avail, want, X1 + npairs, X2 + npairs, R2 + npairs, &got);
          if (u8 > 0) {
            npairs += got;
            s.pos += 4 * u8;
            continue;
          }
        }
#endif
        int used = 0;
        int64_t np_ = npairs;
        // stop as soon as enough pairs are accepted: later attempts belong to the next consumer
        for (; used < avail && np_ < need_pairs; ++used) {
          const double d1 = ((int32_t)(t[4 * used] >> 5) * 67108864.0 + (int32_t)(t[4 * used + 1] >> 6)) / 9007199254740992.0;
          const double d2 = ((int32_t)(t[4 * used + 2] >> 5) * 67108864.0 + (int32_t)(t[4 * used + 3] >> 6)) / 9007199254740992.0;
          const double a = 2.0 * d1 - 1.0, b = 2.0 * d2 - 1.0;
          const double c = a * a + b * b;
          X1[np_] = a;
          X2[np_] = b;
          R2[np_] = c;
          np_ += (c < 1.0) & (c != 0.0);
        }
        (void)want;
        npairs = np_;
        s.pos += 4 * used;
        continue;
      }
      double a, b, c;
      if (!polar_attempt(s, a, b, c)) continue;  // attempt straddling a block boundary
      X1[npairs] = a;
      X2[npairs] = b;
      R2[npairs] = c;
      ++npairs;
    }
    return first;
  }
  inline double slot_value(int64_t t) const {
    if (t < slot_shift) return carried;
    const int64_t u = t - slot_shift;
    const int64_t p = u >> 1;
    return (u & 1) ? f[p] * x1[p] : f[p] * x2[p];
  }
};

void write_rotation(const double* q, double* m);

void run_phase_b(GaussPlan& plan, int nthreads) {
  const int64_t np = plan.npairs;
  plan.f.resize(np);
  auto work_f = [&](int64_t lo, int64_t hi) {
    for (int64_t p = lo; p < hi; ++p) plan.f[p] = std::sqrt(-2.0 * std::log(plan.r2[p]) / plan.r2[p]);
  };
  auto work_seg = [&](size_t lo, size_t hi) {
    for (size_t k = lo; k < hi; ++k) {
      const Segment& sg = plan.segs[k];
      if (sg.is_rotation) {
        double q[4];
        for (int i = 0; i < 4; ++i) q[i] = 0.0 + 1.0 * plan.slot_value(sg.first_slot + i);
        write_rotation(q, sg.dst);
      } else {
        for (int64_t i = 0; i < sg.n; ++i) sg.dst[i] = 0.0 + sg.scale * plan.slot_value(sg.first_slot + i);
      }
    }
  };
  if (nthreads <= 1 || np < 4096) {
    work_f(0, np);
    work_seg(0, plan.segs.size());
    return;
  }
  std::vector<std::thread> th;
  for (int t = 0; t < nthreads; ++t)
    th.emplace_back([&, t]() { work_f(np * t / nthreads, np * (t + 1) / nthreads); });
  for (auto& x : th) x.join();
  th.clear();
  const size_t ns = plan.segs.size();
  for (int t = 0; t < nthreads; ++t)
    th.emplace_back([&, t]() { work_seg(ns * t / nthreads, ns * (t + 1) / nthreads); });
  for (auto& x : th) x.join();
}

// scipy.spatial.transform.Rotation.random(): q = normal(size=4); q /= |q|; as_matrix()
void write_rotation(const double* q, double* m) {
  const double norm = std::sqrt(q[0] * q[0] + q[1] * q[1] + q[2] * q[2] + q[3] * q[3]);
  const double x = q[0] / norm, y = q[1] / norm, z = q[2] / norm, w = q[3] / norm;
  const double x2 = x * x, y2 = y * y, z2 = z * z, w2 = w * w;
  const double xy = x * y, zw = z * w, xz = x * z, yw = y * w, yz = y * z, xw = x * w;
  m[0] = x2 - y2 - z2 + w2;
  m[3] = 2 * (xy + zw);
  m[6] = 2 * (xz - yw);
  m[1] = 2 * (xy - zw);
  m[4] = -x2 + y2 - z2 + w2;
  m[7] = 2 * (yz + xw);
  m[2] = 2 * (xz + yw);
  m[5] = 2 * (yz - xw);
  m[8] = -x2 - y2 + z2 + w2;
}

}  // namespace

struct RngPlan {
  PairBuffers buffers;
  GaussPlan plan;
  RngPlan() : plan(buffers) {}
  void reset() {
    plan.segs.clear();
    plan.npairs = 0;
    plan.nslots = 0;
    plan.slot_shift = 0;
    plan.carried = 0.0;
  }
};

// Phase A: walks the MT19937 stream (sequential), fills the uniform outputs, records the accepted
// polar pairs and the destination segments of every Gaussian; advances the state.  `body(s, plan)`
// issues the draws in the caller's order through fill_uniform / plan.take.
template <class Body>
int rng_run_phase_a(RngPlan& rp, uint32_t* key, int32_t* pos, int32_t* has_gauss, double* cached_gauss, int64_t est_pairs,
                    int nthreads, Body body) {
  if (*pos < 0 || *pos > 624) return -1;
  std::unique_ptr<Ring> ring(new Ring());
  std::memcpy(ring->key, key, sizeof(ring->key));
  temper_raw(key, ring->tb[0]);
  ring->produced.store(1);
  // A separate producer thread for the recurrence pays off only where handing 2.5 KB blocks from core to
  // core is cheap; on the B200 hosts (measured, profiles/rng_host_timing.py) generating the blocks inline
  // is 1.5x faster (0.22 vs 0.33 ms per C2 step), so the thread is opt-in.
  const bool use_producer = std::getenv("QMCB_RNG_PRODUCER") != nullptr;
  ring->threaded = nthreads > 1 && use_producer;
  std::thread producer;
  if (ring->threaded) producer = std::thread(producer_loop, ring.get());
  MT s{ring->tb[0], *pos, 0, ring.get()};
  rp.reset();
  GaussPlan& plan = rp.plan;
  if (*has_gauss) {
    plan.slot_shift = 1;
    plan.carried = *cached_gauss;
  }
  plan.x1.resize(est_pairs);
  plan.x2.resize(est_pairs);
  plan.r2.resize(est_pairs);
  body(s, plan);
  if (ring->threaded) {
    ring->stop.store(true, std::memory_order_release);
    producer.join();
  }
  // state of the legacy Gaussian cache after the last slot handed out: the second value of the
  // last pair stays cached when an odd number of values was taken from the generated pairs
  const int64_t used = plan.nslots - plan.slot_shift;
  if (plan.nslots == 0) {
    // nothing consumed: cache unchanged
  } else if (used <= 0) {
    *has_gauss = 0;
    *cached_gauss = 0.0;
  } else if (used & 1) {
    const int64_t p = used >> 1;
    const double f = std::sqrt(-2.0 * std::log(plan.r2[p]) / plan.r2[p]);
    *has_gauss = 1;
    *cached_gauss = f * plan.x1[p];
  } else {
    *has_gauss = 0;
    *cached_gauss = 0.0;
  }
  untemper_block(ring->tb[s.blk % Ring::K], key);
  *pos = s.pos;
  return 0;
}

// one VMC block: per step and electron normal(N,3), rand(N); then per electron and ECP atom
// random(N) and a rotation (mc.py:119,132; eval_ecp.py:145,263)
int rng_phase_a(RngPlan& rp, uint32_t* key, int32_t* pos, int32_t* has_gauss, double* cached_gauss, int nsteps, int ne,
                int64_t N, int necp, double scale, double* gauss, double* unif, double* ecp_u, double* ecp_rot,
                int nthreads) {
  const int64_t est = (int64_t)nsteps * ne * (N * 3 / 2 + 1 + 2 * (ecp_u ? necp : 0)) + 64;
  return rng_run_phase_a(rp, key, pos, has_gauss, cached_gauss, est, nthreads, [&](MT& s, GaussPlan& plan) {
    for (int step = 0; step < nsteps; ++step) {
      for (int e = 0; e < ne; ++e) {
        double* go = gauss + ((int64_t)step * ne + e) * N * 3;
        plan.segs.push_back(Segment{go, N * 3, plan.take(s, N * 3), scale, 0});
        fill_uniform(s, unif + ((int64_t)step * ne + e) * N, N);
      }
      if (ecp_u) {
        for (int e = 0; e < ne; ++e)
          for (int a = 0; a < necp; ++a) {
            const int64_t ea = ((int64_t)step * ne + e) * necp + a;
            fill_uniform(s, ecp_u + ea * N, N);
            plan.segs.push_back(Segment{ecp_rot + ea * 9, 4, plan.take(s, 4), 1.0, 1});
          }
      }
    }
  });
}

extern "C" {

// handles for the two-stage host pipeline: phase A of block b+1 may run while phase B of block b
// (the log/sqrt of its accepted pairs) is still in flight on other threads
void* qmcb_rng_plan_create(void) { return new RngPlan(); }
void qmcb_rng_plan_destroy(void* p) { delete static_cast<RngPlan*>(p); }

int qmcb_rng_phase_a(void* plan, uint32_t* key, int32_t* pos, int32_t* has_gauss, double* cached_gauss, int nsteps,
                     int ne, int64_t N, int necp, double scale, double* gauss, double* unif, double* ecp_u,
                     double* ecp_rot, int nthreads) {
  return rng_phase_a(*static_cast<RngPlan*>(plan), key, pos, has_gauss, cached_gauss, nsteps, ne, N, necp, scale, gauss,
                     unif, ecp_u, ecp_rot, nthreads);
}

int qmcb_rng_phase_b(void* plan, int nthreads) {
  run_phase_b(static_cast<RngPlan*>(plan)->plan, nthreads);
  return 0;
}

int qmcb_rng_vmc_block(uint32_t* key, int32_t* pos, int32_t* has_gauss, double* cached_gauss, int nsteps,
                       int ne, int64_t N, int necp, double scale, double* gauss, double* unif, double* ecp_u,
                       double* ecp_rot, int nthreads) {
  static thread_local RngPlan rp;  // reused across blocks by the drawing thread
  const bool dbg = std::getenv("QMCB_RNG_DEBUG") != nullptr;
  auto t0 = std::chrono::steady_clock::now();
  const int rc = rng_phase_a(rp, key, pos, has_gauss, cached_gauss, nsteps, ne, N, necp, scale, gauss, unif, ecp_u,
                             ecp_rot, nthreads);
  if (rc) return rc;
  auto tA = std::chrono::steady_clock::now();
  run_phase_b(rp.plan, nthreads);
  if (dbg) {
    auto tB = std::chrono::steady_clock::now();
    std::fprintf(stderr, "[rng] phase A %.3f ms, phase B %.3f ms (%d threads, %lld pairs)\n",
                 std::chrono::duration<double, std::milli>(tA - t0).count(),
                 std::chrono::duration<double, std::milli>(tB - tA).count(), nthreads, (long long)rp.plan.npairs);
  }
  return 0;
}

// Generic draw program on the same generator: op i is kind[i] = 0 uniform doubles (np.random.random /
// rand), 1 normals scaled by scale[i] (np.random.normal / randn), 2 one scipy Rotation.random()
// matrix (4 normals -> dst[9]); count[i] values go to dst[i].  Used for the DMC block, whose draw
// order (dmc.py:150-198) interleaves T-move, diffusion and energy variates.
int qmcb_rng_program(uint32_t* key, int32_t* pos, int32_t* has_gauss, double* cached_gauss, int64_t nops,
                     const int32_t* kind, const int64_t* count, const uint64_t* dst, const double* scale, int nthreads) {
  static thread_local RngPlan rp;
  int64_t est = 64;
  for (int64_t i = 0; i < nops; ++i) est += kind[i] == 1 ? count[i] / 2 + 1 : (kind[i] == 2 ? 3 : 0);
  const int rc = rng_run_phase_a(rp, key, pos, has_gauss, cached_gauss, est, nthreads, [&](MT& s, GaussPlan& plan) {
    for (int64_t i = 0; i < nops; ++i) {
      double* d = reinterpret_cast<double*>(static_cast<uintptr_t>(dst[i]));
      if (kind[i] == 0)
        fill_uniform(s, d, count[i]);
      else if (kind[i] == 1)
        plan.segs.push_back(Segment{d, count[i], plan.take(s, count[i]), scale[i], 0});
      else
        plan.segs.push_back(Segment{d, 4, plan.take(s, 4), 1.0, 1});
    }
  });
  if (rc) return rc;
  run_phase_b(rp.plan, nthreads);
  return 0;
}

// Stochastic comb of the DMC branching step (dmc.py:358-366) in the reference's arithmetic: ladder = cumsum(weights)
// (sequential adds, as numpy), total = ladder[-1], teeth = (offset * total + linspace(0, total, n, endpoint=False)) %
// total with linspace = i * (total / n) + 0.0, and picked = searchsorted(ladder, teeth) (side = left).  The teeth wrap
// at most once (0 <= offset < 1), so x % total is x - total for x >= total (exact, like fmod) and both runs are ascending:
// one two-pointer pass per run instead of n binary searches -- on 16384 walkers every rank of an 8-GPU run spent ~1 ms
// per block in the numpy version.  picked[i] = the walker slot i of the new population copies.
int qmcb_comb_indices(int64_t n, const double* weights, double offset, int64_t* picked, double* total_out) {
  if (n <= 0) return -1;
  std::vector<double> ladder((size_t)n);
  double acc = 0.0;
  for (int64_t i = 0; i < n; ++i) {
    acc += weights[i];
    ladder[(size_t)i] = acc;
  }
  const double total = ladder[(size_t)n - 1];
  const double step = total / (double)n;
  const double shift = offset * total;
  int64_t j = 0;
  bool wrapped = false;
  for (int64_t i = 0; i < n; ++i) {
    const double x = shift + ((double)i * step + 0.0);
    double tooth = x;
    if (x >= total) {
      tooth = std::fmod(x, total);
      if (!wrapped) {  // second ascending run starts: restart the ladder pointer
        wrapped = true;
        j = 0;
      }
    }
    while (j < n && ladder[(size_t)j] < tooth) ++j;
    picked[i] = j;
  }
  if (total_out) *total_out = total;
  return 0;
}

}  // extern "C"
