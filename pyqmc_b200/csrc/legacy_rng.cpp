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
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstdlib>
#include <chrono>
#include <cstdio>
#include <thread>
#include <vector>

namespace {

// MT19937 state block + its tempered image; outputs are consumed from the tempered block.
struct MT {
  uint32_t* key;
  int pos;
  uint32_t tb[624];
};

__attribute__((target_clones("avx2", "default"))) void mt_reload(uint32_t* mt) {
  const int N = 624, M = 397;
  const uint32_t UPPER = 0x80000000u, LOWER = 0x7fffffffu, MATRIX_A = 0x9908b0dfu;
  int kk;
  uint32_t y;
  for (kk = 0; kk < N - M; kk++) {
    y = (mt[kk] & UPPER) | (mt[kk + 1] & LOWER);
    mt[kk] = mt[kk + M] ^ (y >> 1) ^ (-(int32_t)(y & 1) & MATRIX_A);
  }
  for (; kk < N - 1; kk++) {
    y = (mt[kk] & UPPER) | (mt[kk + 1] & LOWER);
    mt[kk] = mt[kk + (M - N)] ^ (y >> 1) ^ (-(int32_t)(y & 1) & MATRIX_A);
  }
  y = (mt[N - 1] & UPPER) | (mt[0] & LOWER);
  mt[N - 1] = mt[M - 1] ^ (y >> 1) ^ (-(int32_t)(y & 1) & MATRIX_A);
}

__attribute__((target_clones("avx2", "default"))) void temper_block(MT& s) {
  for (int i = 0; i < 624; ++i) {
    uint32_t y = s.key[i];
    y ^= (y >> 11);
    y ^= (y << 7) & 0x9d2c5680u;
    y ^= (y << 15) & 0xefc60000u;
    y ^= (y >> 18);
    s.tb[i] = y;
  }
}

inline uint32_t next32(MT& s) {
  if (s.pos == 624) {
    mt_reload(s.key);
    temper_block(s);
    s.pos = 0;
  }
  return s.tb[s.pos++];
}

inline double next_double(MT& s) {
  const int32_t a = next32(s) >> 5, b = next32(s) >> 6;
  return (a * 67108864.0 + b) / 9007199254740992.0;
}

inline void fill_uniform(MT& s, double* out, int64_t n) {
  int64_t i = 0;
  while (i < n) {
    if (s.pos >= 623) {  // block boundary: scalar path
      out[i++] = next_double(s);
      continue;
    }
    const int64_t m = std::min<int64_t>(n - i, (624 - s.pos) / 2);
    const uint32_t* t = s.tb + s.pos;
    for (int64_t k = 0; k < m; ++k) {
      const int32_t a = t[2 * k] >> 5, b = t[2 * k + 1] >> 6;
      out[i + k] = (a * 67108864.0 + b) / 9007199254740992.0;
    }
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

struct GaussPlan {
  std::vector<double> x1, x2, r2, f;
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
    if ((int64_t)r2.size() < need_pairs + 1) {  // +1: the branch-free loop stores before it tests
      x1.resize(need_pairs + 1);
      x2.resize(need_pairs + 1);
      r2.resize(need_pairs + 1);
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

extern "C" {

// Fills gauss [nsteps][ne][N][3], unif [nsteps][ne][N], and (if necp >= 0 and ecp_u != NULL)
// ecp_u [nsteps][ne][necp][N], ecp_rot [nsteps][ne][necp][9]; advances the MT19937 state in place.
int qmcb_rng_vmc_block(uint32_t* key, int32_t* pos, int32_t* has_gauss, double* cached_gauss, int nsteps,
                       int ne, int64_t N, int necp, double scale, double* gauss, double* unif, double* ecp_u,
                       double* ecp_rot, int nthreads) {
  if (*pos < 0 || *pos > 624) return -1;
  auto t0 = std::chrono::steady_clock::now();
  static thread_local MT s;
  s.key = key;
  s.pos = *pos;
  temper_block(s);
  GaussPlan plan;
  if (*has_gauss) {
    plan.slot_shift = 1;
    plan.carried = *cached_gauss;
  }
  const int64_t est = (int64_t)nsteps * ne * (N * 3 / 2 + 1 + 2 * (ecp_u ? necp : 0)) + 16;
  plan.x1.resize(est);
  plan.x2.resize(est);
  plan.r2.resize(est);
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
  const bool dbg = std::getenv("QMCB_RNG_DEBUG") != nullptr;
  auto tA = std::chrono::steady_clock::now();
  run_phase_b(plan, nthreads);
  if (dbg) {
    auto tB = std::chrono::steady_clock::now();
    std::fprintf(stderr, "[rng] phase A %.3f ms, phase B %.3f ms (%d threads, %lld pairs)\n",
                 std::chrono::duration<double, std::milli>(tA - t0).count(),
                 std::chrono::duration<double, std::milli>(tB - tA).count(), nthreads, (long long)plan.npairs);
  }
  // state of the legacy Gaussian cache after the last slot handed out
  const int64_t used = plan.nslots - plan.slot_shift;  // slots taken from generated pairs
  if (plan.nslots == 0) {
    // nothing consumed: cache unchanged
  } else if (used <= 0) {
    *has_gauss = 0;  // only the carried value was consumed
    *cached_gauss = 0.0;
  } else if (used & 1) {
    const int64_t p = used >> 1;
    *has_gauss = 1;
    *cached_gauss = plan.f[p] * plan.x1[p];
  } else {
    *has_gauss = 0;
    *cached_gauss = 0.0;
  }
  *pos = s.pos;
  return 0;
}

}  // extern "C"
