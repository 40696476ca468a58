// pbc.cuh -- periodic-boundary kernels of libqmcb200.
//
// Reference statements (relative to /root/reference):
//   lattice-summed GTOs            pyqmc/wf/numba/pbcgto.py:98-515 (_pbc_eval_gto{,_grad,_lap})
//   primitive-cell wrap, wrap phase, per-k MO contraction   pyqmc/wf/orbitals.py:192-239
//   minimal-image Jastrow          pyqmc/wf/jastrowspin.py:56-419 with configs.dist (distance.py:83-159)
//   VMC move with make_irreducible pyqmc/method/mc.py:115-137, configurations/coord.py:168-194
//   Ewald sums                     pyqmc/observables/ewald.py:242-354
#pragma once
#include "coop.cuh"

__device__ __forceinline__ void limdrift3(double (&g)[3]);
__device__ __forceinline__ double sgn(double x);

// =========================================================================================
// k_pbc_mo: MO rows (value [, gradient [, Laplacian]]) of lattice-summed Bloch orbitals at a list
// of points.  G lanes per point.  Per point:
//   0  wrap into the PRIMITIVE cell (orbitals.py:201), zero the AO accumulators ao[k][mu][c]
//   1  lanes over candidate (basis atom, image) pairs: r^2 test against the atom cutoff, survivors
//      are compacted into a list (ballot + prefix count)
//   A  lanes over surviving pairs: solid harmonics, radial sums with the per-shell cutoff, chi,
//      grad chi, lap chi of that atom's AOs -> staging slot in shared memory
//   B  lanes over (k, mu, c): ao[k][mu][c] += phase[L][k] * staged                 (pbcgto.py:216-227)
//   C  lanes over (c, MO j): mo = wrapphase[k(j)] * sum_mu ao[k(j)][mu][c] C[mu][j] (orbitals.py:204-229)
// =========================================================================================
struct PbcMoArgs {
  long long npoints;
  const int* count;  // optional device counter: npoints = *count * per_item
  int per_item;
  const double* pos;   // [..][3] indexed by posidx
  const double* wrap;  // [..][3] simulation-cell wrap vectors of the points (nullptr: zero)
  const int* idx;      // optional compacted walker list: posidx = idx[p / naip] * naip + p % naip
  int naip;
  int spin_mode;       // 0: `spin`; 1: electron = posidx % ne; 2: electron of ECP work item p / per_item
  int spin;
  const int* work;     // spin_mode 2: es.work
  int workN, necp, e_only;
  const uint8_t* mask; // optional per-point-walker mask (posidx / naip)
  double* out;         // out[p * stride_p + c * stride_c + j * stride_j]
  long long stride_p, stride_c, stride_j;
  double* out_val;     // optional second copy of the value row: out_val[p * stride_vp + j]
  long long stride_vp;
  double* ao_out;      // DERIV 0, CTA kernel only: AO values with the wrap phase, ao_out[(p * nk + k) * nao + mu]
                       // (Slater._aovals of the reference, slater.py:233) instead of the MO contraction
};

__host__ __device__ inline int pbc_mo_scratch_doubles(const Sys& S, int nc, int G) {
  return S.nkp * S.nao * nc + G * S.maxao_atom * nc + G + 2;  // accumulators, staging, candidate list (2G ints)
}

// Complex Bloch orbital j (component c) from the cos / sin accumulator planes of its k-point, complex MO
// coefficients [Re C | Im C] and the wrap phase exp(i k.R) (orbitals.py:38-39, 204-229).
__device__ __forceinline__ void pbc_mo_cx(const Sys& S, const double* __restrict__ sd, const double* __restrict__ ak_re,
                                          const double* __restrict__ ak_im, int nc, const double* __restrict__ C, int ldc,
                                          int j, int off, int k, const double (&wt)[3], double& re, double& im) {
  double rr = 0.0, ii = 0.0, ri = 0.0, ir = 0.0;
  for (int mu = 0; mu < S.nao; ++mu) {
    const double a = ak_re[mu * nc], b = ak_im[mu * nc];
    const double cr = C[mu * ldc + j], ci = C[mu * ldc + off + j];
    rr = fma(a, cr, rr);
    ii = fma(b, ci, ii);
    ri = fma(a, ci, ri);
    ir = fma(b, cr, ir);
  }
  re = rr - ii;
  im = ri + ir;
  if (!S.isgamma) {
    const double* __restrict__ kl = sd + S.o_kl + 3 * k;
    const double kd = kl[0] * wt[0] + kl[1] * wt[1] + kl[2] * wt[2];
    double sn, cs;
    sincos(kd, &sn, &cs);
    const double t = re * cs - im * sn;
    im = re * sn + im * cs;
    re = t;
  }
}

template <int DERIV, int G>
__global__ void __launch_bounds__(128) k_pbc_mo(const Sys S, const State st, const PbcMoArgs a) {
  constexpr int NC = NComp<DERIV>::value;
  const double* sd;
  const int* si;
  stage_tables(S, sd, si);
  extern __shared__ __align__(128) unsigned char qmcb_smem[];
  const int lane32 = threadIdx.x & 31;
  const int lane = lane32 & (G - 1);
  const unsigned gm = group_mask<G>(lane32);
  const int gshift = lane32 & ~(G - 1);
  const int gslot = threadIdx.x / G, gper = blockDim.x / G;
  const size_t tab = (16 + (size_t)S.dwords * 8 + (size_t)S.iwords * 4 + 15) & ~(size_t)15;
  const int per = pbc_mo_scratch_doubles(S, NC, G);
  double* ws = reinterpret_cast<double*>(qmcb_smem + tab) + (size_t)gslot * per;
  double* __restrict__ ao = ws;
  double* __restrict__ stg = ws + S.nkp * S.nao * NC;
  int* __restrict__ lst = reinterpret_cast<int*>(stg + G * S.maxao_atom * NC);
  const int stg_stride = S.maxao_atom * NC;
  const long long np = a.count ? (long long)(*a.count) * a.per_item : a.npoints;
  const double* __restrict__ Ls = sd + S.o_Ls;
  const double* __restrict__ prim = sd + S.o_prim;
  const double* __restrict__ phase = sd + S.o_phase;
  for (long long p = (long long)blockIdx.x * gper + gslot; p < np; p += (long long)gridDim.x * gper) {
    long long posidx = p;
    if (a.idx) posidx = (long long)a.idx[p / a.naip] * a.naip + p % a.naip;
    if (a.mask && !a.mask[posidx / a.naip]) continue;
    const double px = a.pos[3 * posidx], py = a.pos[3 * posidx + 1], pz = a.pos[3 * posidx + 2];
    if (isnan(px)) continue;
    int spin = a.spin;
    if (a.spin_mode == 1) spin = (int)(posidx % S.ne) >= S.nup ? 1 : 0;
    if (a.spin_mode == 2) {
      const int t = a.work[p / a.per_item];
      const int e = a.e_only >= 0 ? a.e_only : (t / a.workN) / a.necp;
      spin = e >= S.nup ? 1 : 0;
    }
    double q[3], pw[3];
    wrap_cell(sd + S.o_lprim, sd + S.o_lpriminv, px, py, pz, q, pw);
    for (int i = lane; i < S.nkp * S.nao * NC; i += G) ao[i] = 0.0;
    __syncwarp(gm);

    // ---- phases A and B on `cnt` compacted (atom, image) pairs lst[0..cnt)
    auto process = [&](int cnt) {
      if (lane < cnt) {
        const int c = lst[lane];
        int at = 0;
        while (c >= si[S.o_candoff + at + 1]) ++at;
        const int j = c - si[S.o_candoff + at];
        const double x = (q[0] - sd[S.o_bxyz + 3 * at]) - Ls[3 * j], y = (q[1] - sd[S.o_bxyz + 3 * at + 1]) - Ls[3 * j + 1],
                     z = (q[2] - sd[S.o_bxyz + 3 * at + 2]) - Ls[3 * j + 2];
        const double r2 = x * x + y * y + z * z;
        double* __restrict__ o = stg + lane * stg_stride;
        const int ao0 = si[S.o_shao + si[S.o_atsh + at]];
        double sph[44];
        int lastl = -1;
        for (int sh = si[S.o_atsh + at]; sh < si[S.o_atsh + at + 1]; ++sh) {
          const int l = si[S.o_shl + sh];
          const int m0 = si[S.o_shao + sh] - ao0;
          const double lcut = sd[S.o_lcut + sh];
          // value kernel keeps r2 < cutoff (pbcgto.py:215); the derivative kernels skip r2 > cutoff (348, 490)
          const bool in = DERIV == 0 ? (r2 < lcut) : !(r2 > lcut);
          if (!in) {
            for (int m = 0; m < (2 * l + 1) * NC; ++m) o[m0 * NC + m] = 0.0;
            continue;
          }
          if (l != lastl) {
            switch (l) {
              case 0: sph_store<0, (DERIV > 0)>(x, y, z, sph); break;
              case 1: sph_store<1, (DERIV > 0)>(x, y, z, sph); break;
              case 2: sph_store<2, (DERIV > 0)>(x, y, z, sph); break;
              case 3: sph_store<3, (DERIV > 0)>(x, y, z, sph); break;
              case 4: sph_store<4, (DERIV > 0)>(x, y, z, sph); break;
              default: sph_store<5, (DERIV > 0)>(x, y, z, sph); break;
            }
            lastl = l;
          }
          double R = 0.0, Rp = 0.0, Rl = 0.0;
          for (int pp = si[S.o_shprim + sh]; pp < si[S.o_shprim + sh + 1]; ++pp) {
            const double al = prim[2 * pp], cf = prim[2 * pp + 1];
            const double g = cf * exp(-al * r2);
            R += g;
            if (DERIV > 0) {
              const double t = 2.0 * al * g;
              Rp -= t;
              if (DERIV > 1) Rl = fma(t, 2.0 * al * r2 - 3.0, Rl);
            }
          }
          const double dRx = Rp * x, dRy = Rp * y, dRz = Rp * z;
          for (int m = 0; m < 2 * l + 1; ++m) {
            const double s = sph[4 * m];
            double* __restrict__ om = o + (m0 + m) * NC;
            om[0] = s * R;
            if (DERIV > 0) {
              const double gx = sph[4 * m + 1], gy = sph[4 * m + 2], gz = sph[4 * m + 3];
              om[1] = gx * R + s * dRx;
              om[2] = gy * R + s * dRy;
              om[3] = gz * R + s * dRz;
              if (DERIV > 1) om[4] = s * Rl + 2.0 * (gx * dRx + gy * dRy + gz * dRz);
            }
          }
        }
      }
      __syncwarp(gm);
      for (int s = 0; s < cnt; ++s) {
        const int c = lst[s];
        int at = 0;
        while (c >= si[S.o_candoff + at + 1]) ++at;
        const int j = c - si[S.o_candoff + at];
        const int ao0 = si[S.o_shao + si[S.o_atsh + at]];
        const int nloc = (si[S.o_shao + si[S.o_atsh + at + 1]] - ao0) * NC;
        const double* __restrict__ src = stg + s * stg_stride;
        for (int i = lane; i < S.nkp * nloc; i += G) {
          const int k = i / nloc, rem = i - k * nloc;
          ao[(k * S.nao + ao0) * NC + rem] = fma(phase[j * S.nkp + k], src[rem], ao[(k * S.nao + ao0) * NC + rem]);
        }
      }
      __syncwarp(gm);
    };

    int pending = 0;
    for (int base = 0; base < S.ncand; base += G) {
      const int c = base + lane;
      bool ok = false;
      if (c < S.ncand) {
        int at = 0;
        while (c >= si[S.o_candoff + at + 1]) ++at;
        const int j = c - si[S.o_candoff + at];
        const double x = (q[0] - sd[S.o_bxyz + 3 * at]) - Ls[3 * j], y = (q[1] - sd[S.o_bxyz + 3 * at + 1]) - Ls[3 * j + 1],
                     z = (q[2] - sd[S.o_bxyz + 3 * at + 2]) - Ls[3 * j + 2];
        ok = !(x * x + y * y + z * z > sd[S.o_atomcut + at]);  // pbcgto.py:207
      }
      unsigned b = __ballot_sync(gm, ok);
      b = G == 32 ? b : ((b >> gshift) & ((1u << G) - 1u));
      if (ok) lst[pending + __popc(b & ((1u << lane) - 1u))] = c;
      pending += __popc(b);
      __syncwarp(gm);
      if (pending >= G) {
        process(G);
        const int rest = pending - G;
        int tmp = 0;
        if (lane < rest) tmp = lst[G + lane];
        __syncwarp(gm);
        if (lane < rest) lst[lane] = tmp;
        __syncwarp(gm);
        pending = rest;
      }
    }
    if (pending > 0) process(pending);

    // ---- phase C: wrap phase and per-k MO contraction
    const int ldc = S.ldc[spin], nmo = S.nmo[spin];
    const double* __restrict__ C = sd + S.o_mo[spin];
    const int* __restrict__ mok = si + S.o_mok[spin];
    double wt[3] = {pw[0], pw[1], pw[2]};
    if (!S.isgamma && a.wrap) {
      const double* __restrict__ Sm = sd + S.o_smat;
      const double w0 = a.wrap[3 * posidx], w1 = a.wrap[3 * posidx + 1], w2 = a.wrap[3 * posidx + 2];
#pragma unroll
      for (int k = 0; k < 3; ++k) wt[k] = (w0 * Sm[k] + w1 * Sm[3 + k] + w2 * Sm[6 + k]) + pw[k];
    }
    if (S.cplx) {
      const int nmt = S.nmo_t[spin], off = S.cxoff[spin];
      for (int t = lane; t < NC * ldc; t += G) {
        const int c = t / ldc, j = t - c * ldc;
        if (j >= nmt && j < off + nmt) continue;  // imaginary columns: written with their real partner
        double re = 0.0, im = 0.0;
        if (j < nmt) {
          const int k = mok[j];
          pbc_mo_cx(S, sd, ao + (size_t)k * S.nao * NC + c, ao + (size_t)(S.nk + k) * S.nao * NC + c, NC, C, ldc, j, off, k,
                    wt, re, im);
          a.out[p * a.stride_p + c * a.stride_c + (off + j) * a.stride_j] = im;
          if (c == 0 && a.out_val) a.out_val[p * a.stride_vp + off + j] = im;
        }
        a.out[p * a.stride_p + c * a.stride_c + j * a.stride_j] = re;
        if (c == 0 && a.out_val) a.out_val[p * a.stride_vp + j] = re;
      }
      __syncwarp(gm);
      continue;
    }
    for (int t = lane; t < NC * ldc; t += G) {
      const int c = t / ldc, j = t - c * ldc;
      double acc = 0.0;
      if (j < nmo) {
        const int k = mok[j];
        const double* __restrict__ ak = ao + (size_t)k * S.nao * NC + c;
        for (int mu = 0; mu < S.nao; ++mu) acc = fma(ak[mu * NC], C[mu * ldc + j], acc);
        if (!S.isgamma) {  // (-1) ** round(k.R / pi)  (orbitals.py:34-35, 204-213)
          const double* __restrict__ kl = sd + S.o_kl + 3 * k;
          const double kd = kl[0] * wt[0] + kl[1] * wt[1] + kl[2] * wt[2];
          const double n = rint(kd / 3.141592653589793);
          if (fmod(fabs(n), 2.0) == 1.0) acc = -acc;
        }
      }
      a.out[p * a.stride_p + c * a.stride_c + j * a.stride_j] = acc;
      if (c == 0 && a.out_val) a.out_val[p * a.stride_vp + j] = acc;
    }
    __syncwarp(gm);
  }
}

// chi, grad chi, lap chi of one shell at displacement (x,y,z) -> om[m * NC + c]
template <int L, int DERIV>
__device__ __forceinline__ void pbc_stage_shell(double x, double y, double z, double R, double Rp, double Rl,
                                                double* __restrict__ om) {
  constexpr int NC = NComp<DERIV>::value;
  constexpr int NF = 2 * L + 1;
  double s[NF], gx[NF], gy[NF], gz[NF];
  if constexpr (L == 0) sph_l0<(DERIV > 0)>(x, y, z, s, gx, gy, gz);
  if constexpr (L == 1) sph_l1<(DERIV > 0)>(x, y, z, s, gx, gy, gz);
  if constexpr (L == 2) sph_l2<(DERIV > 0)>(x, y, z, s, gx, gy, gz);
  if constexpr (L == 3) sph_l3<(DERIV > 0)>(x, y, z, s, gx, gy, gz);
  if constexpr (L == 4) sph_l4<(DERIV > 0)>(x, y, z, s, gx, gy, gz);
  if constexpr (L == 5) sph_l5<(DERIV > 0)>(x, y, z, s, gx, gy, gz);
  const double dRx = Rp * x, dRy = Rp * y, dRz = Rp * z;
#pragma unroll
  for (int m = 0; m < NF; ++m) {
    om[m * NC] = s[m] * R;
    if (DERIV > 0) {
      om[m * NC + 1] = gx[m] * R + s[m] * dRx;
      om[m * NC + 2] = gy[m] * R + s[m] * dRy;
      om[m * NC + 3] = gz[m] * R + s[m] * dRz;
      if (DERIV > 1) om[m * NC + 4] = s[m] * Rl + 2.0 * (gx[m] * dRx + gy[m] * dRy + gz[m] * dRz);
    }
  }
}

// -----------------------------------------------------------------------------------------
// k_pbc_mo_cta: the same evaluation with ONE CTA PER POINT (T = blockDim.x threads), for launches
// whose point count alone cannot fill the machine (the per-electron proposals of the periodic VMC
// block: N points) -- and measured faster per point than the lane-group form in general.
//   A  threads over (candidate pair, shell of its atom) tasks: r^2 cutoffs, radial sums, solid
//      harmonics of that shell, chi / grad chi / lap chi -> staging slot of the pair (no compaction)
//   B  thread r owns the (mu, c) column r of the accumulators for ALL k in registers:
//      acc[k] += phase[L][k] * staged[r]                                    (pbcgto.py:216-227)
//   C  threads over (c, MO)                                                 (orbitals.py:204-229)
// Limits (else the lane-group kernel runs): nk <= 8, nao * NC <= 2 T, ncand * stride fits smem.
// -----------------------------------------------------------------------------------------
#define QMCB_PBC_NKMAX 8
#define QMCB_PBC_RU 2

__host__ __device__ inline int pbc_mo_cta_chunk(const Sys& S) { return S.ncand < 64 ? S.ncand : 64; }
__host__ __device__ inline size_t pbc_mo_cta_scratch_bytes(const Sys& S, int nc) {
  const int chunk = pbc_mo_cta_chunk(S);
  return ((size_t)S.nkp * S.nao * nc + (size_t)chunk * S.maxao_atom * nc) * 8 + (size_t)chunk * 2 * 4 + 16;
}

// LMAX: highest angular momentum the instantiation dispatches.  The l <= 4 form is the tuned one (register budget
// of the 4-CTAs-per-SM variant); bases that reach l = 5 use the LMAX = 5 instantiations.
template <int DERIV, int MAXT = 256, int MINB = 2, int LMAX = 4>
__global__ void __launch_bounds__(MAXT, MINB) k_pbc_mo_cta(const Sys S, const State st, const PbcMoArgs a) {
  constexpr int NC = NComp<DERIV>::value;
  const int T = blockDim.x;
  const double* sd;
  const int* si;
  stage_tables(S, sd, si);
  extern __shared__ __align__(128) unsigned char qmcb_smem[];
  const size_t tab = (16 + (size_t)S.dwords * 8 + (size_t)S.iwords * 4 + 15) & ~(size_t)15;
  const int chunk = pbc_mo_cta_chunk(S);
  const int ncol = S.nao * NC;
  const int stg_stride = S.maxao_atom * NC;
  double* __restrict__ ao = reinterpret_cast<double*>(qmcb_smem + tab);  // [nk][nao][NC]
  double* __restrict__ stg = ao + (size_t)S.nkp * ncol;
  int* __restrict__ meta = reinterpret_cast<int*>(stg + (size_t)chunk * stg_stride);  // [chunk][2]: atom (-1 invalid), image
  const int tid = threadIdx.x;
  const double* __restrict__ Ls = sd + S.o_Ls;
  const double* __restrict__ prim = sd + S.o_prim;
  const double* __restrict__ phase = sd + S.o_phase;
  // this thread's accumulator columns r = tid + T u: atom of mu and offset inside the atom's block
  int cat[QMCB_PBC_RU], coff[QMCB_PBC_RU];
#pragma unroll
  for (int u = 0; u < QMCB_PBC_RU; ++u) {
    const int r = tid + T * u;
    cat[u] = -2;
    coff[u] = 0;
    if (r < ncol) {
      const int sh = si[S.o_aoshell + r / NC];
      const int at = si[S.o_shatom + sh];
      cat[u] = at;
      coff[u] = r - si[S.o_shao + si[S.o_atsh + at]] * NC;
    }
  }
  int maxsh = 1;
  for (int at = 0; at < S.nbatom; ++at) maxsh = max(maxsh, si[S.o_atsh + at + 1] - si[S.o_atsh + at]);
  const long long np = a.count ? (long long)(*a.count) * a.per_item : a.npoints;
  for (long long p = blockIdx.x; p < np; p += gridDim.x) {
    long long posidx = p;
    if (a.idx) posidx = (long long)a.idx[p / a.naip] * a.naip + p % a.naip;
    if (a.mask && !a.mask[posidx / a.naip]) continue;
    const double px = a.pos[3 * posidx], py = a.pos[3 * posidx + 1], pz = a.pos[3 * posidx + 2];
    if (isnan(px)) continue;
    int spin = a.spin;
    if (a.spin_mode == 1) spin = (int)(posidx % S.ne) >= S.nup ? 1 : 0;
    if (a.spin_mode == 2) {
      const int tt = a.work[p / a.per_item];
      const int e = a.e_only >= 0 ? a.e_only : (tt / a.workN) / a.necp;
      spin = e >= S.nup ? 1 : 0;
    }
    double q[3], pw[3];
    wrap_cell(sd + S.o_lprim, sd + S.o_lpriminv, px, py, pz, q, pw);
    double acc[QMCB_PBC_RU][QMCB_PBC_NKMAX];
#pragma unroll
    for (int u = 0; u < QMCB_PBC_RU; ++u)
#pragma unroll
      for (int k = 0; k < QMCB_PBC_NKMAX; ++k) acc[u][k] = 0.0;
    for (int base = 0; base < S.ncand; base += chunk) {
      const int cnt = (S.ncand - base) < chunk ? (S.ncand - base) : chunk;
      // ---- phase A: (pair, shell) tasks
      for (int task = tid; task < cnt * maxsh; task += T) {
        const int slot = task / maxsh, ish = task - slot * maxsh;
        const int c = base + slot;
        int at = 0;
        while (c >= si[S.o_candoff + at + 1]) ++at;
        const int j = c - si[S.o_candoff + at];
        const double x = (q[0] - sd[S.o_bxyz + 3 * at]) - Ls[3 * j], y = (q[1] - sd[S.o_bxyz + 3 * at + 1]) - Ls[3 * j + 1],
                     z = (q[2] - sd[S.o_bxyz + 3 * at + 2]) - Ls[3 * j + 2];
        const double r2 = x * x + y * y + z * z;
        const bool near = !(r2 > sd[S.o_atomcut + at]);  // pbcgto.py:207
        if (ish == 0) {
          meta[2 * slot] = near ? at : -1;
          meta[2 * slot + 1] = j;
        }
        const int sh = si[S.o_atsh + at] + ish;
        if (!near || sh >= si[S.o_atsh + at + 1]) continue;
        const int l = si[S.o_shl + sh];
        double* __restrict__ om = stg + (size_t)slot * stg_stride + (si[S.o_shao + sh] - si[S.o_shao + si[S.o_atsh + at]]) * NC;
        const double lcut = sd[S.o_lcut + sh];
        // value kernel keeps r2 < cutoff (pbcgto.py:215); the derivative kernels skip r2 > cutoff (348, 490)
        const bool in = DERIV == 0 ? (r2 < lcut) : !(r2 > lcut);
        if (!in) {
          for (int m = 0; m < (2 * l + 1) * NC; ++m) om[m] = 0.0;
          continue;
        }
        double R = 0.0, Rp = 0.0, Rl = 0.0;
        for (int pp = si[S.o_shprim + sh]; pp < si[S.o_shprim + sh + 1]; ++pp) {
          const double al = prim[2 * pp], cf = prim[2 * pp + 1];
          const double g = cf * exp(-al * r2);
          R += g;
          if (DERIV > 0) {
            const double t2 = 2.0 * al * g;
            Rp -= t2;
            if (DERIV > 1) Rl = fma(t2, 2.0 * al * r2 - 3.0, Rl);
          }
        }
        switch (l) {
          case 0: pbc_stage_shell<0, DERIV>(x, y, z, R, Rp, Rl, om); break;
          case 1: pbc_stage_shell<1, DERIV>(x, y, z, R, Rp, Rl, om); break;
          case 2: pbc_stage_shell<2, DERIV>(x, y, z, R, Rp, Rl, om); break;
          case 3: pbc_stage_shell<3, DERIV>(x, y, z, R, Rp, Rl, om); break;
          default:
            if (LMAX >= 5 && l == 5)
              pbc_stage_shell<LMAX >= 5 ? 5 : 4, DERIV>(x, y, z, R, Rp, Rl, om);
            else
              pbc_stage_shell<4, DERIV>(x, y, z, R, Rp, Rl, om);
            break;
        }
      }
      __syncthreads();
      // ---- phase B
      for (int s2 = 0; s2 < cnt; ++s2) {
        const int at = meta[2 * s2];
        if (at < 0) continue;
        const double* __restrict__ ph = phase + meta[2 * s2 + 1] * S.nkp;
        const double* __restrict__ src = stg + (size_t)s2 * stg_stride;
#pragma unroll
        for (int u = 0; u < QMCB_PBC_RU; ++u) {
          if (cat[u] == at) {
            const double f = src[coff[u]];
#pragma unroll
            for (int k = 0; k < QMCB_PBC_NKMAX; ++k)
              if (k < S.nkp) acc[u][k] = fma(ph[k], f, acc[u][k]);
          }
        }
      }
      __syncthreads();
    }
#pragma unroll
    for (int u = 0; u < QMCB_PBC_RU; ++u) {
      const int r = tid + T * u;
      if (r < ncol) {
#pragma unroll
        for (int k = 0; k < QMCB_PBC_NKMAX; ++k)
          if (k < S.nkp) ao[k * ncol + r] = acc[u][k];
      }
    }
    __syncthreads();
    // ---- phase C: wrap phase and per-k MO contraction
    const int ldc = S.ldc[spin], nmo = S.nmo[spin];
    const double* __restrict__ C = sd + S.o_mo[spin];
    const int* __restrict__ mok = si + S.o_mok[spin];
    double wt[3] = {pw[0], pw[1], pw[2]};
    if (!S.isgamma && a.wrap) {
      const double* __restrict__ Sm = sd + S.o_smat;
      const double w0 = a.wrap[3 * posidx], w1 = a.wrap[3 * posidx + 1], w2 = a.wrap[3 * posidx + 2];
#pragma unroll
      for (int k = 0; k < 3; ++k) wt[k] = (w0 * Sm[k] + w1 * Sm[3 + k] + w2 * Sm[6 + k]) + pw[k];
    }
    if (DERIV == 0 && a.ao_out && S.cplx) {
      // complex AO values with the wrap phase: ao_out[((p * 2 + re/im) * nk + k) * nao + mu]
      for (int t = tid; t < S.nk * S.nao; t += T) {
        const int k = t / S.nao;
        double re = ao[t], im = ao[(size_t)S.nk * S.nao + t];
        if (!S.isgamma) {
          const double* __restrict__ kl = sd + S.o_kl + 3 * k;
          const double kd = kl[0] * wt[0] + kl[1] * wt[1] + kl[2] * wt[2];
          double sn, cs;
          sincos(kd, &sn, &cs);
          const double tr = re * cs - im * sn;
          im = re * sn + im * cs;
          re = tr;
        }
        a.ao_out[(size_t)p * 2 * S.nk * S.nao + t] = re;
        a.ao_out[((size_t)p * 2 + 1) * S.nk * S.nao + t] = im;
      }
      __syncthreads();
      continue;
    }
    if (S.cplx) {
      const int nmt = S.nmo_t[spin], off = S.cxoff[spin];
      for (int t = tid; t < NC * ldc; t += T) {
        const int c = t / ldc, j = t - c * ldc;
        if (j >= nmt && j < off + nmt) continue;  // imaginary columns: written with their real partner
        double re = 0.0, im = 0.0;
        if (j < nmt) {
          const int k = mok[j];
          pbc_mo_cx(S, sd, ao + (size_t)k * ncol + c, ao + (size_t)(S.nk + k) * ncol + c, NC, C, ldc, j, off, k, wt, re, im);
          a.out[p * a.stride_p + c * a.stride_c + (off + j) * a.stride_j] = im;
          if (c == 0 && a.out_val) a.out_val[p * a.stride_vp + off + j] = im;
        }
        a.out[p * a.stride_p + c * a.stride_c + j * a.stride_j] = re;
        if (c == 0 && a.out_val) a.out_val[p * a.stride_vp + j] = re;
      }
      __syncthreads();
      continue;
    }
    if (DERIV == 0 && a.ao_out) {
      for (int t = tid; t < S.nk * S.nao; t += T) {
        const int k = t / S.nao;
        double v = ao[t];
        if (!S.isgamma) {
          const double* __restrict__ kl = sd + S.o_kl + 3 * k;
          const double kd = kl[0] * wt[0] + kl[1] * wt[1] + kl[2] * wt[2];
          const double n = rint(kd / 3.141592653589793);
          if (fmod(fabs(n), 2.0) == 1.0) v = -v;
        }
        a.ao_out[(size_t)p * S.nk * S.nao + t] = v;
      }
      __syncthreads();
      continue;
    }
    for (int t = tid; t < NC * ldc; t += T) {
      const int c = t / ldc, j = t - c * ldc;
      double v = 0.0;
      if (j < nmo) {
        const int k = mok[j];
        const double* __restrict__ ak = ao + (size_t)k * ncol + c;
        for (int mu = 0; mu < S.nao; ++mu) v = fma(ak[mu * NC], C[mu * ldc + j], v);
        if (!S.isgamma) {
          const double* __restrict__ kl = sd + S.o_kl + 3 * k;
          const double kd = kl[0] * wt[0] + kl[1] * wt[1] + kl[2] * wt[2];
          const double n = rint(kd / 3.141592653589793);
          if (fmod(fabs(n), 2.0) == 1.0) v = -v;
        }
      }
      a.out[p * a.stride_p + c * a.stride_c + j * a.stride_j] = v;
      if (c == 0 && a.out_val) a.out_val[p * a.stride_vp + j] = v;
    }
    __syncthreads();
  }
}

// =========================================================================================
// Minimal-image Jastrow with lanes over PARTNERS (other electrons, then atoms): one 27-shift search
// per partner, all radial functions of that partner on the same lane.
// WANT 1: du (log ratio vs cached partial sums) and grad U;  WANT 2: grad U and laplacian U.
// =========================================================================================
template <int WANT, int G>
__device__ __forceinline__ void coop_jastrow_pbc(const Sys& S, const double* __restrict__ sd,
                                                 const int* __restrict__ si, const State& st, int w, int e, double px,
                                                 double py, double pz, int lane, unsigned gm, double& du,
                                                 double (&g)[3], double& lap, double* vb_store = nullptr,
                                                 double* va_store = nullptr, double* gp_store = nullptr,
                                                 double* ga_out = nullptr) {
  // vb_store [(ne-1)][nb] / va_store [natom][na]: the radial values at this point (0 beyond the cutoff), kept
  // by the fused periodic move kernel for the cache update of an accepted move; gp_store [(ne-1)][3]: each partner's
  // term of grad U (the pair-gradient cache entry, coop.cuh GPAIR), ga_out [3]: the electron-ion part of grad U (AGRAD)
  const int s = e >= S.nup ? 1 : 0;
  double unew = 0.0, uold = 0.0, g0 = 0.0, g1 = 0.0, g2 = 0.0, lp = 0.0;
  double ga0 = 0.0, ga1 = 0.0, ga2 = 0.0;
  const int npart = (S.nb > 0 ? S.ne - 1 : 0), nat = (S.na > 0 ? S.natom : 0);
#pragma unroll 1
  for (int t = lane; t < npart + nat; t += G) {
    double dx, dy, dz, rcut;
    int nfun;
    const bool isb = t < npart;
    int j = 0, I = 0;
    if (isb) {
      j = t < e ? t : t + 1;
      dx = px - CONF(st, S, w, j, 0);
      dy = py - CONF(st, S, w, j, 1);
      dz = pz - CONF(st, S, w, j, 2);
      rcut = S.rcut_b;
      nfun = S.nb;
    } else {
      I = t - npart;
      dx = px - sd[S.o_xyz + 3 * I];
      dy = py - sd[S.o_xyz + 3 * I + 1];
      dz = pz - sd[S.o_xyz + 3 * I + 2];
      rcut = S.rcut_a;
      nfun = S.na;
    }
    if (S.pbc) min_image(S, sd, dx, dy, dz);
    const double r = sqrt(dx * dx + dy * dy + dz * dz);
    double* vst = isb ? (vb_store ? vb_store + t * S.nb : nullptr) : (va_store ? va_store + I * S.na : nullptr);
    if (vst != nullptr && !(r < rcut))
      for (int l = 0; l < nfun; ++l) vst[l] = 0.0;
    double gsp = 0.0;  // sum_l c_l b_l'(r) / r of this partner
    if (r < rcut) {
      const int sj = j >= S.nup ? 1 : 0;
      for (int l = 0; l < nfun; ++l) {
        double v, gg, ll;
        double c;
        if (isb) {
          c = sd[S.o_bcoef + l * 3 + s + sj];
          radial_ool<WANT>(si[S.o_bkind + l], sd[S.o_bpar + l], rcut, r, v, gg, ll);
        } else {
          c = sd[S.o_acoef + (I * S.na + l) * 2 + s];
          radial_ool<WANT>(si[S.o_akind + l], sd[S.o_apar + l], rcut, r, v, gg, ll);
        }
        if (vst != nullptr) vst[l] = v;
        unew = fma(c, v, unew);
        const double cg = c * gg;
        g0 = fma(cg, dx, g0);
        g1 = fma(cg, dy, g1);
        g2 = fma(cg, dz, g2);
        if (WANT == 2) lp = fma(c, ll, lp);
        if (isb) {
          gsp += cg;
        } else if (ga_out != nullptr) {
          ga0 = fma(cg, dx, ga0);
          ga1 = fma(cg, dy, ga1);
          ga2 = fma(cg, dz, ga2);
        }
      }
    }
    if (isb && gp_store != nullptr) {
      gp_store[t * 3] = gsp * dx;
      gp_store[t * 3 + 1] = gsp * dy;
      gp_store[t * 3 + 2] = gsp * dz;
    }
  }
  if (ga_out != nullptr) {
    ga_out[0] = group_sum<G>(ga0, gm);
    ga_out[1] = group_sum<G>(ga1, gm);
    ga_out[2] = group_sum<G>(ga2, gm);
  }
  if (WANT != 2) {
    const int na_items = S.natom * S.na, nb_items = S.nb * 2;
    for (int t = lane; t < na_items + nb_items; t += G) {
      if (t < na_items) {
        const int I = t / S.na, k = t - I * S.na;
        uold = fma(sd[S.o_acoef + (I * S.na + k) * 2 + s], APART(st, S, w, e, I, k), uold);
      } else {
        const int u = t - na_items;
        const int l = u >> 1, tt = u & 1;
        uold = fma(sd[S.o_bcoef + l * 3 + s + tt], BPART(st, S, w, e, l, tt), uold);
      }
    }
  }
  du = group_sum<G>(unew, gm) - group_sum<G>(uold, gm);
  g[0] = group_sum<G>(g0, gm);
  g[1] = group_sum<G>(g1, gm);
  g[2] = group_sum<G>(g2, gm);
  lap = WANT == 2 ? group_sum<G>(lp, gm) : 0.0;
}

// Jastrow caches of walker w after electron e moved from its current position to (nx,ny,nz)
// (jastrowspin.py:111-137, 221-249), lanes over partners; jtmp: (ne - 1) * nb doubles per walker.
template <int G>
__device__ __forceinline__ void coop_jastrow_update_pbc(const Sys& S, const double* __restrict__ sd,
                                                        const int* __restrict__ si, const State& st, int w, int e,
                                                        double nx, double ny, double nz, int lane, unsigned gm,
                                                        double* __restrict__ jtmp) {
  const int s = e >= S.nup ? 1 : 0;
  if (S.na > 0) {
    for (int I = lane; I < S.natom; I += G) {
      double dx = nx - sd[S.o_xyz + 3 * I], dy = ny - sd[S.o_xyz + 3 * I + 1], dz = nz - sd[S.o_xyz + 3 * I + 2];
      if (S.pbc) min_image(S, sd, dx, dy, dz);
      const double r = sqrt(dx * dx + dy * dy + dz * dz);
      for (int k = 0; k < S.na; ++k) {
        double v = 0.0, gg, ll;
        if (r < S.rcut_a) radial_ool<0>(si[S.o_akind + k], sd[S.o_apar + k], S.rcut_a, r, v, gg, ll);
        AVAL(st, S, w, I, k, s) += v - APART(st, S, w, e, I, k);
        APART(st, S, w, e, I, k) = v;
      }
    }
  }
  if (S.nb > 0) {
    const double ox = CONF(st, S, w, e, 0), oy = CONF(st, S, w, e, 1), oz = CONF(st, S, w, e, 2);
#pragma unroll 1
    for (int jj = lane; jj < S.ne - 1; jj += G) {
      const int j = jj < e ? jj : jj + 1;
      const double jx = CONF(st, S, w, j, 0), jy = CONF(st, S, w, j, 1), jz = CONF(st, S, w, j, 2);
      double dx = nx - jx, dy = ny - jy, dz = nz - jz;
      if (S.pbc) min_image(S, sd, dx, dy, dz);
      const double rn = sqrt(dx * dx + dy * dy + dz * dz);
      dx = ox - jx;
      dy = oy - jy;
      dz = oz - jz;
      if (S.pbc) min_image(S, sd, dx, dy, dz);
      const double ro = sqrt(dx * dx + dy * dy + dz * dz);
      for (int l = 0; l < S.nb; ++l) {
        double vn = 0.0, vo = 0.0, gg, ll;
        if (rn < S.rcut_b) radial_ool<0>(si[S.o_bkind + l], sd[S.o_bpar + l], S.rcut_b, rn, vn, gg, ll);
        if (ro < S.rcut_b) radial_ool<0>(si[S.o_bkind + l], sd[S.o_bpar + l], S.rcut_b, ro, vo, gg, ll);
        BPART(st, S, w, j, l, s) += vn - vo;
        jtmp[jj * S.nb + l] = vn;
      }
    }
    __syncwarp(gm);
#pragma unroll 1
    for (int t = lane; t < S.nb * 2; t += G) {
      const int l = t >> 1, tt = t & 1;
      double bn = 0.0;
      for (int jj = 0; jj < S.ne - 1; ++jj) {
        const int j = jj < e ? jj : jj + 1;
        if ((j >= S.nup ? 1 : 0) == tt) bn += jtmp[jj * S.nb + l];
      }
      BVAL(st, S, w, l, s + tt) += bn - BPART(st, S, w, e, l, tt);
      BPART(st, S, w, e, l, tt) = bn;
    }
  }
  __syncwarp(gm);
}

// updateinternals of the Jastrow caches with G lanes per walker (lanes over partners), open or periodic;
// same contract as the thread-per-walker k_jastrow_update: new position = st.saved_pos[w]
template <int G>
__global__ void __launch_bounds__(128) k_jastrow_update_coop(const Sys S, const State st, int e, int do_jastrow,
                                                             int move_conf, const uint8_t* mask) {
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
  const int jper = ((S.ne > 1 ? S.ne - 1 : 0) * S.nb + 1) & ~1;
  double* jtmp = reinterpret_cast<double*>(qmcb_smem + tab) + (size_t)slot * jper;
  const double nx = st.saved_pos[(size_t)w * 3], ny = st.saved_pos[(size_t)w * 3 + 1], nz = st.saved_pos[(size_t)w * 3 + 2];
  if (do_jastrow) coop_jastrow_update_pbc<G>(S, sd, si, st, w, e, nx, ny, nz, lane, gm, jtmp);
  __syncwarp(gm);
  if (move_conf && lane < 3) {
    CONF(st, S, w, e, lane) = st.saved_pos[(size_t)w * 3 + lane];
    if (S.pbc) st.wrap[((size_t)w * S.ne + e) * 3 + lane] = st.saved_wrap[(size_t)w * 3 + lane];
  }
}

// =========================================================================================
// Device-resident VMC move for periodic single-determinant wave functions, one warp per walker,
// around k_pbc_mo and the Sherman-Morrison kernel (mc.py:115-137):
//   k_pbc_propose : drift at the current position (cached MO rows . inverse column + Jastrow),
//                   proposal, wrap into the simulation cell (make_irreducible)
//   k_pbc_mo<2>   : MO rows at the proposed positions -> st.monew
//   k_pbc_accept  : ratio + drift at the proposed position, Metropolis test; accepted walkers update
//                   the Jastrow caches, the cached MO rows, the coordinates and wrap vectors
//   k_sm_warp     : masked Sherman-Morrison update of the inverse from the value row in st.monew
// =========================================================================================
// Cache update of an accepted move from stored radial values: jnew_b / jnew_a hold b_l / a_k at the new
// position (coop_jastrow_pbc's vb_store / va_store), jold_b the b_l at the old one (stored by the proposal).
// Same bookkeeping as coop_jastrow_update_pbc (jastrowspin.py:221-249) without any distance or radial work.
// gp_new [(ne-1)][3] / ga_new [3]: pair-gradient and electron-ion gradient terms at the new position -> GPAIR / AGRAD;
// the b_l at the old position come from BPAIR, which takes the new ones (the pair caches of coop.cuh, kept for the
// proposal of the following electrons: their drift is a sum of cached pair terms, no minimal-image or radial work)
template <int G>
__device__ __forceinline__ void coop_jastrow_commit_pbc(const Sys& S, const State& st, int w, int e, int lane,
                                                        unsigned gm, const double* jnew_b, const double* jnew_a,
                                                        const double* gp_new, const double* ga_new) {
  const int s = e >= S.nup ? 1 : 0;
  if (S.na > 0) {
    for (int t = lane; t < S.natom * S.na; t += G) {
      const int I = t / S.na, k = t - I * S.na;
      const double v = jnew_a[t];
      AVAL(st, S, w, I, k, s) += v - APART(st, S, w, e, I, k);
      APART(st, S, w, e, I, k) = v;
    }
  }
  if (S.nb > 0) {
    for (int t = lane; t < (S.ne - 1) * S.nb; t += G) {
      const int jj = t / S.nb, l = t - jj * S.nb;
      const int j = jj < e ? jj : jj + 1;
      const int p = e < j ? pair_index(S.ne, e, j) : pair_index(S.ne, j, e);
      BPART(st, S, w, j, l, s) += jnew_b[t] - BPAIR(st, S, w, p, l);
      BPAIR(st, S, w, p, l) = jnew_b[t];
    }
#pragma unroll 1
    for (int t = lane; t < S.nb * 2; t += G) {
      const int l = t >> 1, tt = t & 1;
      double bn = 0.0;
      for (int jj = 0; jj < S.ne - 1; ++jj) {
        const int j = jj < e ? jj : jj + 1;
        if ((j >= S.nup ? 1 : 0) == tt) bn += jnew_b[jj * S.nb + l];
      }
      BVAL(st, S, w, l, s + tt) += bn - BPART(st, S, w, e, l, tt);
      BPART(st, S, w, e, l, tt) = bn;
    }
    for (int t = lane; t < (S.ne - 1) * 3; t += G) {
      const int jj = t / 3, x = t - jj * 3;
      const int j = jj < e ? jj : jj + 1;
      if (e < j)
        GPAIR(st, S, w, pair_index(S.ne, e, j), x) = gp_new[t];
      else
        GPAIR(st, S, w, pair_index(S.ne, j, e), x) = -gp_new[t];
    }
  }
  if (lane < 3) AGRAD(st, S, w, e, lane) = ga_new[lane];
  __syncwarp(gm);
}

// doubles of shared-memory scratch per warp of k_pbc_accept: b_l / a_k at the proposed point, pair-gradient terms
__host__ __device__ inline int pbc_accept_jper(const Sys& S) {
  return ((S.ne > 1 ? S.ne - 1 : 0) * (S.nb + 3) + S.natom * S.na + 3 + 1) & ~1;
}

// Sherman-Morrison row replacement by one warp, lane j owns column j (n <= NPAD <= 32): the arithmetic of
// k_sm_thread / k_sm_warp (kernels.cuh), shared by the stand-alone kernel and the fused periodic move kernel.
// v = this lane's entry of the new row (0 for lanes >= n).  Returns the determinant ratio on every lane.
template <int NPAD>
__device__ __forceinline__ double sm_warp_apply(double* inv, int n, int e, int lane, double vk) {
  const bool act = lane < n;
  double A[NPAD];
#pragma unroll
  for (int k = 0; k < NPAD; ++k) A[k] = (k < n && act) ? inv[k * n + lane] : 0.0;
  double t = 0.0;
#pragma unroll
  for (int k = 0; k < NPAD; ++k) {
    const double v = __shfl_sync(0xffffffffu, vk, k);
    t = fma(v, A[k], t);
  }
  const double ratio = __shfl_sync(0xffffffffu, t, e);
  // lane k fetches inv[k][e] (just read by lane e -> L1/L2 hit) and does one division
  double col = 0.0;
  if (act) col = inv[lane * n + e] / ratio;
#pragma unroll
  for (int k = 0; k < NPAD; ++k) {
    const double ck = __shfl_sync(0xffffffffu, col, k);
    if (k < n && act) inv[k * n + lane] = (lane == e) ? ck : fma(-ck, t, A[k]);
  }
  return ratio;
}

struct PbcMoveArgs {
  int e;
  double tstep;
  const double* gauss;  // [N][3]
  const double* unif;   // [N]
  uint8_t* accept;      // [N]
  unsigned long long* nacc;
  const double* gauss_next;  // fused kernel: variates of electron e + 1 (nullptr after the last electron)
  int w0, wn;                // walker range [w0, w0 + wn) of this launch (wn = 0: all walkers)
};

template <int G>
__device__ __forceinline__ void slater_row_ratio4(const Sys& S, const int* __restrict__ si, const State& st, int w,
                                                  int s, int eeff, const double* __restrict__ rows, int ldmax, int lane,
                                                  unsigned gm, double (&r)[4]) {
  const int n = s ? S.ndn : S.nup;
  const int* __restrict__ occ = si + S.o_occ[s];
  const double* __restrict__ inv = st.inv[s] + (size_t)w * n * n + eeff;
  double a0 = 0.0, a1 = 0.0, a2 = 0.0, a3 = 0.0;
  for (int k = lane; k < n; k += G) {
    const double c = inv[k * n];
    const int o = occ[k];
    a0 = fma(rows[o], c, a0);
    a1 = fma(rows[ldmax + o], c, a1);
    a2 = fma(rows[2 * ldmax + o], c, a2);
    a3 = fma(rows[3 * ldmax + o], c, a3);
  }
  r[0] = group_sum<G>(a0, gm);
  r[1] = group_sum<G>(a1, gm);
  r[2] = group_sum<G>(a2, gm);
  r[3] = group_sum<G>(a3, gm);
}

// drift at the current position of electron e, proposal, wrap: one warp per walker
__device__ __forceinline__ void pbc_propose_warp(const Sys& S, const double* sd, const int* si, const State& st, int w,
                                                 int e, double tstep, const double* __restrict__ gauss_e, int lane) {
  constexpr int G = 32;
  const unsigned gm = 0xffffffffu;
  const int s = e >= S.nup ? 1 : 0;
  const int eeff = e - s * S.nup;
  const int ldmax = S.ldc[0] > S.ldc[1] ? S.ldc[0] : S.ldc[1];
  const double ox = CONF(st, S, w, e, 0), oy = CONF(st, S, w, e, 1), oz = CONF(st, S, w, e, 2);
  double grad[3] = {0.0, 0.0, 0.0};
  if (S.nmo[0] + S.nmo[1] > 0) {
    double r[4];
    slater_row_ratio4<G>(S, si, st, w, s, eeff, st.mocache + ((size_t)w * S.ne + e) * 5 * ldmax, ldmax, lane, gm, r);
#pragma unroll
    for (int i = 0; i < 3; ++i) {
      double gs = r[1 + i] / r[0];
      if (!isfinite(gs)) gs = 0.0;
      grad[i] = gs;
    }
  }
  if (S.na + S.nb > 0) {
    double du, gj[3], lj;
    if (st.jold != nullptr)  // fused chain: the pair caches are kept current (k_pair_cache_build, coop_jastrow_commit_pbc)
      coop_jastrow_cached_grad<G>(S, st, w, e, lane, gm, gj);
    else
      coop_jastrow_pbc<1, G>(S, sd, si, st, w, e, ox, oy, oz, lane, gm, du, gj, lj);
#pragma unroll
    for (int i = 0; i < 3; ++i) grad[i] = grad[i] + gj[i];
  }
  limdrift3(grad);
  if (lane == 0) {
    const double* __restrict__ gauss = gauss_e + (size_t)w * 3;
    const double nx = __dadd_rn(__dadd_rn(ox, gauss[0]), __dmul_rn(grad[0], tstep));
    const double ny = __dadd_rn(__dadd_rn(oy, gauss[1]), __dmul_rn(grad[1], tstep));
    const double nz = __dadd_rn(__dadd_rn(oz, gauss[2]), __dmul_rn(grad[2], tstep));
    double o[3] = {nx, ny, nz}, ww[3] = {0.0, 0.0, 0.0};
    if (S.pbc) wrap_cell(sd + S.o_lat, sd + S.o_latinv, nx, ny, nz, o, ww);
#pragma unroll
    for (int i = 0; i < 3; ++i) {
      st.saved_pos[(size_t)w * 3 + i] = o[i];
      st.saved_wrap[(size_t)w * 3 + i] = st.wrap[((size_t)w * S.ne + e) * 3 + i] + ww[i];
      st.gold[(size_t)w * 3 + i] = grad[i];
    }
  }
}

__global__ void __launch_bounds__(128) k_pbc_propose(const Sys S, const State st, const PbcMoveArgs a) {
  const double* sd;
  const int* si;
  stage_tables(S, sd, si);
  const int lane = threadIdx.x & 31;
  const int w = ((blockIdx.x * blockDim.x + threadIdx.x) >> 5) + a.w0;
  if (w >= (a.wn > 0 ? a.w0 + a.wn : st.N)) return;
  pbc_propose_warp(S, sd, si, st, w, a.e, a.tstep, a.gauss, lane);
}

// FUSED: the accepted walkers' inverse is updated by the same warp (sm_warp_apply, n <= 32, one determinant)
// and the proposal of electron e + 1 follows in the same launch -- two launches per electron move
// (orbitals at the proposed points, this kernel) instead of four.
template <bool FUSED>
__global__ void __launch_bounds__(128) k_pbc_accept(const Sys S, const State st, const PbcMoveArgs a) {
  constexpr int G = 32;
  const double* sd;
  const int* si;
  stage_tables(S, sd, si);
  extern __shared__ __align__(128) unsigned char qmcb_smem[];
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  const unsigned gm = 0xffffffffu;
  const int w = ((blockIdx.x * blockDim.x + threadIdx.x) >> 5) + a.w0;
  const int N = st.N;
  if (w >= (a.wn > 0 ? a.w0 + a.wn : N)) return;
  const size_t tab = (16 + (size_t)S.dwords * 8 + (size_t)S.iwords * 4 + 15) & ~(size_t)15;
  const int npb = (S.ne > 1 ? S.ne - 1 : 0) * S.nb;
  const int jper = pbc_accept_jper(S);
  double* jtmp = reinterpret_cast<double*>(qmcb_smem + tab) + (size_t)wib * jper;
  double* gps = jtmp + npb + S.natom * S.na;  // [(ne-1)][3] pair-gradient terms, then [3] electron-ion gradient
  double* gas = gps + (S.ne > 1 ? S.ne - 1 : 0) * 3;
  const bool cached = FUSED && st.jold != nullptr;  // radial values and pair terms kept for the cache update
  const int e = a.e;
  const int s = e >= S.nup ? 1 : 0;
  const int eeff = e - s * S.nup;
  const int ldmax = S.ldc[0] > S.ldc[1] ? S.ldc[0] : S.ldc[1];
  const double nx = st.saved_pos[(size_t)w * 3], ny = st.saved_pos[(size_t)w * 3 + 1], nz = st.saved_pos[(size_t)w * 3 + 2];
  const double* __restrict__ rows = st.monew + (size_t)w * 5 * ldmax;
  const bool has_s = S.nmo[0] + S.nmo[1] > 0, has_j = S.na + S.nb > 0;
  double ngrad[3] = {0.0, 0.0, 0.0}, val = 1.0;
  if (has_s) {
    double r[4];
    slater_row_ratio4<G>(S, si, st, w, s, eeff, rows, ldmax, lane, gm, r);
#pragma unroll
    for (int i = 0; i < 3; ++i) {
      double gs = r[1 + i] / r[0];
      if (!isfinite(gs)) gs = 0.0;
      ngrad[i] = gs;
    }
    val = isfinite(r[0]) ? r[0] : 1.0;
  }
  if (has_j) {
    double du, gj[3], lj;
    double ga[3] = {0.0, 0.0, 0.0};
    coop_jastrow_pbc<1, G>(S, sd, si, st, w, e, nx, ny, nz, lane, gm, du, gj, lj, cached ? jtmp : nullptr,
                           cached ? jtmp + npb : nullptr, cached ? gps : nullptr, cached ? ga : nullptr);
    if (cached && lane < 3) gas[lane] = ga[lane];
#pragma unroll
    for (int i = 0; i < 3; ++i) ngrad[i] = ngrad[i] + gj[i];
    val = val * exp(du);
  }
  limdrift3(ngrad);
  const double* __restrict__ gauss = a.gauss + (size_t)w * 3;
  double fwd = 0.0, bwd = 0.0;
#pragma unroll
  for (int i = 0; i < 3; ++i) {
    fwd = __dadd_rn(fwd, __dmul_rn(gauss[i], gauss[i]));
    const double b = __dadd_rn(gauss[i], __dmul_rn(a.tstep, __dadd_rn(st.gold[(size_t)w * 3 + i], ngrad[i])));
    bwd = __dadd_rn(bwd, __dmul_rn(b, b));
  }
  const double tprob = exp(__dmul_rn(1.0 / (2.0 * a.tstep), __dadd_rn(fwd, -bwd)));
  const double aval = fabs(val);
  const double ratio = __dmul_rn(__dmul_rn(aval, aval), tprob);
  const bool acc = __shfl_sync(gm, (ratio > a.unif[w]) ? 1 : 0, 0) != 0;
  if (lane == 0) {
    a.accept[w] = acc ? 1 : 0;
    if (acc) atomicAdd(a.nacc, 1ULL);
  }
  if (!acc && !FUSED) return;
  if (acc) {
    if (has_j && cached) {
      __syncwarp(gm);
      coop_jastrow_commit_pbc<G>(S, st, w, e, lane, gm, jtmp, jtmp + npb, gps, gas);
    } else if (has_j) {
      coop_jastrow_update_pbc<G>(S, sd, si, st, w, e, nx, ny, nz, lane, gm, jtmp);
    }
    if (has_s) {
      double* __restrict__ mc = st.mocache + ((size_t)w * S.ne + e) * 5 * ldmax;
      for (int i = lane; i < 5 * ldmax; i += G) mc[i] = rows[i];
      if (FUSED) {
        const int n = s ? S.ndn : S.nup;
        double* inv = st.inv[s] + (size_t)w * n * n;
        const double vk = lane < n ? rows[si[S.o_occ[s] + lane]] : 0.0;
        double dr;
        if (n <= 8)
          dr = sm_warp_apply<8>(inv, n, eeff, lane, vk);
        else if (n <= 16)
          dr = sm_warp_apply<16>(inv, n, eeff, lane, vk);
        else
          dr = sm_warp_apply<32>(inv, n, eeff, lane, vk);
        if (lane == 0) {
          st.dsign[s][w] *= sgn(dr);
          st.dlog[s][w] += log(fabs(dr));
        }
      }
    }
    __syncwarp(gm);
    if (lane < 3) {
      CONF(st, S, w, e, lane) = st.saved_pos[(size_t)w * 3 + lane];
      st.wrap[((size_t)w * S.ne + e) * 3 + lane] = st.saved_wrap[(size_t)w * 3 + lane];
    }
  }
  if (FUSED && a.gauss_next != nullptr) {
    __syncwarp(gm);  // the walker's caches, inverse and position written above are read by all lanes below
    pbc_propose_warp(S, sd, si, st, w, e + 1, a.tstep, a.gauss_next, lane);
  }
}

// =========================================================================================
// Ewald energy, one CTA per walker (ewald.py:242-354): real-space electron-ion and electron-
// electron sums over the minimal image plus the (2 nlatvec + 1)^3 displacements, reciprocal sums
// over the selected G points, and the self / charged-system constants.  out: [2][N] = ee, ei.
// Terms with alpha r > 6.5 (erfc < 4e-20) are skipped.
// =========================================================================================
__global__ void __launch_bounds__(128) k_ewald(const Sys S, const State st, double* __restrict__ out) {
  const double* sd;
  const int* si;
  stage_tables(S, sd, si);
  extern __shared__ __align__(128) unsigned char qmcb_smem[];
  const size_t tab = (16 + (size_t)S.dwords * 8 + (size_t)S.iwords * 4 + 15) & ~(size_t)15;
  double* conf = reinterpret_cast<double*>(qmcb_smem + tab);  // [ne][3]
  double* red = conf + 3 * S.ne;                               // [4 warps][4]
  const int w = blockIdx.x;
  const int N = st.N;
  const int ne = S.ne;
  for (int i = threadIdx.x; i < 3 * ne; i += blockDim.x) conf[i] = st.conf[(size_t)w * ne * 3 + i];
  __syncthreads();
  const double alpha = S.ew_alpha;
  double ei_real = 0.0, ee_real = 0.0, ee_rec = 0.0, ei_rec = 0.0;
  auto cij = [&](double dx, double dy, double dz) {
    min_image(S, sd, dx, dy, dz);
    double acc = 0.0;
    for (int d = 0; d < S.ew_ndisp; ++d) {
      const double x = dx + S.ew_disp[3 * d], y = dy + S.ew_disp[3 * d + 1], z = dz + S.ew_disp[3 * d + 2];
      const double r = sqrt(x * x + y * y + z * z);
      const double ar = alpha * r;
      if (ar < 6.5) acc += erfc(ar) / r;
    }
    return acc;
  };
  for (int t = threadIdx.x; t < S.natom * ne; t += blockDim.x) {
    const int I = t / ne, i = t - I * ne;
    ei_real -= sd[S.o_chg + I] * cij(conf[3 * i] - sd[S.o_xyz + 3 * I], conf[3 * i + 1] - sd[S.o_xyz + 3 * I + 1],
                                     conf[3 * i + 2] - sd[S.o_xyz + 3 * I + 2]);
  }
  const int npair = ne * (ne - 1) / 2;
  for (int t = threadIdx.x; t < npair; t += blockDim.x) {
    int i = 0, rem = t;
    while (rem >= ne - 1 - i) {
      rem -= ne - 1 - i;
      ++i;
    }
    const int j = i + 1 + rem;
    ee_real += cij(conf[3 * i] - conf[3 * j], conf[3 * i + 1] - conf[3 * j + 1], conf[3 * i + 2] - conf[3 * j + 2]);
  }
  for (int gI = threadIdx.x; gI < S.ew_nG; gI += blockDim.x) {
    const double gx = S.ew_g[4 * gI], gy = S.ew_g[4 * gI + 1], gz = S.ew_g[4 * gI + 2], wgt = S.ew_g[4 * gI + 3];
    double ssin = 0.0, scos = 0.0;
    for (int i = 0; i < ne; ++i) {
      double sn, cs;
      sincos(conf[3 * i] * gx + conf[3 * i + 1] * gy + conf[3 * i + 2] * gz, &sn, &cs);
      ssin += sn;
      scos += cs;
    }
    ee_rec = fma(ssin * ssin + scos * scos, wgt, ee_rec);
    ei_rec = fma(-S.ew_ion[2 * gI] * scos - S.ew_ion[2 * gI + 1] * ssin, wgt, ei_rec);
  }
  double v[4] = {ei_real, ee_real, ee_rec, ei_rec};
#pragma unroll
  for (int k = 0; k < 4; ++k) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v[k] += __shfl_xor_sync(0xffffffffu, v[k], o);
  }
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  if (lane == 0)
    for (int k = 0; k < 4; ++k) red[wid * 4 + k] = v[k];
  __syncthreads();
  if (threadIdx.x == 0) {
    double tot[4] = {0.0, 0.0, 0.0, 0.0};
    for (int q = 0; q < (int)(blockDim.x >> 5); ++q)
      for (int k = 0; k < 4; ++k) tot[k] += red[q * 4 + k];
    const double dne = (double)ne;
    const double ee = (tot[1] + tot[2]) + (dne * (dne - 1.0) / 2.0 * S.ew_ijconst + dne * S.ew_sqconst);  // ee_const
    const double ei = (tot[0] + 2.0 * tot[3]) + (-dne * S.ew_isum * S.ew_ijconst);                        // ei_const
    out[w] = ee;
    out[(size_t)N + w] = ei;
  }
}
