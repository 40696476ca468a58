// coop.cuh -- warp-per-walker building blocks for the device-resident sweep.
//
// One warp owns one walker.  A single-electron evaluation is split across the 32 lanes in
// phases that communicate through a per-warp shared-memory scratch:
//   0  lanes over atoms      displacement + r^2
//   1  lanes over primitives c exp(-a r^2) and its radial derivative factors      (gto.py:257-321)
//   S  lanes over (atom, l)  real solid harmonics of one l and their gradients
//   1b lanes over shells     contraction sums R, R', R''
//   2  lanes over AOs        chi, grad chi, lap chi                              (gto.py:89-254)
//   3  lanes over (component, MO)  AO -> MO contraction                          (orbitals.py:95-96)
// Jastrow sums run with lanes over partner electrons / atoms and a warp-shuffle reduction
// (jastrowspin.py:296-419); the Sherman-Morrison update stages the inverse in shared memory with
// lanes over matrix elements (slater.py:88-94).
#pragma once
#include "device_common.cuh"

struct CoopLayout {  // offsets (doubles) into the per-warp scratch
  int at, pv, sv, sph, comp, mo, minv, tvec, colv, jtmp, total;
};

__host__ __device__ inline CoopLayout coop_layout(const Sys& S) {
  CoopLayout L;
  const int ldmax = S.ldc[0] > S.ldc[1] ? S.ldc[0] : S.ldc[1];
  const int nmax = S.nup > S.ndn ? S.nup : S.ndn;
  int o = 0;
  L.at = o;
  o += 4 * (S.nbatom > S.natom ? S.nbatom : S.natom);
  L.pv = o;
  o += 3 * S.nprim;
  L.sv = o;
  o += 3 * S.nshell;
  L.sph = o;
  o += 4 * S.nsph;
  L.comp = o;
  o += 5 * S.nao;
  L.mo = o;
  o += 5 * ldmax;
  L.minv = o;
  o += nmax * nmax;
  L.tvec = o;
  o += nmax;
  L.colv = o;
  o += nmax;
  L.jtmp = o;
  o += 3 * (S.ne > 1 ? S.ne - 1 : 0) * S.nb + S.natom * S.na;
  L.total = (o + 1) & ~1;
  return L;
}

// A walker is owned by a group of G consecutive lanes (G = 8, 16 or 32); several walkers share a
// warp and may diverge (accept / reject), so every sync and shuffle is scoped to the group mask.
template <int G>
__device__ __forceinline__ unsigned group_mask(int lane) {
  return G == 32 ? 0xffffffffu : (((1u << G) - 1u) << (lane & ~(G - 1)));
}

template <int G>
__device__ __forceinline__ double group_sum(double v, unsigned gm) {
#pragma unroll
  for (int o = G / 2; o > 0; o >>= 1) v += __shfl_xor_sync(gm, v, o);
  return v;
}

template <int L, bool D>
__device__ __forceinline__ void sph_store(double x, double y, double z, double* __restrict__ out) {
  constexpr int NF = 2 * L + 1;
  double s[NF], gx[NF], gy[NF], gz[NF];
  if constexpr (L == 0) sph_l0<D>(x, y, z, s, gx, gy, gz);
  if constexpr (L == 1) sph_l1<D>(x, y, z, s, gx, gy, gz);
  if constexpr (L == 2) sph_l2<D>(x, y, z, s, gx, gy, gz);
  if constexpr (L == 3) sph_l3<D>(x, y, z, s, gx, gy, gz);
  if constexpr (L == 4) sph_l4<D>(x, y, z, s, gx, gy, gz);
  if constexpr (L == 5) sph_l5<D>(x, y, z, s, gx, gy, gz);
#pragma unroll
  for (int m = 0; m < NF; ++m) {
    out[4 * m] = s[m];
    if (D) {
      out[4 * m + 1] = gx[m];
      out[4 * m + 2] = gy[m];
      out[4 * m + 3] = gz[m];
    }
  }
}

// MO rows (value [, gradient [, Laplacian]]) of spin `spin` at (px,py,pz) -> ws[L.mo + c*ldmax + j]
template <int DERIV, int G>
__device__ __forceinline__ void coop_eval_mo(const Sys& S, const CoopLayout& L, const double* __restrict__ sd,
                                          const int* __restrict__ si, int spin, double px, double py, double pz,
                                          double* __restrict__ ws, int lane, unsigned gm) {
  constexpr int NC = NComp<DERIV>::value;
  const int ldmax = S.ldc[0] > S.ldc[1] ? S.ldc[0] : S.ldc[1];
  double* __restrict__ at = ws + L.at;
  double* __restrict__ pv = ws + L.pv;
  double* __restrict__ sv = ws + L.sv;
  double* __restrict__ sph = ws + L.sph;
  double* __restrict__ comp = ws + L.comp;
  double* __restrict__ mo = ws + L.mo;
#pragma unroll 1
  for (int a = lane; a < S.nbatom; a += G) {
    const double x = px - sd[S.o_bxyz + 3 * a], y = py - sd[S.o_bxyz + 3 * a + 1], z = pz - sd[S.o_bxyz + 3 * a + 2];
    at[4 * a] = x;
    at[4 * a + 1] = y;
    at[4 * a + 2] = z;
    at[4 * a + 3] = x * x + y * y + z * z;
  }
  __syncwarp(gm);
  const double* __restrict__ prim = sd + S.o_prim;
#pragma unroll 1
  for (int p = lane; p < S.nprim; p += G) {
    const double r2 = at[4 * si[S.o_primatom + p] + 3];
    const double al = prim[2 * p], cf = prim[2 * p + 1];
    const double g = cf * exp(-al * r2);
    pv[3 * p] = g;
    if (DERIV > 0) {
      const double t = 2.0 * al * g;
      pv[3 * p + 1] = t;
      if (DERIV > 1) pv[3 * p + 2] = t * (2.0 * al * r2 - 3.0);
    }
  }
#pragma unroll 1
  for (int t = lane; t < S.nsphtask; t += G) {
    const int a = si[S.o_sphtask + 2 * t], l = si[S.o_sphtask + 2 * t + 1];
    const double x = at[4 * a], y = at[4 * a + 1], z = at[4 * a + 2];
    double* out = sph + 4 * (si[S.o_sphoff + a] + l * l);
    switch (l) {
      case 0: sph_store<0, (DERIV > 0)>(x, y, z, out); break;
      case 1: sph_store<1, (DERIV > 0)>(x, y, z, out); break;
      case 2: sph_store<2, (DERIV > 0)>(x, y, z, out); break;
      case 3: sph_store<3, (DERIV > 0)>(x, y, z, out); break;
      case 4: sph_store<4, (DERIV > 0)>(x, y, z, out); break;
      default: sph_store<5, (DERIV > 0)>(x, y, z, out); break;
    }
  }
  __syncwarp(gm);
#pragma unroll 1
  for (int sh = lane; sh < S.nshell; sh += G) {
    double R = 0.0, Rp = 0.0, Rl = 0.0;
    const int p1 = si[S.o_shprim + sh + 1];
    for (int p = si[S.o_shprim + sh]; p < p1; ++p) {
      R += pv[3 * p];
      if (DERIV > 0) Rp -= pv[3 * p + 1];
      if (DERIV > 1) Rl += pv[3 * p + 2];
    }
    sv[3 * sh] = R;
    sv[3 * sh + 1] = Rp;
    sv[3 * sh + 2] = Rl;
  }
  __syncwarp(gm);
#pragma unroll 1
  for (int mu = lane; mu < S.nao; mu += G) {
    const int sh = si[S.o_aoshell + mu];
    const int a = si[S.o_shatom + sh];
    const int l = si[S.o_shl + sh];
    const double* __restrict__ sp = sph + 4 * (si[S.o_sphoff + a] + l * l + (mu - si[S.o_shao + sh]));
    const double R = sv[3 * sh];
    const double s = sp[0];
    comp[mu] = s * R;
    if (DERIV > 0) {
      const double Rp = sv[3 * sh + 1];
      const double dRx = Rp * at[4 * a], dRy = Rp * at[4 * a + 1], dRz = Rp * at[4 * a + 2];
      const double gx = sp[1], gy = sp[2], gz = sp[3];
      comp[S.nao + mu] = gx * R + s * dRx;
      comp[2 * S.nao + mu] = gy * R + s * dRy;
      comp[3 * S.nao + mu] = gz * R + s * dRz;
      if (DERIV > 1) comp[4 * S.nao + mu] = s * sv[3 * sh + 2] + 2.0 * (gx * dRx + gy * dRy + gz * dRz);
    }
  }
  __syncwarp(gm);
  const int ldc = S.ldc[spin];
  const double* __restrict__ C = sd + S.o_mo[spin];
  if (ldc == 4 && G == 8) {
    // lanes over components, the four MOs of a row in registers (two 16-byte shared loads per AO)
    if (lane < NC) {
      const double* __restrict__ cp = comp + lane * S.nao;
      double a0 = 0.0, a1 = 0.0, a2 = 0.0, a3 = 0.0;
#pragma unroll 4
      for (int mu = 0; mu < S.nao; ++mu) {
        const double x = cp[mu];
        const double2 c01 = *reinterpret_cast<const double2*>(C + mu * 4);
        const double2 c23 = *reinterpret_cast<const double2*>(C + mu * 4 + 2);
        a0 = fma(x, c01.x, a0);
        a1 = fma(x, c01.y, a1);
        a2 = fma(x, c23.x, a2);
        a3 = fma(x, c23.y, a3);
      }
      mo[lane * ldmax] = a0;
      mo[lane * ldmax + 1] = a1;
      mo[lane * ldmax + 2] = a2;
      mo[lane * ldmax + 3] = a3;
    }
  } else if (ldc == 4 && 2 * NC <= G) {
    // lanes over (component, MO pair): two MOs of a row in registers, one 16-byte load per AO
    if (lane < 2 * NC) {
      const int c = lane >> 1, jh = (lane & 1) * 2;
      const double* __restrict__ cp = comp + c * S.nao;
      double a0 = 0.0, a1 = 0.0;
#pragma unroll 4
      for (int mu = 0; mu < S.nao; ++mu) {
        const double x = cp[mu];
        const double2 cc = *reinterpret_cast<const double2*>(C + mu * 4 + jh);
        a0 = fma(x, cc.x, a0);
        a1 = fma(x, cc.y, a1);
      }
      mo[c * ldmax + jh] = a0;
      mo[c * ldmax + jh + 1] = a1;
    }
  } else {
#pragma unroll 1
    for (int t = lane; t < NC * ldc; t += G) {
      const int c = t / ldc, j = t - c * ldc;
      const double* __restrict__ cp = comp + c * S.nao;
      double acc = 0.0;
      for (int mu = 0; mu < S.nao; ++mu) acc = fma(cp[mu], C[mu * ldc + j], acc);
      mo[c * ldmax + j] = acc;
    }
  }
  __syncwarp(gm);
}

// Jastrow terms of electron e of walker w at (px,py,pz); every lane returns the full sums.
// WANT 1: du (log ratio vs cached partial sums) and grad U;  WANT 2: grad U and laplacian U.
template <int WANT, int G>
__device__ __forceinline__ void coop_jastrow(const Sys& S, const double* __restrict__ sd, const int* __restrict__ si,
                                          const State& st, int w, int e, double px, double py, double pz, int lane,
                                          unsigned gm, double& du, double (&g)[3], double& lap) {
  const int s = e >= S.nup ? 1 : 0;
  double unew = 0.0, uold = 0.0, g0 = 0.0, g1 = 0.0, g2 = 0.0, lp = 0.0;
  const int ntb = (S.ne - 1) * S.nb, nta = S.natom * S.na;
#pragma unroll 1
  for (int t = lane; t < ntb + nta; t += G) {
    double dx, dy, dz, c, rcut, par;
    int kind;
    if (t < ntb) {
      const int jj = t / S.nb, l = t - jj * S.nb;
      const int j = jj < e ? jj : jj + 1;
      dx = px - CONF(st, S, w, j, 0);
      dy = py - CONF(st, S, w, j, 1);
      dz = pz - CONF(st, S, w, j, 2);
      c = sd[S.o_bcoef + l * 3 + s + (j >= S.nup ? 1 : 0)];
      rcut = S.rcut_b;
      par = sd[S.o_bpar + l];
      kind = si[S.o_bkind + l];
    } else {
      const int u = t - ntb;
      const int I = u / S.na, k = u - I * S.na;
      dx = px - sd[S.o_xyz + 3 * I];
      dy = py - sd[S.o_xyz + 3 * I + 1];
      dz = pz - sd[S.o_xyz + 3 * I + 2];
      c = sd[S.o_acoef + (I * S.na + k) * 2 + s];
      rcut = S.rcut_a;
      par = sd[S.o_apar + k];
      kind = si[S.o_akind + k];
    }
    const double r = sqrt(dx * dx + dy * dy + dz * dz);
    if (r < rcut) {
      double v, gg = 0.0, ll = 0.0;
      radial_ool<WANT>(kind, par, rcut, r, v, gg, ll);
      unew = fma(c, v, unew);
      if (WANT >= 1) {
        const double cg = c * gg;
        g0 = fma(cg, dx, g0);
        g1 = fma(cg, dy, g1);
        g2 = fma(cg, dz, g2);
      }
      if (WANT == 2) lp = fma(c, ll, lp);
    }
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

// Jastrow cache update of walker w after electron e moved from its current position to
// (nx,ny,nz)  (jastrowspin.py:111-137, 221-249); also moves the walker coordinate.
template <int G>
__device__ __forceinline__ void coop_jastrow_update(const Sys& S, const double* __restrict__ sd,
                                                 const int* __restrict__ si, const State& st, int w, int e,
                                                 double nx, double ny, double nz, int lane, unsigned gm,
                                                 bool has_jastrow, double* __restrict__ jtmp) {
  const int s = e >= S.nup ? 1 : 0;
  if (has_jastrow) {
    for (int t = lane; t < S.natom * S.na; t += G) {
      const int I = t / S.na, k = t - I * S.na;
      const double dx = nx - sd[S.o_xyz + 3 * I], dy = ny - sd[S.o_xyz + 3 * I + 1], dz = nz - sd[S.o_xyz + 3 * I + 2];
      const double r = sqrt(dx * dx + dy * dy + dz * dz);
      double v = 0.0, gg, ll;
      if (r < S.rcut_a) radial_ool<0>(si[S.o_akind + k], sd[S.o_apar + k], S.rcut_a, r, v, gg, ll);
      AVAL(st, S, w, I, k, s) += v - APART(st, S, w, e, I, k);
      APART(st, S, w, e, I, k) = v;
    }
    const double ox = CONF(st, S, w, e, 0), oy = CONF(st, S, w, e, 1), oz = CONF(st, S, w, e, 2);
    // lanes over (partner, basis function): patch the partners' partial sums, park the new values
    const int ntb = (S.ne - 1) * S.nb;
#pragma unroll 1
    for (int t = lane; t < ntb; t += G) {
      const int jj = t / S.nb, l = t - jj * S.nb;
      const int j = jj < e ? jj : jj + 1;
      const double jx = CONF(st, S, w, j, 0), jy = CONF(st, S, w, j, 1), jz = CONF(st, S, w, j, 2);
      double dx = nx - jx, dy = ny - jy, dz = nz - jz;
      const double rn = sqrt(dx * dx + dy * dy + dz * dz);
      dx = ox - jx;
      dy = oy - jy;
      dz = oz - jz;
      const double ro = sqrt(dx * dx + dy * dy + dz * dz);
      double vn = 0.0, vo = 0.0, gg, ll;
      if (rn < S.rcut_b) radial_ool<0>(si[S.o_bkind + l], sd[S.o_bpar + l], S.rcut_b, rn, vn, gg, ll);
      if (ro < S.rcut_b) radial_ool<0>(si[S.o_bkind + l], sd[S.o_bpar + l], S.rcut_b, ro, vo, gg, ll);
      BPART(st, S, w, j, l, s) += vn - vo;
      jtmp[t] = vn;
    }
    __syncwarp(gm);
    // new partial sums of electron e, accumulated in partner order as _b_update does
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
  if (lane == 0) {
    CONF(st, S, w, e, 0) = nx;
    CONF(st, S, w, e, 1) = ny;
    CONF(st, S, w, e, 2) = nz;
  }
  __syncwarp(gm);
}

// Sherman-Morrison row replacement of the single determinant of spin s of walker w: the new row
// is vec[k] = mo[occ[k]] (values at the accepted position).  slater.py:88-94, 290-291.
template <int G>
__device__ __forceinline__ void coop_sherman_morrison(const Sys& S, const CoopLayout& L, const int* __restrict__ si,
                                                   const State& st, int w, int s, int eeff,
                                                   double* __restrict__ ws, int lane, unsigned gm) {
  const int n = s ? S.ndn : S.nup;
  double* __restrict__ inv = st.inv[s] + (size_t)w * n * n;
  double* __restrict__ minv = ws + L.minv;
  double* __restrict__ tvec = ws + L.tvec;
  double* __restrict__ colv = ws + L.colv;
  const double* __restrict__ mo = ws + L.mo;
  const int* __restrict__ occ = si + S.o_occ[s];
  for (int i = lane; i < n * n; i += G) minv[i] = inv[i];
  __syncwarp(gm);
  for (int j = lane; j < n; j += G) {
    double t = 0.0;
    for (int k = 0; k < n; ++k) t = fma(mo[occ[k]], minv[k * n + j], t);
    tvec[j] = t;
  }
  __syncwarp(gm);
  const double ratio = tvec[eeff];
  for (int k = lane; k < n; k += G) colv[k] = minv[k * n + eeff] / ratio;
  __syncwarp(gm);
  for (int i = lane; i < n * n; i += G) {
    const int k = i / n, j = i - k * n;
    inv[i] = (j == eeff) ? colv[k] : fma(-colv[k], tvec[j], minv[i]);
  }
  if (lane == 0) {
    st.dsign[s][w] *= (ratio > 0.0 ? 1.0 : (ratio < 0.0 ? -1.0 : 0.0));
    st.dlog[s][w] += log(fabs(ratio));
  }
  __syncwarp(gm);
}

// ---------------------------------------------------------------------------------------
// Pair-cached Jastrow for the sweep kernel.  Per walker the context keeps, for every electron
// pair (i<j), the basis values b_l(r_ij) and the gradient term  sum_l c_l g_l(r_ij) (r_i - r_j),
// and for every electron the electron-ion gradient term.  The drift at the CURRENT position is
// then a sum of cached terms, and an accepted move needs no radial-function evaluation beyond
// the ones already done for the proposed position (same arithmetic as jastrowspin.py:296-340,
// 111-137, 221-249; only the order of the partner sum differs).
// ---------------------------------------------------------------------------------------
__device__ __forceinline__ int pair_index(int ne, int i, int j) {  // i < j
  return i * ne - (i * (i + 1)) / 2 + (j - i - 1);
}
#define BPAIR(st, S, w, p, l) (st).bpair[((size_t)(w) * (S).npair + (p)) * (S).nb + (l)]
#define GPAIR(st, S, w, p, x) (st).gpair[((size_t)(w) * (S).npair + (p)) * 3 + (x)]
#define AGRAD(st, S, w, e, x) (st).agrad[((size_t)(w) * (S).ne + (e)) * 3 + (x)]
#define LPAIR(st, S, w, p) (st).lpair[(size_t)(w) * (S).npair + (p)]
#define ALAP(st, S, w, e) (st).alap[(size_t)(w) * (S).ne + (e)]

template <int G>
__device__ __forceinline__ void coop_jastrow_cached_grad(const Sys& S, const State& st, int w, int e, int lane,
                                                         unsigned gm, double (&g)[3]) {
  double g0 = 0.0, g1 = 0.0, g2 = 0.0;
#pragma unroll 1
  for (int t = lane; t < S.ne - 1; t += G) {
    const int j = t < e ? t : t + 1;
    const int p = e < j ? pair_index(S.ne, e, j) : pair_index(S.ne, j, e);
    const double sg = e < j ? 1.0 : -1.0;
    g0 += sg * GPAIR(st, S, w, p, 0);
    g1 += sg * GPAIR(st, S, w, p, 1);
    g2 += sg * GPAIR(st, S, w, p, 2);
  }
  if (lane == 0) {
    g0 += AGRAD(st, S, w, e, 0);
    g1 += AGRAD(st, S, w, e, 1);
    g2 += AGRAD(st, S, w, e, 2);
  }
  g[0] = group_sum<G>(g0, gm);
  g[1] = group_sum<G>(g1, gm);
  g[2] = group_sum<G>(g2, gm);
}

// Laplacian of U with respect to electron e from the cached pair / electron-ion terms
template <int G>
__device__ __forceinline__ double coop_jastrow_cached_lap(const Sys& S, const State& st, int w, int e, int lane,
                                                          unsigned gm) {
  double l = 0.0;
#pragma unroll 1
  for (int t = lane; t < S.ne - 1; t += G) {
    const int j = t < e ? t : t + 1;
    l += LPAIR(st, S, w, e < j ? pair_index(S.ne, e, j) : pair_index(S.ne, j, e));
  }
  if (lane == 0) l += ALAP(st, S, w, e);
  return group_sum<G>(l, gm);
}

// value + gradient at the proposed position; parks per-task values for a possible commit:
//   jt[t]            = b_l(r)                 t = jj * nb + l
//   jt[ntb + t]      = c * g_l(r)             (gradient factor of that task)
//   jt[2 ntb + u]    = a_k(r_eI)              u = I * na + k
//   jt[2 ntb + nta + t] = c * lap_l(r)        (Laplacian term of that task)
// ga = electron-ion part of the gradient, la = electron-ion part of the Laplacian (all lanes).
template <int G>
__device__ __forceinline__ void coop_jastrow_propose(const Sys& S, const double* __restrict__ sd,
                                                     const int* __restrict__ si, const State& st, int w, int e,
                                                     double px, double py, double pz, int lane, unsigned gm,
                                                     double* __restrict__ jt, double& du, double (&g)[3],
                                                     double (&ga)[3], double& la) {
  const int s = e >= S.nup ? 1 : 0;
  const int ntb = (S.ne - 1) * S.nb, nta = S.natom * S.na;
  double unew = 0.0, uold = 0.0, b0 = 0.0, b1 = 0.0, b2 = 0.0, a0 = 0.0, a1 = 0.0, a2 = 0.0, al = 0.0;
#pragma unroll 1
  for (int t = lane; t < ntb + nta; t += G) {
    double dx, dy, dz, c, rcut, par;
    int kind;
    const bool isb = t < ntb;
    if (isb) {
      const int jj = t / S.nb, l = t - jj * S.nb;
      const int j = jj < e ? jj : jj + 1;
      dx = px - CONF(st, S, w, j, 0);
      dy = py - CONF(st, S, w, j, 1);
      dz = pz - CONF(st, S, w, j, 2);
      c = sd[S.o_bcoef + l * 3 + s + (j >= S.nup ? 1 : 0)];
      rcut = S.rcut_b;
      par = sd[S.o_bpar + l];
      kind = si[S.o_bkind + l];
    } else {
      const int u = t - ntb;
      const int I = u / S.na, k = u - I * S.na;
      dx = px - sd[S.o_xyz + 3 * I];
      dy = py - sd[S.o_xyz + 3 * I + 1];
      dz = pz - sd[S.o_xyz + 3 * I + 2];
      c = sd[S.o_acoef + (I * S.na + k) * 2 + s];
      rcut = S.rcut_a;
      par = sd[S.o_apar + k];
      kind = si[S.o_akind + k];
    }
    const double r = sqrt(dx * dx + dy * dy + dz * dz);
    double v = 0.0, gg = 0.0, ll = 0.0;
    if (r < rcut) radial_ool<2>(kind, par, rcut, r, v, gg, ll);
    unew = fma(c, v, unew);
    const double cg = c * gg;
    if (isb) {
      jt[t] = v;
      jt[ntb + t] = cg;
      jt[2 * ntb + nta + t] = c * ll;
      b0 = fma(cg, dx, b0);
      b1 = fma(cg, dy, b1);
      b2 = fma(cg, dz, b2);
    } else {
      jt[2 * ntb + (t - ntb)] = v;
      a0 = fma(cg, dx, a0);
      a1 = fma(cg, dy, a1);
      a2 = fma(cg, dz, a2);
      al = fma(c, ll, al);
    }
  }
  const int na_items = nta, nb_items = S.nb * 2;
#pragma unroll 1
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
  du = group_sum<G>(unew, gm) - group_sum<G>(uold, gm);
  ga[0] = group_sum<G>(a0, gm);
  ga[1] = group_sum<G>(a1, gm);
  ga[2] = group_sum<G>(a2, gm);
  la = group_sum<G>(al, gm);
  g[0] = group_sum<G>(b0, gm) + ga[0];
  g[1] = group_sum<G>(b1, gm) + ga[1];
  g[2] = group_sum<G>(b2, gm) + ga[2];
  __syncwarp(gm);
}

// accepted move: every cache is refreshed from the parked values (no radial evaluations)
template <int G>
__device__ __forceinline__ void coop_jastrow_commit(const Sys& S, const double* __restrict__ sd,
                                                    const int* __restrict__ si, const State& st, int w, int e,
                                                    double nx, double ny, double nz, int lane, unsigned gm,
                                                    bool has_jastrow, const double* __restrict__ jt,
                                                    const double (&ga)[3], double la) {
  const int s = e >= S.nup ? 1 : 0;
  if (has_jastrow) {
    const int ntb = (S.ne - 1) * S.nb;
#pragma unroll 1
    for (int u = lane; u < S.natom * S.na; u += G) {
      const int I = u / S.na, k = u - I * S.na;
      const double v = jt[2 * ntb + u];
      AVAL(st, S, w, I, k, s) += v - APART(st, S, w, e, I, k);
      APART(st, S, w, e, I, k) = v;
    }
#pragma unroll 1
    for (int t = lane; t < ntb; t += G) {
      const int jj = t / S.nb, l = t - jj * S.nb;
      const int j = jj < e ? jj : jj + 1;
      const int p = e < j ? pair_index(S.ne, e, j) : pair_index(S.ne, j, e);
      const double vn = jt[t];
      BPART(st, S, w, j, l, s) += vn - BPAIR(st, S, w, p, l);
      BPAIR(st, S, w, p, l) = vn;
    }
#pragma unroll 1
    for (int t = lane; t < S.nb * 2; t += G) {
      const int l = t >> 1, tt = t & 1;
      double bn = 0.0;
      for (int jj = 0; jj < S.ne - 1; ++jj) {
        const int j = jj < e ? jj : jj + 1;
        if ((j >= S.nup ? 1 : 0) == tt) bn += jt[jj * S.nb + l];
      }
      BVAL(st, S, w, l, s + tt) += bn - BPART(st, S, w, e, l, tt);
      BPART(st, S, w, e, l, tt) = bn;
    }
#pragma unroll 1
    for (int jj = lane; jj < S.ne - 1; jj += G) {
      const int j = jj < e ? jj : jj + 1;
      const int p = e < j ? pair_index(S.ne, e, j) : pair_index(S.ne, j, e);
      double gs = 0.0, ls = 0.0;
      const int nta = S.natom * S.na;
      for (int l = 0; l < S.nb; ++l) {
        gs += jt[ntb + jj * S.nb + l];
        ls += jt[2 * ntb + nta + jj * S.nb + l];
      }
      LPAIR(st, S, w, p) = ls;
      const double sg = e < j ? 1.0 : -1.0;
      // NOTE: sum_l (c g_l) * d differs from sum_l (c g_l d) only by rounding
      GPAIR(st, S, w, p, 0) = sg * gs * (nx - CONF(st, S, w, j, 0));
      GPAIR(st, S, w, p, 1) = sg * gs * (ny - CONF(st, S, w, j, 1));
      GPAIR(st, S, w, p, 2) = sg * gs * (nz - CONF(st, S, w, j, 2));
    }
    if (lane == 0) {
      AGRAD(st, S, w, e, 0) = ga[0];
      AGRAD(st, S, w, e, 1) = ga[1];
      AGRAD(st, S, w, e, 2) = ga[2];
      ALAP(st, S, w, e) = la;
    }
  }
  __syncwarp(gm);
  if (lane == 0) {
    CONF(st, S, w, e, 0) = nx;
    CONF(st, S, w, e, 1) = ny;
    CONF(st, S, w, e, 2) = nz;
  }
  __syncwarp(gm);
}
