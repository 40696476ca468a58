// device_common.cuh -- system/state descriptors, TMA table staging, GTO->MO and Jastrow
// device functions shared by every kernel of libqmcb200.
//
// Arithmetic restated from the reference (citations relative to /root/reference):
//   GTO value/grad/Laplacian      pyqmc/wf/numba/gto.py:89-254, 257-321
//   AO -> MO contraction          pyqmc/wf/orbitals.py:95-96
//   Jastrow radial functions      pyqmc/wf/func3d.py:25-49 (PolyPade), 112-181 (CutoffCusp)
//   r < rcut selection            pyqmc/wf/func3d.py:299-333
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

#include "sph_gen.cuh"

#define QMCB_MAX_ATOM_L 5

// ---------------------------------------------------------------------------------------
// System description: small POD passed by value to every kernel.  The tables themselves live
// in two packed blobs in global memory (doubles / int32) that each CTA stages into shared
// memory with one TMA bulk copy per blob (cp.async.bulk + mbarrier).
// ---------------------------------------------------------------------------------------
struct Sys {
  int natom, nshell, nprim, nao;
  int nup, ndn, ne;
  int nmo[2], ldc[2];  // MOs per spin; padded leading dimension of the MO matrix (mult. of 8)
  int nds[2], ndet;    // unique spin determinants, total determinants
  int fast;            // single determinant with identity occupation and n_s <= 8
  int na, nb;
  int na3, nb3;  // three-body Jastrow basis sizes (0: factor absent)
  int npair;     // ne (ne - 1) / 2
  int necp, nchan, nterm, max_naip, tot_naip;
  double rcut_a, rcut_b, ecp_threshold, e_ii;
  double rcut_a3, rcut_b3;
  int o_c3, o_a3par, o_b3par, o_a3kind, o_b3kind;  // C[I][k][l][m][3] symmetrised; basis tables
  // offsets into the double blob
  int o_xyz, o_chg, o_prim, o_mo[2], o_apar, o_bpar, o_acoef, o_bcoef, o_talpha, o_tcoef;
  // offsets into the int blob
  int o_atsh, o_shl, o_shprim, o_shao, o_occ[2], o_akind, o_bkind;
  int o_primatom, o_shatom, o_aoshell, o_sphoff, o_sphtask;  // cooperative (warp-per-walker) tables
  int nsph, nsphtask;  // sum over atoms of (lmax+1)^2; number of (atom, l) tasks
  int o_ecpatom, o_chanoff, o_termoff, o_tpow, o_naip, o_aipoff;
  int dwords, iwords;  // padded blob lengths (elements)
  const double* dblob;
  const int* iblob;
  // per-determinant tables stay in global memory (too large for shared memory at CAS sizes)
  const int* map[2];      // [ndet] total determinant -> unique spin determinant
  const double* detc;     // [ndet]
  const int* grp_off[2];  // [nds+1] CSR: total determinants that use unique spin det d
  const int* grp_det[2];  // [ndet]
  const double* grp_coef[2];  // [ndet] c_D in group order
  const int* grp_other[2];    // [ndet] map_other(D) in group order
  // dense determinant-coefficient matrix (CAS-like expansions, nds[0] * nds[1] <= 65536, else null):
  // dense[s][j * nds[s] + d] = sum of c_D over the determinants D with map_s(D) = d and map_other(D) = j, so that
  // lanes over d read consecutive addresses in the W update (k_det_cache)
  const double* dense[2];
  // ---- periodic boundary conditions (pbc == 0: open).  pbc is the minimal-image mode of
  // MinimalImageDistance (distance.py:97-110): 1 diagonal, 2 orthogonal, 3 general (27 shifts).
  int pbc;
  int nbatom, o_bxyz;  // atoms that carry basis shells: the PRIMITIVE cell when periodic, else == natom / o_xyz
  int nk, nL, isgamma, ncand, maxao_atom;
  int o_lat, o_latinv, o_shifts;                 // simulation cell rows, inverse, 27 image shifts
  int o_lprim, o_lpriminv, o_smat, o_kl;         // primitive cell, supercell matrix S, k . a_i  [nk][3]
  int o_Ls, o_atomcut, o_lcut, o_phase;          // sorted images, r^2 cutoffs (pbcgto.py:551-591), exp(i L.k) [nL][nk]
  int o_numLs, o_candoff, o_mok[2];              // int blob: images per atom, prefix sum, MO -> k-point index
  // Ewald tables stay in global memory (ewald.py:93-200)
  int ew_ndisp, ew_nG;
  double ew_alpha, ew_ijconst, ew_sqconst, ew_isum;
  const double* ew_disp;  // [ndisp][3]
  const double* ew_g;     // [nG][4]  G vector, weight
  const double* ew_ion;   // [nG][2]  Re, Im of sum_I Z_I exp(i G.R_I)
  // ---- complex wave functions (slater.py:212-216, orbitals.py:34-39,61-65,160-165): complex MO coefficients and /
  // or complex Bloch phases.  An MO row then holds nmo_t real parts followed (at column cxoff) by nmo_t imaginary
  // parts, so nmo[s] = 2 nmo_t[s] and every row buffer keeps its real layout; for open boundaries the coefficient
  // table is [Re C | Im C] and the orbital evaluation itself is unchanged.  Inverses, determinant phases and the
  // multi-determinant caches carry a separate imaginary array (State::*_im).
  int cplx;
  int nkp;                 // phase-table columns / AO accumulator planes: nk (real), 2 nk (cos | sin) when complex
  int nmo_t[2], cxoff[2];  // true orbital count per spin; column offset of the imaginary parts in an MO row
  const double* detc_im;       // [ndet]
  const double* grp_coef_im[2];  // [ndet] Im c_D in group order
};

// Walker state (device pointers).  All arrays are walker-major; Slater arrays keep the
// reference's layout (inverse[s] (N, D_s, n, n) indexed [orbital, electron]; slater.py:254-259).
struct State {
  int N;
  double* inv[2];    // [N][D_s][n][n]
  double* dsign[2];  // [N][D_s]
  double* dlog[2];   // [N][D_s]
  double* dv[2];     // [N][D_s]  sign * exp(log - ref[w])          (multi-determinant cache)
  double* W[2];      // [N][D_s]  sum_{D: map_s(D)=d} c_D dv_other  (multi-determinant cache)
  double* ref[2];    // [N]
  double* conf;      // [N][ne][3]  current walker coordinates (Jastrow._configscurrent)
  double* a_partial; // [N][ne][I][na]
  double* b_partial; // [N][ne][nb][2]
  double* avalues;   // [N][I][na][2]
  double* bvalues;   // [N][nb][3]
  double* mocache;   // [N][ne][5][ldmax]  MO value/grad/Laplacian rows at the current positions
  double* bpair;     // [N][npair][nb]  b_l(r_ij) of every electron pair        (sweep kernel cache)
  double* gpair;     // [N][npair][3]   sum_l c_l g_l(r_ij) (r_i - r_j)          (sweep kernel cache)
  double* agrad;     // [N][ne][3]      electron-ion part of grad_e U            (sweep kernel cache)
  double* lpair;     // [N][npair]      sum_l c_l lap_l(r_ij)                    (sweep kernel cache)
  double* alap;      // [N][ne]         electron-ion part of lap_e U             (sweep kernel cache)
  double* a3v;       // [N][ne][I][na3]  three-body a_k(r_eI)      (three_body_jastrow.py:103)
  double* P3;        // [N][ne]          P_i                       (three_body_jastrow.py:98-101)
  double* val3;      // [N]              U = 1/2 sum_i P_i
  double* saved_mo;  // [N][ldc]  MO row at the last gradient_value/testvalue position
  double* saved_pos; // [N][3]
  double* mo_all;    // [N][ne][ldcmax]  recompute scratch
  double* wrap;      // [N][ne][3]  periodic: integer wrap vectors of the walkers (coord.py:137-152)
  double* saved_wrap;// [N][3]      wrap of the position in saved_pos
  double* monew;     // [N][5][ldmax]  periodic VMC: MO rows at the proposed position
  double* gold;      // [N][3]      periodic VMC: limited drift at the old position
  double* jold;      // [N][ne-1][nb] periodic block driver: b_l(r_ej) at the old position of the proposed electron (or null)
  // complex wave functions: imaginary parts (dsign / dphs_im = Re / Im of the unit phase of each determinant)
  double* inv_im[2];
  double* dphs_im[2];
  double* dv_im[2];
  double* W_im[2];
};

// walker-major accessors: everything one walker owns is contiguous, so a warp that works on one
// walker touches a handful of cache lines
#define CONF(st, S, w, e, x) (st).conf[((size_t)(w) * (S).ne + (e)) * 3 + (x)]
#define APART(st, S, w, e, I, k) (st).a_partial[(((size_t)(w) * (S).ne + (e)) * (S).natom + (I)) * (S).na + (k)]
#define BPART(st, S, w, e, l, t) (st).b_partial[(((size_t)(w) * (S).ne + (e)) * (S).nb + (l)) * 2 + (t)]
#define AVAL(st, S, w, I, k, t) (st).avalues[(((size_t)(w) * (S).natom + (I)) * (S).na + (k)) * 2 + (t)]
#define BVAL(st, S, w, l, t) (st).bvalues[((size_t)(w) * (S).nb + (l)) * 3 + (t)]
#define A3V(st, S, w, e, I, k) (st).a3v[(((size_t)(w) * (S).ne + (e)) * (S).natom + (I)) * (S).na3 + (k)]
#define QMCB_J3_MAXA 128  // natom * na3 held per thread
#define QMCB_J3_MAXB 8

// ---------------------------------------------------------------------------------------
// TMA staging of the table blobs:  [mbarrier | double blob | int blob] in dynamic smem.
// ---------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

__device__ __forceinline__ void stage_tables(const Sys& S, const double*& sd, const int*& si) {
  extern __shared__ __align__(128) unsigned char qmcb_smem[];
  uint64_t* mbar = reinterpret_cast<uint64_t*>(qmcb_smem);
  double* d = reinterpret_cast<double*>(qmcb_smem + 16);
  int* i = reinterpret_cast<int*>(qmcb_smem + 16 + (size_t)S.dwords * 8);
  const uint32_t mb = smem_u32(mbar);
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(mb), "r"(1));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    const uint32_t dbytes = (uint32_t)S.dwords * 8u, ibytes = (uint32_t)S.iwords * 4u;
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(mb),
                 "r"(dbytes + ibytes)
                 : "memory");
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::
            "r"(smem_u32(d)),
        "l"(S.dblob), "r"(dbytes), "r"(mb)
        : "memory");
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::
            "r"(smem_u32(i)),
        "l"(S.iblob), "r"(ibytes), "r"(mb)
        : "memory");
  }
  uint32_t done = 0;
  while (!done) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done)
        : "r"(mb), "r"(0)
        : "memory");
  }
  sd = d;
  si = i;
}

// ---------------------------------------------------------------------------------------
// GTO shells -> MO accumulators.  One thread evaluates one point.
//   DERIV 0: value            NC = 1
//   DERIV 1: value + gradient NC = 4
//   DERIV 2: + Laplacian      NC = 5
// acc[c][j] += chi_c(mu) * C[mu][mo0 + j],  j < NMOT.
// ---------------------------------------------------------------------------------------
template <int DERIV>
struct NComp {
  static constexpr int value = DERIV == 0 ? 1 : (DERIV == 1 ? 4 : 5);
};

template <int L, int DERIV, int NMOT>
__device__ __forceinline__ void shell_accumulate(double x, double y, double z, double R, double Rp,
                                                 double Rl, const double* __restrict__ Crow, int ldc,
                                                 double (&acc)[NComp<DERIV>::value][NMOT]) {
  constexpr int NF = 2 * L + 1;
  double s[NF], gx[NF], gy[NF], gz[NF];
  if constexpr (L == 0) sph_l0<(DERIV > 0)>(x, y, z, s, gx, gy, gz);
  if constexpr (L == 1) sph_l1<(DERIV > 0)>(x, y, z, s, gx, gy, gz);
  if constexpr (L == 2) sph_l2<(DERIV > 0)>(x, y, z, s, gx, gy, gz);
  if constexpr (L == 3) sph_l3<(DERIV > 0)>(x, y, z, s, gx, gy, gz);
  if constexpr (L == 4) sph_l4<(DERIV > 0)>(x, y, z, s, gx, gy, gz);
  if constexpr (L == 5) sph_l5<(DERIV > 0)>(x, y, z, s, gx, gy, gz);
  const double dRx = Rp * x, dRy = Rp * y, dRz = Rp * z;  // dR/dx_i (gto.py:290-296)
#pragma unroll
  for (int m = 0; m < NF; ++m) {
    double comp[NComp<DERIV>::value];
    comp[0] = s[m] * R;
    if (DERIV > 0) {
      comp[1] = gx[m] * R + s[m] * dRx;
      comp[2] = gy[m] * R + s[m] * dRy;
      comp[3] = gz[m] * R + s[m] * dRz;
    }
    if (DERIV > 1) {
      // lap chi = S * lap-radial + 2 grad S . grad R   (gto.py:241-250)
      comp[4] = s[m] * Rl + 2.0 * (gx[m] * dRx + gy[m] * dRy + gz[m] * dRz);
    }
    const double* __restrict__ c = Crow + m * ldc;
#pragma unroll
    for (int j = 0; j < NMOT; ++j) {
      const double cj = c[j];
#pragma unroll
      for (int k = 0; k < NComp<DERIV>::value; ++k) acc[k][j] = fma(comp[k], cj, acc[k][j]);
    }
  }
}

// ---------------------------------------------------------------------------------------
// Periodic boundary conditions: numpy's float divmod (npy_divmod) and Python's %, enforce_pbc
// (pbc/pbc.py:17-49) and the minimal-image conventions of distance.py:133-159.
// ---------------------------------------------------------------------------------------
__device__ __forceinline__ double py_mod(double a, double b) {
  double m = fmod(a, b);
  if (m != 0.0) {
    if ((b < 0.0) != (m < 0.0)) m += b;
  } else {
    m = copysign(0.0, b);
  }
  return m;
}

__device__ __forceinline__ void np_divmod1(double a, double& q, double& r) {  // np.divmod(a, 1)
  double mod = fmod(a, 1.0);
  double div = a - mod;
  if (mod != 0.0) {
    if (mod < 0.0) {
      mod += 1.0;
      div -= 1.0;
    }
  } else {
    mod = 0.0;
  }
  double fl;
  if (div != 0.0) {
    fl = floor(div);
    if (div - fl > 0.5) fl += 1.0;
  } else {
    fl = copysign(0.0, a);
  }
  q = fl;
  r = mod;
}

// position -> (position inside the cell, integer wrap): lat / inv are row-major 3x3 (rows = lattice vectors)
__device__ __forceinline__ void wrap_cell(const double* __restrict__ lat, const double* __restrict__ inv, double x,
                                          double y, double z, double (&o)[3], double (&w)[3]) {
  double rem[3];
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    const double f = __dadd_rn(__dadd_rn(__dmul_rn(x, inv[k]), __dmul_rn(y, inv[3 + k])), __dmul_rn(z, inv[6 + k]));
    np_divmod1(f, w[k], rem[k]);
  }
#pragma unroll
  for (int k = 0; k < 3; ++k)
    o[k] = __dadd_rn(__dadd_rn(__dmul_rn(rem[0], lat[k]), __dmul_rn(rem[1], lat[3 + k])), __dmul_rn(rem[2], lat[6 + k]));
}

__device__ __forceinline__ void min_image(const Sys& S, const double* __restrict__ sd, double& x, double& y, double& z) {
  if (S.pbc == 3) {
    // argmin_i |d + s_i|^2 = argmin_i (2 d.s_i + |s_i|^2): three FMAs per shift; |s_i|^2 is tabulated
    // behind the shifts.  First minimum in shift order, as np.argmin (distance.py:133-142); a different
    // pick than the reference's direct evaluation is only possible at an exact tie of two images, i.e. on
    // the Wigner-Seitz boundary, beyond every cutoff of this path.
    const double* __restrict__ sh = sd + S.o_shifts;
    const double* __restrict__ sh2 = sh + 81;
    const double tx = 2.0 * x, ty = 2.0 * y, tz = 2.0 * z;
    double best = INFINITY;
    int bi = 0;
#pragma unroll
    for (int i = 0; i < 27; ++i) {
      const double sc = fma(tx, sh[3 * i], fma(ty, sh[3 * i + 1], fma(tz, sh[3 * i + 2], sh2[i])));
      if (sc < best) {
        best = sc;
        bi = i;
      }
    }
    x = x + sh[3 * bi];
    y = y + sh[3 * bi + 1];
    z = z + sh[3 * bi + 2];
  } else if (S.pbc == 1) {
    const double* __restrict__ lat = sd + S.o_lat;
    x = py_mod(x + lat[0] / 2, lat[0]) - lat[0] / 2;
    y = py_mod(y + lat[4] / 2, lat[4]) - lat[4] / 2;
    z = py_mod(z + lat[8] / 2, lat[8]) - lat[8] / 2;
  } else if (S.pbc == 2) {
    const double* __restrict__ lat = sd + S.o_lat;
    const double* __restrict__ inv = sd + S.o_latinv;
    double f[3];
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      const double v = __dadd_rn(__dadd_rn(__dmul_rn(x, inv[k]), __dmul_rn(y, inv[3 + k])), __dmul_rn(z, inv[6 + k]));
      f[k] = py_mod(v + 0.5, 1.0) - 0.5;
    }
    x = __dadd_rn(__dadd_rn(__dmul_rn(f[0], lat[0]), __dmul_rn(f[1], lat[3])), __dmul_rn(f[2], lat[6]));
    y = __dadd_rn(__dadd_rn(__dmul_rn(f[0], lat[1]), __dmul_rn(f[1], lat[4])), __dmul_rn(f[2], lat[7]));
    z = __dadd_rn(__dadd_rn(__dmul_rn(f[0], lat[2]), __dmul_rn(f[1], lat[5])), __dmul_rn(f[2], lat[8]));
  }
}

template <int DERIV, int NMOT>
__device__ __forceinline__ void eval_mo(const Sys& S, const double* __restrict__ sd,
                                        const int* __restrict__ si, int spin, double px, double py,
                                        double pz, int mo0,
                                        double (&acc)[NComp<DERIV>::value][NMOT]) {
#pragma unroll
  for (int k = 0; k < NComp<DERIV>::value; ++k)
#pragma unroll
    for (int j = 0; j < NMOT; ++j) acc[k][j] = 0.0;
  const double* __restrict__ prim = sd + S.o_prim;  // (alpha, coef) pairs
  const double* __restrict__ C = sd + S.o_mo[spin] + mo0;
  const int ldc = S.ldc[spin];
  for (int a = 0; a < S.nbatom; ++a) {
    const double x = px - sd[S.o_bxyz + 3 * a], y = py - sd[S.o_bxyz + 3 * a + 1],
                 z = pz - sd[S.o_bxyz + 3 * a + 2];
    const double r2 = x * x + y * y + z * z;
    const int sh1 = si[S.o_atsh + a + 1];
    for (int sh = si[S.o_atsh + a]; sh < sh1; ++sh) {
      const int p1 = si[S.o_shprim + sh + 1];
      double R = 0.0, Rp = 0.0, Rl = 0.0;
      for (int p = si[S.o_shprim + sh]; p < p1; ++p) {
        const double al = prim[2 * p], cf = prim[2 * p + 1];
        const double g = cf * exp(-al * r2);  // gto.py:257-269
        R += g;
        if (DERIV > 0) {
          const double t = 2.0 * al * g;
          Rp -= t;
          if (DERIV > 1) Rl = fma(t, 2.0 * al * r2 - 3.0, Rl);  // gto.py:313-320
        }
      }
      const double* Crow = C + (size_t)si[S.o_shao + sh] * ldc;
      switch (si[S.o_shl + sh]) {
        case 0: shell_accumulate<0, DERIV, NMOT>(x, y, z, R, Rp, Rl, Crow, ldc, acc); break;
        case 1: shell_accumulate<1, DERIV, NMOT>(x, y, z, R, Rp, Rl, Crow, ldc, acc); break;
        case 2: shell_accumulate<2, DERIV, NMOT>(x, y, z, R, Rp, Rl, Crow, ldc, acc); break;
        case 3: shell_accumulate<3, DERIV, NMOT>(x, y, z, R, Rp, Rl, Crow, ldc, acc); break;
        case 4: shell_accumulate<4, DERIV, NMOT>(x, y, z, R, Rp, Rl, Crow, ldc, acc); break;
        default: shell_accumulate<5, DERIV, NMOT>(x, y, z, R, Rp, Rl, Crow, ldc, acc); break;
      }
    }
  }
}

// ---------------------------------------------------------------------------------------
// Jastrow radial functions.  WANT 0: value (func3d.py:25-29 / 125-131); 1: value + gradient
// factor g (grad = g * rvec; func3d.py:32-39 / 147-160); 2: g + Laplacian (42-49 / 165-181).
// Caller guarantees r < rcut.
// ---------------------------------------------------------------------------------------
template <int WANT>
__device__ __forceinline__ void radial_func(int kind, double par, double rcut, double r, double& v,
                                            double& g, double& lap) {
  if (kind == 0) {  // PolyPade, par = beta
    if (WANT == 0) {
      const double z = r / rcut;
      const double p = ((3.0 * z - 8.0) * z + 6.0) * (z * z);
      v = (1.0 - p) / (1.0 + par * p);
    } else {
      const double z1 = r / rcut - 1.0;
      const double z12 = z1 * z1;
      const double p = (3.0 * z12 + 4.0 * z1) * z12 + 1.0;
      const double obp = 1.0 / (1.0 + par * p);
      v = (1.0 - p) * obp;
      g = -(1.0 + par) * 12.0 / (rcut * rcut) * obp * obp * z12;
      if (WANT == 2) {
        const double zp = z1 + 1.0;
        lap = g * (5.0 + 2.0 / z1 - 24.0 * par * (zp * zp) * z12 * obp);
      }
    }
  } else {  // CutoffCusp, par = gamma
    const double y = r / rcut;
    const double y1 = y - 1.0;
    const double a = y1 * y1;
    const double b = (a * y1 + 1.0) / 3.0;
    const double ogb = 1.0 / (1.0 + par * b);
    v = (-b * ogb + 1.0 / (3.0 + par)) * rcut;
    if (WANT >= 1) {
      const double c = ogb * ogb / r;
      g = -a * c;
      if (WANT == 2) lap = -c * 2.0 * ((y1 - a * a * par * ogb) * y + a);
    }
  }
}

// out-of-line copy shared by the warp-cooperative kernels (keeps their code footprint small)
template <int WANT>
__device__ __noinline__ void radial_ool(int kind, double par, double rcut, double r, double& v, double& g, double& lap) {
  radial_func<WANT>(kind, par, rcut, r, v, g, lap);
}

// Jastrow terms for electron e of walker w placed at (px,py,pz).
//   du  = sum_c coef*(new - cached partial sums)  (log of the ratio; jastrowspin.py:404-415)
//   g   = grad U,  lap = laplacian U              (jastrowspin.py:296-385)
template <int WANT>
__device__ __forceinline__ void jastrow_point(const Sys& S, const double* __restrict__ sd,
                                              const int* __restrict__ si, const State& st, int w,
                                              int e, double px, double py, double pz, double& du,
                                              double (&g)[3], double& lap) {
  const int s = e >= S.nup ? 1 : 0;
  double ua = 0.0, ub = 0.0, ua_old = 0.0, ub_old = 0.0;
  g[0] = g[1] = g[2] = 0.0;
  lap = 0.0;
  for (int I = 0; I < S.natom; ++I) {
    double dx = px - sd[S.o_xyz + 3 * I], dy = py - sd[S.o_xyz + 3 * I + 1], dz = pz - sd[S.o_xyz + 3 * I + 2];
    if (S.pbc) min_image(S, sd, dx, dy, dz);
    const double r = sqrt(dx * dx + dy * dy + dz * dz);
    const bool in = r < S.rcut_a;
    for (int k = 0; k < S.na; ++k) {
      const double c = sd[S.o_acoef + (I * S.na + k) * 2 + s];
      if (WANT != 2) ua_old = fma(c, APART(st, S, w, e, I, k), ua_old);
      if (in) {
        double v, gg, ll;
        radial_func<WANT>(si[S.o_akind + k], sd[S.o_apar + k], S.rcut_a, r, v, gg, ll);
        ua = fma(c, v, ua);
        if (WANT >= 1) {
          const double cg = c * gg;
          g[0] = fma(cg, dx, g[0]);
          g[1] = fma(cg, dy, g[1]);
          g[2] = fma(cg, dz, g[2]);
        }
        if (WANT == 2) lap = fma(c, ll, lap);
      }
    }
  }
  for (int j = 0; j < S.ne; ++j) {
    if (j == e) continue;
    const int sj = j >= S.nup ? 1 : 0;
    double dx = px - CONF(st, S, w, j, 0), dy = py - CONF(st, S, w, j, 1), dz = pz - CONF(st, S, w, j, 2);
    if (S.pbc) min_image(S, sd, dx, dy, dz);
    const double r = sqrt(dx * dx + dy * dy + dz * dz);
    if (r < S.rcut_b) {
      for (int l = 0; l < S.nb; ++l) {
        const double c = sd[S.o_bcoef + l * 3 + s + sj];
        double v, gg, ll;
        radial_func<WANT>(si[S.o_bkind + l], sd[S.o_bpar + l], S.rcut_b, r, v, gg, ll);
        ub = fma(c, v, ub);
        if (WANT >= 1) {
          const double cg = c * gg;
          g[0] = fma(cg, dx, g[0]);
          g[1] = fma(cg, dy, g[1]);
          g[2] = fma(cg, dz, g[2]);
        }
        if (WANT == 2) lap = fma(c, ll, lap);
      }
    }
  }
  if (WANT != 2) {
    for (int l = 0; l < S.nb; ++l)
      for (int t = 0; t < 2; ++t)
        ub_old = fma(sd[S.o_bcoef + l * 3 + s + t], BPART(st, S, w, e, l, t),
                     ub_old);
  }
  du = (ub - ub_old) + (ua - ua_old);
}

// ---------------------------------------------------------------------------------------
// Three-body Jastrow (three_body_jastrow.py): pair term of electron e (at pos, with a-values
// av/ag/al over (I,k)) and partner j:
//   P_ej = sum_{Iklm} C[I,k,l,m,sp] a_k(r_eI) a_l(r_jI) b_m(r_ej)
// plus gradient / Laplacian contributions w.r.t. the position of e (454-655).
// ---------------------------------------------------------------------------------------
template <int WANT>
__device__ __forceinline__ void j3_a_values(const Sys& S, const double* __restrict__ sd, const int* __restrict__ si,
                                            double px, double py, double pz, double* __restrict__ av,
                                            double* __restrict__ ag, double* __restrict__ al) {
  for (int I = 0; I < S.natom; ++I) {
    double dx = px - sd[S.o_xyz + 3 * I], dy = py - sd[S.o_xyz + 3 * I + 1], dz = pz - sd[S.o_xyz + 3 * I + 2];
    if (S.pbc) min_image(S, sd, dx, dy, dz);  // MinimalImageDistance (distance.py:133-159)
    const double r = sqrt(dx * dx + dy * dy + dz * dz);
    for (int k = 0; k < S.na3; ++k) {
      double v = 0.0, g = 0.0, l = 0.0;
      if (r < S.rcut_a3) radial_ool<WANT>(si[S.o_a3kind + k], sd[S.o_a3par + k], S.rcut_a3, r, v, g, l);
      av[I * S.na3 + k] = v;
      if (WANT >= 1) ag[I * S.na3 + k] = g;
      if (WANT >= 2) al[I * S.na3 + k] = l;
    }
  }
}

template <int WANT>
__device__ __forceinline__ void j3_pair(const Sys& S, const double* __restrict__ sd, const int* __restrict__ si,
                                        const State& st, int w, int e, int j, double px, double py, double pz,
                                        const double* __restrict__ av, const double* __restrict__ ag,
                                        const double* __restrict__ al, double jx, double jy, double jz, double& P,
                                        double (&g)[3], double& lap) {
  double dx = px - jx, dy = py - jy, dz = pz - jz;
  if (S.pbc) min_image(S, sd, dx, dy, dz);
  const double r = sqrt(dx * dx + dy * dy + dz * dz);
  if (!(r < S.rcut_b3)) return;
  double bv[QMCB_J3_MAXB], bg[QMCB_J3_MAXB], bl[QMCB_J3_MAXB];
  for (int m = 0; m < S.nb3; ++m) {
    double v, gg = 0.0, ll = 0.0;
    radial_ool<WANT>(si[S.o_b3kind + m], sd[S.o_b3par + m], S.rcut_b3, r, v, gg, ll);
    bv[m] = v;
    bg[m] = gg;
    bl[m] = ll;
  }
  const int sp = (e >= S.nup ? 1 : 0) + (j >= S.nup ? 1 : 0);
  const int na = S.na3, nb = S.nb3;
  for (int I = 0; I < S.natom; ++I) {
    double ax = px - sd[S.o_xyz + 3 * I], ay = py - sd[S.o_xyz + 3 * I + 1], az = pz - sd[S.o_xyz + 3 * I + 2];
    if (WANT >= 1 && S.pbc) min_image(S, sd, ax, ay, az);
    const double dot = ax * dx + ay * dy + az * dz;
    const double* __restrict__ C = sd + S.o_c3 + (size_t)I * na * na * nb * 3 + sp;
    for (int m = 0; m < nb; ++m) {
      double s0 = 0.0, s1 = 0.0, s2 = 0.0;
      for (int k = 0; k < na; ++k) {
        double t = 0.0;  // sum_l C[I,k,l,m,sp] a_l(r_jI)
        for (int l = 0; l < na; ++l) t = fma(C[((k * na + l) * nb + m) * 3], A3V(st, S, w, j, I, l), t);
        s0 = fma(av[I * na + k], t, s0);
        if (WANT >= 1) s1 = fma(ag[I * na + k], t, s1);
        if (WANT >= 2) s2 = fma(al[I * na + k], t, s2);
      }
      P = fma(s0, bv[m], P);
      if (WANT >= 1) {
        const double ca = s1 * bv[m], cb = s0 * bg[m];
        g[0] += ca * ax + cb * dx;
        g[1] += ca * ay + cb * dy;
        g[2] += ca * az + cb * dz;
      }
      if (WANT >= 2) lap += s2 * bv[m] + 2.0 * s1 * bg[m] * dot + s0 * bl[m];
    }
  }
}

// adds the three-body terms of electron e at (px,py,pz) to du, g, lap (raw Laplacian of U)
template <int WANT>
__device__ __forceinline__ void jastrow3_point(const Sys& S, const double* __restrict__ sd, const int* __restrict__ si,
                                               const State& st, int w, int e, double px, double py, double pz,
                                               double& du, double (&g)[3], double& lap) {
  double av[QMCB_J3_MAXA], ag[QMCB_J3_MAXA], al[QMCB_J3_MAXA];
  j3_a_values<WANT>(S, sd, si, px, py, pz, av, ag, al);
  double P = 0.0;
  for (int j = 0; j < S.ne; ++j) {
    if (j == e) continue;
    j3_pair<WANT>(S, sd, si, st, w, e, j, px, py, pz, av, ag, al, CONF(st, S, w, j, 0), CONF(st, S, w, j, 1),
                  CONF(st, S, w, j, 2), P, g, lap);
  }
  if (WANT != 2) du += P - st.P3[(size_t)w * S.ne + e];
}
