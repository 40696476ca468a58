// qmcb200.cu -- host side of libqmcb200.so (C ABI declared in include/qmcb200.h).
// Owns device memory, packs the system tables, launches the kernels of kernels.cuh.
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/qmcb200.h"
#undef QMCB_SLATER
#undef QMCB_JASTROW
#undef QMCB_JASTROW3
#include "kernels.cuh"
#include "device_rng.cuh"
#include "mt_jump_poly.h"

namespace {

thread_local std::string g_err;

int fail(const std::string& msg) {
  g_err = msg;
  return -1;
}

#define CK(call)                                                                              \
  do {                                                                                        \
    cudaError_t e_ = (call);                                                                  \
    if (e_ != cudaSuccess)                                                                    \
      return fail(std::string(#call) + " failed: " + cudaGetErrorString(e_) + " (" __FILE__ ":" + \
                  std::to_string(__LINE__) + ")");                                            \
  } while (0)

template <class T>
struct DBuf {
  T* p = nullptr;
  size_t n = 0;
  int ensure(size_t count) {
    if (count <= n && p) return 0;
    if (p) cudaFree(p);
    p = nullptr;
    n = 0;
    if (count == 0) count = 1;
    cudaError_t e = cudaMalloc(&p, count * sizeof(T));
    if (e != cudaSuccess) return fail(std::string("cudaMalloc: ") + cudaGetErrorString(e));
    n = count;
    return 0;
  }
  void release() {
    if (p) cudaFree(p);
    p = nullptr;
    n = 0;
  }
};

struct Pinned {
  void* p = nullptr;
  size_t n = 0;
  int ensure(size_t bytes) {
    if (bytes <= n && p) return 0;
    if (p) cudaFreeHost(p);
    p = nullptr;
    n = 0;
    bytes = std::max<size_t>(bytes, 1 << 16);
    cudaError_t e = cudaMallocHost(&p, bytes);
    if (e != cudaSuccess) return fail(std::string("cudaMallocHost: ") + cudaGetErrorString(e));
    n = bytes;
    return 0;
  }
  void release() {
    if (p) cudaFreeHost(p);
    p = nullptr;
  }
};

}  // namespace

struct qmcb_ctx {
  int device = 0;
  cudaStream_t stream = nullptr;
  // ---- host description
  std::vector<double> xyz, chg;
  std::vector<int> sh_atom, sh_l, prim_off;
  std::vector<double> pexp, pcoef;
  bool have_slater = false, have_jastrow = false;
  int nup = 0, ndn = 0;
  int nmo[2] = {0, 0}, nds[2] = {0, 0}, ndet = 0;
  std::vector<double> mo[2], detc;
  std::vector<int> occ[2], dmap[2];
  // complex wave functions (qmcb_set_slater_cx / qmcb_set_pbc_phases_imag): imaginary parts of the MO and determinant
  // coefficients and of the Bloch phase table; nmo[] stays the true orbital count
  bool cplx = false;
  std::vector<double> mo_im[2], detc_im, phases_im;
  DBuf<double> b_gemm;  // split partial tiles of the SR product
  DBuf<double> b_inv_im[2], b_dphs_im[2], b_dv_im[2], b_W_im[2], d_detc_im, d_grp_coef_im[2], b_cxwork, e_contrib_im;
  int na = 0, nb = 0;
  std::vector<int> akind, bkind;
  std::vector<double> apar, bpar, acoef, bcoef;
  double rcut_a = 1.0, rcut_b = 1.0;
  bool have_j3 = false;
  int na3 = 0, nb3 = 0;
  std::vector<int> a3kind, b3kind;
  std::vector<double> a3par, b3par, c3;
  double rcut_a3 = 1.0, rcut_b3 = 1.0;
  int necp = 0;
  std::vector<int> ecp_atom, chan_off, term_off, term_pow, naip;
  std::vector<double> term_alpha, term_coef, quad;
  double threshold = 10.0;
  // periodic systems
  int pbc_mode = 0;
  std::vector<double> lat, shifts;
  bool have_pbc_orb = false;
  int nk = 0, nL = 0, isgamma = 1;
  std::vector<double> bxyz, lprim, smat, kpts, Ls, atomcut, lcut, phases;
  std::vector<int> numLs, mok[2];
  bool have_ewald = false;
  double ew_alpha = 0, ew_ij = 0, ew_sq = 0, ew_isum = 0, ew_eii = 0;
  int ew_ndisp = 0, ew_nG = 0;
  DBuf<double> d_ewdisp, d_ewg, d_ewion;
  DBuf<double> b_wrap, b_swrap, b_monew, b_gold, b_jold, d_pwrap, e_ewald, e_ecppos, e_ecpwrap;
  bool pending_wrap = false;  // d_pwrap holds the wrap vectors of the next point call
  // DMC block scratch (kept across blocks: a cudaMalloc / cudaFree pair costs more than a DMC step)
  DBuf<double> m_tmu, m_tmrot, m_tmsel, m_tmacc, m_w, m_eold, m_v2old, m_r2p, m_r2a, m_prod, m_ws;
  DBuf<unsigned long long> m_ntacc;
  bool dirty = true;
  std::vector<unsigned char> key_slater, key_jastrow, key_j3;  // last parameter sets (ParamKey)
  // ---- device tables
  Sys S{};
  DBuf<double> d_dblob, d_detc, d_quad;
  DBuf<int> d_iblob, d_map[2], d_grp_off[2], d_grp_det[2], d_grp_other[2];
  DBuf<double> d_grp_coef[2], d_dense[2];
  size_t smem_bytes = 0;
  int nmot = 0;  // 4 / 8: register fast path, 0: general path
  // ---- walker state
  State st{};
  int N = 0;
  DBuf<double> b_inv[2], b_dsign[2], b_dlog[2], b_dv[2], b_W[2], b_ref[2];
  DBuf<double> b_conf, b_ap, b_bp, b_av, b_bv, b_smo, b_spos, b_moall, b_lu, b_mocache, b_a3v, b_P3, b_val3, b_bpair, b_gpair, b_agrad, b_lpair, b_alap;
  // ---- staging / scratch
  DBuf<double> d_in, d_out, d_scr, d_u, d_rot, d_gauss, d_unif, d_energy, d_esum;
  DBuf<uint8_t> d_mask, d_accept;
  DBuf<int> d_idx;
  DBuf<unsigned long long> d_nacc;
  Pinned h_in, h_out;
  // energy scratch
  DBuf<double> e_ke, e_g2, e_loc, e_vls, e_contrib;
  DBuf<int> e_item, e_work, e_count;
  EnergyScratch es{};
  // saved slot
  int64_t slot_counter = 0, saved_slot = -1;
  int saved_e = -1, saved_which = 0;
  int64_t nlaunch = 0;
  std::vector<int> shape_sig;
  bool mocache_valid = false;
  bool paircache_valid = false;  // Jastrow pair caches of the sweep kernel match the walkers
  bool kinetic_valid = false;    // es.ke_e / es.g2_e were written by the sweep kernel for the current walkers
  // double-buffered device copies of a block's variates, filled on a copy stream by the host
  // thread that draws them (qmcb_vmc_upload) while the previous block computes
  static constexpr int NSLOT = 3;
  DBuf<double> s_gauss[NSLOT], s_unif[NSLOT], s_u[NSLOT], s_rot[NSLOT];
  cudaStream_t copy_stream = nullptr;
  cudaEvent_t slot_ready[NSLOT] = {nullptr, nullptr, nullptr};
  cudaEvent_t done_event = nullptr;  // blocking-sync event: the host thread sleeps while a block runs
  cudaEvent_t block_done[NSLOT] = {nullptr, nullptr, nullptr};  // qmcb_vmc_block_slot_begin / _end
  bool block_pending[NSLOT] = {false, false, false};
  void* devrng = nullptr;            // DevRng (devrng_api.cuh): device-resident legacy generator
  // overlapped step loop of qmcb_vmc_block_device: the energy accumulator of step s runs on its own stream from a
  // snapshot of the arrays it reads while the sweep of step s+1 already moves the walkers
  cudaStream_t energy_stream = nullptr;
  cudaEvent_t ev_snap = nullptr, ev_edone = nullptr;
  // periodic fused move chain: walker ranges 1..3 run on their own streams next to range 0 on the block's stream
  cudaStream_t split_stream[3] = {nullptr, nullptr, nullptr};
  cudaEvent_t ev_fork = nullptr, ev_join[3] = {nullptr, nullptr, nullptr};
  DBuf<double> sn_inv[2], sn_conf, sn_ap, sn_bp, sn_wrap, e_ke2, e_g22;
  // grid cap of the persistent periodic orbital kernel while it runs as overlapped background work: its CTAs stride
  // over the points and hold their SM's shared memory until the launch ends, so a full-machine grid would stall the
  // short kernels of the walker moves it is supposed to hide behind (0 = no cap)
  unsigned pbc_mo_grid_cap = 0;
  // variates of a DMC block generated on the device (qmcb_devrng_dmc_block), two sets: the generator fills one
  // while qmcb_dmc_block_slot consumes the other
  struct DmcSlot {
    DBuf<double> gauss, unif, u, rot, tmu, tmrot, tmsel, tmacc, branch;
    cudaEvent_t ready = nullptr;
    bool filled = false;
  } dmc_slot[2];
  int dmc_use_slot = -1;
  // per-slot result staging of qmcb_vmc_block_slot_begin: the device->host copies of block b run on their own
  // stream (copy engine) while block b+1 computes
  cudaStream_t d2h_stream = nullptr;
  cudaEvent_t ev_block = nullptr;
  DBuf<double> r_energy[NSLOT], r_conf[NSLOT], r_esum[NSLOT];
  DBuf<unsigned long long> r_nacc[NSLOT];
};

extern "C" {
static void devrng_free(qmcb_ctx* c);
}

namespace {

int round_up(int x, int m) { return (x + m - 1) / m * m; }

struct Guard {
  explicit Guard(qmcb_ctx* c) { cudaSetDevice(c->device); }
};

template <class K>
int prep_kernel(K kernel, size_t smem) {
  if (smem > 48 * 1024) CK(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  return 0;
}

// block size: spread small problems over many SMs (one warp per CTA), else 128 threads
int pick_block(long long nthreads) { return nthreads <= 148LL * 64 ? 32 : (nthreads <= 148LL * 256 ? 64 : 128); }

int build_tables(qmcb_ctx* c) {
  if (!c->dirty) return 0;
  Sys& S = c->S;
  std::memset(&S, 0, sizeof(S));
  const int natom = (int)c->chg.size();
  if (natom == 0) return fail("qmcb_set_atoms has not been called");
  const int nshell = (int)c->sh_l.size();
  S.natom = natom;
  S.nshell = nshell;
  S.nprim = (int)c->pexp.size();
  std::vector<int> atsh(natom + 1, 0), shao(nshell + 1, 0);
  const int nbatom_host = (c->pbc_mode && c->have_pbc_orb) ? (int)c->bxyz.size() / 3 : natom;
  atsh.assign(nbatom_host + 1, 0);
  for (int s = 0; s < nshell; ++s) {
    if (c->sh_atom[s] >= nbatom_host) return fail("shell table refers to an atom outside the basis-atom list");
    if (c->sh_l[s] > QMCB_MAX_ATOM_L || c->sh_l[s] < 0) return fail("angular momentum l > 5 is not supported (the reference dispatches l <= 5, gto.py:107-118)");
    if (s > 0 && c->sh_atom[s] < c->sh_atom[s - 1]) return fail("shells must be grouped by atom");
    atsh[c->sh_atom[s] + 1]++;
    shao[s + 1] = shao[s] + 2 * c->sh_l[s] + 1;
  }
  for (int a = 0; a < nbatom_host; ++a) atsh[a + 1] += atsh[a];
  S.nao = shao[nshell];
  S.nup = c->nup;
  S.ndn = c->ndn;
  S.ne = c->nup + c->ndn;
  S.npair = S.ne * (S.ne - 1) / 2;
  S.ndet = c->have_slater ? c->ndet : 0;
  bool ident = c->have_slater && c->ndet == 1;
  const bool cx = c->have_slater && c->cplx;
  S.cplx = cx ? 1 : 0;
  for (int s = 0; s < 2; ++s) {
    S.nmo_t[s] = c->have_slater ? c->nmo[s] : 0;
    S.cxoff[s] = cx ? S.nmo_t[s] : 0;
    S.nmo[s] = (cx ? 2 : 1) * S.nmo_t[s];
    S.nds[s] = c->have_slater ? c->nds[s] : 0;
    const int n = s ? c->ndn : c->nup;
    if (c->have_slater) {
      if (c->nds[s] != 1) ident = false;
      for (int k = 0; k < n && ident; ++k)
        if (c->occ[s][k] != k) ident = false;
    }
  }
  const int nmax = std::max(S.nmo[0], S.nmo[1]);
  if (c->pbc_mode) ident = false;  // periodic orbitals always go through k_pbc_mo + the general path
  if (cx) ident = false;           // complex wave functions: general path + the kernels of cplx.cuh
  if (ident && nmax <= 4)
    c->nmot = 4;
  else if (ident && nmax <= 8)
    c->nmot = 8;
  else
    c->nmot = 0;
  S.fast = c->nmot;
  for (int s = 0; s < 2; ++s) S.ldc[s] = c->nmot ? c->nmot : round_up(std::max(S.nmo[s], 1), 8);
  S.na = c->have_jastrow ? c->na : 0;
  S.nb = c->have_jastrow ? c->nb : 0;
  S.rcut_a = c->rcut_a;
  S.rcut_b = c->rcut_b;
  S.na3 = c->have_j3 ? c->na3 : 0;
  S.nb3 = c->have_j3 ? c->nb3 : 0;
  S.rcut_a3 = c->rcut_a3;
  S.rcut_b3 = c->rcut_b3;
  S.necp = c->necp;
  S.ecp_threshold = c->threshold;
  double eii = 0.0;
  for (int i = 0; i < natom; ++i)
    for (int j = i + 1; j < natom; ++j) {
      const double dx = c->xyz[3 * i] - c->xyz[3 * j], dy = c->xyz[3 * i + 1] - c->xyz[3 * j + 1],
                   dz = c->xyz[3 * i + 2] - c->xyz[3 * j + 2];
      eii += c->chg[i] * c->chg[j] / std::sqrt(dx * dx + dy * dy + dz * dz);
    }
  S.e_ii = eii;

  std::vector<double> db;
  std::vector<int> ib;
  auto dpush = [&](const double* p, size_t n) {
    int o = (int)db.size();
    db.insert(db.end(), p, p + n);
    return o;
  };
  auto ipush = [&](const int* p, size_t n) {
    int o = (int)ib.size();
    ib.insert(ib.end(), p, p + n);
    return o;
  };
  S.o_xyz = dpush(c->xyz.data(), c->xyz.size());
  S.o_chg = dpush(c->chg.data(), c->chg.size());
  S.pbc = c->pbc_mode;
  S.nbatom = natom;
  S.o_bxyz = S.o_xyz;
  S.isgamma = 1;
  if (c->pbc_mode) {
    auto inv3 = [](const std::vector<double>& m) {
      std::vector<double> r(9);
      const double det = m[0] * (m[4] * m[8] - m[5] * m[7]) - m[1] * (m[3] * m[8] - m[5] * m[6]) + m[2] * (m[3] * m[7] - m[4] * m[6]);
      r[0] = (m[4] * m[8] - m[5] * m[7]) / det;
      r[1] = (m[2] * m[7] - m[1] * m[8]) / det;
      r[2] = (m[1] * m[5] - m[2] * m[4]) / det;
      r[3] = (m[5] * m[6] - m[3] * m[8]) / det;
      r[4] = (m[0] * m[8] - m[2] * m[6]) / det;
      r[5] = (m[2] * m[3] - m[0] * m[5]) / det;
      r[6] = (m[3] * m[7] - m[4] * m[6]) / det;
      r[7] = (m[1] * m[6] - m[0] * m[7]) / det;
      r[8] = (m[0] * m[4] - m[1] * m[3]) / det;
      return r;
    };
    S.o_lat = dpush(c->lat.data(), 9);
    std::vector<double> li = inv3(c->lat);
    S.o_latinv = dpush(li.data(), 9);
    {
      std::vector<double> sh(c->shifts);  // 27 shifts followed by their squared norms
      for (int i = 0; i < 27; ++i) sh.push_back(c->shifts[3 * i] * c->shifts[3 * i] + c->shifts[3 * i + 1] * c->shifts[3 * i + 1] + c->shifts[3 * i + 2] * c->shifts[3 * i + 2]);
      S.o_shifts = dpush(sh.data(), sh.size());
    }
    if (c->have_slater && !c->have_pbc_orb) return fail("periodic Slater factor: qmcb_set_pbc_orbitals has not been called");
    if (c->have_pbc_orb) {
      S.nbatom = (int)c->bxyz.size() / 3;
      S.o_bxyz = dpush(c->bxyz.data(), c->bxyz.size());
      S.nk = c->nk;
      S.nkp = cx ? 2 * c->nk : c->nk;
      S.nL = c->nL;
      S.isgamma = c->isgamma;
      S.o_lprim = dpush(c->lprim.data(), 9);
      std::vector<double> lpi = inv3(c->lprim);
      S.o_lpriminv = dpush(lpi.data(), 9);
      S.o_smat = dpush(c->smat.data(), 9);
      std::vector<double> kl((size_t)c->nk * 3);
      for (int k = 0; k < c->nk; ++k)
        for (int i = 0; i < 3; ++i)
          kl[3 * k + i] = c->kpts[3 * k] * c->lprim[3 * i] + c->kpts[3 * k + 1] * c->lprim[3 * i + 1] + c->kpts[3 * k + 2] * c->lprim[3 * i + 2];
      S.o_kl = dpush(kl.data(), kl.size());
      S.o_Ls = dpush(c->Ls.data(), c->Ls.size());
      S.o_atomcut = dpush(c->atomcut.data(), c->atomcut.size());
      if ((int)c->lcut.size() != nshell) return fail("qmcb_set_pbc_orbitals: l_cutoff must have one entry per shell");
      S.o_lcut = dpush(c->lcut.data(), c->lcut.size());
      if (cx) {  // [nL][2 nk]: cos | sin planes of exp(i L.k)  (pbcgto.py:620)
        std::vector<double> ph((size_t)c->nL * 2 * c->nk, 0.0);
        const bool have_im = c->phases_im.size() == c->phases.size();
        for (int L = 0; L < c->nL; ++L)
          for (int k = 0; k < c->nk; ++k) {
            ph[((size_t)L * 2) * c->nk + k] = c->phases[(size_t)L * c->nk + k];
            ph[((size_t)L * 2 + 1) * c->nk + k] = have_im ? c->phases_im[(size_t)L * c->nk + k] : 0.0;
          }
        S.o_phase = dpush(ph.data(), ph.size());
      } else
        S.o_phase = dpush(c->phases.data(), c->phases.size());
    }
  }
  if (db.size() % 2) db.push_back(0.0);
  {
    std::vector<double> prim(2 * c->pexp.size());
    for (size_t p = 0; p < c->pexp.size(); ++p) {
      prim[2 * p] = c->pexp[p];
      prim[2 * p + 1] = c->pcoef[p];
    }
    S.o_prim = dpush(prim.data(), prim.size());
  }
  for (int s = 0; s < 2; ++s) {
    std::vector<double> cm((size_t)std::max(S.nao, 1) * S.ldc[s], 0.0);
    if (c->have_slater)
      for (int a = 0; a < S.nao; ++a)
        for (int j = 0; j < S.nmo_t[s]; ++j) {
          cm[(size_t)a * S.ldc[s] + j] = c->mo[s][(size_t)a * S.nmo_t[s] + j];
          if (cx) cm[(size_t)a * S.ldc[s] + S.cxoff[s] + j] = c->mo_im[s][(size_t)a * S.nmo_t[s] + j];
        }
    S.o_mo[s] = dpush(cm.data(), cm.size());
  }
  S.o_apar = dpush(c->apar.data(), S.na);
  S.o_bpar = dpush(c->bpar.data(), S.nb);
  S.o_acoef = dpush(c->acoef.data(), (size_t)natom * S.na * 2);
  S.o_bcoef = dpush(c->bcoef.data(), (size_t)S.nb * 3);
  S.o_a3par = dpush(c->a3par.data(), S.na3);
  S.o_b3par = dpush(c->b3par.data(), S.nb3);
  S.o_c3 = dpush(c->c3.data(), c->have_j3 ? c->c3.size() : 0);
  S.o_talpha = dpush(c->term_alpha.data(), c->term_alpha.size());
  S.o_tcoef = dpush(c->term_coef.data(), c->term_coef.size());
  while (db.size() % 2) db.push_back(0.0);
  if (db.empty()) db.resize(2, 0.0);

  S.o_atsh = ipush(atsh.data(), atsh.size());
  S.o_shl = ipush(c->sh_l.data(), nshell);
  S.o_shprim = ipush(c->prim_off.data(), c->prim_off.size());
  S.o_shao = ipush(shao.data(), shao.size());
  for (int s = 0; s < 2; ++s) S.o_occ[s] = ipush(c->occ[s].data(), c->have_slater ? c->occ[s].size() : 0);
  {
    const int nba = nbatom_host;
    std::vector<int> primatom(c->pexp.size()), aoshell(S.nao), sphoff(nba + 1, 0), lmax(nba, -1), task;
    for (int sh = 0; sh < nshell; ++sh) {
      for (int p = c->prim_off[sh]; p < c->prim_off[sh + 1]; ++p) primatom[p] = c->sh_atom[sh];
      for (int m = shao[sh]; m < shao[sh + 1]; ++m) aoshell[m] = sh;
      lmax[c->sh_atom[sh]] = std::max(lmax[c->sh_atom[sh]], c->sh_l[sh]);
    }
    for (int a = 0; a < nba; ++a) {
      sphoff[a + 1] = sphoff[a] + (lmax[a] + 1) * (lmax[a] + 1);
      for (int l = 0; l <= lmax[a]; ++l) {
        task.push_back(a);
        task.push_back(l);
      }
    }
    S.nsph = sphoff[nba];
    S.nsphtask = (int)task.size() / 2;
    S.o_primatom = ipush(primatom.data(), primatom.size());
    S.o_shatom = ipush(c->sh_atom.data(), nshell);
    S.o_aoshell = ipush(aoshell.data(), aoshell.size());
    S.o_sphoff = ipush(sphoff.data(), sphoff.size());
    S.o_sphtask = ipush(task.data(), task.size());
  }
  if (c->pbc_mode && c->have_pbc_orb) {
    S.o_numLs = ipush(c->numLs.data(), c->numLs.size());
    std::vector<int> candoff(S.nbatom + 1, 0);
    int mx = 0;
    for (int a = 0; a < S.nbatom; ++a) {
      candoff[a + 1] = candoff[a] + c->numLs[a];
      mx = std::max(mx, shao[atsh[a + 1]] - shao[atsh[a]]);
    }
    S.ncand = candoff[S.nbatom];
    S.maxao_atom = mx;
    S.o_candoff = ipush(candoff.data(), candoff.size());
    for (int s = 0; s < 2; ++s) {
      std::vector<int> mk(S.ldc[s], 0);
      if (c->have_slater) {
        if ((int)c->mok[s].size() != S.nmo_t[s]) return fail("qmcb_set_pbc_orbitals: MO -> k-point map does not match the MO count");
        for (int j = 0; j < S.nmo_t[s]; ++j) mk[j] = c->mok[s][j];
      }
      S.o_mok[s] = ipush(mk.data(), mk.size());
    }
  }
  S.o_akind = ipush(c->akind.data(), S.na);
  S.o_bkind = ipush(c->bkind.data(), S.nb);
  S.o_a3kind = ipush(c->a3kind.data(), S.na3);
  S.o_b3kind = ipush(c->b3kind.data(), S.nb3);
  S.o_ecpatom = ipush(c->ecp_atom.data(), c->ecp_atom.size());
  S.o_chanoff = ipush(c->chan_off.data(), c->chan_off.size());
  S.o_termoff = ipush(c->term_off.data(), c->term_off.size());
  S.o_tpow = ipush(c->term_pow.data(), c->term_pow.size());
  S.o_naip = ipush(c->naip.data(), c->naip.size());
  {
    std::vector<int> aipoff(c->necp + 1, 0);
    int mx = 0;
    for (int a = 0; a < c->necp; ++a) {
      aipoff[a + 1] = aipoff[a] + c->naip[a];
      mx = std::max(mx, c->naip[a]);
    }
    S.o_aipoff = ipush(aipoff.data(), aipoff.size());
    S.max_naip = mx;
    S.tot_naip = aipoff[c->necp];
  }
  S.nchan = c->chan_off.empty() ? 0 : c->chan_off.back();
  S.nterm = (int)c->term_pow.size();
  while (ib.size() % 4) ib.push_back(0);
  if (ib.empty()) ib.resize(4, 0);
  S.dwords = (int)db.size();
  S.iwords = (int)ib.size();
  c->smem_bytes = 16 + db.size() * 8 + ib.size() * 4;
  if (c->smem_bytes > 200 * 1024) return fail("system tables exceed the shared-memory staging budget (200 KB)");
  if (c->d_dblob.ensure(db.size())) return -1;
  if (c->d_iblob.ensure(ib.size())) return -1;
  CK(cudaMemcpy(c->d_dblob.p, db.data(), db.size() * 8, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(c->d_iblob.p, ib.data(), ib.size() * 4, cudaMemcpyHostToDevice));
  S.dblob = c->d_dblob.p;
  S.iblob = c->d_iblob.p;
  // per-determinant tables (global memory)
  if (c->have_slater) {
    if (c->d_detc.ensure(c->ndet)) return -1;
    CK(cudaMemcpy(c->d_detc.p, c->detc.data(), (size_t)c->ndet * 8, cudaMemcpyHostToDevice));
    S.detc = c->d_detc.p;
    for (int s = 0; s < 2; ++s) {
      if (c->d_map[s].ensure(c->ndet)) return -1;
      CK(cudaMemcpy(c->d_map[s].p, c->dmap[s].data(), (size_t)c->ndet * 4, cudaMemcpyHostToDevice));
      S.map[s] = c->d_map[s].p;
      std::vector<int> off(c->nds[s] + 1, 0), lst(c->ndet);
      for (int D = 0; D < c->ndet; ++D) off[c->dmap[s][D] + 1]++;
      for (int d = 0; d < c->nds[s]; ++d) off[d + 1] += off[d];
      std::vector<int> cur(off.begin(), off.end() - 1);
      for (int D = 0; D < c->ndet; ++D) lst[cur[c->dmap[s][D]]++] = D;
      if (c->d_grp_off[s].ensure(off.size())) return -1;
      if (c->d_grp_det[s].ensure(lst.size())) return -1;
      CK(cudaMemcpy(c->d_grp_off[s].p, off.data(), off.size() * 4, cudaMemcpyHostToDevice));
      CK(cudaMemcpy(c->d_grp_det[s].p, lst.data(), lst.size() * 4, cudaMemcpyHostToDevice));
      S.grp_off[s] = c->d_grp_off[s].p;
      S.grp_det[s] = c->d_grp_det[s].p;
      std::vector<double> gcoef(lst.size());
      std::vector<int> gother(lst.size());
      for (size_t k = 0; k < lst.size(); ++k) {
        gcoef[k] = c->detc[lst[k]];
        gother[k] = c->dmap[1 - s][lst[k]];
      }
      if (c->d_grp_coef[s].ensure(lst.size()) || c->d_grp_other[s].ensure(lst.size())) return -1;
      CK(cudaMemcpy(c->d_grp_coef[s].p, gcoef.data(), gcoef.size() * 8, cudaMemcpyHostToDevice));
      CK(cudaMemcpy(c->d_grp_other[s].p, gother.data(), gother.size() * 4, cudaMemcpyHostToDevice));
      S.grp_coef[s] = c->d_grp_coef[s].p;
      S.grp_other[s] = c->d_grp_other[s].p;
      S.dense[s] = nullptr;
      if (!cx && c->ndet > 1 && (long long)c->nds[0] * c->nds[1] <= 65536 && std::getenv("QMCB_NO_DENSE_DET") == nullptr) {
        const int nd = c->nds[s], no = c->nds[1 - s];
        std::vector<double> dm((size_t)nd * no, 0.0);
        for (int D = 0; D < c->ndet; ++D) dm[(size_t)c->dmap[1 - s][D] * nd + c->dmap[s][D]] += c->detc[D];
        if (c->d_dense[s].ensure(dm.size())) return -1;
        CK(cudaMemcpy(c->d_dense[s].p, dm.data(), dm.size() * 8, cudaMemcpyHostToDevice));
        S.dense[s] = c->d_dense[s].p;
      }
      if (cx) {
        std::vector<double> gim(lst.size());
        for (size_t k = 0; k < lst.size(); ++k) gim[k] = c->detc_im[lst[k]];
        if (c->d_grp_coef_im[s].ensure(lst.size())) return -1;
        CK(cudaMemcpy(c->d_grp_coef_im[s].p, gim.data(), gim.size() * 8, cudaMemcpyHostToDevice));
        S.grp_coef_im[s] = c->d_grp_coef_im[s].p;
      }
    }
    if (cx) {
      if (c->d_detc_im.ensure(c->ndet)) return -1;
      CK(cudaMemcpy(c->d_detc_im.p, c->detc_im.data(), (size_t)c->ndet * 8, cudaMemcpyHostToDevice));
      S.detc_im = c->d_detc_im.p;
    }
  }
  if (c->necp > 0) {
    if (c->d_quad.ensure(c->quad.size())) return -1;
    CK(cudaMemcpy(c->d_quad.p, c->quad.data(), c->quad.size() * 8, cudaMemcpyHostToDevice));
  }
  if (c->pbc_mode && c->have_ewald) {
    S.ew_ndisp = c->ew_ndisp;
    S.ew_nG = c->ew_nG;
    S.ew_alpha = c->ew_alpha;
    S.ew_ijconst = c->ew_ij;
    S.ew_sqconst = c->ew_sq;
    S.ew_isum = c->ew_isum;
    S.ew_disp = c->d_ewdisp.p;
    S.ew_g = c->d_ewg.p;
    S.ew_ion = c->d_ewion.p;
    S.e_ii = c->ew_eii;
  }
  c->dirty = false;
  // walker state survives a table rebuild unless a shape it depends on changed
  std::vector<int> sig = {S.natom, S.nup, S.ndn, S.nds[0], S.nds[1], S.ldc[0], S.ldc[1], S.na, S.nb, S.ndet, S.na3, S.nb3, S.pbc, S.cplx};
  if (sig != c->shape_sig) {
    c->shape_sig = sig;
    c->N = 0;
  }
  return 0;
}

int ensure_state(qmcb_ctx* c, int N) {
  if (build_tables(c)) return -1;
  if (N == c->N) return 0;
  const Sys& S = c->S;
  State& st = c->st;
  st.N = N;
  for (int s = 0; s < 2; ++s) {
    const int n = s ? S.ndn : S.nup;
    const size_t nd = (size_t)N * std::max(S.nds[s], 1);
    if (c->b_inv[s].ensure(nd * std::max(n * n, 1))) return -1;
    if (c->b_dsign[s].ensure(nd) || c->b_dlog[s].ensure(nd) || c->b_dv[s].ensure(nd) || c->b_W[s].ensure(nd) ||
        c->b_ref[s].ensure(N))
      return -1;
    st.inv[s] = c->b_inv[s].p;
    st.dsign[s] = c->b_dsign[s].p;
    st.dlog[s] = c->b_dlog[s].p;
    st.dv[s] = c->b_dv[s].p;
    st.W[s] = c->b_W[s].p;
    st.ref[s] = c->b_ref[s].p;
    if (S.cplx) {
      if (c->b_inv_im[s].ensure(nd * std::max(n * n, 1)) || c->b_dphs_im[s].ensure(nd) || c->b_dv_im[s].ensure(nd) ||
          c->b_W_im[s].ensure(nd))
        return -1;
    }
    st.inv_im[s] = c->b_inv_im[s].p;
    st.dphs_im[s] = c->b_dphs_im[s].p;
    st.dv_im[s] = c->b_dv_im[s].p;
    st.W_im[s] = c->b_W_im[s].p;
  }
  const int ldmax = std::max(S.ldc[0], S.ldc[1]);
  if (c->b_conf.ensure((size_t)N * S.ne * 3) || c->b_ap.ensure((size_t)N * S.ne * S.natom * std::max(S.na, 1)) ||
      c->b_bp.ensure((size_t)N * S.ne * std::max(S.nb, 1) * 2) || c->b_av.ensure((size_t)N * S.natom * std::max(S.na, 1) * 2) ||
      c->b_bv.ensure((size_t)N * std::max(S.nb, 1) * 3) || c->b_smo.ensure((size_t)N * ldmax) ||
      c->b_spos.ensure((size_t)N * 3) || c->b_moall.ensure((size_t)N * S.ne * ldmax) ||
      c->b_mocache.ensure((size_t)N * S.ne * 5 * ldmax) ||
      c->b_a3v.ensure((size_t)N * S.ne * S.natom * std::max(S.na3, 1)) || c->b_P3.ensure((size_t)N * std::max(S.ne, 1)) ||
      c->b_val3.ensure(N) || c->b_bpair.ensure((size_t)N * std::max(S.npair, 1) * std::max(S.nb, 1)) ||
      c->b_gpair.ensure((size_t)N * std::max(S.npair, 1) * 3) || c->b_agrad.ensure((size_t)N * std::max(S.ne, 1) * 3) ||
      c->b_lpair.ensure((size_t)N * std::max(S.npair, 1)) || c->b_alap.ensure((size_t)N * std::max(S.ne, 1)))
    return -1;
  st.conf = c->b_conf.p;
  st.a_partial = c->b_ap.p;
  st.b_partial = c->b_bp.p;
  st.avalues = c->b_av.p;
  st.bvalues = c->b_bv.p;
  st.saved_mo = c->b_smo.p;
  st.saved_pos = c->b_spos.p;
  st.mo_all = c->b_moall.p;
  st.mocache = c->b_mocache.p;
  st.a3v = c->b_a3v.p;
  st.P3 = c->b_P3.p;
  st.val3 = c->b_val3.p;
  st.bpair = c->b_bpair.p;
  st.gpair = c->b_gpair.p;
  st.agrad = c->b_agrad.p;
  st.lpair = c->b_lpair.p;
  st.alap = c->b_alap.p;
  if (S.pbc) {
    if (c->b_wrap.ensure((size_t)N * S.ne * 3) || c->b_swrap.ensure((size_t)N * 3) || c->b_monew.ensure((size_t)N * 5 * ldmax) ||
        c->b_gold.ensure((size_t)N * 3) || c->b_jold.ensure((size_t)N * std::max(1, (S.ne - 1) * S.nb)))
      return -1;
    cudaMemset(c->b_wrap.p, 0, (size_t)N * S.ne * 3 * 8);
    cudaMemset(c->b_swrap.p, 0, (size_t)N * 3 * 8);
  }
  st.wrap = c->b_wrap.p;
  st.saved_wrap = c->b_swrap.p;
  st.monew = c->b_monew.p;
  st.gold = c->b_gold.p;
  st.jold = nullptr;  // set by the fused periodic block driver only
  c->N = N;
  c->saved_slot = -1;
  return 0;
}

int ensure_scratch(qmcb_ctx* c, size_t npoints, int ncomp) {
  if (c->nmot) return c->d_scr.ensure(1);
  const int ldmax = std::max(c->S.ldc[0], c->S.ldc[1]);
  return c->d_scr.ensure(npoints * (size_t)ldmax * ncomp);
}

template <int MODE>
int launch_point(qmcb_ctx* c, const PointArgs& pa, cudaStream_t stream) {
  const int block = pick_block(pa.npoints);
  const int grid = (pa.npoints + block - 1) / block;
  if (grid == 0) return 0;
  const size_t sm = c->smem_bytes;
  if (c->S.cplx && ((pa.which & 1) || MODE == PV_MOSAVE)) {
    if (prep_kernel(k_cx_point<MODE>, sm)) return -1;
    k_cx_point<MODE><<<grid, block, sm, stream>>>(c->S, c->st, pa);
  } else if (c->nmot == 4) {
    if (prep_kernel(k_point<MODE, 4>, sm)) return -1;
    k_point<MODE, 4><<<grid, block, sm, stream>>>(c->S, c->st, pa);
  } else if (c->nmot == 8) {
    if (prep_kernel(k_point<MODE, 8>, sm)) return -1;
    k_point<MODE, 8><<<grid, block, sm, stream>>>(c->S, c->st, pa);
  } else {
    if (prep_kernel(k_point<MODE, 0>, sm)) return -1;
    k_point<MODE, 0><<<grid, block, sm, stream>>>(c->S, c->st, pa);
  }
  c->nlaunch++;
  CK(cudaGetLastError());
  return 0;
}

int launch_sm(qmcb_ctx* c, const SmArgs& a, cudaStream_t stream, int64_t* nlaunch) {
  if (a.nmat == 0 || a.n == 0) return 0;
  if (a.e < 0 || a.e >= a.n) return fail("sherman-morrison: electron index out of range");
#define SM_T(NN)                                                                     \
  case NN: {                                                                         \
    const int block = 128;                                                           \
    const long long grid = (a.nmat + block - 1) / block;                             \
    k_sm_thread<NN><<<(unsigned)grid, block, 0, stream>>>(a);                        \
  } break;
  if (a.n <= 8) {
    switch (a.n) {
      SM_T(1) SM_T(2) SM_T(3) SM_T(4) SM_T(5) SM_T(6) SM_T(7) SM_T(8)
    }
  } else if (a.n == 32 && a.nmat >= 148LL * 12 && (reinterpret_cast<uintptr_t>(a.inv) & 15) == 0 &&
             std::getenv("QMCB_SM_NO_TMA") == nullptr) {
    // large batches of 32 x 32 matrices: bulk-copy (TMA) staged kernel, three CTAs of four warps per SM
    const size_t smem = (size_t)QMCB_SM_TMA_WARPS * 2 * 32 * 32 * 8 + QMCB_SM_TMA_WARPS * 2 * 8;
    if (prep_kernel(k_sm_tma32, smem)) return -1;
    const long long want = (a.nmat + QMCB_SM_TMA_WARPS - 1) / QMCB_SM_TMA_WARPS;
    k_sm_tma32<<<(unsigned)std::min<long long>(want, 148LL * 3), QMCB_SM_TMA_WARPS * 32, smem, stream>>>(a);
  } else if (a.n <= 32) {
    const int block = 256;
    const long long grid = (a.nmat * 32 + block - 1) / block;
    if (a.n <= 16)
      k_sm_warp<16><<<(unsigned)grid, block, 0, stream>>>(a);
    else
      k_sm_warp<32><<<(unsigned)grid, block, 0, stream>>>(a);
  } else {
    return fail("Sherman-Morrison kernels support n <= 32 electrons per spin in this build");
  }
#undef SM_T
  if (nlaunch) (*nlaunch)++;
  (void)c;
  CK(cudaGetLastError());
  return 0;
}

// C[P][P] = A^T B for A, B [N][P] (dp^T (w f dp), stochastic_reconfiguration.py:110-113).  variant 0: the FP64-FMA
// tile kernel; 1: DMMA (mma.sync m8n8k4) with the walker range split over gridDim.z; -1: default (QMCB_GEMM_FMA=1
// selects 0, else 1).
int launch_gemm_tn(const double* A, const double* B, int N, int P, double* C, DBuf<double>& work, int variant,
                   cudaStream_t stream) {
  if (variant < 0) variant = std::getenv("QMCB_GEMM_FMA") ? 0 : 1;
  const unsigned tiles = (unsigned)((P + 63) / 64);
  if (variant == 0) {
    k_gemm_tn<<<dim3(tiles, tiles), 256, 0, stream>>>(A, B, N, P, C);
    CK(cudaGetLastError());
    return 0;
  }
  // enough splits for ~4 CTAs per SM, at least 64 walkers per split
  int nsplit = (int)std::max<long long>(1, std::min<long long>((148LL * 4 + tiles * tiles - 1) / ((long long)tiles * tiles), (N + 63) / 64));
  int rows = ((N + nsplit - 1) / nsplit + 15) / 16 * 16;
  nsplit = (N + rows - 1) / rows;
  if (work.ensure((size_t)nsplit * P * P)) return -1;
  k_gemm_tn_dmma<<<dim3(tiles, tiles, (unsigned)nsplit), 128, 0, stream>>>(A, B, N, P, rows, work.p);
  CK(cudaGetLastError());
  const size_t n = (size_t)P * P;
  k_gemm_reduce<<<(unsigned)((n + 255) / 256), 256, 0, stream>>>(work.p, nsplit, n, C);
  CK(cudaGetLastError());
  return 0;
}

int which_ok(qmcb_ctx* c, int which) {
  if ((which & 1) && !c->have_slater) return fail("context has no Slater factor");
  if ((which & 2) && !c->have_jastrow) return fail("context has no Jastrow factor");
  if ((which & 4) && !c->have_j3) return fail("context has no three-body Jastrow factor");
  if (which == 0) return fail("which == 0");
  return 0;
}

// host <-> device through pinned staging
int h2d(qmcb_ctx* c, void* dst, const void* src, size_t bytes) {
  if (bytes == 0) return 0;
  if (c->h_in.ensure(bytes)) return -1;
  std::memcpy(c->h_in.p, src, bytes);
  CK(cudaMemcpyAsync(dst, c->h_in.p, bytes, cudaMemcpyHostToDevice, c->stream));
  // the staging buffer is reused by the next call: make sure the copy engine is done with it
  CK(cudaStreamSynchronize(c->stream));
  return 0;
}
int d2h(qmcb_ctx* c, void* dst, const void* src, size_t bytes) {
  if (bytes == 0) return 0;
  if (c->h_out.ensure(bytes)) return -1;
  CK(cudaMemcpyAsync(c->h_out.p, src, bytes, cudaMemcpyDeviceToHost, c->stream));
  CK(cudaStreamSynchronize(c->stream));
  std::memcpy(dst, c->h_out.p, bytes);
  return 0;
}

// lattice-summed MO rows at a point list (periodic systems): 16 lanes per point, 4 points per CTA
int launch_pbc_mo(qmcb_ctx* c, int deriv, const PbcMoArgs& a, long long max_points, cudaStream_t stream) {
  if (max_points <= 0) return 0;
  constexpr int G = 16, BLOCK = 64;
  const int nc = deriv == 0 ? 1 : (deriv == 1 ? 4 : 5);
  const size_t tab = (c->smem_bytes + 15) & ~(size_t)15;
  // one CTA per point (k_pbc_mo_cta) whenever the accumulator columns fit its registers
  const size_t csm = tab + pbc_mo_cta_scratch_bytes(c->S, nc);
  const int ncol = c->S.nao * nc;
  if (c->S.nkp <= QMCB_PBC_NKMAX && ncol <= QMCB_PBC_RU * 256 && csm <= 100 * 1024 &&
      std::getenv("QMCB_PBC_NO_CTA") == nullptr) {
    int T = std::max(std::max(ncol, nc * std::max(c->S.ldc[0], c->S.ldc[1])), 64);
    T = std::min((T + 31) / 32 * 32, 256);
    if (ncol > T) T = std::min(((ncol + 1) / 2 + 31) / 32 * 32, 256);
    if (deriv == 0) {
      // value rows (ECP quadrature points): the accumulator columns need few threads (nao), the (pair, shell) tasks of
      // phase A do not -- more threads per point shorten the CTA's dependent chain (results do not depend on T)
      static const int t0 = std::getenv("QMCB_PBC_T0") ? std::atoi(std::getenv("QMCB_PBC_T0")) : 0;
      if (t0 >= 32 && t0 <= 256) T = std::max(T, t0 / 32 * 32);
    }
    unsigned grid = (unsigned)std::min<long long>(max_points, 148LL * 64);
    if (c->pbc_mo_grid_cap) grid = std::min(grid, c->pbc_mo_grid_cap);
    int maxl = 0;
    for (int l : c->sh_l) maxl = std::max(maxl, l);
    const bool tuned4 = T <= 160 && 4 * csm <= 220 * 1024 && std::getenv("QMCB_PBC_MO_OCC2") == nullptr;
    // 96 registers: four CTAs per SM, so the 1024 points of a C4 move are two waves instead of three
    const unsigned g4 = std::min(grid, 148u * 4u);  // one resident wave, CTAs stride over the points
#define QMCB_PBC_CTA(D, MAXT_, MINB_, LM, GRID)                                  \
  do {                                                                           \
    if (prep_kernel(k_pbc_mo_cta<D, MAXT_, MINB_, LM>, csm)) return -1;          \
    k_pbc_mo_cta<D, MAXT_, MINB_, LM><<<GRID, T, csm, stream>>>(c->S, c->st, a); \
  } while (0)
    if (maxl <= 4) {
      if (deriv == 0) QMCB_PBC_CTA(0, 256, 2, 4, grid);
      else if (deriv == 1) QMCB_PBC_CTA(1, 256, 2, 4, grid);
      else if (tuned4) QMCB_PBC_CTA(2, 160, 4, 4, g4);
      else QMCB_PBC_CTA(2, 256, 2, 4, grid);
    } else {
      if (deriv == 0) QMCB_PBC_CTA(0, 256, 2, 5, grid);
      else if (deriv == 1) QMCB_PBC_CTA(1, 256, 2, 5, grid);
      else QMCB_PBC_CTA(2, 256, 2, 5, grid);
    }
    c->nlaunch++;
    CK(cudaGetLastError());
    return 0;
  }
  const size_t sm = tab + (size_t)(BLOCK / G) * pbc_mo_scratch_doubles(c->S, nc, G) * 8;
  if (sm > 200 * 1024) return fail("periodic orbital evaluation: shared-memory scratch exceeds 200 KB (nk * nao too large)");
  const long long grid = std::min<long long>((max_points + (BLOCK / G) - 1) / (BLOCK / G), 148LL * 16);
  if (deriv == 0) {
    if (prep_kernel(k_pbc_mo<0, G>, sm)) return -1;
    k_pbc_mo<0, G><<<(unsigned)grid, BLOCK, sm, stream>>>(c->S, c->st, a);
  } else if (deriv == 1) {
    if (prep_kernel(k_pbc_mo<1, G>, sm)) return -1;
    k_pbc_mo<1, G><<<(unsigned)grid, BLOCK, sm, stream>>>(c->S, c->st, a);
  } else {
    if (prep_kernel(k_pbc_mo<2, G>, sm)) return -1;
    k_pbc_mo<2, G><<<(unsigned)grid, BLOCK, sm, stream>>>(c->S, c->st, a);
  }
  c->nlaunch++;
  CK(cudaGetLastError());
  return 0;
}

// MO value / gradient / Laplacian rows of every electron at its current position -> mocache
// (and the value rows -> mo_all when write_values)
int launch_mo_all(qmcb_ctx* c, int write_values, cudaStream_t stream) {
  const Sys& S = c->S;
  const long long np = (long long)c->N * S.ne;
  if (S.pbc) {
    const int ldmax = std::max(S.ldc[0], S.ldc[1]);
    PbcMoArgs a{};
    a.npoints = np;
    a.pos = c->st.conf;
    a.wrap = c->st.wrap;
    a.naip = 1;
    a.spin_mode = 1;
    a.out = c->st.mocache;
    a.stride_p = 5 * ldmax;
    a.stride_c = ldmax;
    a.stride_j = 1;
    if (write_values) {
      a.out_val = c->st.mo_all;
      a.stride_vp = ldmax;
    }
    if (launch_pbc_mo(c, 2, a, np, stream)) return -1;
  } else {
    const int block = pick_block(np);
    if (prep_kernel(k_mo_all, c->smem_bytes)) return -1;
    k_mo_all<<<(unsigned)((np + block - 1) / block), block, c->smem_bytes, stream>>>(S, c->st, write_values);
    c->nlaunch++;
    CK(cudaGetLastError());
  }
  c->mocache_valid = true;
  return 0;
}

// wait for the context's stream without spinning: the block drivers run next to the host RNG threads
int sync_blocking(qmcb_ctx* c) {
  CK(cudaEventRecord(c->done_event, c->stream));
  CK(cudaEventSynchronize(c->done_event));
  return 0;
}

// AO values of every electron at its current position: [N][ne][A], periodic [N][ne][nk][A] with the
// wrap phase (the reference's _aovals, slater.py:233)
int launch_ao_all(qmcb_ctx* c, double* d_ao, cudaStream_t stream) {
  const Sys& S = c->S;
  const long long np = (long long)c->N * S.ne;
  if (S.pbc) {
    const size_t tab = (c->smem_bytes + 15) & ~(size_t)15;
    const size_t csm = tab + pbc_mo_cta_scratch_bytes(S, 1);
    if (S.nkp > QMCB_PBC_NKMAX || S.nao > QMCB_PBC_RU * 256 || csm > 100 * 1024)
      return fail("periodic parameter gradients: k-point / AO count beyond the CTA orbital kernel's limits");
    PbcMoArgs a{};
    a.npoints = np;
    a.pos = c->st.conf;
    a.wrap = c->st.wrap;
    a.naip = 1;
    a.spin_mode = 1;
    a.ao_out = d_ao;
    a.out = d_ao;  // unused in AO mode
    int T = std::min((std::max(S.nao, 64) + 31) / 32 * 32, 256);
    if (S.nao > T) T = std::min(((S.nao + 1) / 2 + 31) / 32 * 32, 256);
    int maxl = 0;
    for (int l : c->sh_l) maxl = std::max(maxl, l);
    if (maxl <= 4) {
      if (prep_kernel(k_pbc_mo_cta<0>, csm)) return -1;
      k_pbc_mo_cta<0><<<(unsigned)std::min<long long>(np, 148LL * 64), T, csm, stream>>>(S, c->st, a);
    } else {
      if (prep_kernel(k_pbc_mo_cta<0, 256, 2, 5>, csm)) return -1;
      k_pbc_mo_cta<0, 256, 2, 5><<<(unsigned)std::min<long long>(np, 148LL * 64), T, csm, stream>>>(S, c->st, a);
    }
  } else {
    if (prep_kernel(k_ao_all, c->smem_bytes)) return -1;
    k_ao_all<<<(unsigned)((np + 127) / 128), 128, c->smem_bytes, stream>>>(S, c->st, d_ao);
  }
  c->nlaunch++;
  CK(cudaGetLastError());
  return 0;
}

int slater_rebuild(qmcb_ctx* c, cudaStream_t stream) {
  const Sys& S = c->S;
  const int N = c->N;
  if (launch_mo_all(c, 1, stream)) return -1;
  if (S.cplx) {
    for (int s = 0; s < 2; ++s) {
      const int n = s ? S.ndn : S.nup;
      const long long nt = (long long)N * S.nds[s];
      if (n > QMCB_CX_NMAX) return fail("complex wave functions: more than 64 electrons per spin is not supported");
      if (c->b_cxwork.ensure((size_t)nt * std::max(n * n, 1) * 2)) return -1;
      k_cx_invert<<<(unsigned)((nt * 32 + 127) / 128), 128, 0, stream>>>(S, c->st, s, reinterpret_cast<cd*>(c->b_cxwork.p));
      c->nlaunch++;
      CK(cudaGetLastError());
    }
    if (S.ndet > 1) {
      k_cx_det_cache<<<(unsigned)(((long long)N * 32 + 127) / 128), 128, 0, stream>>>(S, c->st, nullptr, -1);
      c->nlaunch++;
      CK(cudaGetLastError());
    }
    return 0;
  }
  for (int s = 0; s < 2; ++s) {
    const int n = s ? S.ndn : S.nup;
    const long long nt = (long long)N * S.nds[s];
    const int block = 64;
    const unsigned grid = (unsigned)((nt + block - 1) / block);
    if (n <= 8)
      k_invert<8><<<grid, block, 0, stream>>>(S, c->st, s, nullptr);
    else if (n <= 32 && std::getenv("QMCB_NO_WARP_INVERT") == nullptr) {
      const size_t ism = (size_t)4 * (32 * 33 + 32) * 8;
      k_invert_warp<<<(unsigned)((nt * 32 + 127) / 128), 128, ism, stream>>>(S, c->st, s);
    } else if (n <= 16)
      k_invert<16><<<grid, block, 0, stream>>>(S, c->st, s, nullptr);
    else if (n <= 64) {
      if (c->b_lu.ensure((size_t)nt * n * n)) return -1;
      k_invert<0><<<grid, block, 0, stream>>>(S, c->st, s, c->b_lu.p);
    } else
      return fail("more than 64 electrons per spin is not supported");
    c->nlaunch++;
    CK(cudaGetLastError());
  }
  if (S.ndet > 1) {
    k_det_cache<<<(unsigned)(((long long)N * 32 + 127) / 128), 128, 0, stream>>>(S, c->st, nullptr, -1);
    c->nlaunch++;
    CK(cudaGetLastError());
  }
  return 0;
}

int launch_value(qmcb_ctx* c, int which, double* d_sign, double* d_log, cudaStream_t stream) {
  const int block = pick_block(c->N);
  if (c->S.cplx && (which & 1)) {  // d_sign: [N] complex phases (interleaved), d_log: [N]
    if (prep_kernel(k_cx_value, c->smem_bytes)) return -1;
    k_cx_value<<<(c->N + block - 1) / block, block, c->smem_bytes, stream>>>(c->S, c->st, which, d_sign, d_log);
    c->nlaunch++;
    CK(cudaGetLastError());
    return 0;
  }
  if (prep_kernel(k_value, c->smem_bytes)) return -1;
  k_value<<<(c->N + block - 1) / block, block, c->smem_bytes, stream>>>(c->S, c->st, which, d_sign, d_log);
  c->nlaunch++;
  CK(cudaGetLastError());
  return 0;
}

// Slater + Jastrow internal update for the walkers flagged in d_mask (nullptr = all); the MO row
// and the new position are in st.saved_mo / st.saved_pos.
int launch_update(qmcb_ctx* c, int which, int e, const uint8_t* d_mask, cudaStream_t stream) {
  const Sys& S = c->S;
  if ((which & 1) && c->have_slater && S.cplx) {
    const int s = e >= S.nup ? 1 : 0;
    const long long nt = (long long)c->N * S.nds[s];
    if ((s ? S.ndn : S.nup) > 0) {
      k_cx_sm<<<(unsigned)((nt * 32 + 127) / 128), 128, 0, stream>>>(S, c->st, s, e - s * S.nup, d_mask);
      c->nlaunch++;
      CK(cudaGetLastError());
    }
    if (S.ndet > 1) {
      k_cx_det_cache<<<(unsigned)(((long long)c->N * 32 + 127) / 128), 128, 0, stream>>>(S, c->st, d_mask, s);
      c->nlaunch++;
      CK(cudaGetLastError());
    }
  } else if ((which & 1) && c->have_slater && S.ndet > 1 && (e >= S.nup ? S.ndn : S.nup) >= 1 &&
             (e >= S.nup ? S.ndn : S.nup) <= 8 && std::getenv("QMCB_NO_DET_UPDATE_FUSION") == nullptr) {
    // multi-determinant, n <= 8: Sherman-Morrison of every spin determinant + the dv / W caches in one launch
    const int s = e >= S.nup ? 1 : 0, n = s ? S.ndn : S.nup, ee = e - s * S.nup;
    const unsigned grid = (unsigned)(((long long)c->N * 32 + 127) / 128);
    switch (n) {
#define DU_T(NN) case NN: k_det_update<NN><<<grid, 128, 0, stream>>>(S, c->st, s, ee, d_mask); break;
      DU_T(1) DU_T(2) DU_T(3) DU_T(4) DU_T(5) DU_T(6) DU_T(7) DU_T(8)
#undef DU_T
    }
    c->nlaunch++;
    CK(cudaGetLastError());
  } else if ((which & 1) && c->have_slater) {
    const int s = e >= S.nup ? 1 : 0;
    SmArgs a{};
    a.n = s ? S.ndn : S.nup;
    a.e = e - s * S.nup;
    a.nds = S.nds[s];
    a.vec_stride = S.ldc[s];
    a.nmat = (long long)c->N * S.nds[s];
    a.inv = c->st.inv[s];
    a.vec = c->st.saved_mo;
    a.occ = S.iblob + S.o_occ[s];
    a.mask = d_mask;
    a.ratio = nullptr;
    a.dsign = c->st.dsign[s];
    a.dlog = c->st.dlog[s];
    if (launch_sm(c, a, stream, &c->nlaunch)) return -1;
    if (S.ndet > 1) {
      k_det_cache<<<(unsigned)(((long long)c->N * 32 + 127) / 128), 128, 0, stream>>>(S, c->st, d_mask, s);
      c->nlaunch++;
      CK(cudaGetLastError());
    }
  }
  // The walker coordinates are moved by the LAST factor of the context in the canonical order
  // (Slater, Jastrow, three-body), so a context driven factor by factor still sees the old
  // position in every cache update.
  const int owner = c->have_j3 ? 4 : (c->have_jastrow ? 2 : 1);
  const bool do_j = (which & 2) && c->have_jastrow;
  const bool do_j3 = (which & 4) && c->have_j3;
  if (do_j || (owner == 1 && (which & 1)) || (owner == 2 && (which & 2))) {
    const int mv = (which & owner) && owner != 4 ? 1 : 0;
    constexpr int GJ = 8;
    const int jper = ((S.ne > 1 ? S.ne - 1 : 0) * S.nb + 1) & ~1;
    const size_t jsm = ((c->smem_bytes + 15) & ~(size_t)15) + (size_t)(128 / GJ) * jper * 8;
    if (do_j && jsm <= 100 * 1024 && std::getenv("QMCB_NO_COOP_JUPDATE") == nullptr) {
      if (prep_kernel(k_jastrow_update_coop<GJ>, jsm)) return -1;
      k_jastrow_update_coop<GJ><<<(unsigned)(((long long)c->N * GJ + 127) / 128), 128, jsm, stream>>>(S, c->st, e, 1, mv, d_mask);
    } else {
      const int block = pick_block(c->N);
      if (prep_kernel(k_jastrow_update, c->smem_bytes)) return -1;
      k_jastrow_update<<<(c->N + block - 1) / block, block, c->smem_bytes, stream>>>(S, c->st, e, do_j ? 1 : 0, mv, d_mask);
    }
    c->nlaunch++;
    CK(cudaGetLastError());
  }
  if (do_j3) {
    constexpr int G3 = 16;
    const size_t sm3 = ((c->smem_bytes + 15) & ~(size_t)15) + (size_t)(128 / G3) * j3_update_scratch_doubles(S) * 8;
    if (prep_kernel(k_jastrow3_update_coop<G3>, sm3)) return -1;
    k_jastrow3_update_coop<G3><<<(unsigned)(((long long)c->N * G3 + 127) / 128), 128, sm3, stream>>>(S, c->st, e, 1, d_mask);
    c->nlaunch++;
    CK(cudaGetLastError());
  }
  return 0;
}

int ensure_energy_scratch(qmcb_ctx* c) {
  const Sys& S = c->S;
  const size_t N = c->N;
  const size_t nea = (size_t)S.ne * std::max(S.necp, 1) * N;
  int maxchan = 1;
  for (int a = 0; a < S.necp; ++a) maxchan = std::max(maxchan, c->chan_off[a + 1] - c->chan_off[a] - 1);
  if (c->e_ke.ensure(S.ne * N) || c->e_g2.ensure(S.ne * N) || c->e_loc.ensure(nea) || c->e_item.ensure(nea) ||
      c->e_work.ensure(nea) || c->e_vls.ensure(nea * maxchan) || c->e_contrib.ensure(nea * std::max(S.max_naip, 1)) ||
      c->e_count.ensure(1))
    return -1;
  EnergyScratch& es = c->es;
  es.ke_e = c->e_ke.p;
  es.g2_e = c->e_g2.p;
  es.ecp_loc = c->e_loc.p;
  es.item_of = c->e_item.p;
  es.work = c->e_work.p;
  es.vls = c->e_vls.p;
  es.contrib = c->e_contrib.p;
  es.ratio = nullptr;
  es.count = c->e_count.p;
  es.maxchan = maxchan;
  if (S.pbc) {
    const size_t npts = nea * std::max(S.max_naip, 1);
    if (c->e_ewald.ensure(2 * N) || c->e_ecppos.ensure(npts * 3) || c->e_ecpwrap.ensure(npts * 3)) return -1;
  }
  es.ewald = c->e_ewald.p;
  es.ecp_pos = c->e_ecppos.p;
  es.ecp_wrap = c->e_ecpwrap.p;
  return 0;
}

// periodic systems: pass 0 of k_ecp_points writes the wrapped quadrature points, k_pbc_mo evaluates
// the orbitals there into the scratch that pass 1 reads (ea is completed for pass 1)
template <int NMOT>
int ecp_points_pbc_prepass(qmcb_ctx* c, EcpPointArgs& ea, long long maxpts, long long grid, cudaStream_t stream) {
  const Sys& S = c->S;
  ea.pos_out = c->es.ecp_pos;
  ea.wrap_out = c->es.ecp_wrap;
  ea.pos_in = nullptr;
  k_ecp_points<NMOT><<<(unsigned)grid, 128, c->smem_bytes, stream>>>(S, c->st, c->es, ea);
  c->nlaunch++;
  CK(cudaGetLastError());
  if (c->have_slater) {
    PbcMoArgs a{};
    a.count = c->es.count;
    a.per_item = S.max_naip;
    a.pos = c->es.ecp_pos;
    a.wrap = c->es.ecp_wrap;
    a.naip = 1;
    a.spin_mode = 2;
    a.work = c->es.work;
    a.workN = c->N;
    a.necp = S.necp;
    a.e_only = ea.e_only;
    a.out = ea.scr;
    a.stride_p = 1;
    a.stride_c = 0;  // value rows only
    a.stride_j = (long long)ea.scr_stride;
    if (launch_pbc_mo(c, 0, a, maxpts, stream)) return -1;
  }
  ea.pos_out = nullptr;
  ea.wrap_out = nullptr;
  ea.pos_in = c->es.ecp_pos;
  return 0;
}

// `st` / `es`: the walker state and energy scratch the kernels read -- the live ones, or the snapshot the
// overlapped step loop hands over (qmcb_vmc_block_device)
// kinetic-energy pieces (es.ke_e / es.g2_e) from the cached MO rows, unless the sweep kernel already wrote them
int launch_kinetic(qmcb_ctx* c, const State& st, const EnergyScratch& es, cudaStream_t stream) {
  const Sys& S = c->S;
  const size_t sm = c->smem_bytes;
  const long long np = (long long)c->N * S.ne;
  if (c->kinetic_valid) return 0;
  if (!c->mocache_valid && c->have_slater) {  // protocol-path updates do not maintain the cache
    if (launch_mo_all(c, 0, stream)) return -1;
  }
  // three-body factor: per-group a-value scratch behind the tables
  const size_t ksm = c->have_j3 ? ((sm + 15) & ~(size_t)15) + (size_t)(128 / 8) * j3_scratch_doubles(S) * 8 : sm;
  if (S.cplx) {
    if (prep_kernel(k_cx_kinetic, sm)) return -1;
    k_cx_kinetic<<<(unsigned)((np + 63) / 64), 64, sm, stream>>>(S, st, es);
  } else {
    if (prep_kernel(k_kinetic<8>, ksm)) return -1;
    k_kinetic<8><<<(unsigned)((np * 8 + 127) / 128), 128, ksm, stream>>>(S, st, es);
  }
  c->nlaunch++;
  CK(cudaGetLastError());
  return 0;
}

template <int NMOT>
int launch_energy_t(qmcb_ctx* c, const State& st, const EnergyScratch& es, const double* d_u, const double* d_rot,
                    double* d_out, cudaStream_t stream) {
  const Sys& S = c->S;
  const int N = c->N;
  const size_t sm = c->smem_bytes;
  if (launch_kinetic(c, st, es, stream)) return -1;
  c->kinetic_valid = false;
  if (S.necp > 0) {
    CK(cudaMemsetAsync(es.count, 0, sizeof(int), stream));
    const long long nt = (long long)N * S.ne * S.necp;
    const int block = 128;
    if (prep_kernel(k_ecp_prepare, sm)) return -1;
    k_ecp_prepare<<<(unsigned)((nt + block - 1) / block), block, sm, stream>>>(S, st, es, d_u, -1);
    c->nlaunch++;
    CK(cudaGetLastError());
    EcpPointArgs ea{};
    ea.rot = d_rot;
    ea.quad = c->d_quad.p;
    ea.e_only = -1;
    ea.tmove_tau = 0.0;
    ea.scr = c->d_scr.p;
    ea.scr_stride = (size_t)nt * S.max_naip;
    const long long maxpts = nt * S.max_naip;
    const long long grid = std::min<long long>((maxpts + 127) / 128, 148LL * 16);
    if (prep_kernel(k_ecp_points<NMOT>, sm)) return -1;
    if (S.pbc) {
      if (ecp_points_pbc_prepass<NMOT>(c, ea, maxpts, grid, stream)) return -1;
    }
    if (S.cplx) {
      if (c->e_contrib_im.ensure((size_t)nt * std::max(S.max_naip, 1))) return -1;
      EcpCxArgs cxa{c->e_contrib_im.p, nullptr};
      if (prep_kernel(k_cx_ecp_points, sm)) return -1;
      k_cx_ecp_points<<<(unsigned)grid, 128, sm, stream>>>(S, st, es, ea, cxa);
    } else
    k_ecp_points<NMOT><<<(unsigned)grid, 128, sm, stream>>>(S, st, es, ea);
    c->nlaunch++;
    CK(cudaGetLastError());
  }
  if (S.pbc) {
    if (!c->have_ewald) return fail("periodic energy: qmcb_set_ewald has not been called");
    const size_t tab = (sm + 15) & ~(size_t)15;
    const size_t esm = tab + (size_t)(3 * S.ne + 16) * 8;
    if (prep_kernel(k_ewald, esm)) return -1;
    k_ewald<<<(unsigned)N, 128, esm, stream>>>(S, st, es.ewald);
    c->nlaunch++;
    CK(cudaGetLastError());
  }
  {
    const int ne = S.ne;
    int scr = std::max(std::max(ne * std::max(S.necp, 1), ne * (ne - 1) / 2), S.natom * ne);
    if (S.pbc) scr = ne * std::max(S.necp, 1);  // the Coulomb terms come from k_ewald
    scr = (scr + 1) & ~1;
    const size_t tab = (sm + 15) & ~(size_t)15;
    const size_t fsm = tab + (size_t)(128 / 8) * scr * 8;
    if (prep_kernel(k_energy_finalize<8>, fsm)) return -1;
    k_energy_finalize<8><<<(unsigned)(((long long)N * 8 + 127) / 128), 128, fsm, stream>>>(S, st, es, d_out, scr);
    c->nlaunch++;
    CK(cudaGetLastError());
  }
  if (S.cplx) {  // rows 6, 7 of the output: Im ecp, Im total
    if (S.necp > 0) {
      k_cx_ecp_imag<<<(unsigned)(((long long)N * 32 + 127) / 128), 128, (size_t)4 * S.ne * S.necp * 8, stream>>>(
          S, st, es, c->e_contrib_im.p, d_out + (size_t)6 * N);
      c->nlaunch++;
      CK(cudaGetLastError());
    } else
      CK(cudaMemsetAsync(d_out + (size_t)6 * N, 0, (size_t)2 * N * 8, stream));
  }
  return 0;
}

int launch_energy_on(qmcb_ctx* c, const State& st, const EnergyScratch& es, const double* d_u, const double* d_rot,
                     double* d_out, cudaStream_t stream) {
  if (c->nmot == 4) return launch_energy_t<4>(c, st, es, d_u, d_rot, d_out, stream);
  if (c->nmot == 8) return launch_energy_t<8>(c, st, es, d_u, d_rot, d_out, stream);
  return launch_energy_t<0>(c, st, es, d_u, d_rot, d_out, stream);
}

int launch_energy(qmcb_ctx* c, const double* d_u, const double* d_rot, double* d_out, cudaStream_t stream) {
  return launch_energy_on(c, c->st, c->es, d_u, d_rot, d_out, stream);
}

int energy_scratch_points(qmcb_ctx* c) {
  const Sys& S = c->S;
  // value rows only (one component) at the quadrature points of the general / periodic path
  const size_t pts = std::max<size_t>((size_t)c->N * S.ne * std::max(S.necp, 1) * std::max(S.max_naip, 1), (size_t)c->N * S.ne);
  return ensure_scratch(c, pts, 1);
}

}  // namespace

// =========================================================================================
extern "C" {

const char* qmcb_last_error(void) { return g_err.c_str(); }

int qmcb_device_count(void) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) return 0;
  return n;
}

int qmcb_create(int device, qmcb_ctx** out) {
  int n = 0;
  cudaError_t e = cudaGetDeviceCount(&n);
  if (e != cudaSuccess || n == 0) return fail("no CUDA device available: libqmcb200 has no CPU path");
  if (device < 0 || device >= n) return fail("device index out of range");
  CK(cudaSetDevice(device));
  qmcb_ctx* c = new qmcb_ctx();
  c->device = device;
  {
    int lo = 0, hi = 0;  // the copy stream also runs the draw-program kernels of the device generator: ahead of compute
    CK(cudaDeviceGetStreamPriorityRange(&lo, &hi));
    CK(cudaStreamCreateWithPriority(&c->copy_stream, cudaStreamNonBlocking, hi));
    // the walker moves are a dependent chain of short kernels: their CTAs go ahead of the overlapped energy
    // accumulator's (energy_stream, lowest priority), which only needs to finish before the next snapshot
    CK(cudaStreamCreateWithPriority(&c->stream, cudaStreamNonBlocking, lo > hi ? lo - 1 : lo));
  }
  for (int i = 0; i < qmcb_ctx::NSLOT; ++i) CK(cudaEventCreateWithFlags(&c->slot_ready[i], cudaEventDisableTiming));
  CK(cudaEventCreateWithFlags(&c->done_event, cudaEventDisableTiming | cudaEventBlockingSync));
  for (int i = 0; i < qmcb_ctx::NSLOT; ++i)
    CK(cudaEventCreateWithFlags(&c->block_done[i], cudaEventDisableTiming | cudaEventBlockingSync));
  CK(cudaStreamCreateWithFlags(&c->energy_stream, cudaStreamNonBlocking));
  CK(cudaStreamCreateWithFlags(&c->d2h_stream, cudaStreamNonBlocking));
  CK(cudaEventCreateWithFlags(&c->ev_block, cudaEventDisableTiming));
  CK(cudaEventCreateWithFlags(&c->ev_snap, cudaEventDisableTiming));
  CK(cudaEventCreateWithFlags(&c->ev_edone, cudaEventDisableTiming));
  *out = c;
  return 0;
}

void qmcb_destroy(qmcb_ctx* c) {
  if (!c) return;
  cudaSetDevice(c->device);
  cudaStreamSynchronize(c->stream);
  DBuf<double>* dd[] = {&c->d_dblob, &c->d_detc, &c->d_quad, &c->b_conf, &c->b_ap, &c->b_bp, &c->b_av, &c->b_bv,
                        &c->b_smo, &c->b_spos, &c->b_moall, &c->b_lu, &c->b_mocache, &c->b_a3v, &c->b_P3, &c->b_val3, &c->b_bpair, &c->b_gpair, &c->b_agrad, &c->b_lpair, &c->b_alap, &c->d_in, &c->d_out, &c->d_scr, &c->d_u,
                        &c->d_rot, &c->d_gauss, &c->d_unif, &c->d_energy, &c->d_esum, &c->e_ke, &c->e_g2, &c->e_loc,
                        &c->e_vls, &c->e_contrib};
  for (auto* b : dd) b->release();
  DBuf<double>* mb[] = {&c->m_tmu, &c->m_tmrot, &c->m_tmsel, &c->m_tmacc, &c->m_w, &c->m_eold, &c->m_v2old, &c->m_r2p, &c->m_r2a, &c->m_prod, &c->m_ws};
  for (auto* b : mb) b->release();
  c->m_ntacc.release();
  DBuf<double>* pb[] = {&c->d_ewdisp, &c->d_ewg, &c->d_ewion, &c->b_wrap, &c->b_swrap, &c->b_monew, &c->b_gold, &c->b_jold, &c->d_pwrap,
                        &c->e_ewald, &c->e_ecppos, &c->e_ecpwrap};
  for (auto* b : pb) b->release();
  for (int s = 0; s < 2; ++s) {
    c->b_inv[s].release();
    c->b_dsign[s].release();
    c->b_dlog[s].release();
    c->b_dv[s].release();
    c->b_W[s].release();
    c->b_ref[s].release();
    c->d_map[s].release();
    c->d_grp_off[s].release();
    c->d_grp_det[s].release();
    c->d_grp_coef[s].release();
    c->d_dense[s].release();
    c->d_grp_other[s].release();
  }
  c->d_iblob.release();
  c->d_idx.release();
  c->e_item.release();
  c->e_work.release();
  c->e_count.release();
  c->d_mask.release();
  c->d_accept.release();
  c->d_nacc.release();
  c->h_in.release();
  c->h_out.release();
  for (int i = 0; i < qmcb_ctx::NSLOT; ++i) {
    c->s_gauss[i].release();
    c->s_unif[i].release();
    c->s_u[i].release();
    c->s_rot[i].release();
    if (c->slot_ready[i]) cudaEventDestroy(c->slot_ready[i]);
    if (c->block_done[i]) cudaEventDestroy(c->block_done[i]);
  }
  if (c->done_event) cudaEventDestroy(c->done_event);
  if (c->energy_stream) cudaStreamSynchronize(c->energy_stream);
  for (auto* b : {&c->sn_inv[0], &c->sn_inv[1], &c->sn_conf, &c->sn_ap, &c->sn_bp, &c->sn_wrap, &c->e_ke2, &c->e_g22}) b->release();
  if (c->ev_snap) cudaEventDestroy(c->ev_snap);
  if (c->ev_edone) cudaEventDestroy(c->ev_edone);
  if (c->ev_fork) cudaEventDestroy(c->ev_fork);
  for (int i = 0; i < 3; ++i) {
    if (c->ev_join[i]) cudaEventDestroy(c->ev_join[i]);
    if (c->split_stream[i]) {
      cudaStreamSynchronize(c->split_stream[i]);
      cudaStreamDestroy(c->split_stream[i]);
    }
  }
  if (c->energy_stream) cudaStreamDestroy(c->energy_stream);
  if (c->d2h_stream) {
    cudaStreamSynchronize(c->d2h_stream);
    cudaStreamDestroy(c->d2h_stream);
  }
  if (c->ev_block) cudaEventDestroy(c->ev_block);
  for (auto& sl : c->dmc_slot) {
    for (auto* b : {&sl.gauss, &sl.unif, &sl.u, &sl.rot, &sl.tmu, &sl.tmrot, &sl.tmsel, &sl.tmacc, &sl.branch}) b->release();
    if (sl.ready) cudaEventDestroy(sl.ready);
  }
  for (int i = 0; i < qmcb_ctx::NSLOT; ++i) {
    c->r_energy[i].release();
    c->r_conf[i].release();
    c->r_esum[i].release();
    c->r_nacc[i].release();
  }
  devrng_free(c);
  if (c->copy_stream) cudaStreamDestroy(c->copy_stream);
  cudaStreamDestroy(c->stream);
  delete c;
}

int qmcb_set_atoms(qmcb_ctx* c, int natom, const double* xyz, const double* charges) {
  c->xyz.assign(xyz, xyz + 3 * natom);
  c->chg.assign(charges, charges + natom);
  c->dirty = true;
  return 0;
}

int qmcb_set_basis(qmcb_ctx* c, int nshell, const int32_t* shell_atom, const int32_t* shell_l,
                   const int32_t* prim_off, const double* prim_exp, const double* prim_coef) {
  c->sh_atom.assign(shell_atom, shell_atom + nshell);
  c->sh_l.assign(shell_l, shell_l + nshell);
  c->prim_off.assign(prim_off, prim_off + nshell + 1);
  const int np = prim_off[nshell];
  c->pexp.assign(prim_exp, prim_exp + np);
  c->pcoef.assign(prim_coef, prim_coef + np);
  c->dirty = true;
  return 0;
}

// Parameter setters are called before every recompute (the host objects push their current parameters);
// an unchanged parameter set must not invalidate the packed device tables, or every block would re-pack and
// re-upload them.  Each setter keeps the raw bytes of its last inputs and returns early on an exact match.
extern "C++" {
struct ParamKey {
  std::vector<unsigned char> b;
  template <class T>
  void add(const T* p, size_t n) {
    const size_t old = b.size(), bytes = n * sizeof(T);
    b.resize(old + bytes);
    if (bytes) std::memcpy(b.data() + old, p, bytes);
  }
  template <class T>
  void val(T v) {
    add(&v, 1);
  }
};
}  // extern "C++"

static thread_local bool g_setting_cx = false;

int qmcb_set_slater(qmcb_ctx* c, int nup, int ndn, int nmo_up, const double* mo_up, int nmo_dn,
                    const double* mo_dn, int ndet_up, const int32_t* occ_up, int ndet_dn,
                    const int32_t* occ_dn, int ndet, const int32_t* map_up, const int32_t* map_dn,
                    const double* det_coeff) {
  if (!g_setting_cx && c->cplx) {  // a real parameter set replaces a complex one
    c->cplx = false;
    c->dirty = true;
    c->key_slater.clear();
  }
  int nao = 0;
  for (size_t s = 0; s < c->sh_l.size(); ++s) nao += 2 * c->sh_l[s] + 1;
  if (nao == 0) return fail("qmcb_set_basis must be called before qmcb_set_slater");
  if (c->have_jastrow && (nup != c->nup || ndn != c->ndn)) return fail("electron counts differ from the Jastrow factor");
  ParamKey key;
  for (int v : {nao, nup, ndn, nmo_up, nmo_dn, ndet_up, ndet_dn, ndet}) key.val(v);
  key.add(mo_up, (size_t)nao * nmo_up);
  key.add(mo_dn, (size_t)nao * nmo_dn);
  key.add(occ_up, (size_t)ndet_up * nup);
  key.add(occ_dn, (size_t)ndet_dn * ndn);
  key.add(map_up, (size_t)ndet);
  key.add(map_dn, (size_t)ndet);
  key.add(det_coeff, (size_t)ndet);
  if (c->have_slater && key.b == c->key_slater) return 0;
  c->nup = nup;
  c->ndn = ndn;
  c->nmo[0] = nmo_up;
  c->nmo[1] = nmo_dn;
  c->mo[0].assign(mo_up, mo_up + (size_t)nao * nmo_up);
  c->mo[1].assign(mo_dn, mo_dn + (size_t)nao * nmo_dn);
  c->nds[0] = ndet_up;
  c->nds[1] = ndet_dn;
  c->occ[0].assign(occ_up, occ_up + (size_t)ndet_up * nup);
  c->occ[1].assign(occ_dn, occ_dn + (size_t)ndet_dn * ndn);
  for (int v : c->occ[0])
    if (v < 0 || v >= nmo_up) return fail("occupation index out of range (up)");
  for (int v : c->occ[1])
    if (v < 0 || v >= nmo_dn) return fail("occupation index out of range (down)");
  c->ndet = ndet;
  c->dmap[0].assign(map_up, map_up + ndet);
  c->dmap[1].assign(map_dn, map_dn + ndet);
  c->detc.assign(det_coeff, det_coeff + ndet);
  c->have_slater = true;
  c->dirty = true;
  c->key_slater.swap(key.b);
  return 0;
}

int qmcb_set_slater_cx(qmcb_ctx* c, int nup, int ndn, int nmo_up, const double* mo_up_re, const double* mo_up_im,
                       int nmo_dn, const double* mo_dn_re, const double* mo_dn_im, int ndet_up, const int32_t* occ_up,
                       int ndet_dn, const int32_t* occ_dn, int ndet, const int32_t* map_up, const int32_t* map_dn,
                       const double* det_re, const double* det_im) {
  int nao = 0;
  for (size_t s = 0; s < c->sh_l.size(); ++s) nao += 2 * c->sh_l[s] + 1;
  if (nao == 0) return fail("qmcb_set_basis must be called before qmcb_set_slater_cx");
  // the imaginary parts are part of the parameter key: keep them before the real setter compares / swaps its key
  std::vector<double> im_up(mo_up_im, mo_up_im + (size_t)nao * nmo_up), im_dn(mo_dn_im, mo_dn_im + (size_t)nao * nmo_dn),
      im_det(det_im, det_im + ndet);
  const bool same_im = c->cplx && im_up == c->mo_im[0] && im_dn == c->mo_im[1] && im_det == c->detc_im;
  g_setting_cx = true;
  const int rc = qmcb_set_slater(c, nup, ndn, nmo_up, mo_up_re, nmo_dn, mo_dn_re, ndet_up, occ_up, ndet_dn, occ_dn, ndet,
                                 map_up, map_dn, det_re);
  g_setting_cx = false;
  if (rc) return -1;
  if (!same_im) {
    c->mo_im[0].swap(im_up);
    c->mo_im[1].swap(im_dn);
    c->detc_im.swap(im_det);
    c->cplx = true;
    c->dirty = true;
  }
  return 0;
}

int qmcb_set_pbc_phases_imag(qmcb_ctx* c, int nL, int nk, const double* phases_im) {
  if (!c->have_pbc_orb || nL != c->nL || nk != c->nk) return fail("qmcb_set_pbc_phases_imag: call qmcb_set_pbc_orbitals first (same nL, nk)");
  c->phases_im.assign(phases_im, phases_im + (size_t)nL * nk);
  c->dirty = true;
  return 0;
}

int qmcb_is_complex(qmcb_ctx* c) { return c->cplx && c->have_slater ? 1 : 0; }

int qmcb_set_jastrow(qmcb_ctx* c, int nup, int ndn, int na, const int32_t* a_kind, const double* a_par,
                     double rcut_a, int nb, const int32_t* b_kind, const double* b_par, double rcut_b,
                     const double* acoeff, const double* bcoeff) {
  if (c->have_slater && (nup != c->nup || ndn != c->ndn)) return fail("electron counts differ from the Slater factor");
  const int natom = (int)c->chg.size();
  if (natom == 0) return fail("qmcb_set_atoms must be called before qmcb_set_jastrow");
  ParamKey key;
  for (int v : {natom, nup, ndn, na, nb}) key.val(v);
  key.val(rcut_a);
  key.val(rcut_b);
  key.add(a_kind, (size_t)na);
  key.add(a_par, (size_t)na);
  key.add(b_kind, (size_t)nb);
  key.add(b_par, (size_t)nb);
  key.add(acoeff, (size_t)natom * na * 2);
  key.add(bcoeff, (size_t)nb * 3);
  if (c->have_jastrow && key.b == c->key_jastrow) return 0;
  c->nup = nup;
  c->ndn = ndn;
  c->na = na;
  c->nb = nb;
  c->akind.assign(a_kind, a_kind + na);
  c->apar.assign(a_par, a_par + na);
  c->bkind.assign(b_kind, b_kind + nb);
  c->bpar.assign(b_par, b_par + nb);
  c->rcut_a = rcut_a;
  c->rcut_b = rcut_b;
  c->acoef.assign(acoeff, acoeff + (size_t)natom * na * 2);
  c->bcoef.assign(bcoeff, bcoeff + (size_t)nb * 3);
  c->have_jastrow = true;
  c->dirty = true;
  c->key_jastrow.swap(key.b);
  return 0;
}

int qmcb_set_jastrow3(qmcb_ctx* c, int nup, int ndn, int na, const int32_t* a_kind, const double* a_par,
                      double rcut_a, int nb, const int32_t* b_kind, const double* b_par, double rcut_b,
                      const double* ccoeff) {
  if ((c->have_slater || c->have_jastrow) && (nup != c->nup || ndn != c->ndn))
    return fail("electron counts differ from the other factors");
  const int natom = (int)c->chg.size();
  if (natom == 0) return fail("qmcb_set_atoms must be called before qmcb_set_jastrow3");
  if (nb > QMCB_J3_MAXB) return fail("three-body Jastrow: more than 8 b functions");
  if (natom * na > QMCB_J3_MAXA) return fail("three-body Jastrow: natom * na > 128");
  ParamKey key;
  for (int v : {natom, nup, ndn, na, nb}) key.val(v);
  key.val(rcut_a);
  key.val(rcut_b);
  key.add(a_kind, (size_t)na);
  key.add(a_par, (size_t)na);
  key.add(b_kind, (size_t)nb);
  key.add(b_par, (size_t)nb);
  key.add(ccoeff, (size_t)natom * na * na * nb * 3);
  if (c->have_j3 && key.b == c->key_j3) return 0;
  c->nup = nup;
  c->ndn = ndn;
  c->na3 = na;
  c->nb3 = nb;
  c->a3kind.assign(a_kind, a_kind + na);
  c->a3par.assign(a_par, a_par + na);
  c->b3kind.assign(b_kind, b_kind + nb);
  c->b3par.assign(b_par, b_par + nb);
  c->rcut_a3 = rcut_a;
  c->rcut_b3 = rcut_b;
  // C = (ccoeff + ccoeff^T(k,l)) / 2   (three_body_jastrow.py:94-96)
  c->c3.resize((size_t)natom * na * na * nb * 3);
  for (int I = 0; I < natom; ++I)
    for (int k = 0; k < na; ++k)
      for (int l = 0; l < na; ++l)
        for (int m = 0; m < nb; ++m)
          for (int t = 0; t < 3; ++t) {
            const size_t a = ((((size_t)I * na + k) * na + l) * nb + m) * 3 + t;
            const size_t b = ((((size_t)I * na + l) * na + k) * nb + m) * 3 + t;
            c->c3[a] = (ccoeff[a] + ccoeff[b]) / 2;
          }
  c->have_j3 = true;
  c->dirty = true;
  c->key_j3.swap(key.b);
  return 0;
}

int qmcb_set_ecp(qmcb_ctx* c, int necp, const int32_t* ecp_atom, const int32_t* chan_off,
                 const int32_t* term_off, const int32_t* term_power, const double* term_alpha,
                 const double* term_coef, const int32_t* naip, const double* quad, double threshold) {
  c->necp = necp;
  c->ecp_atom.assign(ecp_atom, ecp_atom + necp);
  c->chan_off.assign(chan_off, chan_off + necp + 1);
  const int nchan = necp ? chan_off[necp] : 0;
  c->term_off.assign(term_off, term_off + nchan + 1);
  const int nterm = nchan ? term_off[nchan] : 0;
  c->term_pow.assign(term_power, term_power + nterm);
  c->term_alpha.assign(term_alpha, term_alpha + nterm);
  c->term_coef.assign(term_coef, term_coef + nterm);
  c->naip.assign(naip, naip + necp);
  size_t nq = 0;
  for (int a = 0; a < necp; ++a) {
    nq += (size_t)naip[a] * 4;
    if (chan_off[a + 1] - chan_off[a] > 8) return fail("more than 8 ECP channels per atom");
    if (chan_off[a + 1] - chan_off[a] - 2 > 4) return fail("ECP channels with l > 4 are not supported (eval_ecp.py:203-225)");
  }
  c->quad.assign(quad, quad + nq);
  c->threshold = threshold;
  c->dirty = true;
  return 0;
}

int qmcb_set_lattice(qmcb_ctx* c, const double* lat, int mode, const double* shifts) {
  if (mode < 1 || mode > 3) return fail("minimal-image mode must be 1 (diagonal), 2 (orthogonal) or 3 (general)");
  c->lat.assign(lat, lat + 9);
  c->shifts.assign(shifts, shifts + 81);
  c->pbc_mode = mode;
  c->dirty = true;
  return 0;
}

int qmcb_set_pbc_orbitals(qmcb_ctx* c, int nbatom, const double* bxyz, const double* lprim, const double* smat, int nk,
                          const double* kpts, int nL, const double* Ls, const int32_t* num_Ls,
                          const double* atom_cutoff, int nshell, const double* l_cutoff, const double* phases,
                          int nmo_up, const int32_t* mo_k_up, int nmo_dn, const int32_t* mo_k_dn, int isgamma) {
  if (!c->pbc_mode) return fail("qmcb_set_lattice must be called before qmcb_set_pbc_orbitals");
  if (nk < 1 || nL < 1 || nbatom < 1) return fail("qmcb_set_pbc_orbitals: empty k-point / image / atom list");
  c->bxyz.assign(bxyz, bxyz + 3 * nbatom);
  c->lprim.assign(lprim, lprim + 9);
  c->smat.assign(smat, smat + 9);
  c->nk = nk;
  c->kpts.assign(kpts, kpts + 3 * nk);
  c->nL = nL;
  c->Ls.assign(Ls, Ls + 3 * nL);
  c->numLs.assign(num_Ls, num_Ls + nbatom);
  for (int v : c->numLs)
    if (v < 1 || v > nL) return fail("qmcb_set_pbc_orbitals: num_Ls out of range");
  c->atomcut.assign(atom_cutoff, atom_cutoff + nbatom);
  c->lcut.assign(l_cutoff, l_cutoff + nshell);
  c->phases.assign(phases, phases + (size_t)nL * nk);
  c->mok[0].assign(mo_k_up, mo_k_up + nmo_up);
  c->mok[1].assign(mo_k_dn, mo_k_dn + nmo_dn);
  for (int s = 0; s < 2; ++s)
    for (int v : c->mok[s])
      if (v < 0 || v >= nk) return fail("qmcb_set_pbc_orbitals: k-point index out of range");
  c->isgamma = isgamma ? 1 : 0;
  c->have_pbc_orb = true;
  c->dirty = true;
  return 0;
}

int qmcb_set_ewald(qmcb_ctx* c, double alpha, int ndisp, const double* disp, int nG, const double* gpoints,
                   const double* gweight, const double* ion_re, const double* ion_im, double ijconst,
                   double squareconst, double i_sum, double e_ii) {
  Guard g(c);
  if (!c->pbc_mode) return fail("qmcb_set_lattice must be called before qmcb_set_ewald");
  std::vector<double> gt((size_t)std::max(nG, 1) * 4, 0.0), ion((size_t)std::max(nG, 1) * 2, 0.0);
  for (int i = 0; i < nG; ++i) {
    gt[4 * i] = gpoints[3 * i];
    gt[4 * i + 1] = gpoints[3 * i + 1];
    gt[4 * i + 2] = gpoints[3 * i + 2];
    gt[4 * i + 3] = gweight[i];
    ion[2 * i] = ion_re[i];
    ion[2 * i + 1] = ion_im[i];
  }
  if (c->d_ewdisp.ensure((size_t)std::max(ndisp, 1) * 3) || c->d_ewg.ensure(gt.size()) || c->d_ewion.ensure(ion.size())) return -1;
  CK(cudaMemcpy(c->d_ewdisp.p, disp, (size_t)ndisp * 3 * 8, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(c->d_ewg.p, gt.data(), gt.size() * 8, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(c->d_ewion.p, ion.data(), ion.size() * 8, cudaMemcpyHostToDevice));
  c->ew_alpha = alpha;
  c->ew_ndisp = ndisp;
  c->ew_nG = nG;
  c->ew_ij = ijconst;
  c->ew_sq = squareconst;
  c->ew_isum = i_sum;
  c->ew_eii = e_ii;
  c->have_ewald = true;
  c->dirty = true;
  return 0;
}

int qmcb_set_point_wrap(qmcb_ctx* c, const double* wrap, int64_t count) {
  Guard g(c);
  if (count <= 0 || !wrap) {
    c->pending_wrap = false;
    return 0;
  }
  if (c->d_pwrap.ensure((size_t)count * 3)) return -1;
  if (h2d(c, c->d_pwrap.p, wrap, (size_t)count * 3 * 8)) return -1;
  c->pending_wrap = true;
  return 0;
}

// ---------------------------------------------------------------------------------------
// recompute kernels of the selected factors from the device-resident coordinates (no copies, no sync)
static int recompute_from_resident(qmcb_ctx* c, int which, int nconf) {
  const Sys& S = c->S;
  if (which & 1)
    if (slater_rebuild(c, c->stream)) return -1;
  if (which & 2) {
    if (S.ne >= 16 && S.nb <= 8 && std::getenv("QMCB_NO_COOP_JRECOMPUTE") == nullptr) {
      // many electrons: CTA per walker, threads over electrons
      const int block = S.ne <= 32 ? 32 : (S.ne <= 64 ? 64 : 128);
      if (prep_kernel(k_jastrow_recompute_coop, c->smem_bytes)) return -1;
      k_jastrow_recompute_coop<<<nconf, block, c->smem_bytes, c->stream>>>(S, c->st);
    } else if (S.ne >= 2 && S.ne < 16 && S.npair * S.nb <= 512 && std::getenv("QMCB_NO_COOP_JRECOMPUTE") == nullptr) {
      // few electrons: 8 lanes per walker, pair values through shared memory, sums in the one-thread order
      constexpr int GR = 8;
      const int per = (S.npair * S.nb + 1) & ~1;
      const size_t rsm = ((c->smem_bytes + 15) & ~(size_t)15) + (size_t)(128 / GR) * per * 8;
      if (prep_kernel(k_jastrow_recompute_group<GR>, rsm)) return -1;
      k_jastrow_recompute_group<GR><<<(unsigned)(((long long)nconf * GR + 127) / 128), 128, rsm, c->stream>>>(S, c->st);
    } else {
      const int block = pick_block(nconf);
      if (prep_kernel(k_jastrow_recompute, c->smem_bytes)) return -1;
      k_jastrow_recompute<<<(nconf + block - 1) / block, block, c->smem_bytes, c->stream>>>(S, c->st);
    }
    c->nlaunch++;
    CK(cudaGetLastError());
  }
  if ((which & 4) && S.ne >= 2 && std::getenv("QMCB_NO_COOP_JRECOMPUTE") == nullptr) {
    constexpr int G3R = 8;  // lanes over electrons
    if (prep_kernel(k_jastrow3_recompute_group<G3R>, c->smem_bytes)) return -1;
    k_jastrow3_recompute_group<G3R><<<(unsigned)(((long long)nconf * G3R + 127) / 128), 128, c->smem_bytes, c->stream>>>(S, c->st);
    c->nlaunch++;
    CK(cudaGetLastError());
  } else if (which & 4) {
    const int block = pick_block(nconf);
    if (prep_kernel(k_jastrow3_recompute, c->smem_bytes)) return -1;
    k_jastrow3_recompute<<<(nconf + block - 1) / block, block, c->smem_bytes, c->stream>>>(S, c->st);
    c->nlaunch++;
    CK(cudaGetLastError());
  }
  c->saved_slot = -1;
  c->paircache_valid = false;
  c->kinetic_valid = false;
  return 0;
}

int qmcb_recompute(qmcb_ctx* c, int which, int nconf, const double* configs, double* sign, double* logval) {
  return qmcb_recompute_pbc(c, which, nconf, configs, nullptr, sign, logval);
}

int qmcb_recompute_pbc(qmcb_ctx* c, int which, int nconf, const double* configs, const double* wrap, double* sign,
                       double* logval) {
  Guard g(c);
  if (which_ok(c, which)) return -1;
  if (ensure_state(c, nconf)) return -1;
  const Sys& S = c->S;
  const size_t nel = (size_t)nconf * S.ne * 3;
  if (c->d_in.ensure(nel) || c->d_out.ensure((size_t)nconf * 8)) return -1;
  if (S.pbc) {
    if (wrap) {
      if (h2d(c, c->st.wrap, wrap, nel * 8)) return -1;
    } else
      CK(cudaMemsetAsync(c->st.wrap, 0, nel * 8, c->stream));
  }
  if (h2d(c, c->d_in.p, configs, nel * 8)) return -1;
  k_conf_in<<<(unsigned)((nel + 255) / 256), 256, 0, c->stream>>>(c->d_in.p, c->st.conf, nconf, S.ne);
  c->nlaunch++;
  CK(cudaGetLastError());
  if (recompute_from_resident(c, which, nconf)) return -1;
  if (sign || logval) return qmcb_value(c, which, sign, logval);
  CK(cudaStreamSynchronize(c->stream));
  return 0;
}

// wf.recompute for coordinates that are ALREADY on the device (the walkers a device-resident block just
// returned): same kernels as qmcb_recompute, no upload, no value read-back, asynchronous on the context's
// stream.  The host driver uses it at the start of block b+1 when nothing touched the state since block b.
int qmcb_recompute_resident(qmcb_ctx* c, int which) {
  Guard g(c);
  if (which_ok(c, which)) return -1;
  if (c->N == 0) return fail("recompute has not been called");
  if (build_tables(c)) return -1;
  if (c->N == 0) return fail("system shapes changed: call recompute again");
  return recompute_from_resident(c, which, (int)c->N);
}

// the same on a caller-supplied stream (bench.py times whole blocks -- recompute + steps -- with events on its own stream)
int qmcb_recompute_resident_on(qmcb_ctx* c, int which, void* stream_) {
  if (!stream_) return qmcb_recompute_resident(c, which);
  cudaStream_t keep = c->stream;
  c->stream = (cudaStream_t)stream_;
  const int rc = qmcb_recompute_resident(c, which);
  c->stream = keep;
  return rc;
}

int qmcb_value(qmcb_ctx* c, int which, double* sign, double* logval) {
  Guard g(c);
  if (which_ok(c, which)) return -1;
  if (c->N == 0) return fail("recompute has not been called");
  if (build_tables(c)) return -1;
  if (c->N == 0) return fail("system shapes changed: call recompute again");
  const size_t N = c->N;
  const size_t cw = (c->S.cplx && (which & 1)) ? 2 : 1;  // complex context: `sign` is the complex unit phase
  if (c->d_out.ensure(N * 8)) return -1;
  if (launch_value(c, which, c->d_out.p, c->d_out.p + cw * N, c->stream)) return -1;
  if (c->h_out.ensure((cw + 1) * N * 8)) return -1;
  CK(cudaMemcpyAsync(c->h_out.p, c->d_out.p, (cw + 1) * N * 8, cudaMemcpyDeviceToHost, c->stream));
  CK(cudaStreamSynchronize(c->stream));
  if (sign) std::memcpy(sign, c->h_out.p, cw * N * 8);
  if (logval) std::memcpy(logval, (double*)c->h_out.p + cw * N, N * 8);
  return 0;
}

// periodic systems: lattice-summed MO rows of electron e's spin at the points of a protocol call,
// in the scratch layout slater_point_general reads (scr[(c * ldc + j) * stride + p])
static int pbc_point_rows(qmcb_ctx* c, int deriv, int e, const double* d_pos, const double* d_wrap, const int* d_idx,
                          int naip, long long npoints, size_t stride, cudaStream_t stream) {
  const Sys& S = c->S;
  const int s = e >= S.nup ? 1 : 0;
  PbcMoArgs a{};
  a.npoints = npoints;
  a.pos = d_pos;
  a.wrap = d_wrap;
  a.idx = d_idx;
  a.naip = naip;
  a.spin_mode = 0;
  a.spin = s;
  a.out = c->d_scr.p;
  a.stride_p = 1;
  a.stride_c = (long long)S.ldc[s] * (long long)stride;
  a.stride_j = (long long)stride;
  return launch_pbc_mo(c, deriv, a, npoints, stream);
}

static int point_call(qmcb_ctx* c, int mode, int which, int e, const double* epos, int naip, const uint8_t* mask,
                      double* o1, double* o2, int64_t* slot) {
  Guard g(c);
  if (which_ok(c, which)) return -1;
  if (c->N == 0) return fail("recompute has not been called");
  if (build_tables(c)) return -1;
  if (c->N == 0) return fail("system shapes changed: call recompute again");
  const Sys& S = c->S;
  if (e < 0 || e >= S.ne) return fail("electron index out of range");
  const size_t N = c->N;
  const size_t cw = (S.cplx && (which & 1)) ? 2 : 1;  // complex context: wave-function-valued outputs are complex128
  if (c->d_in.ensure(N * naip * 3) || c->d_out.ensure(cw * std::max<size_t>(N * 8, N * naip))) return -1;
  if (h2d(c, c->d_in.p, epos, N * naip * 3 * 8)) return -1;
  PointArgs pa{};
  pa.which = which;
  pa.e = e;
  pa.naip = naip;
  pa.pos = c->d_in.p;
  size_t nm = N;
  if (mask) {
    std::vector<int> idx;
    idx.reserve(N);
    for (size_t w = 0; w < N; ++w)
      if (mask[w]) idx.push_back((int)w);
    nm = idx.size();
    if (c->d_idx.ensure(N)) return -1;
    if (nm) CK(cudaMemcpyAsync(c->d_idx.p, idx.data(), nm * 4, cudaMemcpyHostToDevice, c->stream));
    CK(cudaStreamSynchronize(c->stream));
    pa.idx = c->d_idx.p;
  }
  pa.npoints = (int)(nm * naip);
  pa.o_val = c->d_out.p + cw * 3 * N;
  pa.o_grad = c->d_out.p;
  pa.o_lap = c->d_out.p + cw * 3 * N;
  if (mode == PV_VALUE) pa.o_val = c->d_out.p;
  const bool save = (mode == PV_GRADVAL) || (mode == PV_VALUE && naip == 1 && !mask);
  pa.save = save ? 1 : 0;
  if (ensure_scratch(c, pa.npoints, 5)) return -1;
  pa.scr = c->d_scr.p;
  pa.scr_stride = std::max(pa.npoints, 1);
  const double* d_pwrap = c->pending_wrap ? c->d_pwrap.p : nullptr;
  c->pending_wrap = false;
  if (S.pbc && (which & 1)) {
    const int deriv = mode == PV_VALUE ? 0 : (mode == PV_GRADLAP ? 2 : 1);
    if (pbc_point_rows(c, deriv, e, c->d_in.p, d_pwrap, pa.idx, naip, pa.npoints, pa.scr_stride, c->stream)) return -1;
  }
  int rc = 0;
  switch (mode) {
    case PV_VALUE: rc = launch_point<PV_VALUE>(c, pa, c->stream); break;
    case PV_GRAD: rc = launch_point<PV_GRAD>(c, pa, c->stream); break;
    case PV_GRADVAL: rc = launch_point<PV_GRADVAL>(c, pa, c->stream); break;
    default: rc = launch_point<PV_GRADLAP>(c, pa, c->stream); break;
  }
  if (rc) return rc;
  if (save && S.pbc) {
    if (d_pwrap)
      CK(cudaMemcpyAsync(c->st.saved_wrap, d_pwrap, N * 3 * 8, cudaMemcpyDeviceToDevice, c->stream));
    else
      CK(cudaMemsetAsync(c->st.saved_wrap, 0, N * 3 * 8, c->stream));
  }
  if (save) {
    c->saved_slot = ++c->slot_counter;
    c->saved_e = e;
    c->saved_which = which;
    if (slot) *slot = c->saved_slot;
  } else if (slot)
    *slot = -1;
  if (mode == PV_VALUE) {
    if (d2h(c, o1, c->d_out.p, cw * nm * naip * 8)) return -1;
  } else {
    const size_t nout = cw * ((mode == PV_GRAD) ? 3 * N : 4 * N);
    if (c->h_out.ensure(nout * 8)) return -1;
    CK(cudaMemcpyAsync(c->h_out.p, c->d_out.p, nout * 8, cudaMemcpyDeviceToHost, c->stream));
    CK(cudaStreamSynchronize(c->stream));
    std::memcpy(o1, c->h_out.p, cw * 3 * N * 8);
    if (o2) std::memcpy(o2, (double*)c->h_out.p + cw * 3 * N, cw * N * 8);
  }
  return 0;
}

int qmcb_gradient(qmcb_ctx* c, int which, int e, const double* epos, double* grad) {
  return point_call(c, PV_GRAD, which, e, epos, 1, nullptr, grad, nullptr, nullptr);
}
int qmcb_gradient_value(qmcb_ctx* c, int which, int e, const double* epos, double* grad, double* val, int64_t* slot) {
  return point_call(c, PV_GRADVAL, which, e, epos, 1, nullptr, grad, val, slot);
}
int qmcb_gradient_laplacian(qmcb_ctx* c, int which, int e, const double* epos, double* grad, double* lap) {
  return point_call(c, PV_GRADLAP, which, e, epos, 1, nullptr, grad, lap, nullptr);
}
int qmcb_testvalue(qmcb_ctx* c, int which, int e, const double* epos, int naip, const uint8_t* mask, double* ratio,
                   int64_t* slot) {
  return point_call(c, PV_VALUE, which, e, epos, naip, mask, ratio, nullptr, slot);
}

int qmcb_testvalue_many(qmcb_ctx* c, int which, int ne_list, const int32_t* elist, const double* epos,
                        const uint8_t* mask, double* ratio) {
  Guard g(c);
  if (which_ok(c, which)) return -1;
  if (c->N == 0) return fail("recompute has not been called");
  if (build_tables(c)) return -1;
  if (c->N == 0) return fail("system shapes changed: call recompute again");
  const Sys& S = c->S;
  const size_t N = c->N;
  const size_t cw = (S.cplx && (which & 1)) ? 2 : 1;
  if (c->d_in.ensure(N * 3) || c->d_out.ensure(cw * std::max<size_t>(N * ne_list, 8 * N))) return -1;
  if (h2d(c, c->d_in.p, epos, N * 3 * 8)) return -1;
  size_t nm = N;
  PointArgs pa{};
  if (mask) {
    std::vector<int> idx;
    for (size_t w = 0; w < N; ++w)
      if (mask[w]) idx.push_back((int)w);
    nm = idx.size();
    if (c->d_idx.ensure(N)) return -1;
    if (nm) CK(cudaMemcpyAsync(c->d_idx.p, idx.data(), nm * 4, cudaMemcpyHostToDevice, c->stream));
    CK(cudaStreamSynchronize(c->stream));
    pa.idx = c->d_idx.p;
  }
  if (ensure_scratch(c, nm, 5)) return -1;
  const double* d_pwrap = c->pending_wrap ? c->d_pwrap.p : nullptr;
  c->pending_wrap = false;
  for (int i = 0; i < ne_list; ++i) {
    if (elist[i] < 0 || elist[i] >= S.ne) return fail("electron index out of range");
    if (S.pbc && (which & 1) &&
        pbc_point_rows(c, 0, elist[i], c->d_in.p, d_pwrap, pa.idx, 1, (long long)nm, std::max<size_t>(nm, 1), c->stream))
      return -1;
    pa.which = which;
    pa.e = elist[i];
    pa.naip = 1;
    pa.pos = c->d_in.p;
    pa.npoints = (int)nm;
    pa.o_val = c->d_out.p + cw * (size_t)i * nm;
    pa.save = 0;
    pa.scr = c->d_scr.p;
    pa.scr_stride = std::max<size_t>(nm, 1);
    if (launch_point<PV_VALUE>(c, pa, c->stream)) return -1;
  }
  std::vector<double> tmp(cw * nm * ne_list);
  if (d2h(c, tmp.data(), c->d_out.p, tmp.size() * 8)) return -1;
  for (size_t m = 0; m < nm; ++m)
    for (int i = 0; i < ne_list; ++i)
      for (size_t r = 0; r < cw; ++r) ratio[(m * ne_list + i) * cw + r] = tmp[((size_t)i * nm + m) * cw + r];
  c->saved_slot = -1;
  return 0;
}

int qmcb_updateinternals(qmcb_ctx* c, int which, int e, const double* epos, const uint8_t* mask, int64_t slot) {
  Guard g(c);
  if (which_ok(c, which)) return -1;
  if (c->N == 0) return fail("recompute has not been called");
  if (build_tables(c)) return -1;
  if (c->N == 0) return fail("system shapes changed: call recompute again");
  const Sys& S = c->S;
  if (e < 0 || e >= S.ne) return fail("electron index out of range");
  const size_t N = c->N;
  const uint8_t* d_mask = nullptr;
  if (mask) {
    if (c->d_mask.ensure(N)) return -1;
    if (h2d(c, c->d_mask.p, mask, N)) return -1;
    d_mask = c->d_mask.p;
  }
  const bool have_saved = slot >= 0 && slot == c->saved_slot && e == c->saved_e && ((c->saved_which & which) == which);
  if (!have_saved) {
    if (c->d_in.ensure(N * 3)) return -1;
    if (h2d(c, c->d_in.p, epos, N * 3 * 8)) return -1;
    if (which & 1) {
      PointArgs pa{};
      pa.which = 1;
      pa.e = e;
      pa.naip = 1;
      pa.pos = c->d_in.p;
      pa.npoints = (int)N;
      pa.mask = d_mask;
      pa.save = 1;
      if (ensure_scratch(c, N, 5)) return -1;
      pa.scr = c->d_scr.p;
      pa.scr_stride = N;
      if (S.pbc && pbc_point_rows(c, 0, e, c->d_in.p, c->pending_wrap ? c->d_pwrap.p : nullptr, nullptr, 1, (long long)N, N, c->stream))
        return -1;
      if (launch_point<PV_MOSAVE>(c, pa, c->stream)) return -1;
    }
    CK(cudaMemcpyAsync(c->st.saved_pos, c->d_in.p, N * 3 * 8, cudaMemcpyDeviceToDevice, c->stream));
    if (S.pbc) {
      if (c->pending_wrap)
        CK(cudaMemcpyAsync(c->st.saved_wrap, c->d_pwrap.p, N * 3 * 8, cudaMemcpyDeviceToDevice, c->stream));
      else
        CK(cudaMemsetAsync(c->st.saved_wrap, 0, N * 3 * 8, c->stream));
    }
  }
  c->pending_wrap = false;
  if (launch_update(c, which, e, d_mask, c->stream)) return -1;
  if (which & 1) c->mocache_valid = false;
  c->paircache_valid = false;
  c->kinetic_valid = false;
  CK(cudaStreamSynchronize(c->stream));
  // a shared context may be driven factor by factor (Slater call, then Jastrow call with the
  // same token), so the slot stays valid until the next query overwrites it
  return 0;
}

int qmcb_get_state(qmcb_ctx* c, const char* name, double* out) {
  Guard g(c);
  if (c->N == 0) return fail("recompute has not been called");
  if (build_tables(c)) return -1;
  if (c->N == 0) return fail("system shapes changed: call recompute again");
  const Sys& S = c->S;
  const size_t N = c->N;
  const std::string k(name);
  CK(cudaStreamSynchronize(c->stream));
  auto fetch = [&](const double* d, size_t n, std::vector<double>& h) -> int {
    h.resize(n);
    CK(cudaMemcpy(h.data(), d, n * 8, cudaMemcpyDeviceToHost));
    return 0;
  };
  std::vector<double> h;
  if (k == "inverse_up" || k == "inverse_dn") {
    const int s = k == "inverse_dn";
    const int n = s ? S.ndn : S.nup;
    if (fetch(c->st.inv[s], N * S.nds[s] * n * n, h)) return -1;
    if (S.cplx) {  // complex128 (interleaved)
      std::vector<double> hi;
      if (fetch(c->st.inv_im[s], h.size(), hi)) return -1;
      for (size_t i = 0; i < h.size(); ++i) {
        out[2 * i] = h[i];
        out[2 * i + 1] = hi[i];
      }
    } else
      std::memcpy(out, h.data(), h.size() * 8);
  } else if (k == "dets_up" || k == "dets_dn") {
    const int s = k == "dets_dn";
    const size_t nd = N * S.nds[s];
    if (fetch(c->st.dsign[s], nd, h)) return -1;
    if (S.cplx) {  // [2][N][D_s] complex128: phase, then log (imaginary part 0), like the reference's _dets
      std::vector<double> hi, hl;
      if (fetch(c->st.dphs_im[s], nd, hi) || fetch(c->st.dlog[s], nd, hl)) return -1;
      for (size_t i = 0; i < nd; ++i) {
        out[2 * i] = h[i];
        out[2 * i + 1] = hi[i];
        out[2 * (nd + i)] = hl[i];
        out[2 * (nd + i) + 1] = 0.0;
      }
      return 0;
    }
    std::memcpy(out, h.data(), nd * 8);
    if (fetch(c->st.dlog[s], nd, h)) return -1;
    std::memcpy(out + nd, h.data(), nd * 8);
  } else if (k == "wrap") {
    if (!S.pbc) return fail("wrap vectors exist only for periodic systems");
    if (fetch(c->st.wrap, N * S.ne * 3, h)) return -1;
    std::memcpy(out, h.data(), h.size() * 8);
  } else if (k == "configs") {
    if (fetch(c->st.conf, N * S.ne * 3, h)) return -1;
    std::memcpy(out, h.data(), h.size() * 8);
  } else if (k == "a_partial" || k == "b_partial") {  // device (N, ne, M) -> (ne, N, M)
    const bool a = k == "a_partial";
    const int M = a ? S.natom * S.na : S.nb * 2;
    if (fetch(a ? c->st.a_partial : c->st.b_partial, N * S.ne * M, h)) return -1;
    for (int e = 0; e < S.ne; ++e)
      for (size_t w = 0; w < N; ++w)
        for (int m = 0; m < M; ++m) out[((size_t)e * N + w) * M + m] = h[(w * S.ne + e) * M + m];
  } else if (k == "a3_values" || k == "P_i") {  // device (N, ne, M) -> (ne, N, M)
    const bool a = k == "a3_values";
    const int M = a ? S.natom * S.na3 : 1;
    if (fetch(a ? c->st.a3v : c->st.P3, N * S.ne * M, h)) return -1;
    for (int e = 0; e < S.ne; ++e)
      for (size_t w = 0; w < N; ++w)
        for (int m = 0; m < M; ++m) out[((size_t)e * N + w) * M + m] = h[(w * S.ne + e) * M + m];
  } else if (k == "avalues" || k == "bvalues") {
    const bool a = k == "avalues";
    const int M = a ? S.natom * S.na * 2 : S.nb * 3;
    if (fetch(a ? c->st.avalues : c->st.bvalues, N * M, h)) return -1;
    std::memcpy(out, h.data(), h.size() * 8);
  } else
    return fail("unknown state array: " + k);
  return 0;
}

int qmcb_pgradient(qmcb_ctx* c, const char* name, double* out) {
  const std::string k(name);
  if (k == "acoeff") return qmcb_get_state(c, "avalues", out);
  if (k == "bcoeff") return qmcb_get_state(c, "bvalues", out);
  if (k == "ccoeff") {
    Guard g3(c);
    if (c->N == 0) return fail("recompute has not been called");
    if (build_tables(c)) return -1;
    if (!c->have_j3 || c->N == 0) return fail("context has no three-body Jastrow state");
    const Sys& S3 = c->S;
    const size_t nt = (size_t)c->N * S3.natom * S3.na3 * S3.na3;
    DBuf<double> d;
    if (d.ensure(nt * S3.nb3 * 3)) return -1;
    if (prep_kernel(k_jastrow3_pgrad, c->smem_bytes)) return -1;
    k_jastrow3_pgrad<<<(unsigned)((nt + 127) / 128), 128, c->smem_bytes, c->stream>>>(S3, c->st, d.p);
    c->nlaunch++;
    int rc3 = cudaGetLastError() == cudaSuccess ? d2h(c, out, d.p, nt * S3.nb3 * 3 * 8) : fail("k_jastrow3_pgrad launch failed");
    d.release();
    return rc3;
  }
  Guard g(c);
  if (c->N == 0) return fail("recompute has not been called");
  if (build_tables(c)) return -1;
  if (c->N == 0) return fail("system shapes changed: call recompute again");
  if (!c->have_slater) return fail("context has no Slater factor");
  const Sys& S = c->S;
  const size_t N = c->N;
  const int gstride = std::max(std::max(S.nds[0], S.nds[1]), 1);
  DBuf<double> d_det, d_G, d_ao, d_mo;
  int rc = 0;
  if (S.cplx) {  // complex128 outputs (slater.py:462-542 with complex determinants / orbitals)
    do {
      if (d_det.ensure(2 * N * S.ndet) || d_G.ensure(4 * N * gstride)) { rc = -1; break; }
      k_cx_pgrad_det<<<(unsigned)((N + 127) / 128), 128, 0, c->stream>>>(S, c->st, reinterpret_cast<cd*>(d_det.p),
                                                                          reinterpret_cast<cd*>(d_G.p), gstride);
      c->nlaunch++;
      if (cudaGetLastError() != cudaSuccess) { rc = fail("k_cx_pgrad_det launch failed"); break; }
      if (k == "det_coeff") {
        rc = d2h(c, out, d_det.p, 2 * N * S.ndet * 8);
        break;
      }
      int s = -1;
      if (k == "mo_coeff_alpha") s = 0;
      if (k == "mo_coeff_beta") s = 1;
      if (s < 0) { rc = fail("unknown parameter: " + k); break; }
      const size_t nout = N * S.nao * S.nmo_t[s];
      if (nout == 0) break;
      if (d_ao.ensure(N * S.ne * S.nao * (S.pbc ? 2 * S.nk : 1)) || d_mo.ensure(2 * nout)) { rc = -1; break; }
      if (launch_ao_all(c, d_ao.p, c->stream)) { rc = -1; break; }
      k_cx_pgrad_mo<<<(unsigned)((nout + 127) / 128), 128, 0, c->stream>>>(S, c->st, s, d_ao.p, reinterpret_cast<cd*>(d_G.p),
                                                                           gstride, reinterpret_cast<cd*>(d_mo.p));
      c->nlaunch++;
      if (cudaGetLastError() != cudaSuccess) { rc = fail("k_cx_pgrad_mo launch failed"); break; }
      rc = d2h(c, out, d_mo.p, 2 * nout * 8);
    } while (0);
    cudaStreamSynchronize(c->stream);
    d_det.release();
    d_G.release();
    d_ao.release();
    d_mo.release();
    return rc;
  }
  do {
    if (d_det.ensure(N * S.ndet) || d_G.ensure(2 * N * gstride)) { rc = -1; break; }
    k_pgrad_det<<<(unsigned)((N + 127) / 128), 128, 0, c->stream>>>(S, c->st, d_det.p, d_G.p, gstride);
    c->nlaunch++;
    if (cudaGetLastError() != cudaSuccess) { rc = fail("k_pgrad_det launch failed"); break; }
    if (k == "det_coeff") {
      rc = d2h(c, out, d_det.p, N * S.ndet * 8);
      break;
    }
    int s = -1;
    if (k == "mo_coeff_alpha") s = 0;
    if (k == "mo_coeff_beta") s = 1;
    if (s < 0) { rc = fail("unknown parameter: " + k); break; }
    const size_t nout = N * S.nao * S.nmo[s];
    if (nout == 0) break;
    if (d_ao.ensure(N * S.ne * S.nao * (S.pbc ? S.nk : 1)) || d_mo.ensure(nout)) { rc = -1; break; }
    if (launch_ao_all(c, d_ao.p, c->stream)) { rc = -1; break; }
    k_pgrad_mo<<<(unsigned)((nout + 127) / 128), 128, 0, c->stream>>>(S, c->st, s, d_ao.p, d_G.p, gstride, d_mo.p);
    c->nlaunch++;
    if (cudaGetLastError() != cudaSuccess) { rc = fail("k_pgrad_mo launch failed"); break; }
    rc = d2h(c, out, d_mo.p, nout * 8);
  } while (0);
  cudaStreamSynchronize(c->stream);
  d_det.release();
  d_G.release();
  d_ao.release();
  d_mo.release();
  return rc;
}

// ---------------------------------------------------------------------------------------
int qmcb_energy(qmcb_ctx* c, const double* ecp_u, const double* ecp_rot, double* out) {
  Guard g(c);
  if (c->N == 0) return fail("recompute has not been called");
  if (build_tables(c)) return -1;
  if (c->N == 0) return fail("system shapes changed: call recompute again");
  const Sys& S = c->S;
  const size_t N = c->N;
  if (ensure_energy_scratch(c) || energy_scratch_points(c)) return -1;
  const size_t nu = (size_t)S.ne * S.necp * N, nr = (size_t)S.ne * S.necp * 9;
  const size_t rows = S.cplx ? 8 : 6;  // complex wave functions: + Im ecp, Im total
  if (c->d_u.ensure(nu) || c->d_rot.ensure(nr) || c->d_energy.ensure(rows * N)) return -1;
  if (S.necp > 0) {
    if (!ecp_u || !ecp_rot) return fail("ECP random variates missing");
    if (h2d(c, c->d_u.p, ecp_u, nu * 8) || h2d(c, c->d_rot.p, ecp_rot, nr * 8)) return -1;
  }
  if (launch_energy(c, c->d_u.p, c->d_rot.p, c->d_energy.p, c->stream)) return -1;
  return d2h(c, out, c->d_energy.p, rows * N * 8);
}

int qmcb_tmoves(qmcb_ctx* c, int e, double tau, const double* ecp_u, const double* ecp_rot, double* ratio,
                double* weight, double* epos) {
  Guard g(c);
  if (c->N == 0) return fail("recompute has not been called");
  if (build_tables(c)) return -1;
  if (c->N == 0) return fail("system shapes changed: call recompute again");
  const Sys& S = c->S;
  if (e < 0 || e >= S.ne) return fail("electron index out of range");
  if (!(tau > 0.0)) return fail("T-moves need tau > 0");
  const size_t N = c->N;
  if (S.necp == 0) return 0;  // ratio/weight have zero columns
  if (ensure_energy_scratch(c)) return -1;
  const size_t npts = N * S.necp * S.max_naip;
  if (ensure_scratch(c, npts, 1)) return -1;
  const size_t M = (size_t)S.tot_naip;
  if (c->d_u.ensure((size_t)S.necp * N) || c->d_rot.ensure((size_t)S.necp * 9) || c->d_out.ensure(N * M * 6)) return -1;
  if (h2d(c, c->d_u.p, ecp_u, (size_t)S.necp * N * 8) || h2d(c, c->d_rot.p, ecp_rot, (size_t)S.necp * 9 * 8)) return -1;
  double* d_ratio = c->d_out.p;
  double* d_weight = d_ratio + N * M;
  double* d_pos = d_weight + N * M;
  k_tmove_init<<<(unsigned)((N * M + 255) / 256), 256, 0, c->stream>>>(S, c->st, e, d_ratio, d_weight, d_pos);
  c->nlaunch++;
  CK(cudaGetLastError());
  double* d_ratio_im = d_pos + N * M * 3;  // complex wave functions: Im ratio (0 for the masked-out walkers)
  if (S.cplx) CK(cudaMemsetAsync(d_ratio_im, 0, N * M * 8, c->stream));
  CK(cudaMemsetAsync(c->es.count, 0, sizeof(int), c->stream));
  const long long nt = (long long)N * S.necp;
  if (prep_kernel(k_ecp_prepare, c->smem_bytes)) return -1;
  k_ecp_prepare<<<(unsigned)((nt + 127) / 128), 128, c->smem_bytes, c->stream>>>(S, c->st, c->es, c->d_u.p, e);
  c->nlaunch++;
  CK(cudaGetLastError());
  EcpPointArgs ea{};
  ea.rot = c->d_rot.p;
  ea.quad = c->d_quad.p;
  ea.e_only = e;
  ea.tmove_tau = tau;
  ea.tm_ratio = d_ratio;
  ea.tm_weight = d_weight;
  ea.tm_pos = d_pos;
  ea.scr = c->d_scr.p;
  ea.scr_stride = npts;
  const long long grid = std::min<long long>(((long long)npts + 127) / 128, 148LL * 16);
  int rc = 0;
  const size_t sm = c->smem_bytes;
  if (S.cplx) {
    EcpCxArgs cxa{nullptr, d_ratio_im};
    rc = prep_kernel(k_cx_ecp_points, sm);
    if (!rc && S.pbc) rc = ecp_points_pbc_prepass<0>(c, ea, (long long)npts, grid, c->stream);
    if (!rc) k_cx_ecp_points<<<(unsigned)grid, 128, sm, c->stream>>>(S, c->st, c->es, ea, cxa);
  } else if (c->nmot == 4) {
    rc = prep_kernel(k_ecp_points<4>, sm);
    if (!rc) k_ecp_points<4><<<(unsigned)grid, 128, sm, c->stream>>>(S, c->st, c->es, ea);
  } else if (c->nmot == 8) {
    rc = prep_kernel(k_ecp_points<8>, sm);
    if (!rc) k_ecp_points<8><<<(unsigned)grid, 128, sm, c->stream>>>(S, c->st, c->es, ea);
  } else {
    rc = prep_kernel(k_ecp_points<0>, sm);
    if (!rc && S.pbc) rc = ecp_points_pbc_prepass<0>(c, ea, (long long)npts, grid, c->stream);
    if (!rc) k_ecp_points<0><<<(unsigned)grid, 128, sm, c->stream>>>(S, c->st, c->es, ea);
  }
  if (rc) return rc;
  c->nlaunch++;
  CK(cudaGetLastError());
  std::vector<double> h(N * M * 6);
  if (d2h(c, h.data(), c->d_out.p, h.size() * 8)) return -1;
  if (S.cplx) {  // `ratio` is complex128
    for (size_t i = 0; i < N * M; ++i) {
      ratio[2 * i] = h[i];
      ratio[2 * i + 1] = h[N * M * 5 + i];
    }
  } else
    std::memcpy(ratio, h.data(), N * M * 8);
  std::memcpy(weight, h.data() + N * M, N * M * 8);
  std::memcpy(epos, h.data() + 2 * N * M, N * M * 3 * 8);
  return 0;
}

}  // extern "C"
// ---------------------------------------------------------------------------------------
template <int NMOT>
static int launch_move_t(qmcb_ctx* c, const MoveArgs& ma, cudaStream_t stream) {
  const int block = pick_block(c->N);
  if (prep_kernel(k_vmc_move<NMOT>, c->smem_bytes)) return -1;
  k_vmc_move<NMOT><<<(c->N + block - 1) / block, block, c->smem_bytes, stream>>>(c->S, c->st, ma);
  c->nlaunch++;
  CK(cudaGetLastError());
  return 0;
}

extern "C" {

int qmcb_vmc_block_device(qmcb_ctx* c, int nsteps, double tstep, int with_energy, const double* d_gauss,
                          const double* d_unif, const double* d_ecp_u, const double* d_ecp_rot, uint8_t* d_accept,
                          double* d_energy, double* d_esum, int64_t* d_nacc, void* stream_) {
  Guard g(c);
  if (c->N == 0) return fail("recompute has not been called");
  if (build_tables(c)) return -1;
  if (c->N == 0) return fail("system shapes changed: call recompute again");
  const Sys& S = c->S;
  const size_t N = c->N;
  cudaStream_t stream = stream_ ? (cudaStream_t)stream_ : c->stream;
  // complex wave functions: the query kernels of cplx.cuh chained on the device (k_cx_chain), 8 energy rows per step
  const bool chain = S.cplx != 0;
  const size_t rows = S.cplx ? 8 : 6;
  const int which = (c->have_slater ? 1 : 0) | (c->have_jastrow ? 2 : 0) | (c->have_j3 ? 4 : 0);
  if (with_energy && (ensure_energy_scratch(c) || energy_scratch_points(c))) return -1;
  if (ensure_scratch(c, N, 5)) return -1;
  if (c->d_accept.ensure(N) || c->d_nacc.ensure((size_t)nsteps * S.ne)) return -1;
  if (!d_energy && with_energy) {
    if (c->d_energy.ensure(rows * N)) return -1;
  }
  unsigned long long* nacc = d_nacc ? (unsigned long long*)d_nacc : c->d_nacc.p;
  CK(cudaMemsetAsync(nacc, 0, (size_t)nsteps * S.ne * 8, stream));
  // warp-per-walker sweep kernel: single determinant, inverse staged in shared memory
  const CoopLayout CL = coop_layout(S);
  const size_t tab = (c->smem_bytes + 15) & ~(size_t)15;
  int G = 16;
  if (const char* env = std::getenv("QMCB_SWEEP_G")) G = std::atoi(env);
  if (G != 8 && G != 16 && G != 32) G = 16;
  int sweep_warps = 4;
  if (const char* env = std::getenv("QMCB_SWEEP_WARPS")) sweep_warps = std::max(1, std::min(4, std::atoi(env)));
  const int sweep_walkers = sweep_warps * (32 / G);
  const size_t sweep_smem = tab + (size_t)sweep_walkers * CL.total * 8;
  const bool use_sweep = (!c->have_slater || S.ndet == 1) && !c->have_j3 && sweep_smem <= 200 * 1024 && !S.pbc && !chain &&
                         std::getenv("QMCB_NO_SWEEP") == nullptr;
  const bool use_pbc = S.pbc != 0 && !chain;
  if (chain) {
    if (c->b_gold.ensure(N * 3) || c->d_in.ensure(N * 3) || c->d_pwrap.ensure(N * 3) || c->d_out.ensure(16 * N)) return -1;
    c->st.gold = c->b_gold.p;
  }
  // periodic multi-determinant and / or three-body wave functions: k_pbc_move_general around the orbital kernel + launch_update
  const bool pbc_general = use_pbc && ((c->have_slater && S.ndet != 1) || c->have_j3);
  if (use_pbc && c->have_slater && !c->mocache_valid)
    if (launch_mo_all(c, 0, stream)) return -1;
  if (use_sweep && c->have_slater && !c->mocache_valid) {
    const long long np = (long long)N * S.ne;
    const int block = pick_block(np);
    if (prep_kernel(k_mo_all, c->smem_bytes)) return -1;
    k_mo_all<<<(unsigned)((np + block - 1) / block), block, c->smem_bytes, stream>>>(S, c->st, 0);
    c->nlaunch++;
    CK(cudaGetLastError());
    c->mocache_valid = true;
  }
  if (use_sweep && c->have_jastrow && !c->paircache_valid) {  // pair caches, from the current positions
    const long long nt = (long long)N * (S.npair + S.ne);
    if (prep_kernel(k_pair_cache_build, c->smem_bytes)) return -1;
    k_pair_cache_build<<<(unsigned)((nt + 127) / 128), 128, c->smem_bytes, stream>>>(S, c->st);
    c->nlaunch++;
    CK(cudaGetLastError());
    c->paircache_valid = true;
  }
  // Overlap (single-determinant fast path): the energy kernels read the inverse, the coordinates and the Jastrow
  // partial sums only (slater_point_fast / jastrow_point), ~1.7 KB per walker.  After the sweep of step s those
  // arrays are copied aside (device-to-device, a few microseconds) and the accumulator of step s runs from the
  // copy on the energy stream while the sweep of step s+1 proceeds; the kinetic pieces alternate between two buffers.
  const bool overlap = use_sweep && with_energy && c->have_slater && c->have_jastrow && (c->nmot == 4 || c->nmot == 8) &&
                       nsteps > 1 && std::getenv("QMCB_NO_ENERGY_OVERLAP") == nullptr;
  // Periodic chain: the same overlap with the kinetic pieces computed on the main stream first (they read the cached
  // MO rows, 80 KB per walker at C4 -- not worth copying); the ECP, Ewald and finalize kernels (3 of the 8 ms of a C4
  // step) then run from the copy of inverse / coordinates / wrap / Jastrow partial sums while the next step's moves,
  // which leave most of the machine idle at 1024 walkers, proceed.
  const bool overlap_pbc = use_pbc && !pbc_general && with_energy && c->have_slater && c->have_jastrow && nsteps > 1 &&
                           std::getenv("QMCB_NO_ENERGY_OVERLAP") == nullptr;
  State snap = c->st;
  EnergyScratch es_alt[2] = {c->es, c->es};
  if (overlap_pbc) {
    if (c->sn_wrap.ensure(N * S.ne * 3)) return -1;
    snap.wrap = c->sn_wrap.p;
  }
  if (overlap || overlap_pbc) {
    const size_t ninv0 = N * (size_t)S.nup * S.nup, ninv1 = N * (size_t)S.ndn * S.ndn;
    const size_t nap = N * (size_t)S.ne * S.natom * S.na, nbp = N * (size_t)S.ne * S.nb * 2;
    if (c->sn_inv[0].ensure(ninv0) || c->sn_inv[1].ensure(ninv1) || c->sn_conf.ensure(N * S.ne * 3) || c->sn_ap.ensure(nap) ||
        c->sn_bp.ensure(nbp) || c->e_ke2.ensure(S.ne * N) || c->e_g22.ensure(S.ne * N))
      return -1;
    snap.inv[0] = c->sn_inv[0].p;
    snap.inv[1] = c->sn_inv[1].p;
    snap.conf = c->sn_conf.p;
    snap.a_partial = c->sn_ap.p;
    snap.b_partial = c->sn_bp.p;
    es_alt[1].ke_e = c->e_ke2.p;
    es_alt[1].g2_e = c->e_g22.p;
  }
  for (int step = 0; step < nsteps; ++step) {
    if (use_sweep) {
      const size_t se = (size_t)step * S.ne;
      SweepArgs sa{};
      sa.tstep = tstep;
      sa.gauss = d_gauss + se * N * 3;
      sa.unif = d_unif + se * N;
      sa.accept = d_accept ? d_accept + se * N : nullptr;
      sa.nacc = nacc + se;
      if (with_energy) {  // kinetic pieces of the final positions straight from the sweep's caches
        const EnergyScratch& esw = overlap ? es_alt[step & 1] : c->es;
        sa.ke_e = esw.ke_e;
        sa.g2_e = esw.g2_e;
        c->kinetic_valid = true;
      }
      const unsigned grid = (unsigned)((N + sweep_walkers - 1) / sweep_walkers);
      if (G == 8) {
        if (prep_kernel(k_vmc_sweep<8, false>, sweep_smem)) return -1;
        k_vmc_sweep<8, false><<<grid, sweep_warps * 32, sweep_smem, stream>>>(S, c->st, sa);
      } else if (G == 16) {
        if (prep_kernel(k_vmc_sweep<16, false>, sweep_smem)) return -1;
        k_vmc_sweep<16, false><<<grid, sweep_warps * 32, sweep_smem, stream>>>(S, c->st, sa);
      } else {
        if (prep_kernel(k_vmc_sweep<32, false>, sweep_smem)) return -1;
        k_vmc_sweep<32, false><<<grid, sweep_warps * 32, sweep_smem, stream>>>(S, c->st, sa);
      }
      c->nlaunch++;
      CK(cudaGetLastError());
    }
    // Periodic move chain.  Fused flavour (n <= 32 per spin): [propose(0)] then per electron the orbitals at the
    // proposed points and ONE warp-per-walker kernel doing Metropolis test, cache updates, the Sherman-Morrison
    // update and the proposal of electron e + 1.  QMCB_PBC_UNFUSED=1 keeps the four-launch chain (A/B checks).
    const bool fuse_pbc = use_pbc && !pbc_general && (!c->have_slater || (S.nup <= 32 && S.ndn <= 32)) &&
                          std::getenv("QMCB_PBC_UNFUSED") == nullptr;
    State stf = c->st;  // fused chain: the proposal keeps the pair values at the old position for the cache update
    // non-null = the fused chain keeps the pair caches (BPAIR / GPAIR / AGRAD) current and takes the drift at an
    // electron's current position from them; QMCB_PBC_NO_PAIRCACHE=1 recomputes it (A/B checks)
    if (fuse_pbc && c->have_jastrow && S.nb > 0 && std::getenv("QMCB_PBC_NO_PAIRCACHE") == nullptr) stf.jold = c->b_jold.p;
    // Both kernels of the fused chain leave most of the machine idle at ~1000 walkers per GPU (one warp per walker:
    // 7 warps per SM; orbital kernel: one CTA per point, under two waves), and an electron move is their dependent
    // sequence.  The walkers are independent of each other, so the ensemble is cut into ranges whose chains run
    // concurrently on separate streams: the orbital kernel of one range overlaps the accept kernel of another.
    int nsplit = 1;
    if (fuse_pbc && c->have_slater) {
      nsplit = N >= 512 ? 2 : 1;
      if (const char* env = std::getenv("QMCB_PBC_SPLIT")) nsplit = std::atoi(env);
      nsplit = std::max(1, std::min(4, nsplit));
      if ((size_t)nsplit > N) nsplit = 1;
    }
    if (stf.jold != nullptr && step == 0) {
      // pair caches of the fused chain (b_l per pair, pair-gradient terms, electron-ion gradients) from the current
      // walkers; the accept kernel keeps them current from here on
      const long long nt = (long long)N * (S.npair + S.ne);
      if (prep_kernel(k_pair_cache_build, c->smem_bytes)) return -1;
      k_pair_cache_build<<<(unsigned)((nt + 127) / 128), 128, c->smem_bytes, stream>>>(S, c->st);
      c->nlaunch++;
      CK(cudaGetLastError());
    }
    if (nsplit > 1) {
      if (!c->ev_fork) CK(cudaEventCreateWithFlags(&c->ev_fork, cudaEventDisableTiming));
      for (int k = 0; k < nsplit - 1; ++k) {
        if (!c->split_stream[k]) {
          int lo = 0, hi = 0;
          CK(cudaDeviceGetStreamPriorityRange(&lo, &hi));
          CK(cudaStreamCreateWithPriority(&c->split_stream[k], cudaStreamNonBlocking, lo > hi ? lo - 1 : lo));
          CK(cudaEventCreateWithFlags(&c->ev_join[k], cudaEventDisableTiming));
        }
      }
      CK(cudaEventRecord(c->ev_fork, stream));
      for (int k = 0; k < nsplit - 1; ++k) CK(cudaStreamWaitEvent(c->split_stream[k], c->ev_fork, 0));
    }
    if (nsplit > 1) {
      const int ldmax = std::max(S.ldc[0], S.ldc[1]);
      const int jper = pbc_accept_jper(S);
      const size_t asm_ = ((c->smem_bytes + 15) & ~(size_t)15) + (size_t)4 * jper * 8;
      if (prep_kernel(k_pbc_propose, c->smem_bytes) || prep_kernel(k_pbc_accept<true>, asm_)) return -1;
      // issue order: electron outer, range inner, so that every stream has work queued from the start
      for (int e = 0; e < S.ne; ++e) {
        const size_t se = (size_t)step * S.ne + e;
        const int s = e >= S.nup ? 1 : 0;
        for (int part = 0; part < nsplit; ++part) {
          const size_t w0 = N * part / nsplit, wn = N * (part + 1) / nsplit - w0;
          cudaStream_t ps = part == 0 ? stream : c->split_stream[part - 1];
          const unsigned wgrid = (unsigned)((wn * 32 + 127) / 128);
          PbcMoveArgs ma{};
          ma.e = e;
          ma.tstep = tstep;
          ma.gauss = d_gauss + se * N * 3;
          ma.unif = d_unif + se * N;
          ma.accept = d_accept ? d_accept + se * N : c->d_accept.p;
          ma.nacc = nacc + se;
          ma.gauss_next = e + 1 < S.ne ? d_gauss + (se + 1) * N * 3 : nullptr;
          ma.w0 = (int)w0;
          ma.wn = (int)wn;
          if (e == 0) {
            k_pbc_propose<<<wgrid, 128, c->smem_bytes, ps>>>(S, stf, ma);
            c->nlaunch++;
            CK(cudaGetLastError());
          }
          PbcMoArgs a{};
          a.npoints = (long long)wn;
          a.pos = c->st.saved_pos + w0 * 3;
          a.wrap = c->st.saved_wrap + w0 * 3;
          a.naip = 1;
          a.spin_mode = 0;
          a.spin = s;
          a.out = c->st.monew + w0 * 5 * ldmax;
          a.stride_p = 5 * ldmax;
          a.stride_c = ldmax;
          a.stride_j = 1;
          if (launch_pbc_mo(c, 2, a, (long long)wn, ps)) return -1;
          k_pbc_accept<true><<<wgrid, 128, asm_, ps>>>(S, stf, ma);
          c->nlaunch++;
          CK(cudaGetLastError());
        }
      }
      for (int part = 1; part < nsplit; ++part) {
        CK(cudaEventRecord(c->ev_join[part - 1], c->split_stream[part - 1]));
        CK(cudaStreamWaitEvent(stream, c->ev_join[part - 1], 0));
      }
      c->paircache_valid = false;
    }
    for (int e = 0; e < S.ne && pbc_general; ++e) {
      const size_t se = (size_t)step * S.ne + e;
      const int s = e >= S.nup ? 1 : 0;
      const int ldmax = std::max(S.ldc[0], S.ldc[1]);
      MoveArgs ma{};
      ma.e = e;
      ma.tstep = tstep;
      ma.gauss = d_gauss + se * N * 3;
      ma.unif = d_unif + se * N;
      ma.accept = d_accept ? d_accept + se * N : c->d_accept.p;
      ma.nacc = nacc + se;
      constexpr int GM = 16;
      const size_t msm = tab + (size_t)(128 / GM) * j3_scratch_doubles(S) * 8;
      const unsigned mgrid = (unsigned)(((long long)N * GM + 127) / 128);
      if (prep_kernel(k_pbc_move_general<GM>, msm)) return -1;
      k_pbc_move_general<GM><<<mgrid, 128, msm, stream>>>(S, c->st, ma, 0);
      c->nlaunch++;
      CK(cudaGetLastError());
      if (c->have_slater) {
        PbcMoArgs a{};
        a.npoints = (long long)N;
        a.pos = c->st.saved_pos;
        a.wrap = c->st.saved_wrap;
        a.naip = 1;
        a.spin_mode = 0;
        a.spin = s;
        a.out = c->st.monew;
        a.stride_p = 5 * ldmax;
        a.stride_c = ldmax;
        a.stride_j = 1;
        if (launch_pbc_mo(c, 2, a, (long long)N, stream)) return -1;
      }
      k_pbc_move_general<GM><<<mgrid, 128, msm, stream>>>(S, c->st, ma, 1);
      c->nlaunch++;
      CK(cudaGetLastError());
      if (launch_update(c, which, e, ma.accept, stream)) return -1;
      c->paircache_valid = false;  // the cached MO rows stay valid: accepted walkers refreshed theirs
    }
    for (int e = 0; e < S.ne && use_pbc && !pbc_general && nsplit == 1; ++e) {
      const size_t se = (size_t)step * S.ne + e;
      const int s = e >= S.nup ? 1 : 0;
      const int ldmax = std::max(S.ldc[0], S.ldc[1]);
      PbcMoveArgs ma{};
      ma.e = e;
      ma.tstep = tstep;
      ma.gauss = d_gauss + se * N * 3;
      ma.unif = d_unif + se * N;
      ma.accept = d_accept ? d_accept + se * N : c->d_accept.p;
      ma.nacc = nacc + se;
      ma.gauss_next = (fuse_pbc && e + 1 < S.ne) ? d_gauss + (se + 1) * N * 3 : nullptr;
      const unsigned wgrid = (unsigned)((N * 32 + 127) / 128);
      if (!fuse_pbc || e == 0) {
        if (prep_kernel(k_pbc_propose, c->smem_bytes)) return -1;
        k_pbc_propose<<<wgrid, 128, c->smem_bytes, stream>>>(S, stf, ma);
        c->nlaunch++;
        CK(cudaGetLastError());
      }
      if (c->have_slater) {
        PbcMoArgs a{};
        a.npoints = (long long)N;
        a.pos = c->st.saved_pos;
        a.wrap = c->st.saved_wrap;
        a.naip = 1;
        a.spin_mode = 0;
        a.spin = s;
        a.out = c->st.monew;
        a.stride_p = 5 * ldmax;
        a.stride_c = ldmax;
        a.stride_j = 1;
        if (launch_pbc_mo(c, 2, a, (long long)N, stream)) return -1;
      }
      const int jper = pbc_accept_jper(S);
      const size_t asm_ = ((c->smem_bytes + 15) & ~(size_t)15) + (size_t)4 * jper * 8;
      if (fuse_pbc) {
        if (prep_kernel(k_pbc_accept<true>, asm_)) return -1;
        k_pbc_accept<true><<<wgrid, 128, asm_, stream>>>(S, stf, ma);
      } else {
        if (prep_kernel(k_pbc_accept<false>, asm_)) return -1;
        k_pbc_accept<false><<<wgrid, 128, asm_, stream>>>(S, c->st, ma);
      }
      c->nlaunch++;
      CK(cudaGetLastError());
      if (c->have_slater && !fuse_pbc) {
        SmArgs sa{};
        sa.n = s ? S.ndn : S.nup;
        sa.e = e - s * S.nup;
        sa.nds = 1;
        sa.vec_stride = 5 * ldmax;
        sa.nmat = (long long)N;
        sa.inv = c->st.inv[s];
        sa.vec = c->st.monew;
        sa.occ = S.iblob + S.o_occ[s];
        sa.mask = ma.accept;
        sa.dsign = c->st.dsign[s];
        sa.dlog = c->st.dlog[s];
        if (launch_sm(c, sa, stream, &c->nlaunch)) return -1;
      }
      c->paircache_valid = false;
    }
    for (int e = 0; e < S.ne && chain; ++e) {
      const size_t se = (size_t)step * S.ne + e;
      const unsigned cgrid = (unsigned)((N + 127) / 128);
      CxChainArgs ca{};
      ca.e = e;
      ca.tstep = tstep;
      ca.gauss = d_gauss + se * N * 3;
      ca.unif = d_unif + se * N;
      ca.accept = d_accept ? d_accept + se * N : c->d_accept.p;
      ca.nacc = nacc + se;
      ca.pos = c->d_in.p;
      ca.pwrap = c->d_pwrap.p;
      ca.grad = reinterpret_cast<const cd*>(c->d_out.p);
      ca.val = reinterpret_cast<const cd*>(c->d_out.p + 6 * N);
      PointArgs pa{};
      pa.which = which;
      pa.e = e;
      pa.naip = 1;
      pa.pos = c->d_in.p;
      pa.npoints = (int)N;
      pa.o_grad = c->d_out.p;
      pa.o_val = c->d_out.p + 6 * N;
      pa.o_lap = c->d_out.p + 6 * N;
      pa.scr = c->d_scr.p;
      pa.scr_stride = N;
      k_cx_chain<0><<<cgrid, 128, c->smem_bytes, stream>>>(S, c->st, ca);  // positions (and wrap vectors) of electron e
      c->nlaunch++;
      CK(cudaGetLastError());
      if (S.pbc && pbc_point_rows(c, 1, e, c->d_in.p, c->d_pwrap.p, nullptr, 1, (long long)N, N, stream)) return -1;
      pa.save = 0;
      if (launch_point<PV_GRADVAL>(c, pa, stream)) return -1;
      ca.pwrap = c->st.saved_wrap;
      k_cx_chain<1><<<cgrid, 128, c->smem_bytes, stream>>>(S, c->st, ca);  // limited drift, (wrapped) proposal
      c->nlaunch++;
      CK(cudaGetLastError());
      if (S.pbc && pbc_point_rows(c, 1, e, c->d_in.p, c->st.saved_wrap, nullptr, 1, (long long)N, N, stream)) return -1;
      pa.save = 1;
      if (launch_point<PV_GRADVAL>(c, pa, stream)) return -1;
      k_cx_chain<2><<<cgrid, 128, c->smem_bytes, stream>>>(S, c->st, ca);  // Metropolis test
      c->nlaunch++;
      CK(cudaGetLastError());
      if (launch_update(c, which, e, ca.accept, stream)) return -1;
      c->mocache_valid = false;
      c->paircache_valid = false;
    }
    for (int e = 0; e < S.ne && !use_sweep && !use_pbc && !chain; ++e) {
      const size_t se = (size_t)step * S.ne + e;
      MoveArgs ma{};
      ma.e = e;
      ma.tstep = tstep;
      ma.gauss = d_gauss + se * N * 3;
      ma.unif = d_unif + se * N;
      ma.accept = d_accept ? d_accept + se * N : c->d_accept.p;
      ma.nacc = nacc + se;
      ma.scr = c->d_scr.p;
      ma.scr_stride = N;
      // general wave functions (multi-determinant and / or three-body): G lanes per walker, cached MO rows
      constexpr int GM = 16;
      const size_t msm = tab + (size_t)(128 / GM) * (CL.total + j3_scratch_doubles(S)) * 8;
      if (msm <= 200 * 1024 && std::getenv("QMCB_NO_COOP_MOVE") == nullptr) {
        if (c->have_slater && !c->mocache_valid)
          if (launch_mo_all(c, 0, stream)) return -1;
        if (prep_kernel(k_vmc_move_coop<GM>, msm)) return -1;
        k_vmc_move_coop<GM><<<(unsigned)(((long long)N * GM + 127) / 128), 128, msm, stream>>>(S, c->st, ma);
        c->nlaunch++;
        CK(cudaGetLastError());
        if (launch_update(c, which, e, ma.accept, stream)) return -1;
        c->paircache_valid = false;  // the cached MO rows stay valid: accepted walkers refreshed theirs
        continue;
      }
      int rc = c->nmot == 4 ? launch_move_t<4>(c, ma, stream) : (c->nmot == 8 ? launch_move_t<8>(c, ma, stream) : launch_move_t<0>(c, ma, stream));
      if (rc) return rc;
      if (launch_update(c, which, e, ma.accept, stream)) return -1;
      c->mocache_valid = false;
      c->paircache_valid = false;
    }
    if (with_energy) {
      double* eo = d_energy ? d_energy + (size_t)step * rows * N : c->d_energy.p;
      const size_t ue = (size_t)step * S.ne * S.necp;
      const double* su = d_ecp_u ? d_ecp_u + ue * N : nullptr;
      const double* sr = d_ecp_rot ? d_ecp_rot + ue * 9 : nullptr;
      if (overlap || overlap_pbc) {
        cudaStream_t es_ = c->energy_stream;
        if (overlap_pbc) {
          if (launch_kinetic(c, c->st, es_alt[step & 1], stream)) return -1;
          c->kinetic_valid = true;
        }
        if (step > 0) CK(cudaStreamWaitEvent(stream, c->ev_edone, 0));  // the previous accumulator still reads the copy
        if (overlap_pbc) CK(cudaMemcpyAsync(snap.wrap, c->st.wrap, N * S.ne * 3 * 8, cudaMemcpyDeviceToDevice, stream));
        CK(cudaMemcpyAsync(snap.inv[0], c->st.inv[0], N * (size_t)S.nup * S.nup * 8, cudaMemcpyDeviceToDevice, stream));
        CK(cudaMemcpyAsync(snap.inv[1], c->st.inv[1], N * (size_t)S.ndn * S.ndn * 8, cudaMemcpyDeviceToDevice, stream));
        CK(cudaMemcpyAsync(snap.conf, c->st.conf, N * S.ne * 3 * 8, cudaMemcpyDeviceToDevice, stream));
        CK(cudaMemcpyAsync(snap.a_partial, c->st.a_partial, N * (size_t)S.ne * S.natom * S.na * 8, cudaMemcpyDeviceToDevice, stream));
        CK(cudaMemcpyAsync(snap.b_partial, c->st.b_partial, N * (size_t)S.ne * S.nb * 2 * 8, cudaMemcpyDeviceToDevice, stream));
        CK(cudaEventRecord(c->ev_snap, stream));
        CK(cudaStreamWaitEvent(es_, c->ev_snap, 0));
        if (overlap_pbc) {
          static const unsigned cap = std::getenv("QMCB_PBC_BG_GRID") ? (unsigned)std::atoi(std::getenv("QMCB_PBC_BG_GRID")) : 592u;
          c->pbc_mo_grid_cap = cap;
        }
        const int erc = launch_energy_on(c, snap, es_alt[step & 1], su, sr, eo, es_);
        c->pbc_mo_grid_cap = 0;
        if (erc) return -1;
        if (d_esum) {
          k_colsum<<<(unsigned)rows, 256, 0, es_>>>(eo, (int)N, d_esum + (size_t)step * rows);
          c->nlaunch++;
          CK(cudaGetLastError());
        }
        CK(cudaEventRecord(c->ev_edone, es_));
        c->kinetic_valid = false;
      } else {
        if (launch_energy(c, su, sr, eo, stream)) return -1;
        if (d_esum) {
          k_colsum<<<(unsigned)rows, 256, 0, stream>>>(eo, (int)N, d_esum + (size_t)step * rows);
          c->nlaunch++;
          CK(cudaGetLastError());
        }
      }
    }
  }
  if (overlap || overlap_pbc) CK(cudaStreamWaitEvent(stream, c->ev_edone, 0));  // the caller's stream sees every step's energies
  c->saved_slot = -1;
  return 0;
}

int qmcb_vmc_block(qmcb_ctx* c, int nsteps, double tstep, int with_energy, const double* gauss, const double* unif,
                   const double* ecp_u, const double* ecp_rot, double* configs, uint8_t* accept, double* energy,
                   double* esum, int64_t* nacc) {
  Guard g(c);
  if (c->N == 0) return fail("recompute has not been called");
  if (build_tables(c)) return -1;
  if (c->N == 0) return fail("system shapes changed: call recompute again");
  const Sys& S = c->S;
  const size_t N = c->N;
  const size_t nse = (size_t)nsteps * S.ne;
  if (c->d_gauss.ensure(nse * N * 3) || c->d_unif.ensure(nse * N)) return -1;
  CK(cudaMemcpyAsync(c->d_gauss.p, gauss, nse * N * 3 * 8, cudaMemcpyHostToDevice, c->stream));
  CK(cudaMemcpyAsync(c->d_unif.p, unif, nse * N * 8, cudaMemcpyHostToDevice, c->stream));
  const size_t nu = nse * S.necp * N, nr = nse * S.necp * 9;
  if (with_energy && S.necp > 0) {
    if (!ecp_u || !ecp_rot) return fail("ECP random variates missing");
    if (c->d_u.ensure(nu) || c->d_rot.ensure(nr)) return -1;
    CK(cudaMemcpyAsync(c->d_u.p, ecp_u, nu * 8, cudaMemcpyHostToDevice, c->stream));
    CK(cudaMemcpyAsync(c->d_rot.p, ecp_rot, nr * 8, cudaMemcpyHostToDevice, c->stream));
  }
  DBuf<uint8_t> acc_all;
  if (accept && acc_all.ensure(nse * N)) return -1;
  const size_t rows = S.cplx ? 8 : 6;  // complex wave functions: + Im ecp, Im total per step
  if (with_energy && (c->d_energy.ensure((size_t)nsteps * rows * N) || c->d_esum.ensure((size_t)nsteps * rows))) return -1;
  int rc = qmcb_vmc_block_device(c, nsteps, tstep, with_energy, c->d_gauss.p, c->d_unif.p, c->d_u.p, c->d_rot.p,
                                 accept ? acc_all.p : nullptr, with_energy ? c->d_energy.p : nullptr,
                                 with_energy ? c->d_esum.p : nullptr, nullptr, c->stream);
  if (rc) {
    acc_all.release();
    return rc;
  }
  if (accept) CK(cudaMemcpyAsync(accept, acc_all.p, nse * N, cudaMemcpyDeviceToHost, c->stream));
  if (energy && with_energy)
    CK(cudaMemcpyAsync(energy, c->d_energy.p, (size_t)nsteps * rows * N * 8, cudaMemcpyDeviceToHost, c->stream));
  if (esum && with_energy) CK(cudaMemcpyAsync(esum, c->d_esum.p, (size_t)nsteps * rows * 8, cudaMemcpyDeviceToHost, c->stream));
  if (nacc) CK(cudaMemcpyAsync(nacc, c->d_nacc.p, nse * 8, cudaMemcpyDeviceToHost, c->stream));
  if (configs) {
    const size_t nel = N * S.ne * 3;
    if (c->d_in.ensure(nel)) return -1;
    k_conf_out<<<(unsigned)((nel + 255) / 256), 256, 0, c->stream>>>(c->st.conf, c->d_in.p, (int)N, S.ne);
    c->nlaunch++;
    CK(cudaGetLastError());
    CK(cudaMemcpyAsync(configs, c->d_in.p, nel * 8, cudaMemcpyDeviceToHost, c->stream));
  }
  if (sync_blocking(c)) return -1;
  acc_all.release();
  return 0;
}

// Asynchronous upload of one block's variates (pinned host buffers) into device slot 0/1 on the
// context's copy stream.  May be called from the host thread that draws the variates while another
// thread runs qmcb_vmc_block_slot on the other slot.
int qmcb_vmc_upload(qmcb_ctx* c, int slot, int nsteps, int ne, int64_t N, int necp, const double* gauss,
                    const double* unif, const double* ecp_u, const double* ecp_rot) {
  if (slot < 0 || slot >= qmcb_ctx::NSLOT) return fail("slot out of range");
  cudaSetDevice(c->device);
  const size_t nse = (size_t)nsteps * ne;
  if (c->s_gauss[slot].ensure(nse * N * 3) || c->s_unif[slot].ensure(nse * N)) return -1;
  CK(cudaMemcpyAsync(c->s_gauss[slot].p, gauss, nse * N * 3 * 8, cudaMemcpyHostToDevice, c->copy_stream));
  CK(cudaMemcpyAsync(c->s_unif[slot].p, unif, nse * N * 8, cudaMemcpyHostToDevice, c->copy_stream));
  if (ecp_u && necp > 0) {
    if (c->s_u[slot].ensure(nse * necp * N) || c->s_rot[slot].ensure(nse * necp * 9)) return -1;
    CK(cudaMemcpyAsync(c->s_u[slot].p, ecp_u, nse * necp * N * 8, cudaMemcpyHostToDevice, c->copy_stream));
    CK(cudaMemcpyAsync(c->s_rot[slot].p, ecp_rot, nse * necp * 9 * 8, cudaMemcpyHostToDevice, c->copy_stream));
  }
  CK(cudaEventRecord(c->slot_ready[slot], c->copy_stream));
  return 0;
}

// qmcb_vmc_block on variates previously uploaded into `slot`.
int qmcb_vmc_block_slot(qmcb_ctx* c, int slot, int nsteps, double tstep, int with_energy, double* configs,
                        uint8_t* accept, double* energy, double* esum, int64_t* nacc) {
  Guard g(c);
  if (slot < 0 || slot >= qmcb_ctx::NSLOT) return fail("slot out of range");
  if (c->N == 0) return fail("recompute has not been called");
  if (build_tables(c)) return -1;
  if (c->N == 0) return fail("system shapes changed: call recompute again");
  const Sys& S = c->S;
  const size_t N = c->N;
  const size_t nse = (size_t)nsteps * S.ne;
  if (c->s_gauss[slot].n < nse * N * 3) return fail("slot was not uploaded for this block shape");
  if (S.cplx) return fail("complex wave functions run their device-resident blocks through qmcb_vmc_block (no pipelined slots)");
  CK(cudaStreamWaitEvent(c->stream, c->slot_ready[slot], 0));
  DBuf<uint8_t> acc_all;
  if (accept && acc_all.ensure(nse * N)) return -1;
  if (with_energy && (c->d_energy.ensure((size_t)nsteps * 6 * N) || c->d_esum.ensure((size_t)nsteps * 6))) return -1;
  int rc = qmcb_vmc_block_device(c, nsteps, tstep, with_energy, c->s_gauss[slot].p, c->s_unif[slot].p, c->s_u[slot].p,
                                 c->s_rot[slot].p, accept ? acc_all.p : nullptr, with_energy ? c->d_energy.p : nullptr,
                                 with_energy ? c->d_esum.p : nullptr, nullptr, c->stream);
  if (rc) {
    acc_all.release();
    return rc;
  }
  if (accept) CK(cudaMemcpyAsync(accept, acc_all.p, nse * N, cudaMemcpyDeviceToHost, c->stream));
  if (energy && with_energy)
    CK(cudaMemcpyAsync(energy, c->d_energy.p, (size_t)nsteps * 6 * N * 8, cudaMemcpyDeviceToHost, c->stream));
  if (esum && with_energy) CK(cudaMemcpyAsync(esum, c->d_esum.p, (size_t)nsteps * 6 * 8, cudaMemcpyDeviceToHost, c->stream));
  if (nacc) CK(cudaMemcpyAsync(nacc, c->d_nacc.p, nse * 8, cudaMemcpyDeviceToHost, c->stream));
  if (configs) {
    const size_t nel = N * S.ne * 3;
    if (c->d_in.ensure(nel)) return -1;
    k_conf_out<<<(unsigned)((nel + 255) / 256), 256, 0, c->stream>>>(c->st.conf, c->d_in.p, (int)N, S.ne);
    c->nlaunch++;
    CK(cudaGetLastError());
    CK(cudaMemcpyAsync(configs, c->d_in.p, nel * 8, cudaMemcpyDeviceToHost, c->stream));
  }
  if (sync_blocking(c)) return -1;
  acc_all.release();
  return 0;
}

// Asynchronous pair of qmcb_vmc_block_slot: _begin enqueues (optionally the recompute of the selected factors from
// the resident walkers, then) the block on the variates of `slot` and the copies of its results into the caller's
// (page-locked) buffers, and returns; _end sleeps until that block's results have landed.  The host driver begins
// block b+1 before it ends block b, so result read-back and host bookkeeping leave the critical path.
int qmcb_vmc_block_slot_begin(qmcb_ctx* c, int slot, int nsteps, double tstep, int with_energy, int recompute_which,
                              double* configs, double* energy, int64_t* nacc) {
  Guard g(c);
  if (slot < 0 || slot >= qmcb_ctx::NSLOT) return fail("slot out of range");
  if (c->N == 0) return fail("recompute has not been called");
  if (build_tables(c)) return -1;
  if (c->N == 0) return fail("system shapes changed: call recompute again");
  const Sys& S = c->S;
  const size_t N = c->N;
  const size_t nse = (size_t)nsteps * S.ne;
  if (c->s_gauss[slot].n < nse * N * 3) return fail("slot was not uploaded for this block shape");
  if (S.cplx) return fail("complex wave functions run their device-resident blocks through qmcb_vmc_block (no pipelined slots)");
  if (recompute_which) {
    if (which_ok(c, recompute_which)) return -1;
    if (recompute_from_resident(c, recompute_which, (int)N)) return -1;
  }
  CK(cudaStreamWaitEvent(c->stream, c->slot_ready[slot], 0));
  if (with_energy && (c->r_energy[slot].ensure((size_t)nsteps * 6 * N) || c->r_esum[slot].ensure((size_t)nsteps * 6))) return -1;
  if (c->r_nacc[slot].ensure(nse) || c->r_conf[slot].ensure(N * S.ne * 3)) return -1;
  int rc = qmcb_vmc_block_device(c, nsteps, tstep, with_energy, c->s_gauss[slot].p, c->s_unif[slot].p, c->s_u[slot].p,
                                 c->s_rot[slot].p, nullptr, with_energy ? c->r_energy[slot].p : nullptr,
                                 with_energy ? c->r_esum[slot].p : nullptr, (int64_t*)c->r_nacc[slot].p, c->stream);
  if (rc) return rc;
  // results are staged per slot, so their way to the host (own stream) overlaps the next block
  if (configs) CK(cudaMemcpyAsync(c->r_conf[slot].p, c->st.conf, N * S.ne * 3 * 8, cudaMemcpyDeviceToDevice, c->stream));
  CK(cudaEventRecord(c->ev_block, c->stream));
  CK(cudaStreamWaitEvent(c->d2h_stream, c->ev_block, 0));
  if (energy && with_energy)
    CK(cudaMemcpyAsync(energy, c->r_energy[slot].p, (size_t)nsteps * 6 * N * 8, cudaMemcpyDeviceToHost, c->d2h_stream));
  if (nacc) CK(cudaMemcpyAsync(nacc, c->r_nacc[slot].p, nse * 8, cudaMemcpyDeviceToHost, c->d2h_stream));
  if (configs) CK(cudaMemcpyAsync(configs, c->r_conf[slot].p, N * S.ne * 3 * 8, cudaMemcpyDeviceToHost, c->d2h_stream));
  CK(cudaEventRecord(c->block_done[slot], c->d2h_stream));
  c->block_pending[slot] = true;
  return 0;
}

int qmcb_vmc_block_slot_end(qmcb_ctx* c, int slot) {
  Guard g(c);
  if (slot < 0 || slot >= qmcb_ctx::NSLOT) return fail("slot out of range");
  if (!c->block_pending[slot]) return fail("no block was begun on this slot");
  CK(cudaEventSynchronize(c->block_done[slot]));
  c->block_pending[slot] = false;
  CK(cudaGetLastError());
  return 0;
}

// ---------------------------------------------------------------------------------------
// Device-resident DMC propagation (dmc_propagate, pyqmc/method/dmc.py:123-221) for single-determinant
// open-boundary Slater-Jastrow wave functions: initial local energy, then per step the T-moves of every
// electron, the drift-diffusion sweep (k_vmc_sweep<G, true>), the local energy and the weight update.
static int dmc_tmove_electron(qmcb_ctx* c, int e, double tau, const double* d_u, const double* d_rot,
                              const double* d_sel, const double* d_acc, unsigned long long* ntacc, int which,
                              cudaStream_t stream) {
  const Sys& S = c->S;
  const size_t N = c->N, M = (size_t)S.tot_naip;
  double* d_ratio = c->d_out.p;
  double* d_weight = d_ratio + N * M;
  double* d_pos = d_weight + N * M;
  // fused selection + application (k_tmove_apply): 3 launches per electron instead of 8
  constexpr int GA = 16;
  const CoopLayout CLa = coop_layout(S);
  const size_t apply_sm = ((c->smem_bytes + 15) & ~(size_t)15) + (size_t)(128 / GA) * CLa.total * 8;
  const bool fused = S.ndet == 1 && !c->have_j3 && !S.pbc && apply_sm <= 200 * 1024 && 2 * (int)M <= CLa.total &&
                     std::getenv("QMCB_TMOVE_UNFUSED") == nullptr;
  if (!fused) {
    k_tmove_init<<<(unsigned)((N * M + 255) / 256), 256, 0, stream>>>(S, c->st, e, d_ratio, d_weight, d_pos);
    c->nlaunch++;
    CK(cudaGetLastError());
  }
  if (!fused || e == 0) CK(cudaMemsetAsync(c->es.count, 0, sizeof(int), stream));
  const long long nt = (long long)N * S.necp;
  if (prep_kernel(k_ecp_prepare, c->smem_bytes)) return -1;
  k_ecp_prepare<<<(unsigned)((nt + 127) / 128), 128, c->smem_bytes, stream>>>(S, c->st, c->es, d_u, e);
  c->nlaunch++;
  CK(cudaGetLastError());
  const size_t npts = N * S.necp * S.max_naip;
  EcpPointArgs ea{};
  ea.rot = d_rot;
  ea.quad = c->d_quad.p;
  ea.e_only = e;
  ea.tmove_tau = tau;
  ea.tm_ratio = d_ratio;
  ea.tm_weight = d_weight;
  ea.tm_pos = d_pos;
  ea.scr = c->d_scr.p;
  ea.scr_stride = npts;
  const long long grid = std::min<long long>(((long long)npts + 127) / 128, 148LL * 16);
  const size_t sm = c->smem_bytes;
  constexpr int GE = 8, BE = 64;
  const CoopLayout CLe = coop_layout(S);
  const size_t esm = ((sm + 15) & ~(size_t)15) + (size_t)(BE / GE) * (CLe.total + j3_scratch_doubles(S)) * 8;
  if (S.pbc) {  // wrapped quadrature points, lattice-summed orbitals there, then the ratios (as qmcb_tmoves)
    if (prep_kernel(k_ecp_points<0>, sm)) return -1;
    if (ecp_points_pbc_prepass<0>(c, ea, (long long)npts, grid, stream)) return -1;
    k_ecp_points<0><<<(unsigned)grid, 128, sm, stream>>>(S, c->st, c->es, ea);
  } else if (esm <= 100 * 1024 && std::getenv("QMCB_NO_COOP_ECP") == nullptr) {
    // few points per launch (one electron, masked walkers): lanes cooperate on a point
    const long long cgrid = std::max<long long>(1, std::min<long long>(((long long)npts + (BE / GE) - 1) / (BE / GE), 148LL * 8));
    if (prep_kernel(k_ecp_points_coop<GE>, esm)) return -1;
    k_ecp_points_coop<GE><<<(unsigned)cgrid, BE, esm, stream>>>(S, c->st, c->es, ea);
  } else if (c->nmot == 4) {
    if (prep_kernel(k_ecp_points<4>, sm)) return -1;
    k_ecp_points<4><<<(unsigned)grid, 128, sm, stream>>>(S, c->st, c->es, ea);
  } else if (c->nmot == 8) {
    if (prep_kernel(k_ecp_points<8>, sm)) return -1;
    k_ecp_points<8><<<(unsigned)grid, 128, sm, stream>>>(S, c->st, c->es, ea);
  } else {
    if (prep_kernel(k_ecp_points<0>, sm)) return -1;
    k_ecp_points<0><<<(unsigned)grid, 128, sm, stream>>>(S, c->st, c->es, ea);
  }
  c->nlaunch++;
  CK(cudaGetLastError());
  TmoveSelectArgs ts{};
  ts.e = e;
  ts.M = (int)M;
  ts.ratio = d_ratio;
  ts.weight = d_weight;
  ts.pos = d_pos;
  ts.sel_u = d_sel;
  ts.acc_u = d_acc;
  ts.accept = c->d_accept.p;
  ts.ntacc = ntacc;
  if (fused) {
    ts.item_of = c->es.item_of;
    ts.count = c->es.count;
    if (prep_kernel(k_tmove_apply<GA>, apply_sm)) return -1;
    k_tmove_apply<GA><<<(unsigned)((N + (128 / GA) - 1) / (128 / GA)), 128, apply_sm, stream>>>(S, c->st, ts);
    c->nlaunch++;
    CK(cudaGetLastError());
    c->saved_slot = -1;
  } else {
    k_tmove_select<<<(unsigned)((N + 127) / 128), 128, 0, stream>>>(S, c->st, ts);
    c->nlaunch++;
    CK(cudaGetLastError());
    if (S.pbc) {
      // the selected position wrapped as the reference wraps it, the lattice-summed MO rows there, the rows of the
      // accepted walkers into saved_mo / the MO cache (k_pbc_move_general phases 3 and 2), then the update kernels
      constexpr int GM = 16;
      const size_t msm = ((c->smem_bytes + 15) & ~(size_t)15) + (size_t)(128 / GM) * j3_scratch_doubles(S) * 8;
      const unsigned mgrid = (unsigned)(((long long)N * GM + 127) / 128);
      MoveArgs ma{};
      ma.e = e;
      ma.tstep = tau;
      ma.accept = c->d_accept.p;
      if (prep_kernel(k_pbc_move_general<GM, true>, msm)) return -1;
      k_pbc_move_general<GM, true><<<mgrid, 128, msm, stream>>>(S, c->st, ma, 3);
      c->nlaunch++;
      CK(cudaGetLastError());
      if (c->have_slater) {
        const int s = e >= S.nup ? 1 : 0;
        const int ldmax = std::max(S.ldc[0], S.ldc[1]);
        PbcMoArgs a{};
        a.npoints = (long long)N;
        a.pos = c->st.saved_pos;
        a.wrap = c->st.saved_wrap;
        a.naip = 1;
        a.spin_mode = 0;
        a.spin = s;
        a.out = c->st.monew;
        a.stride_p = 5 * ldmax;
        a.stride_c = ldmax;
        a.stride_j = 1;
        a.mask = c->d_accept.p;  // T-moves are rarely accepted: only those walkers' rows are evaluated
        if (launch_pbc_mo(c, 2, a, (long long)N, stream)) return -1;
        k_pbc_move_general<GM, true><<<mgrid, 128, msm, stream>>>(S, c->st, ma, 2);
        c->nlaunch++;
        CK(cudaGetLastError());
      }
      if (launch_update(c, which, e, c->d_accept.p, stream)) return -1;
      c->paircache_valid = false;  // the cached MO rows stay valid: accepted walkers refreshed theirs
      return 0;
    }
    if (c->have_slater) {  // MO row at the selected position of the accepted walkers (no saved values, dmc.py:176)
      PointArgs pa{};
      pa.which = 1;
      pa.e = e;
      pa.naip = 1;
      pa.pos = c->st.saved_pos;
      pa.npoints = (int)N;
      pa.mask = c->d_accept.p;
      pa.save = 1;
      pa.scr = c->d_scr.p;
      pa.scr_stride = N;
      if (launch_point<PV_MOSAVE>(c, pa, stream)) return -1;
    }
    if (launch_update(c, which, e, c->d_accept.p, stream)) return -1;
  }
  c->mocache_valid = false;
  c->paircache_valid = false;
  return 0;
}

int qmcb_dmc_block(qmcb_ctx* c, int nsteps, double tstep, double branchcut, double e_trial, double e_est,
                   const double* gauss, const double* unif, const double* ecp_u, const double* ecp_rot,
                   const double* tm_u, const double* tm_rot, const double* tm_sel, const double* tm_acc,
                   double* weights, double* configs, double* wsums, int64_t* nacc, int64_t* ntacc) {
  Guard g(c);
  if (c->N == 0) return fail("recompute has not been called");
  if (build_tables(c)) return -1;
  if (c->N == 0) return fail("system shapes changed: call recompute again");
  const Sys& S = c->S;
  const size_t N = c->N;
  cudaStream_t stream = c->stream;
  const int which = (c->have_slater ? 1 : 0) | (c->have_jastrow ? 2 : 0) | (c->have_j3 ? 4 : 0);
  if (S.cplx) return fail("complex wave functions are served by the protocol calls and the energy accumulator; the device-resident block / SR drivers are real-only");
  const CoopLayout CL = coop_layout(S);
  const size_t tab = (c->smem_bytes + 15) & ~(size_t)15;
  const int G = 16, sweep_warps = 4, sweep_walkers = sweep_warps * (32 / G);
  const size_t sweep_smem = tab + (size_t)sweep_walkers * CL.total * 8;
  // single-determinant Slater-Jastrow: one sweep launch per step; multi-determinant and / or three-body wave functions:
  // per electron k_vmc_move_coop<GM, true> + the update kernels, as in the VMC block
  const bool use_sweep = (!c->have_slater || S.ndet == 1) && !c->have_j3 && !S.pbc && sweep_smem <= 200 * 1024;
  constexpr int GM = 16;
  const size_t move_smem = tab + (size_t)(128 / GM) * (CL.total + j3_scratch_doubles(S)) * 8;
  if (!use_sweep && move_smem > 200 * 1024) return fail("device-resident DMC: move scratch exceeds shared memory");
  const bool tmoves = S.necp > 0;
  if (ensure_energy_scratch(c) || energy_scratch_points(c)) return -1;
  const size_t M = (size_t)S.tot_naip;
  const size_t nse = (size_t)nsteps * S.ne;
  const size_t nu1 = (size_t)S.ne * S.necp * N, nr1 = (size_t)S.ne * S.necp * 9;
  DBuf<double>&d_tmu = c->m_tmu, &d_tmrot = c->m_tmrot, &d_tmsel = c->m_tmsel, &d_tmacc = c->m_tmacc, &d_w = c->m_w,
              &d_eold = c->m_eold, &d_v2old = c->m_v2old, &d_r2p = c->m_r2p, &d_r2a = c->m_r2a, &d_prod = c->m_prod,
              &d_ws = c->m_ws;
  DBuf<unsigned long long>& d_ntacc = c->m_ntacc;
  int rc = 0;
  do {
    if (c->d_gauss.ensure(nse * N * 3) || c->d_unif.ensure(nse * N) || c->d_u.ensure((size_t)(nsteps + 1) * nu1) ||
        c->d_rot.ensure((size_t)(nsteps + 1) * nr1) || c->d_energy.ensure(6 * N) || c->d_accept.ensure(N) ||
        c->d_nacc.ensure(nse) || d_ntacc.ensure(nse) || d_w.ensure(N) || d_eold.ensure(N) || d_v2old.ensure(N) ||
        d_r2p.ensure(N) || d_r2a.ensure(N) || d_prod.ensure(7 * N) || d_ws.ensure((size_t)nsteps * 8) ||
        c->d_out.ensure(std::max<size_t>(N * M * 5, N * 8)) || ensure_scratch(c, std::max<size_t>(N * S.necp * S.max_naip, N), 5)) {
      rc = -1;
      break;
    }
    auto up = [&](double* dst, const double* src, size_t n) {
      return n == 0 ? cudaSuccess : cudaMemcpyAsync(dst, src, n * 8, cudaMemcpyHostToDevice, stream);
    };
    // the block's variates: uploaded from the host, or already on the device (qmcb_devrng_dmc_block -> slot)
    const int from_slot = c->dmc_use_slot;
    c->dmc_use_slot = -1;
    double *pg = c->d_gauss.p, *pu = c->d_unif.p, *peu = c->d_u.p, *per = c->d_rot.p;
    double *ptu = nullptr, *ptr_ = nullptr, *pts = nullptr, *pta = nullptr;
    if (from_slot >= 0) {
      qmcb_ctx::DmcSlot& sl = c->dmc_slot[from_slot];
      if (!sl.filled || sl.gauss.n < nse * N * 3) {
        rc = fail("qmcb_dmc_block_slot: the slot holds no variates of this block shape");
        break;
      }
      cudaStreamWaitEvent(stream, sl.ready, 0);
      pg = sl.gauss.p; pu = sl.unif.p; peu = sl.u.p; per = sl.rot.p;
      ptu = sl.tmu.p; ptr_ = sl.tmrot.p; pts = sl.tmsel.p; pta = sl.tmacc.p;
      if (up(d_w.p, weights, N) != cudaSuccess) {
        rc = fail("H2D copy failed");
        break;
      }
    } else if (up(c->d_gauss.p, gauss, nse * N * 3) != cudaSuccess || up(c->d_unif.p, unif, nse * N) != cudaSuccess ||
               up(d_w.p, weights, N) != cudaSuccess) {
      rc = fail("H2D copy failed");
      break;
    }
    if (tmoves && from_slot < 0) {
      if (!ecp_u || !ecp_rot || !tm_u || !tm_rot || !tm_sel || !tm_acc) {
        rc = fail("DMC with ECPs needs the energy and T-move variates");
        break;
      }
      if (d_tmu.ensure(nse * S.necp * N) || d_tmrot.ensure(nse * S.necp * 9) || d_tmsel.ensure(nse * N) || d_tmacc.ensure(nse * N)) {
        rc = -1;
        break;
      }
      if (up(c->d_u.p, ecp_u, (size_t)(nsteps + 1) * nu1) != cudaSuccess || up(c->d_rot.p, ecp_rot, (size_t)(nsteps + 1) * nr1) != cudaSuccess ||
          up(d_tmu.p, tm_u, nse * S.necp * N) != cudaSuccess || up(d_tmrot.p, tm_rot, nse * S.necp * 9) != cudaSuccess ||
          up(d_tmsel.p, tm_sel, nse * N) != cudaSuccess || up(d_tmacc.p, tm_acc, nse * N) != cudaSuccess) {
        rc = fail("H2D copy failed");
        break;
      }
      ptu = d_tmu.p; ptr_ = d_tmrot.p; pts = d_tmsel.p; pta = d_tmacc.p;
    }
    cudaMemsetAsync(c->d_nacc.p, 0, nse * 8, stream);
    cudaMemsetAsync(d_ntacc.p, 0, nse * 8, stream);
    // E_L and v^2 before the first step (dmc.py:150-152)
    if ((rc = launch_energy(c, peu, per, c->d_energy.p, stream))) break;
    DmcWeightArgs wa{};
    wa.tstep = tstep;
    wa.branchcut = branchcut;
    wa.e_trial = e_trial;
    wa.e_est = e_est;
    wa.energy = c->d_energy.p;
    wa.r2prop = d_r2p.p;
    wa.r2acc = d_r2a.p;
    wa.eold = d_eold.p;
    wa.v2old = d_v2old.p;
    wa.weights = d_w.p;
    wa.prod = d_prod.p;
    wa.init = 1;
    k_dmc_weights<<<(unsigned)((N + 127) / 128), 128, 0, stream>>>(S, c->st, wa);
    c->nlaunch++;
    wa.init = 0;
    for (int step = 0; step < nsteps && rc == 0; ++step) {
      if (tmoves) {
        for (int e = 0; e < S.ne && rc == 0; ++e) {
          const size_t se = (size_t)step * S.ne + e;
          rc = dmc_tmove_electron(c, e, tstep, ptu + se * S.necp * N, ptr_ + se * S.necp * 9, pts + se * N, pta + se * N,
                                  d_ntacc.p + se, which, stream);
        }
        if (rc) break;
      }
      if (c->have_slater && !c->mocache_valid) {
        if ((rc = launch_mo_all(c, 0, stream))) break;
      }
      if (use_sweep && c->have_jastrow && !c->paircache_valid) {
        const long long nt = (long long)N * (S.npair + S.ne);
        if ((rc = prep_kernel(k_pair_cache_build, c->smem_bytes))) break;
        k_pair_cache_build<<<(unsigned)((nt + 127) / 128), 128, c->smem_bytes, stream>>>(S, c->st);
        c->nlaunch++;
        c->paircache_valid = true;
      }
      cudaMemsetAsync(d_r2p.p, 0, N * 8, stream);
      cudaMemsetAsync(d_r2a.p, 0, N * 8, stream);
      if (!use_sweep) {
        for (int e = 0; e < S.ne && rc == 0; ++e) {
          const size_t se = (size_t)step * S.ne + e;
          MoveArgs ma{};
          ma.e = e;
          ma.tstep = tstep;
          ma.gauss = pg + se * N * 3;
          ma.unif = pu + se * N;
          ma.accept = c->d_accept.p;
          ma.nacc = c->d_nacc.p + se;
          ma.r2prop = d_r2p.p;
          ma.r2acc = d_r2a.p;
          const unsigned mgrid = (unsigned)(((long long)N * GM + 127) / 128);
          if (S.pbc) {  // drift + wrapped proposal, lattice-summed orbitals, Metropolis test (as the periodic VMC block)
            const size_t msm = tab + (size_t)(128 / GM) * j3_scratch_doubles(S) * 8;
            if ((rc = prep_kernel(k_pbc_move_general<GM, true>, msm))) break;
            k_pbc_move_general<GM, true><<<mgrid, 128, msm, stream>>>(S, c->st, ma, 0);
            c->nlaunch++;
            if (c->have_slater) {
              const int s = e >= S.nup ? 1 : 0;
              const int ldmax = std::max(S.ldc[0], S.ldc[1]);
              PbcMoArgs a{};
              a.npoints = (long long)N;
              a.pos = c->st.saved_pos;
              a.wrap = c->st.saved_wrap;
              a.naip = 1;
              a.spin_mode = 0;
              a.spin = s;
              a.out = c->st.monew;
              a.stride_p = 5 * ldmax;
              a.stride_c = ldmax;
              a.stride_j = 1;
              if ((rc = launch_pbc_mo(c, 2, a, (long long)N, stream))) break;
            }
            k_pbc_move_general<GM, true><<<mgrid, 128, msm, stream>>>(S, c->st, ma, 1);
          } else {
            if ((rc = prep_kernel(k_vmc_move_coop<GM, true>, move_smem))) break;
            k_vmc_move_coop<GM, true><<<mgrid, 128, move_smem, stream>>>(S, c->st, ma);
          }
          c->nlaunch++;
          if (cudaGetLastError() != cudaSuccess) {
            rc = fail("DMC move kernel launch failed");
            break;
          }
          rc = launch_update(c, which, e, ma.accept, stream);
          c->paircache_valid = false;  // the cached MO rows stay valid: accepted walkers refreshed theirs
        }
        if (rc) break;
        c->kinetic_valid = false;
        if ((rc = launch_energy(c, peu + (size_t)(step + 1) * nu1, per + (size_t)(step + 1) * nr1, c->d_energy.p, stream))) break;
        k_dmc_weights<<<(unsigned)((N + 127) / 128), 128, 0, stream>>>(S, c->st, wa);
        c->nlaunch++;
        k_colsum<<<7, 256, 0, stream>>>(d_prod.p, (int)N, d_ws.p + (size_t)step * 8);
        c->nlaunch++;
        continue;
      }
      SweepArgs sa{};
      sa.tstep = tstep;
      sa.gauss = pg + (size_t)step * S.ne * N * 3;
      sa.unif = pu + (size_t)step * S.ne * N;
      sa.accept = nullptr;
      sa.nacc = c->d_nacc.p + (size_t)step * S.ne;
      sa.r2prop = d_r2p.p;
      sa.r2acc = d_r2a.p;
      sa.ke_e = c->es.ke_e;
      sa.g2_e = c->es.g2_e;
      c->kinetic_valid = true;
      if ((rc = prep_kernel(k_vmc_sweep<16, true>, sweep_smem))) break;
      k_vmc_sweep<16, true><<<(unsigned)((N + sweep_walkers - 1) / sweep_walkers), sweep_warps * 32, sweep_smem, stream>>>(S, c->st, sa);
      c->nlaunch++;
      if (cudaGetLastError() != cudaSuccess) {
        rc = fail("k_vmc_sweep<16, true> launch failed");
        break;
      }
      if ((rc = launch_energy(c, peu + (size_t)(step + 1) * nu1, per + (size_t)(step + 1) * nr1, c->d_energy.p, stream))) break;
      k_dmc_weights<<<(unsigned)((N + 127) / 128), 128, 0, stream>>>(S, c->st, wa);
      c->nlaunch++;
      k_colsum<<<7, 256, 0, stream>>>(d_prod.p, (int)N, d_ws.p + (size_t)step * 8);
      c->nlaunch++;
    }
    if (rc) break;
    if (cudaGetLastError() != cudaSuccess) {
      rc = fail("DMC kernel launch failed");
      break;
    }
    cudaMemcpyAsync(weights, d_w.p, N * 8, cudaMemcpyDeviceToHost, stream);
    if (wsums) cudaMemcpyAsync(wsums, d_ws.p, (size_t)nsteps * 8 * 8, cudaMemcpyDeviceToHost, stream);
    if (nacc) cudaMemcpyAsync(nacc, c->d_nacc.p, nse * 8, cudaMemcpyDeviceToHost, stream);
    if (ntacc) cudaMemcpyAsync(ntacc, d_ntacc.p, nse * 8, cudaMemcpyDeviceToHost, stream);
    if (configs) cudaMemcpyAsync(configs, c->st.conf, N * S.ne * 3 * 8, cudaMemcpyDeviceToHost, stream);
    if (sync_blocking(c)) rc = -1;
  } while (0);
  cudaStreamSynchronize(stream);
  c->saved_slot = -1;
  return rc;
}

// qmcb_dmc_block on the variates qmcb_devrng_dmc_block generated into `slot`; *branch_draw receives the uniform
// variate of this block's branching step (dmc.py:361), drawn right after the block's variates.
int qmcb_dmc_block_slot(qmcb_ctx* c, int slot, int nsteps, double tstep, double branchcut, double e_trial, double e_est,
                        double* weights, double* configs, double* wsums, int64_t* nacc, int64_t* ntacc, double* branch_draw) {
  if (slot < 0 || slot > 1) return fail("slot out of range");
  c->dmc_use_slot = slot;
  int rc = qmcb_dmc_block(c, nsteps, tstep, branchcut, e_trial, e_est, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr,
                          nullptr, nullptr, weights, configs, wsums, nacc, ntacc);
  c->dmc_use_slot = -1;
  if (rc) return rc;
  if (branch_draw) CK(cudaMemcpy(branch_draw, c->dmc_slot[slot].branch.p, 8, cudaMemcpyDeviceToHost));
  return 0;
}

// ---------------------------------------------------------------------------------------
// StochasticReconfiguration.avg on the device (stochastic_reconfiguration.py:85-118): local energy,
// parameter gradients, nodal regularisation, the weighted sums and the P x P overlap product.
int qmcb_sr_avg(qmcb_ctx* c, int nparam, const int32_t* src, const int64_t* off, const double* weights,
                const double* ecp_u, const double* ecp_rot, double nodal_cutoff, double* energy_avg, double* dpH,
                double* dppsi, double* dpidpj) {
  Guard g(c);
  if (c->N == 0) return fail("recompute has not been called");
  if (build_tables(c)) return -1;
  if (c->N == 0) return fail("system shapes changed: call recompute again");
  const Sys& S = c->S;
  const size_t N = c->N, P = (size_t)nparam;
  cudaStream_t stream = c->stream;
  if (S.cplx) return fail("complex wave functions are served by the protocol calls and the energy accumulator; the device-resident block / SR drivers are real-only");
  bool need[6] = {false, false, false, false, false, false};
  for (size_t j = 0; j < P; ++j) {
    if (src[j] < 0 || src[j] > 5) return fail("qmcb_sr_avg: unknown parameter source");
    need[src[j]] = true;
  }
  if ((need[0] || need[1] || need[2]) && !c->have_slater) return fail("context has no Slater factor");
  if ((need[3] || need[4]) && !c->have_jastrow) return fail("context has no Jastrow factor");
  if (need[5] && !c->have_j3) return fail("context has no three-body Jastrow factor");
  if (ensure_energy_scratch(c) || energy_scratch_points(c)) return -1;
  const size_t nu = (size_t)S.ne * S.necp * N, nr = (size_t)S.ne * S.necp * 9;
  if (c->d_u.ensure(nu) || c->d_rot.ensure(nr) || c->d_energy.ensure(6 * N)) return -1;
  if (S.necp > 0) {
    if (!ecp_u || !ecp_rot) return fail("ECP random variates missing");
    if (h2d(c, c->d_u.p, ecp_u, nu * 8) || h2d(c, c->d_rot.p, ecp_rot, nr * 8)) return -1;
  }
  if (launch_energy(c, c->d_u.p, c->d_rot.p, c->d_energy.p, stream)) return -1;
  const int gstride = std::max(std::max(S.nds[0], S.nds[1]), 1);
  DBuf<double> d_det, d_G, d_ao, d_mo[2], d_c3, d_w, d_dp, d_wdpr, d_red, d_C;
  DBuf<int> d_src;
  DBuf<long long> d_off;
  int rc = 0;
  do {
    SrArgs a{};
    a.P = (int)P;
    a.cutoff = nodal_cutoff;
    if (need[0] || need[1] || need[2]) {
      if (d_det.ensure(N * std::max(S.ndet, 1)) || d_G.ensure(2 * N * gstride)) { rc = -1; break; }
      k_pgrad_det<<<(unsigned)((N + 127) / 128), 128, 0, stream>>>(S, c->st, d_det.p, d_G.p, gstride);
      c->nlaunch++;
      a.base[0] = d_det.p;
      a.stride[0] = S.ndet;
    }
    if (need[1] || need[2]) {
      if (d_ao.ensure(N * S.ne * S.nao * (S.pbc ? S.nk : 1))) { rc = -1; break; }
      if ((rc = launch_ao_all(c, d_ao.p, stream))) break;
      for (int s = 0; s < 2; ++s) {
        if (!need[1 + s]) continue;
        const size_t nout = N * S.nao * S.nmo[s];
        if (d_mo[s].ensure(nout)) { rc = -1; break; }
        k_pgrad_mo<<<(unsigned)((nout + 127) / 128), 128, 0, stream>>>(S, c->st, s, d_ao.p, d_G.p, gstride, d_mo[s].p);
        c->nlaunch++;
        a.base[1 + s] = d_mo[s].p;
        a.stride[1 + s] = (long long)S.nao * S.nmo[s];
      }
      if (rc) break;
    }
    a.base[3] = c->st.avalues;
    a.stride[3] = (long long)S.natom * S.na * 2;
    a.base[4] = c->st.bvalues;
    a.stride[4] = (long long)S.nb * 3;
    if (need[5]) {
      const size_t nt = N * S.natom * S.na3 * S.na3;
      if (d_c3.ensure(nt * S.nb3 * 3)) { rc = -1; break; }
      if ((rc = prep_kernel(k_jastrow3_pgrad, c->smem_bytes))) break;
      k_jastrow3_pgrad<<<(unsigned)((nt + 127) / 128), 128, c->smem_bytes, stream>>>(S, c->st, d_c3.p);
      c->nlaunch++;
      a.base[5] = d_c3.p;
      a.stride[5] = (long long)S.natom * S.na3 * S.na3 * S.nb3 * 3;
    }
    for (size_t j = 0; j < P; ++j)
      if (off[j] < 0 || off[j] >= a.stride[src[j]]) { rc = fail("qmcb_sr_avg: parameter offset out of range"); break; }
    if (rc) break;
    std::vector<long long> off64(off, off + P);
    if (d_src.ensure(P) || d_off.ensure(P) || d_w.ensure(N) || d_dp.ensure(N * std::max<size_t>(P, 1)) ||
        d_wdpr.ensure(N * std::max<size_t>(P, 1)) || d_red.ensure(2 * (P + 6)) || d_C.ensure(std::max<size_t>(P * P, 1))) { rc = -1; break; }
    cudaMemcpyAsync(d_src.p, src, P * 4, cudaMemcpyHostToDevice, stream);
    cudaMemcpyAsync(d_off.p, off64.data(), P * 8, cudaMemcpyHostToDevice, stream);
    cudaMemcpyAsync(d_w.p, weights, N * 8, cudaMemcpyHostToDevice, stream);
    cudaStreamSynchronize(stream);  // off64 / caller buffers are read by the copies above
    a.src = d_src.p;
    a.off = d_off.p;
    a.weights = d_w.p;
    a.energy = c->d_energy.p;
    a.dp = d_dp.p;
    a.wdpr = d_wdpr.p;
    if (P > 0) {
      k_sr_gather<<<(unsigned)((N * P + 127) / 128), 128, 0, stream>>>(c->st, a);
      c->nlaunch++;
    }
    k_sr_colsum<<<(unsigned)(P + 6), 256, 0, stream>>>(c->st, a, d_red.p);
    c->nlaunch++;
    if (P > 0) {
      if (launch_gemm_tn(d_dp.p, d_wdpr.p, (int)N, (int)P, d_C.p, c->b_gemm, -1, stream)) { rc = -1; break; }
      c->nlaunch += 2;
    }
    if (cudaGetLastError() != cudaSuccess) { rc = fail("stochastic-reconfiguration kernel launch failed"); break; }
    std::vector<double> red(2 * (P + 6));
    if ((rc = d2h(c, red.data(), d_red.p, red.size() * 8))) break;
    for (size_t j = 0; j < P; ++j) {
      dppsi[j] = red[j];
      dpH[j] = red[P + 6 + j];
    }
    for (int k = 0; k < 6; ++k) energy_avg[k] = red[P + k];
    if (P > 0) rc = d2h(c, dpidpj, d_C.p, P * P * 8);
  } while (0);
  cudaStreamSynchronize(stream);
  DBuf<double>* tmp[] = {&d_det, &d_G, &d_ao, &d_mo[0], &d_mo[1], &d_c3, &d_w, &d_dp, &d_wdpr, &d_red, &d_C};
  for (auto* b : tmp) b->release();
  d_src.release();
  d_off.release();
  return rc;
}

int qmcb_pinned_alloc(int64_t bytes, void** out) {
  void* p = nullptr;
  cudaError_t e = cudaMallocHost(&p, (size_t)std::max<int64_t>(bytes, 1));
  if (e != cudaSuccess) return fail(std::string("cudaMallocHost: ") + cudaGetErrorString(e));
  *out = p;
  return 0;
}

int qmcb_pinned_free(void* p) {
  if (p) cudaFreeHost(p);
  return 0;
}

int qmcb_kernel_launches(qmcb_ctx* c, int64_t* count) {
  *count = c->nlaunch;
  return 0;
}

// ---------------------------------------------------------------------------------------
int qmcb_sm_update_device(int n, int e, int64_t nmat, double* d_inv, const double* d_vec, const uint8_t* d_mask,
                          double* d_ratio, void* stream) {
  SmArgs a{};
  a.n = n;
  a.e = e;
  a.nds = 1;
  a.vec_stride = n;
  a.nmat = nmat;
  a.inv = d_inv;
  a.vec = d_vec;
  a.occ = nullptr;
  a.mask = d_mask;
  a.ratio = d_ratio;
  return launch_sm(nullptr, a, (cudaStream_t)stream, nullptr);
}

int qmcb_sm_update(int n, int e, int64_t nmat, double* inv, const double* vec, const uint8_t* mask, double* ratio) {
  DBuf<double> di, dv, dr;
  DBuf<uint8_t> dm;
  const size_t M = (size_t)nmat;
  if (di.ensure(M * n * n) || dv.ensure(M * n) || dr.ensure(M)) return -1;
  int rc = 0;
  do {
    if (cudaMemcpy(di.p, inv, M * n * n * 8, cudaMemcpyHostToDevice) != cudaSuccess ||
        cudaMemcpy(dv.p, vec, M * n * 8, cudaMemcpyHostToDevice) != cudaSuccess) {
      rc = fail("cudaMemcpy H2D failed");
      break;
    }
    if (cudaMemset(dr.p, 0, M * 8) != cudaSuccess) {
      rc = fail("cudaMemset failed");
      break;
    }
    if (mask) {
      if (dm.ensure(M)) {
        rc = -1;
        break;
      }
      cudaMemcpy(dm.p, mask, M, cudaMemcpyHostToDevice);
    }
    rc = qmcb_sm_update_device(n, e, nmat, di.p, dv.p, mask ? dm.p : nullptr, dr.p, nullptr);
    if (rc) break;
    if (cudaDeviceSynchronize() != cudaSuccess) {
      rc = fail(std::string("kernel failed: ") + cudaGetErrorString(cudaGetLastError()));
      break;
    }
    cudaMemcpy(inv, di.p, M * n * n * 8, cudaMemcpyDeviceToHost);
    cudaMemcpy(ratio, dr.p, M * 8, cudaMemcpyDeviceToHost);
  } while (0);
  di.release();
  dv.release();
  dr.release();
  dm.release();
  return rc;
}

// FP64 FMA roof of this device, measured: 8 independent DFMA chains per thread, enough resident warps to saturate
// the FP64 pipe.  bench.py reports the sweep kernel's FP64 rate against it (the step is FP64-latency bound, HBM is
// the wrong roof for it).
__global__ void __launch_bounds__(256) k_fp64_peak(double* out, int iters, double a, double b) {
  double x0 = threadIdx.x * 1e-3, x1 = x0 + 1.0, x2 = x0 + 2.0, x3 = x0 + 3.0, x4 = x0 + 4.0, x5 = x0 + 5.0, x6 = x0 + 6.0,
         x7 = x0 + 7.0;
  for (int i = 0; i < iters; ++i) {
    x0 = fma(x0, a, b);
    x1 = fma(x1, a, b);
    x2 = fma(x2, a, b);
    x3 = fma(x3, a, b);
    x4 = fma(x4, a, b);
    x5 = fma(x5, a, b);
    x6 = fma(x6, a, b);
    x7 = fma(x7, a, b);
  }
  out[(size_t)blockIdx.x * blockDim.x + threadIdx.x] = ((x0 + x1) + (x2 + x3)) + ((x4 + x5) + (x6 + x7));
}

int qmcb_gemm_tn(int device, int64_t N, int P, const double* A, const double* B, double* C, int variant, int reps,
                 double* ms) {
  cudaSetDevice(device);
  DBuf<double> dA, dB, dC, work;
  const size_t nA = (size_t)N * P;
  if (dA.ensure(nA) || dB.ensure(nA) || dC.ensure((size_t)P * P)) return -1;
  CK(cudaMemcpy(dA.p, A, nA * 8, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(dB.p, B, nA * 8, cudaMemcpyHostToDevice));
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0));
  CK(cudaEventCreate(&e1));
  int rc = launch_gemm_tn(dA.p, dB.p, (int)N, P, dC.p, work, variant, nullptr);  // warm-up
  CK(cudaEventRecord(e0, nullptr));
  for (int r = 0; r < reps && !rc; ++r) rc = launch_gemm_tn(dA.p, dB.p, (int)N, P, dC.p, work, variant, nullptr);
  CK(cudaEventRecord(e1, nullptr));
  CK(cudaEventSynchronize(e1));
  float t = 0.f;
  CK(cudaEventElapsedTime(&t, e0, e1));
  if (ms) *ms = reps > 0 ? t / reps : 0.0;
  if (!rc) CK(cudaMemcpy(C, dC.p, (size_t)P * P * 8, cudaMemcpyDeviceToHost));
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  dA.release();
  dB.release();
  dC.release();
  work.release();
  return rc;
}

int qmcb_fp64_peak(int device, double* tflops) {
  CK(cudaSetDevice(device));
  cudaDeviceProp prop;
  CK(cudaGetDeviceProperties(&prop, device));
  const int grid = prop.multiProcessorCount * 8, block = 256, iters = 1 << 14;
  double* d = nullptr;
  CK(cudaMalloc(&d, (size_t)grid * block * 8));
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0));
  CK(cudaEventCreate(&e1));
  double best = 0.0;
  for (int rep = 0; rep < 4; ++rep) {
    CK(cudaEventRecord(e0, 0));
    k_fp64_peak<<<grid, block>>>(d, iters, 0.999999, 1e-6);
    CK(cudaEventRecord(e1, 0));
    CK(cudaEventSynchronize(e1));
    float ms = 0.f;
    CK(cudaEventElapsedTime(&ms, e0, e1));
    const double tf = 2.0 * 8.0 * iters * (double)grid * block / (ms * 1e-3) / 1e12;
    if (rep > 0) best = std::max(best, tf);
  }
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  cudaFree(d);
  *tflops = best;
  return 0;
}

// Orbitals at arbitrary points for the density-matrix accumulators: out[p][j] = sum_mu chi_mu(pos[p]) coeff[mu][j]
// (MoleculeOrbitalEvaluator.aos + mos, orbitals.py:85-96).  Host buffers; needs the basis of the context.
int qmcb_orbitals_at_points(qmcb_ctx* c, int64_t npoints, const double* pos, int norb, const double* coeff, double* out) {
  Guard g(c);
  if (build_tables(c)) return -1;
  const Sys& S = c->S;
  if (S.pbc) return fail("qmcb_orbitals_at_points: periodic orbitals are evaluated through the Slater factor");
  if (S.nao == 0) return fail("qmcb_orbitals_at_points: no basis (qmcb_set_basis)");
  if (npoints == 0 || norb == 0) return 0;
  DBuf<double> d_pos, d_c, d_out;
  int rc = 0;
  do {
    if (d_pos.ensure(npoints * 3) || d_c.ensure((size_t)S.nao * norb) || d_out.ensure((size_t)npoints * norb)) {
      rc = -1;
      break;
    }
    cudaMemcpyAsync(d_pos.p, pos, npoints * 3 * 8, cudaMemcpyHostToDevice, c->stream);
    cudaMemcpyAsync(d_c.p, coeff, (size_t)S.nao * norb * 8, cudaMemcpyHostToDevice, c->stream);
    if (prep_kernel(k_orbitals_points<8>, c->smem_bytes)) {
      rc = -1;
      break;
    }
    const int block = pick_block(npoints);
    k_orbitals_points<8><<<(unsigned)((npoints + block - 1) / block), block, c->smem_bytes, c->stream>>>(S, d_pos.p, npoints, d_c.p, norb,
                                                                                                     d_out.p);
    c->nlaunch++;
    if (cudaGetLastError() != cudaSuccess) {
      rc = fail("k_orbitals_points launch failed");
      break;
    }
    cudaMemcpyAsync(out, d_out.p, (size_t)npoints * norb * 8, cudaMemcpyDeviceToHost, c->stream);
    if (cudaStreamSynchronize(c->stream) != cudaSuccess) rc = fail(std::string("qmcb_orbitals_at_points: ") + cudaGetErrorString(cudaGetLastError()));
  } while (0);
  d_pos.release();
  d_c.release();
  d_out.release();
  return rc;
}

#include "devrng_api.cuh"

}  // extern "C"
