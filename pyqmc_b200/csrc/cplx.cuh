// cplx.cuh -- complex wave functions: complex MO coefficients and / or complex Bloch phases (general twists).
//
// Reference statements (relative to /root/reference):
//   dtype / phase selection        pyqmc/wf/slater.py:32-33, 212-216; orbitals.py:38-43, 61-65, 160-165
//   recompute (slogdet + inverse)  pyqmc/wf/slater.py:227-260 (np.linalg.slogdet / inv on complex matrices)
//   Sherman-Morrison + phase       pyqmc/wf/slater.py:88-94, 262-291 (get_phase = x / |x|)
//   determinant-ratio rows         pyqmc/wf/slater.py:301-380; determinant_tools.py:74-88
//   product with real Jastrows     pyqmc/wf/multiplywf.py:71-132
//   local energy                   pyqmc/observables/energy.py:57-65 (ke = -lap.real / 2, grad2 = sum |grad|^2),
//                                  eval_ecp.py:21-146 (ecp values carry wf.dtype)
//
// Layout: an MO row holds nmo_t real parts followed by nmo_t imaginary parts (Sys::cxoff), so the orbital
// evaluators and every row buffer are the real ones; inverses, determinant phases and the multi-determinant caches
// carry a second (imaginary) array.  Host-visible wave-function-valued outputs are complex128 (interleaved).
// These kernels serve the protocol calls and the energy accumulator; the device-resident block kernels stay real.
#pragma once

struct cd {
  double x, y;
};
__device__ __forceinline__ cd cmul(cd a, cd b) { return {a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x}; }
__device__ __forceinline__ cd cfma(cd a, cd b, cd c) {  // a * b + c
  return {fma(a.x, b.x, fma(-a.y, b.y, c.x)), fma(a.x, b.y, fma(a.y, b.x, c.y))};
}
__device__ __forceinline__ cd cdivc(cd a, cd b) {  // Smith's algorithm (what numpy's complex division uses)
  if (fabs(b.x) >= fabs(b.y)) {
    if (b.x == 0.0 && b.y == 0.0) return {a.x / fabs(b.x), a.y / fabs(b.x)};
    const double rat = b.y / b.x, scl = 1.0 / (b.x + b.y * rat);
    return {(a.x + a.y * rat) * scl, (a.y - a.x * rat) * scl};
  }
  const double rat = b.x / b.y, scl = 1.0 / (b.x * rat + b.y);
  return {(a.x * rat + a.y) * scl, (a.y * rat - a.x) * scl};
}
__device__ __forceinline__ double cabs2(cd a) { return a.x * a.x + a.y * a.y; }
__device__ __forceinline__ bool cfinite(cd a) { return isfinite(a.x) && isfinite(a.y); }

// -----------------------------------------------------------------------------------------
// Slater part of a single-electron query (the complex twin of slater_point_general).
// -----------------------------------------------------------------------------------------
template <int DERIV>
__device__ __forceinline__ void slater_point_cx(const Sys& S, const double* __restrict__ sd, const int* __restrict__ si,
                                                const State& st, int w, int e, double px, double py, double pz,
                                                cd (&rat)[NComp<DERIV>::value], double* __restrict__ mo_save,
                                                double* __restrict__ scr, size_t scr_stride) {
  constexpr int NC = NComp<DERIV>::value;
  const int s = e >= S.nup ? 1 : 0;
  const int n = s ? S.ndn : S.nup;
  const int eeff = e - s * S.nup;
  const int ldc = S.ldc[s], off = S.cxoff[s];
  for (int mo0 = 0; mo0 < ldc && !S.pbc; mo0 += 8) {  // periodic rows were written by k_pbc_mo*
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
  cd num[NC], den = {0.0, 0.0};
#pragma unroll
  for (int c = 0; c < NC; ++c) num[c] = {0.0, 0.0};
  for (int d = 0; d < nds; ++d) {
    const size_t base = ((size_t)w * nds + d) * n * n + eeff;
    const double* __restrict__ ir = st.inv[s] + base;
    const double* __restrict__ ii = st.inv_im[s] + base;
    cd r[NC];
#pragma unroll
    for (int c = 0; c < NC; ++c) r[c] = {0.0, 0.0};
    for (int k = 0; k < n; ++k) {
      const cd a = {ir[k * n], ii[k * n]};
      const int orb = occ[d * n + k];
#pragma unroll
      for (int c = 0; c < NC; ++c) {
        const cd v = {scr[(size_t)(c * ldc + orb) * scr_stride], scr[(size_t)(c * ldc + off + orb) * scr_stride]};
        r[c] = cfma(v, a, r[c]);
      }
    }
    if (S.ndet == 1) {
#pragma unroll
      for (int c = 0; c < NC; ++c) num[c] = r[c];
      den = {1.0, 0.0};
    } else {
      const size_t q = (size_t)w * nds + d;
      const cd wgt = cmul({st.dv[s][q], st.dv_im[s][q]}, {st.W[s][q], st.W_im[s][q]});
      den.x += wgt.x;
      den.y += wgt.y;
#pragma unroll
      for (int c = 0; c < NC; ++c) num[c] = cfma(r[c], wgt, num[c]);
    }
  }
#pragma unroll
  for (int c = 0; c < NC; ++c) rat[c] = cdivc(num[c], den);
}

template <int DERIV>
struct PointEvalCx {
  static constexpr int NC = NComp<DERIV>::value;
  cd rat[NC];
  double du, gj[3], lapj;
  __device__ __forceinline__ void run(const Sys& S, const double* sd, const int* si, const State& st, int which, int w,
                                      int e, double px, double py, double pz, double* mo_save, double* scr,
                                      size_t scr_stride) {
    rat[0] = {1.0, 0.0};
#pragma unroll
    for (int c = 1; c < NC; ++c) rat[c] = {0.0, 0.0};
    du = 0.0;
    gj[0] = gj[1] = gj[2] = 0.0;
    lapj = 0.0;
    if (which & QMCB_SLATER) slater_point_cx<DERIV>(S, sd, si, st, w, e, px, py, pz, rat, mo_save, scr, scr_stride);
    if (which & QMCB_JASTROW) jastrow_point<DERIV>(S, sd, si, st, w, e, px, py, pz, du, gj, lapj);
    if (which & QMCB_JASTROW3) jastrow3_point<DERIV>(S, sd, si, st, w, e, px, py, pz, du, gj, lapj);
  }
};

// wf.testvalue / gradient / gradient_value / gradient_laplacian: one thread per (walker, auxiliary point);
// outputs are complex (interleaved): o_val [npoints], o_grad [3][N], o_lap [N]
template <int MODE>
__global__ void __launch_bounds__(128) k_cx_point(const Sys S, const State st, const PointArgs pa) {
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
  if (pa.save && (pa.which & QMCB_SLATER)) mo_save = st.saved_mo + (size_t)w * S.ldc[pa.e >= S.nup ? 1 : 0];
  PointEvalCx<DERIV> ev;
  ev.run(S, sd, si, st, MODE == PV_MOSAVE ? QMCB_SLATER : pa.which, w, pa.e, px, py, pz, mo_save, pa.scr + p,
         pa.scr_stride);
  if (pa.save) {
    st.saved_pos[(size_t)w * 3 + 0] = px;
    st.saved_pos[(size_t)w * 3 + 1] = py;
    st.saved_pos[(size_t)w * 3 + 2] = pz;
  }
  const int N = st.N;
  cd* o_val = reinterpret_cast<cd*>(pa.o_val);
  cd* o_grad = reinterpret_cast<cd*>(pa.o_grad);
  cd* o_lap = reinterpret_cast<cd*>(pa.o_lap);
  if (MODE == PV_VALUE) {
    const double ex = exp(ev.du);
    o_val[p] = {ev.rat[0].x * ex, ev.rat[0].y * ex};
  } else if (MODE == PV_GRAD || MODE == PV_GRADVAL) {
#pragma unroll
    for (int i = 0; i < 3; ++i) {
      cd g = cdivc(ev.rat[1 + i], ev.rat[0]);
      if (MODE == PV_GRADVAL && !cfinite(g)) g = {0.0, 0.0};  // slater.py:415-417
      o_grad[(size_t)i * N + w] = {g.x + ev.gj[i], g.y};
    }
    if (MODE == PV_GRADVAL) {
      cd v = ev.rat[0];
      if (!cfinite(v)) v = {1.0, 0.0};
      const double ex = exp(ev.du);
      o_val[w] = {v.x * ex, v.y * ex};
    }
  } else if (MODE == PV_GRADLAP) {
    cd gs[3];
#pragma unroll
    for (int i = 0; i < 3; ++i) gs[i] = cdivc(ev.rat[1 + i], ev.rat[0]);
    const cd laps = cdivc(ev.rat[4], ev.rat[0]);
    double lapj = 0.0;
    cd cross = {0.0, 0.0};
    if (pa.which & (QMCB_JASTROW | QMCB_JASTROW3)) {
      lapj = ev.lapj + (ev.gj[0] * ev.gj[0] + ev.gj[1] * ev.gj[1] + ev.gj[2] * ev.gj[2]);
      cross.x = gs[0].x * ev.gj[0] + gs[1].x * ev.gj[1] + gs[2].x * ev.gj[2];
      cross.y = gs[0].y * ev.gj[0] + gs[1].y * ev.gj[1] + gs[2].y * ev.gj[2];
    }
#pragma unroll
    for (int i = 0; i < 3; ++i) o_grad[(size_t)i * N + w] = {gs[i].x + ev.gj[i], gs[i].y};
    o_lap[w] = {(laps.x + lapj) + cross.x * 2.0, laps.y + cross.y * 2.0};  // multiplywf.py:121-129
  }
}

// -----------------------------------------------------------------------------------------
// recompute: complex Gauss-Jordan with partial pivoting (largest |a|), one warp per (walker, spin determinant),
// lanes over columns, the matrix in a global work buffer gw [matrix][n][n] (complex).  Leaves the inverse in
// st.inv / st.inv_im ([orbital][electron]), the unit phase of the determinant in dsign / dphs_im and log|det| in dlog.
// -----------------------------------------------------------------------------------------
#define QMCB_CX_NMAX 64
__global__ void __launch_bounds__(128) k_cx_invert(const Sys S, const State st, int s, cd* __restrict__ gw) {
  __shared__ cd fcol_s[4][QMCB_CX_NMAX];
  __shared__ int piv_s[4][QMCB_CX_NMAX];
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  const long long t = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int nds = S.nds[s];
  const int N = st.N;
  if (t >= (long long)N * nds) return;
  const int w = (int)(t / nds), d = (int)(t - (long long)w * nds);
  const int n = s ? S.ndn : S.nup;
  if (n == 0) {
    if (lane == 0) {
      st.dsign[s][t] = 1.0;
      st.dphs_im[s][t] = 0.0;
      st.dlog[s][t] = 0.0;
    }
    return;
  }
  const int lo = s ? S.nup : 0;
  const int ldmax = S.ldc[0] > S.ldc[1] ? S.ldc[0] : S.ldc[1];
  const int off = S.cxoff[s];
  const int* __restrict__ occ = S.iblob + S.o_occ[s] + d * n;
  cd* __restrict__ a = gw + (size_t)t * n * n;
  cd* fcol = fcol_s[wib];
  int* piv = piv_s[wib];
  // M[i][k] = mo(electron lo + i, orbital occ[k])   (slater.py:239-240)
  for (int k = lane; k < n; k += 32) {
    const int orb = occ[k];
    for (int i = 0; i < n; ++i) {
      const double* row = st.mo_all + ((size_t)w * S.ne + lo + i) * ldmax;
      a[i * n + k] = {row[orb], row[off + orb]};
    }
  }
  __syncwarp();
  cd phase = {1.0, 0.0};
  double logdet = 0.0;
  bool singular = false;
  for (int c = 0; c < n; ++c) {
    // pivot: first row r >= c with the largest |a[r][c]|
    double v = -1.0;
    int p = 0x7fffffff;
    for (int r = c + lane; r < n; r += 32) {
      const double m = cabs2(a[r * n + c]);
      if (m > v) {
        v = m;
        p = r;
      }
    }
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
      phase = {-phase.x, -phase.y};
      for (int k = lane; k < n; k += 32) {
        const cd tmp = a[c * n + k];
        a[c * n + k] = a[p * n + k];
        a[p * n + k] = tmp;
      }
    }
    __syncwarp();
    const cd pv = a[c * n + c];
    const double mag = sqrt(cabs2(pv));
    if (mag == 0.0 || !isfinite(mag)) {
      singular = true;
      break;
    }
    phase = cmul(phase, {pv.x / mag, pv.y / mag});
    logdet += log(mag);
    const cd ipv = cdivc({1.0, 0.0}, pv);
    for (int r = lane; r < n; r += 32) fcol[r] = a[r * n + c];  // column c before the row scaling
    __syncwarp();
    for (int k = lane; k < n; k += 32) {
      const cd x = (k == c) ? cd{1.0, 0.0} : a[c * n + k];
      a[c * n + k] = cmul(x, ipv);
    }
    __syncwarp();
    for (int k = lane; k < n; k += 32) {
      const cd rc = a[c * n + k];
      for (int r = 0; r < n; ++r) {
        if (r == c) continue;
        const cd f = fcol[r];
        const cd x = (k == c) ? cd{0.0, 0.0} : a[r * n + k];
        a[r * n + k] = cfma({-f.x, -f.y}, rc, x);
      }
    }
    __syncwarp();
  }
  double* outr = st.inv[s] + (size_t)t * n * n;
  double* outi = st.inv_im[s] + (size_t)t * n * n;
  if (singular) {
    for (int i = lane; i < n * n; i += 32) {
      outr[i] = 0.0;
      outi[i] = 0.0;
    }
    if (lane == 0) {
      st.dsign[s][t] = 0.0;
      st.dphs_im[s][t] = 0.0;
      st.dlog[s][t] = -INFINITY;
    }
    return;
  }
  for (int c = n - 1; c >= 0; --c) {
    const int p = piv[c];
    if (p != c)
      for (int r = lane; r < n; r += 32) {  // lane = row: swap columns c and p
        const cd tmp = a[r * n + c];
        a[r * n + c] = a[r * n + p];
        a[r * n + p] = tmp;
      }
    __syncwarp();
  }
  for (int i = lane; i < n * n; i += 32) {
    outr[i] = a[i].x;
    outi[i] = a[i].y;
  }
  if (lane == 0) {
    const double m = sqrt(cabs2(phase));  // keep the phase on the unit circle
    st.dsign[s][t] = phase.x / m;
    st.dphs_im[s][t] = phase.y / m;
    st.dlog[s][t] = logdet;
  }
}

// multi-determinant caches (complex twin of k_det_cache): dv = phase exp(log - ref), W = sum c_D dv_other
__global__ void __launch_bounds__(128) k_cx_det_cache(const Sys S, const State st, const uint8_t* mask, int spin) {
  const int lane = threadIdx.x & 31;
  const int w = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (w >= st.N) return;
  if (mask && !mask[w]) return;
  for (int s = 0; s < 2; ++s) {
    if (spin >= 0 && s != spin) continue;
    const int nds = S.nds[s];
    double ref = -INFINITY;
    for (int d = lane; d < nds; d += 32) ref = fmax(ref, st.dlog[s][(size_t)w * nds + d]);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) ref = fmax(ref, __shfl_xor_sync(0xffffffffu, ref, o));
    if (!isfinite(ref)) ref = 0.0;
    if (lane == 0) st.ref[s][w] = ref;
    for (int d = lane; d < nds; d += 32) {
      const size_t q = (size_t)w * nds + d;
      const double ex = exp(st.dlog[s][q] - ref);
      st.dv[s][q] = st.dsign[s][q] * ex;
      st.dv_im[s][q] = st.dphs_im[s][q] * ex;
    }
  }
  __syncwarp();
  for (int s = 0; s < 2; ++s) {
    if (spin >= 0 && s == spin) continue;
    const int nds = S.nds[s], o = 1 - s, ndo = S.nds[o];
    const double* __restrict__ dvr = st.dv[o] + (size_t)w * ndo;
    const double* __restrict__ dvi = st.dv_im[o] + (size_t)w * ndo;
    for (int d = lane; d < nds; d += 32) {
      cd acc = {0.0, 0.0};
      const int k1 = S.grp_off[s][d + 1];
      for (int k = S.grp_off[s][d]; k < k1; ++k) {
        const int od = S.grp_other[s][k];
        acc = cfma({S.grp_coef[s][k], S.grp_coef_im[s][k]}, {dvr[od], dvi[od]}, acc);
      }
      st.W[s][(size_t)w * nds + d] = acc.x;
      st.W_im[s][(size_t)w * nds + d] = acc.y;
    }
  }
}

// wf.value(): unit phase (complex, interleaved) and log|Psi|
__global__ void __launch_bounds__(128) k_cx_value(const Sys S, const State st, int which, double* o_phase, double* o_log) {
  const double* sd;
  const int* si;
  stage_tables(S, sd, si);
  const int w = blockIdx.x * blockDim.x + threadIdx.x;
  const int N = st.N;
  if (w >= N) return;
  cd ph = {1.0, 0.0};
  double lg = 0.0;
  if (which & QMCB_SLATER) {
    if (S.ndet == 1) {
      const cd c = {S.detc[0], S.detc_im[0]};
      const double m = sqrt(cabs2(c));
      ph = cmul({st.dsign[0][w], st.dphs_im[0][w]}, {st.dsign[1][w], st.dphs_im[1][w]});
      ph = cmul(ph, {c.x / m, c.y / m});
      lg = st.dlog[0][w] + st.dlog[1][w] + log(m);
    } else {
      cd val = {0.0, 0.0};
      const int nds = S.nds[0];
      for (int d = 0; d < nds; ++d) {
        const size_t q = (size_t)w * nds + d;
        val = cfma({st.dv[0][q], st.dv_im[0][q]}, {st.W[0][q], st.W_im[0][q]}, val);
      }
      const double m = sqrt(cabs2(val));
      ph = {val.x / m, val.y / m};
      lg = log(m) + st.ref[0][w] + st.ref[1][w];
    }
    if (!cfinite(ph)) ph = {0.0, 0.0};  // np.nan_to_num in compute_value
    if (!isfinite(lg)) {
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
        for (int t = 0; t < 2; ++t) ua = fma(AVAL(st, S, w, I, k, t), sd[S.o_acoef + (I * S.na + k) * 2 + t], ua);
    lg += u + ua;
  }
  if (which & QMCB_JASTROW3) lg += st.val3[w];
  o_phase[2 * w] = ph.x;
  o_phase[2 * w + 1] = ph.y;
  o_log[w] = lg;
}

// Sherman-Morrison row replacement, one warp per (walker, spin determinant), lanes over columns (n <= 64): the new
// row comes from st.saved_mo (complex MO row at the accepted position).  slater.py:88-94, 262-291.
__global__ void __launch_bounds__(128) k_cx_sm(const Sys S, const State st, int s, int e, const uint8_t* mask) {
  __shared__ cd col_s[4][QMCB_CX_NMAX];
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  const long long t = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int nds = S.nds[s];
  if (t >= (long long)st.N * nds) return;
  const int w = (int)(t / nds), d = (int)(t - (long long)w * nds);
  if (mask && !mask[w]) return;
  const int n = s ? S.ndn : S.nup;
  const int off = S.cxoff[s];
  const int* __restrict__ occ = S.iblob + S.o_occ[s] + d * n;
  const double* __restrict__ row = st.saved_mo + (size_t)w * S.ldc[s];
  double* __restrict__ ir = st.inv[s] + (size_t)t * n * n;
  double* __restrict__ ii = st.inv_im[s] + (size_t)t * n * n;
  cd* col = col_s[wib];
  // tmp[j] = sum_k vec[k] inv[k][j]
  cd tmp[2] = {{0.0, 0.0}, {0.0, 0.0}};
  for (int k = 0; k < n; ++k) {
    const cd v = {row[occ[k]], row[off + occ[k]]};
#pragma unroll
    for (int u = 0; u < 2; ++u) {
      const int j = lane + 32 * u;
      if (j < n) tmp[u] = cfma(v, {ir[k * n + j], ii[k * n + j]}, tmp[u]);
    }
  }
  const int src = e & 31, su = e >> 5;
  const cd mine = su ? tmp[1] : tmp[0];
  const cd ratio = {__shfl_sync(0xffffffffu, mine.x, src), __shfl_sync(0xffffffffu, mine.y, src)};
  for (int k = lane; k < n; k += 32) col[k] = cdivc({ir[k * n + e], ii[k * n + e]}, ratio);
  __syncwarp();
  for (int k = 0; k < n; ++k) {
    const cd ck = col[k];
#pragma unroll
    for (int u = 0; u < 2; ++u) {
      const int j = lane + 32 * u;
      if (j < n) {
        cd x;
        if (j == e)
          x = ck;
        else
          x = cfma({-ck.x, -ck.y}, tmp[u], {ir[k * n + j], ii[k * n + j]});
        ir[k * n + j] = x.x;
        ii[k * n + j] = x.y;
      }
    }
  }
  if (lane == 0) {
    const double m = sqrt(cabs2(ratio));
    const cd ph = cmul({st.dsign[s][t], st.dphs_im[s][t]}, {ratio.x / m, ratio.y / m});
    st.dsign[s][t] = ph.x;
    st.dphs_im[s][t] = ph.y;
    st.dlog[s][t] += log(m);
  }
}

// kinetic energy pieces, thread per (walker, electron), from the cached MO rows (energy.py:57-65):
// ke = -Re(lap) / 2, grad2 = sum |grad|^2
__global__ void __launch_bounds__(128) k_cx_kinetic(const Sys S, const State st, const EnergyScratch es) {
  const double* sd;
  const int* si;
  stage_tables(S, sd, si);
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  const int N = st.N;
  if (p >= N * S.ne) return;
  const int w = p / S.ne, e = p - w * S.ne;
  const double px = CONF(st, S, w, e, 0), py = CONF(st, S, w, e, 1), pz = CONF(st, S, w, e, 2);
  const int s = e >= S.nup ? 1 : 0;
  const int n = s ? S.ndn : S.nup;
  const int eeff = e - s * S.nup;
  const int ldmax = S.ldc[0] > S.ldc[1] ? S.ldc[0] : S.ldc[1];
  const int off = S.cxoff[s];
  const double* __restrict__ mc = st.mocache + ((size_t)w * S.ne + e) * 5 * ldmax;
  const int nds = S.nds[s];
  const int* __restrict__ occ = si + S.o_occ[s];
  cd num[5], den = {0.0, 0.0};
  for (int c = 0; c < 5; ++c) num[c] = {0.0, 0.0};
  for (int d = 0; d < nds; ++d) {
    const size_t base = ((size_t)w * nds + d) * n * n + eeff;
    cd r[5];
    for (int c = 0; c < 5; ++c) r[c] = {0.0, 0.0};
    for (int k = 0; k < n; ++k) {
      const cd a = {st.inv[s][base + k * n], st.inv_im[s][base + k * n]};
      const int orb = occ[d * n + k];
      for (int c = 0; c < 5; ++c) r[c] = cfma({mc[c * ldmax + orb], mc[c * ldmax + off + orb]}, a, r[c]);
    }
    if (S.ndet == 1) {
      for (int c = 0; c < 5; ++c) num[c] = r[c];
      den = {1.0, 0.0};
    } else {
      const size_t q = (size_t)w * nds + d;
      const cd wgt = cmul({st.dv[s][q], st.dv_im[s][q]}, {st.W[s][q], st.W_im[s][q]});
      den.x += wgt.x;
      den.y += wgt.y;
      for (int c = 0; c < 5; ++c) num[c] = cfma(r[c], wgt, num[c]);
    }
  }
  const cd r0 = cdivc(num[0], den);
  cd gs[3];
  for (int i = 0; i < 3; ++i) gs[i] = cdivc(cdivc(num[1 + i], den), r0);
  const cd laps = cdivc(cdivc(num[4], den), r0);
  double du = 0.0, gj[3] = {0.0, 0.0, 0.0}, lapj = 0.0;
  if (S.na + S.nb > 0) jastrow_point<2>(S, sd, si, st, w, e, px, py, pz, du, gj, lapj);
  if (S.na3 + S.nb3 > 0) jastrow3_point<2>(S, sd, si, st, w, e, px, py, pz, du, gj, lapj);
  double lj = 0.0, cross = 0.0;
  if (S.na + S.nb + S.na3 + S.nb3 > 0) {
    lj = lapj + (gj[0] * gj[0] + gj[1] * gj[1] + gj[2] * gj[2]);
    cross = gs[0].x * gj[0] + gs[1].x * gj[1] + gs[2].x * gj[2];
  }
  const double lap_re = (laps.x + lj) + cross * 2.0;
  double g2 = 0.0;
  for (int i = 0; i < 3; ++i) {
    const double gr = gs[i].x + gj[i], gi = gs[i].y;
    g2 += gr * gr + gi * gi;
  }
  es.ke_e[(size_t)e * N + w] = -0.5 * lap_re;
  es.g2_e[(size_t)e * N + w] = g2;
}

// ECP quadrature points (complex twin of k_ecp_points<0>): the ratio is complex, so the energy contributions and
// the T-move ratios carry an imaginary array (contrib_im / tm_ratio_im)
struct EcpCxArgs {
  double* contrib_im;   // [items][max_naip]
  double* tm_ratio_im;  // [N][tot_naip]
};

__global__ void __launch_bounds__(128) k_cx_ecp_points(const Sys S, const State st, const EnergyScratch es,
                                                       const EcpPointArgs ea, const EcpCxArgs cx) {
  const double* sd;
  const int* si;
  stage_tables(S, sd, si);
  const int N = st.N;
  const int which = QMCB_SLATER | ((S.na + S.nb) > 0 ? QMCB_JASTROW : 0) | ((S.na3 + S.nb3) > 0 ? QMCB_JASTROW3 : 0);
  const int nitems = *es.count;
  const long long total = (long long)nitems * S.max_naip;
  for (long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x; p < total; p += (long long)gridDim.x * blockDim.x) {
    const int item = (int)(p / S.max_naip), q = (int)(p - (long long)item * S.max_naip);
    const int t = es.work[item];
    const int eai = t / N, w = t - eai * N;
    const int e = ea.e_only >= 0 ? ea.e_only : eai / S.necp;
    const int a = eai % S.necp;
    const int naip = si[S.o_naip + a];
    if (q >= naip) {
      if (ea.pos_out) ea.pos_out[3 * p] = NAN;
      continue;
    }
    const int atom = si[S.o_ecpatom + a];
    const double ex = CONF(st, S, w, e, 0), ey = CONF(st, S, w, e, 1), ez = CONF(st, S, w, e, 2);
    double rx = ex - sd[S.o_xyz + 3 * atom], ry = ey - sd[S.o_xyz + 3 * atom + 1], rz = ez - sd[S.o_xyz + 3 * atom + 2];
    if (S.pbc) min_image(S, sd, rx, ry, rz);
    const double r = sqrt(rx * rx + ry * ry + rz * rz);
    const double* R = ea.rot + (size_t)eai * 9;
    const double* qt = ea.quad + (size_t)si[S.o_aipoff + a] * 4;
    const double qx = qt[q * 3], qy = qt[q * 3 + 1], qz = qt[q * 3 + 2];
    const double wq = qt[naip * 3 + q];
    const double ux = R[0] * qx + R[1] * qy + R[2] * qz, uy = R[3] * qx + R[4] * qy + R[5] * qz,
                 uz = R[6] * qx + R[7] * qy + R[8] * qz;
    const double dx = r * ux, dy = r * uy, dz = r * uz;
    const double cosang = (rx * dx + ry * dy + rz * dz) / (r * sqrt(dx * dx + dy * dy + dz * dz));
    double px = (ex - rx) + dx, py = (ey - ry) + dy, pz = (ez - rz) + dz;
    const double upx = px, upy = py, upz = pz;
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
    PointEvalCx<0> ev;
    ev.run(S, sd, si, st, which, w, e, px, py, pz, nullptr, ea.scr + p, ea.scr_stride);
    const double exu = exp(ev.du);
    const cd ratio = {ev.rat[0].x * exu, ev.rat[0].y * exu};
    const int nlm1 = si[S.o_chanoff + a + 1] - si[S.o_chanoff + a] - 1;
    if (ea.tmove_tau <= 0.0) {
      double acc = 0.0;
      for (int l = 0; l < nlm1; ++l)
        acc += es.vls[(size_t)item * es.maxchan + l] * ((double)(2 * l + 1) * legendre_p(l, cosang) * wq);
      es.contrib[(size_t)item * S.max_naip + q] = ratio.x * acc;
      cx.contrib_im[(size_t)item * S.max_naip + q] = ratio.y * acc;
    } else {
      double wt = 0.0;
      for (int l = 0; l < nlm1; ++l)
        wt += (exp(-ea.tmove_tau * es.vls[(size_t)item * es.maxchan + l]) - 1.0) *
              ((double)(2 * l + 1) * legendre_p(l, cosang) * wq);
      const size_t o = (size_t)w * S.tot_naip + (size_t)si[S.o_aipoff + a] + q;
      ea.tm_ratio[o] = ratio.x;
      cx.tm_ratio_im[o] = ratio.y;
      ea.tm_weight[o] = wt;
      ea.tm_pos[o * 3] = upx;
      ea.tm_pos[o * 3 + 1] = upy;
      ea.tm_pos[o * 3 + 2] = upz;
    }
  }
}

// imaginary part of the ECP energy per walker, summed in the order k_energy_finalize uses for the real part:
// out [2][N] = Im ecp, Im total (the other terms of the local energy are real).  One warp per walker: lanes over the
// (electron, ECP atom) items, lane 0 adds them in the reference's order; dynamic shared memory: ne * necp doubles per warp.
__global__ void __launch_bounds__(128) k_cx_ecp_imag(const Sys S, const State st, const EnergyScratch es,
                                                     const double* __restrict__ contrib_im, double* __restrict__ out) {
  extern __shared__ __align__(128) unsigned char qmcb_smem[];
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  const int w = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int N = st.N;
  if (w >= N) return;
  const int nitem = S.ne * S.necp;
  double* buf = reinterpret_cast<double*>(qmcb_smem) + (size_t)wib * nitem;
  for (int t = lane; t < nitem; t += 32) {
    const int a = t % S.necp;
    const int item = es.item_of[(size_t)t * N + w];
    double nl = 0.0;
    if (item >= 0) {
      const int naip = S.iblob[S.o_naip + a];
      for (int q = 0; q < naip; ++q) nl += contrib_im[(size_t)item * S.max_naip + q];
    }
    buf[t] = nl;
  }
  __syncwarp();
  if (lane == 0) {
    double ecp = 0.0;
    for (int e = 0; e < S.ne; ++e) {
      double ecp_e = 0.0;
      for (int a = 0; a < S.necp; ++a) ecp_e += buf[e * S.necp + a];
      ecp += ecp_e;
    }
    out[w] = ecp;
    out[(size_t)N + w] = ecp;
  }
}

// d ln Psi / d det_coeff [N][ndet] (complex, interleaved) and G_s[d] = sum_{D: map_s(D)=d} c_D dPsi_D (complex)
__global__ void __launch_bounds__(128) k_cx_pgrad_det(const Sys S, const State st, cd* __restrict__ out,
                                                      cd* __restrict__ G, int gstride) {
  const int w = blockIdx.x * blockDim.x + threadIdx.x;
  const int N = st.N;
  if (w >= N) return;
  cd ph;
  double logpsi;
  if (S.ndet == 1) {
    const cd c = {S.detc[0], S.detc_im[0]};
    const double m = sqrt(cabs2(c));
    ph = cmul(cmul({st.dsign[0][w], st.dphs_im[0][w]}, {st.dsign[1][w], st.dphs_im[1][w]}), {c.x / m, c.y / m});
    logpsi = st.dlog[0][w] + st.dlog[1][w] + log(m);
  } else {
    cd val = {0.0, 0.0};
    for (int d = 0; d < S.nds[0]; ++d) {
      const size_t q = (size_t)w * S.nds[0] + d;
      val = cfma({st.dv[0][q], st.dv_im[0][q]}, {st.W[0][q], st.W_im[0][q]}, val);
    }
    const double m = sqrt(cabs2(val));
    ph = {val.x / m, val.y / m};
    logpsi = log(m) + st.ref[0][w] + st.ref[1][w];
  }
  for (int s = 0; s < 2; ++s)
    for (int d = 0; d < S.nds[s]; ++d) G[((size_t)s * N + w) * gstride + d] = {0.0, 0.0};
  const bool ok = cfinite(ph) && (ph.x != 0.0 || ph.y != 0.0);
  for (int D = 0; D < S.ndet; ++D) {
    const int d0 = S.map[0][D], d1 = S.map[1][D];
    cd v = {0.0, 0.0};
    if (ok) {
      const size_t q0 = (size_t)w * S.nds[0] + d0, q1 = (size_t)w * S.nds[1] + d1;
      const double ex = exp(st.dlog[0][q0] + st.dlog[1][q1] - logpsi);
      const cd dd = cmul({st.dsign[0][q0], st.dphs_im[0][q0]}, {st.dsign[1][q1], st.dphs_im[1][q1]});
      v = cdivc({dd.x * ex, dd.y * ex}, ph);
    }
    out[(size_t)w * S.ndet + D] = v;
    const cd cv = cmul({S.detc[D], S.detc_im[D]}, v);
    cd& g0 = G[((size_t)0 * N + w) * gstride + d0];
    cd& g1 = G[((size_t)1 * N + w) * gstride + d1];
    g0 = {g0.x + cv.x, g0.y + cv.y};
    g1 = {g1.x + cv.x, g1.y + cv.y};
  }
}

// d ln Psi / d mo_coeff_s [N][A][nmo_t] (complex): sum_d G_s[d] sum_e ao[e][a] inv[d][col_d(i)][e].
// open: ao real [N][ne][A]; periodic: complex planes ao[((p * 2 + re/im) * nk + k) * A + a]
__global__ void __launch_bounds__(128) k_cx_pgrad_mo(const Sys S, const State st, int s, const double* __restrict__ ao,
                                                     const cd* __restrict__ G, int gstride, cd* __restrict__ out) {
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const int N = st.N, nmo = S.nmo_t[s], A = S.nao;
  if (t >= (long long)N * A * nmo) return;
  const int i = (int)(t % nmo);
  const int a = (int)((t / nmo) % A);
  const int w = (int)(t / ((long long)nmo * A));
  const int n = s ? S.ndn : S.nup, lo = s ? S.nup : 0, nds = S.nds[s];
  const int* __restrict__ occ = S.iblob + S.o_occ[s];
  cd acc = {0.0, 0.0};
  for (int d = 0; d < nds; ++d) {
    int col = -1;
    for (int k = 0; k < n; ++k)
      if (occ[d * n + k] == i) col = k;
    if (col < 0) continue;
    const size_t base = (((size_t)w * nds + d) * n + col) * n;
    cd v = {0.0, 0.0};
    if (S.pbc) {
      const int k = S.iblob[S.o_mok[s] + i];
      for (int e = 0; e < n; ++e) {
        const size_t p = (size_t)w * S.ne + lo + e;
        const cd x = {ao[((p * 2) * S.nk + k) * A + a], ao[((p * 2 + 1) * S.nk + k) * A + a]};
        v = cfma(x, {st.inv[s][base + e], st.inv_im[s][base + e]}, v);
      }
    } else {
      for (int e = 0; e < n; ++e) {
        const double x = ao[((size_t)w * S.ne + lo + e) * A + a];
        v.x = fma(x, st.inv[s][base + e], v.x);
        v.y = fma(x, st.inv_im[s][base + e], v.y);
      }
    }
    acc = cfma(G[((size_t)s * N + w) * gstride + d], v, acc);
  }
  out[t] = acc;
}

// -----------------------------------------------------------------------------------------
// Device-resident VMC block of a COMPLEX wave function (mc.py:112-137): the per-electron loop chains the query kernels
// above without host round trips.  Per electron: k_cx_chain<0> gathers the electron's positions, k_cx_point<GRADVAL>
// gives the drift there, k_cx_chain<1> limits it and writes the (wrapped) proposal, [k_pbc_mo rows +]
// k_cx_point<GRADVAL> with save gives the drift, the ratio and the MO row there, k_cx_chain<2> runs the Metropolis test
// on |ratio|^2 (np.abs of a complex number: hypot), and the update kernels (launch_update: k_cx_sm, k_cx_det_cache,
// Jastrow / three-body caches, coordinates and wrap vectors) commit the accepted walkers.
// -----------------------------------------------------------------------------------------
struct CxChainArgs {
  int e;
  double tstep;
  const double* gauss;  // [N][3]
  const double* unif;   // [N]
  uint8_t* accept;      // [N]
  unsigned long long* nacc;
  double* pos;          // [N][3] staging of the query positions
  double* pwrap;        // [N][3] periodic: wrap vectors of the proposal (= saved_wrap)
  const cd* grad;       // [3][N] output of k_cx_point
  const cd* val;        // [N]
};

template <int PHASE>
__global__ void __launch_bounds__(128) k_cx_chain(const Sys S, const State st, const CxChainArgs a) {
  const double* sd;
  const int* si;
  stage_tables(S, sd, si);
  const int w = blockIdx.x * blockDim.x + threadIdx.x;
  const int N = st.N;
  if (w >= N) return;
  if (PHASE == 0) {
#pragma unroll
    for (int i = 0; i < 3; ++i) a.pos[(size_t)w * 3 + i] = CONF(st, S, w, a.e, i);
    if (S.pbc) {
#pragma unroll
      for (int i = 0; i < 3; ++i) a.pwrap[(size_t)w * 3 + i] = st.wrap[((size_t)w * S.ne + a.e) * 3 + i];
    }
    return;
  }
  double grad[3];
#pragma unroll
  for (int i = 0; i < 3; ++i) grad[i] = a.grad[(size_t)i * N + w].x;  // np.real(g.T)
  limdrift3(grad);
  const double* __restrict__ gauss = a.gauss + (size_t)w * 3;
  if (PHASE == 1) {
    const double nx = __dadd_rn(__dadd_rn(a.pos[(size_t)w * 3], gauss[0]), __dmul_rn(grad[0], a.tstep));
    const double ny = __dadd_rn(__dadd_rn(a.pos[(size_t)w * 3 + 1], gauss[1]), __dmul_rn(grad[1], a.tstep));
    const double nz = __dadd_rn(__dadd_rn(a.pos[(size_t)w * 3 + 2], gauss[2]), __dmul_rn(grad[2], a.tstep));
    double o[3] = {nx, ny, nz}, ww[3] = {0.0, 0.0, 0.0};
    if (S.pbc) wrap_cell(sd + S.o_lat, sd + S.o_latinv, nx, ny, nz, o, ww);
#pragma unroll
    for (int i = 0; i < 3; ++i) {
      a.pos[(size_t)w * 3 + i] = o[i];
      if (S.pbc) a.pwrap[(size_t)w * 3 + i] = st.wrap[((size_t)w * S.ne + a.e) * 3 + i] + ww[i];
      st.gold[(size_t)w * 3 + i] = grad[i];
    }
    return;
  }
  double fwd = 0.0, bwd = 0.0;
#pragma unroll
  for (int i = 0; i < 3; ++i) {
    fwd = __dadd_rn(fwd, __dmul_rn(gauss[i], gauss[i]));
    const double b = __dadd_rn(gauss[i], __dmul_rn(a.tstep, __dadd_rn(st.gold[(size_t)w * 3 + i], grad[i])));
    bwd = __dadd_rn(bwd, __dmul_rn(b, b));
  }
  const double tprob = exp(__dmul_rn(1.0 / (2.0 * a.tstep), __dadd_rn(fwd, -bwd)));
  const cd v = a.val[w];
  const double aval = hypot(v.x, v.y);
  const double ratio = __dmul_rn(__dmul_rn(aval, aval), tprob);
  const bool acc = ratio > a.unif[w];
  a.accept[w] = acc ? 1 : 0;
  const unsigned b = __ballot_sync(__activemask(), acc);
  if (acc && (threadIdx.x & 31) == (__ffs(b) - 1)) atomicAdd(a.nacc, (unsigned long long)__popc(b));
}
