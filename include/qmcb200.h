/*
 * qmcb200.h -- C ABI of libqmcb200.so: B200-native (sm_100a) evaluation of the PyQMC
 * walker-batched trial-wave-function hot path.
 *
 * The reference (WagnerGroup/pyqmc) has no FFI: its boundary for this path is the
 * duck-typed Python wave-function protocol (doc/source/wavefunction.rst:1-37) and the
 * EnergyAccumulator functor (pyqmc/observables/accumulators.py:45-95).  Each entry point
 * below names the reference method it replaces; the ctypes stub a maintainer would add is in
 * INTEGRATION.md.  Conventions:
 *   - every function returns 0 on success, <0 on error; qmcb_last_error() gives the text;
 *   - the caller owns all host buffers (C-contiguous float64 / int32 / uint8); the library
 *     owns device memory; calls are synchronous (stream-synchronised before returning)
 *     unless the name ends in _async;
 *   - one context per (device, wave function); a context is not thread-safe;
 *   - `which` selects the factors a call applies to: QMCB_SLATER | QMCB_JASTROW.  With both
 *     bits the call has MultiplyWF semantics (pyqmc/wf/multiplywf.py:71-132): log-values
 *     and gradients add, ratios multiply, Laplacian gets the 2 grad_i.grad_j cross term.
 *   - walkers: configs[N][nelec][3]; electrons 0..nup-1 are spin up (slater.py:236-239).
 */
#ifndef QMCB200_H
#define QMCB200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct qmcb_ctx qmcb_ctx;

#define QMCB_SLATER 1
#define QMCB_JASTROW 2
#define QMCB_JASTROW3 4

/* Jastrow radial function kinds (pyqmc/wf/func3d.py:52-110 PolyPade, 112-210 CutoffCusp) */
#define QMCB_FUNC_POLYPADE 0
#define QMCB_FUNC_CUSP 1

const char *qmcb_last_error(void);
int qmcb_device_count(void);

int qmcb_create(int device, qmcb_ctx **out);
void qmcb_destroy(qmcb_ctx *ctx);

/* ---- system description ------------------------------------------------------------ */

/* mol.atom_coords() / mol.atom_charges() (energy.py:38-44, jastrowspin.py:52) */
int qmcb_set_atoms(qmcb_ctx *ctx, int natom, const double *xyz /*[natom*3]*/,
                   const double *charges /*[natom]*/);

/* Flattened, NORMALISED shell table: what AtomicOrbitalEvaluator.__init__ builds
 * (pyqmc/wf/numba/gto.py:435-488).  Shells must be grouped by atom, atoms ascending.
 * AO order: shells in order, 2l+1 functions per shell, l <= 4. */
int qmcb_set_basis(qmcb_ctx *ctx, int nshell, const int32_t *shell_atom, const int32_t *shell_l,
                   const int32_t *prim_off /*[nshell+1]*/, const double *prim_exp,
                   const double *prim_coef);

/* Slater.__init__ products (slater.py:193-208; determinant_tools.py:39-71):
 * mo_s [nao][nmo_s] row-major, occ_s [ndet_s][n_s] MO indices of each unique spin
 * determinant, map_s [ndet] -> unique spin-determinant, det_coeff [ndet]. */
int qmcb_set_slater(qmcb_ctx *ctx, int nup, int ndn, int nmo_up, const double *mo_up,
                    int nmo_dn, const double *mo_dn, int ndet_up, const int32_t *occ_up,
                    int ndet_dn, const int32_t *occ_dn, int ndet, const int32_t *map_up,
                    const int32_t *map_dn, const double *det_coeff);

/* COMPLEX wave functions: complex MO and / or determinant coefficients (Slater.__init__ sets
 * dtype = complex when orbitals.mo_dtype == complex or any parameter is complex, slater.py:212-216;
 * MoleculeOrbitalEvaluator / PBCOrbitalEvaluatorKpoints take mo_dtype from the coefficients,
 * orbitals.py:61-65, 160-165).  Same arguments as qmcb_set_slater with the real and imaginary parts
 * separate.  A context set up this way is COMPLEX (qmcb_is_complex): for every call whose `which`
 * includes QMCB_SLATER the wave-function-valued outputs -- sign (the unit phase, get_complex_phase
 * slater.py:32-33), gradient, value / ratio, laplacian, testvalue(_many) ratios, T-move ratios,
 * pgradient of det_coeff / mo_coeff_*, the "inverse_*" / "dets_*" state arrays -- are complex128
 * (interleaved re, im) of the same shape; log values stay real.  qmcb_energy returns 8 rows (see
 * there).  qmcb_vmc_block / qmcb_vmc_block_device run complex contexts too (query kernels chained on
 * the device, 8 energy rows per step); the pipelined slots (qmcb_vmc_block_slot*), qmcb_dmc_block*
 * and qmcb_sr_avg are real-only and fail on a complex context: the reference's drivers run those
 * through the protocol. */
int qmcb_set_slater_cx(qmcb_ctx *ctx, int nup, int ndn, int nmo_up, const double *mo_up_re,
                       const double *mo_up_im, int nmo_dn, const double *mo_dn_re,
                       const double *mo_dn_im, int ndet_up, const int32_t *occ_up, int ndet_dn,
                       const int32_t *occ_dn, int ndet, const int32_t *map_up,
                       const int32_t *map_dn, const double *det_re, const double *det_im);
/* 1 when the context holds a complex Slater factor */
int qmcb_is_complex(qmcb_ctx *ctx);

/* JastrowSpin.__init__ (jastrowspin.py:34-54): a (electron-ion) and b (electron-electron)
 * radial bases with one shared cutoff each, acoeff [natom][na][2], bcoeff [nb][3]. */
int qmcb_set_jastrow(qmcb_ctx *ctx, int nup, int ndn, int na, const int32_t *a_kind,
                     const double *a_par, double rcut_a, int nb, const int32_t *b_kind,
                     const double *b_par, double rcut_b, const double *acoeff,
                     const double *bcoeff);

/* ThreeBodyJastrow.__init__ (three_body_jastrow.py:45-64): a / b radial bases and
 * ccoeff [natom][na][na][nb][3]; the library symmetrises it in (k,l) as recompute does (94-96). */
int qmcb_set_jastrow3(qmcb_ctx *ctx, int nup, int ndn, int na, const int32_t *a_kind,
                      const double *a_par, double rcut_a, int nb, const int32_t *b_kind,
                      const double *b_par, double rcut_b, const double *ccoeff);

/* mol._ecp flattened (eval_ecp.py:160-200): for each ECP atom the channel list in COLUMN
 * order l = 0..lmax then the local channel (l = -1) last; channel c owns terms
 * term_off[c]..term_off[c+1] of v(r) = sum coef * r^power * exp(-alpha r^2). */
int qmcb_set_ecp(qmcb_ctx *ctx, int necp, const int32_t *ecp_atom /*[necp]*/,
                 const int32_t *chan_off /*[necp+1]*/, const int32_t *term_off /*[nchan+1]*/,
                 const int32_t *term_power, const double *term_alpha, const double *term_coef,
                 const int32_t *naip /*[necp]*/,
                 const double *quad /* per ECP atom: naip*3 points then naip weights
                                       (eval_ecp.py:278-336) */,
                 double threshold);

/* ---- periodic systems ------------------------------------------------------------------ */

/* Simulation-cell lattice vectors (rows, Bohr) and the minimal-image convention of
 * MinimalImageDistance (pyqmc/configurations/distance.py:83-121): mode 1 diagonal, 2 orthogonal,
 * 3 general; shifts [27][3] = the reference's point_list . latvec (same order: np.argmin ties). */
int qmcb_set_lattice(qmcb_ctx *ctx, const double *lat /*[9]*/, int mode, const double *shifts /*[81]*/);

/* Tables of PBCOrbitalEvaluatorKpoints / PeriodicAtomicOrbitalEvaluator (pyqmc/wf/orbitals.py:118-186,
 * pyqmc/wf/numba/pbcgto.py:594-621): the shell table of qmcb_set_basis then refers to the nbatom
 * PRIMITIVE-cell atoms bxyz; lprim = primitive lattice, smat = supercell matrix S, kpts [nk][3],
 * Ls [nL][3] images sorted by norm, num_Ls [nbatom], r^2 cutoffs per atom / per shell (max_Ls,
 * pbcgto.py:551-591), phases [nL][nk] = Re exp(i Ls.k) (imaginary parts: qmcb_set_pbc_phases_imag),
 * mo_k_* = k-point index of every MO column of the concatenated coefficient matrices
 * (orbitals.py:155-160). */
int qmcb_set_pbc_orbitals(qmcb_ctx *ctx, int nbatom, const double *bxyz, const double *lprim,
                          const double *smat, int nk, const double *kpts, int nL, const double *Ls,
                          const int32_t *num_Ls, const double *atom_cutoff, int nshell,
                          const double *l_cutoff, const double *phases, int nmo_up,
                          const int32_t *mo_k_up, int nmo_dn, const int32_t *mo_k_dn, int isgamma);

/* General twists: Im exp(i Ls.k) [nL][nk] (PeriodicAtomicOrbitalEvaluator.phases, pbcgto.py:620-621, complex
 * unless every k-point is real) after qmcb_set_pbc_orbitals passed the real parts.  Used by complex
 * contexts (qmcb_set_slater_cx), whose wrap phase is exp(i k.R) (get_wrapphase_complex, orbitals.py:38-39). */
int qmcb_set_pbc_phases_imag(qmcb_ctx *ctx, int nL, int nk, const double *phases_im);

/* Ewald.__init__ products (pyqmc/observables/ewald.py:93-200): alpha, real-space displacements
 * [ndisp][3], selected reciprocal points / weights, ion structure factor, the pair / square
 * constants, sum of charges and the ion-ion energy (ion_ion + ii_const). */
int qmcb_set_ewald(qmcb_ctx *ctx, double alpha, int ndisp, const double *disp, int nG,
                   const double *gpoints, const double *gweight, const double *ion_re,
                   const double *ion_im, double ijconst, double squareconst, double i_sum,
                   double e_ii);

/* Wrap vectors (PeriodicElectron.wrap, pyqmc/configurations/coord.py:115-134) of the positions
 * passed to the NEXT point call (gradient*, testvalue*, updateinternals): [count][3], count = N
 * or N * naip.  Consumed by that call. */
int qmcb_set_point_wrap(qmcb_ctx *ctx, const double *wrap, int64_t count);

/* wf.recompute(configs) for PeriodicConfigs: wrap [N][nelec][3] (NULL = zero). */
int qmcb_recompute_pbc(qmcb_ctx *ctx, int which, int nconf, const double *configs,
                       const double *wrap, double *sign, double *logval);

/* wf.recompute(configs) for the walkers the device ALREADY holds (the configurations a
 * device-resident block just returned; mc.py:110 calls wf.recompute at the start of every block):
 * same kernels as qmcb_recompute without the upload and the value read-back; asynchronous on the
 * context's stream.  Valid only while nothing else changed the walker state. */
int qmcb_recompute_resident(qmcb_ctx *ctx, int which);
/* the same, enqueued on a caller-supplied cudaStream_t (NULL = the context's stream) */
int qmcb_recompute_resident_on(qmcb_ctx *ctx, int which, void *stream);

/* ---- wave-function protocol ---------------------------------------------------------- */

/* wf.recompute(configs) -> (sign, log|psi|)   slater.py:227-260, jastrowspin.py:56-109 */
int qmcb_recompute(qmcb_ctx *ctx, int which, int nconf, const double *configs, double *sign,
                   double *logval);
/* wf.value()   slater.py:293-299, jastrowspin.py:251-255 */
int qmcb_value(qmcb_ctx *ctx, int which, double *sign, double *logval);
/* wf.gradient(e, epos) -> grad[3][N]   slater.py:390-401, jastrowspin.py:258-294 */
int qmcb_gradient(qmcb_ctx *ctx, int which, int e, const double *epos /*[N][3]*/, double *grad);
/* wf.gradient_value(e, epos) -> grad[3][N], val[N]; keeps the MO row / position in the
 * context's saved slot (returned token in *slot) for a following updateinternals
 * slater.py:403-418, jastrowspin.py:296-340 */
int qmcb_gradient_value(qmcb_ctx *ctx, int which, int e, const double *epos, double *grad,
                        double *val, int64_t *slot);
/* wf.gradient_laplacian(e, epos) -> grad[3][N], lap[N] = lap(psi)/psi
 * slater.py:420-427, jastrowspin.py:342-385, multiplywf.py:121-129 */
int qmcb_gradient_laplacian(qmcb_ctx *ctx, int which, int e, const double *epos, double *grad,
                            double *lap);
/* wf.testvalue(e, epos, mask) -> ratio[Nm][naip]; epos is [N][naip][3] (rows of unmasked
 * walkers are ignored); mask NULL = all.  slater.py:429-446, jastrowspin.py:387-419 */
int qmcb_testvalue(qmcb_ctx *ctx, int which, int e, const double *epos, int naip,
                   const uint8_t *mask, double *ratio, int64_t *slot);
/* wf.testvalue_many(e[], epos, mask) -> ratio[Nm][ne_list]  slater.py:448-460 */
int qmcb_testvalue_many(qmcb_ctx *ctx, int which, int ne_list, const int32_t *elist,
                        const double *epos /*[N][3]*/, const uint8_t *mask, double *ratio);
/* wf.updateinternals(e, epos, configs, mask, saved_values): slot = token returned by
 * gradient_value/testvalue for the same (e, epos), or -1 to re-evaluate the orbitals
 * slater.py:262-291 (Sherman-Morrison 88-94), jastrowspin.py:111-137, 221-249 */
int qmcb_updateinternals(qmcb_ctx *ctx, int which, int e, const double *epos,
                         const uint8_t *mask, int64_t slot);
/* wf.pgradient(): name in {"det_coeff","mo_coeff_alpha","mo_coeff_beta","acoeff","bcoeff"}
 * slater.py:462-542, jastrowspin.py:457-464.  out is [N][...param shape]. */
int qmcb_pgradient(qmcb_ctx *ctx, const char *name, double *out);

/* internal state read-back for tests: "inverse_up","inverse_dn" [N][D_s][n][n];
 * "dets_up","dets_dn" [2][N][D_s]; "a_partial" [ne][N][I][na]; "b_partial" [ne][N][nb][2];
 * "avalues" [N][I][na][2]; "bvalues" [N][nb][3]; "configs" [N][ne][3]; "wrap" [N][ne][3] */
int qmcb_get_state(qmcb_ctx *ctx, const char *name, double *out);

/* ---- local energy (EnergyAccumulator.__call__, accumulators.py:60-75) ------------------ */
/* Random variates are drawn by the caller in the reference's order (eval_ecp.py:145, 263):
 * ecp_u [ne][necp][N] uniform numbers for the stochastic channel mask, ecp_rot
 * [ne][necp][3][3] rotation matrices.  out [6][N] = ke, ee, ei, ecp, grad2, total.
 * Complex contexts: out [8][N], rows 6 and 7 = Im ecp, Im total (ke = -Re(lap)/2 and grad2 =
 * sum |grad|^2 are real, energy.py:63-64; the ECP values carry wf.dtype, eval_ecp.py:26). */
int qmcb_energy(qmcb_ctx *ctx, const double *ecp_u, const double *ecp_rot, double *out);

/* eval_ecp.compute_tmoves (eval_ecp.py:43-80) for electron e:
 * ratio[N][M], weight[N][M], epos[N][M][3], M = sum of naip over ECP atoms. */
int qmcb_tmoves(qmcb_ctx *ctx, int e, double tau, const double *ecp_u /*[necp][N]*/,
                const double *ecp_rot /*[necp][9]*/, double *ratio, double *weight,
                double *epos);

/* ---- device-resident VMC block (vmc_worker, pyqmc/method/mc.py:102-153) ---------------- */
/* Runs nsteps sweeps (+ local energy after each sweep when with_energy) on the walkers
 * currently held by the context (after qmcb_recompute).  Random variates in reference
 * order: gauss [nsteps][ne][N][3] ~ N(0, tstep), unif [nsteps][ne][N], ecp_u
 * [nsteps][ne][necp][N], ecp_rot [nsteps][ne][necp][9].
 * Outputs (any may be NULL): configs [N][ne][3] final positions; accept [nsteps][ne][N]
 * (uint8); energy [nsteps][6][N]; esum [nsteps][6] walker sums; nacc [nsteps][ne].
 * Complex contexts: energy [nsteps][8][N], esum [nsteps][8] (rows as qmcb_energy).
 * Every wave function of the path is served: open or periodic boundaries (periodic: the wrap
 * vectors move with the walkers, read them back with qmcb_get_state "wrap"), single- or
 * multi-determinant, with or without the three-body factor, real or complex. */
int qmcb_vmc_block(qmcb_ctx *ctx, int nsteps, double tstep, int with_energy,
                   const double *gauss, const double *unif, const double *ecp_u,
                   const double *ecp_rot, double *configs, uint8_t *accept, double *energy,
                   double *esum, int64_t *nacc);

/* Same block with all inputs/outputs already in device memory and no host sync: used by
 * bench.py to time the HBM-resident path.  Pointers are device pointers. */
int qmcb_vmc_block_device(qmcb_ctx *ctx, int nsteps, double tstep, int with_energy,
                          const double *d_gauss, const double *d_unif, const double *d_ecp_u,
                          const double *d_ecp_rot, uint8_t *d_accept, double *d_energy,
                          double *d_esum, int64_t *d_nacc, void *stream);
/* Pipelined variant: qmcb_vmc_upload copies one block's variates (pinned host buffers) into device
 * slot 0..2 asynchronously on a copy stream -- callable from the host thread that draws them while
 * another thread runs qmcb_vmc_block_slot on the other slot; _slot waits for the upload event. */
int qmcb_vmc_upload(qmcb_ctx *ctx, int slot, int nsteps, int ne, int64_t N, int necp,
                    const double *gauss, const double *unif, const double *ecp_u,
                    const double *ecp_rot);
int qmcb_vmc_block_slot(qmcb_ctx *ctx, int slot, int nsteps, double tstep, int with_energy,
                        double *configs, uint8_t *accept, double *energy, double *esum,
                        int64_t *nacc);
int qmcb_kernel_launches(qmcb_ctx *ctx, int64_t *count); /* launches issued so far */

/* ---- StochasticReconfiguration.avg (pyqmc/observables/stochastic_reconfiguration.py:85-118) ----
 * Parameter j of the serialised gradient (LinearTransform.serialize_gradients, accumulators.py:161-172)
 * is element off[j] of one walker's block of source array src[j]: 0 det_coeff [ndet], 1 / 2
 * mo_coeff_alpha / beta [nao][nmo_s], 3 acoeff [natom][na][2], 4 bcoeff [nb][3], 5 ccoeff.
 * weights [N] normalised to sum 1; ECP variates as qmcb_energy.  Outputs: energy_avg [6] weighted means
 * of ke, ee, ei, ecp, grad2, total; dpH [P], dppsi [P], dpidpj [P][P]. */
int qmcb_sr_avg(qmcb_ctx *ctx, int nparam, const int32_t *src, const int64_t *off,
                const double *weights, const double *ecp_u, const double *ecp_rot,
                double nodal_cutoff, double *energy_avg, double *dpH, double *dppsi, double *dpidpj);

/* ---- device-resident DMC propagation (dmc_propagate, pyqmc/method/dmc.py:123-221) ------------
 * nsteps steps without branching on the walkers held by the context (after qmcb_recompute):
 * initial local energy, then per step T-moves of every electron (propose_tmoves 73-120), the
 * drift-diffusion sweep with Umrigar drift limit and fixed-node rejection (49-70), the local
 * energy and the weight update (compute_S 224-235).  Any REAL wave function of the path: open or
 * periodic boundaries, single- or multi-determinant, with or without the three-body factor
 * (periodic T-moves are wrapped into the cell twice, as propose_tmoves does, dmc.py:110).
 * Random variates in the reference's consumption order:
 *   ecp_u [nsteps+1][ne][necp][N], ecp_rot [nsteps+1][ne][necp][9]  energy evaluations (first = before step 0)
 *   tm_u [nsteps][ne][necp][N], tm_rot [nsteps][ne][necp][9]         nonlocal_tmoves masks / rotations
 *   tm_sel [nsteps][ne][N]  select_walker's rand();  tm_acc [nsteps][ne][N]  T-move acceptance
 *   gauss [nsteps][ne][N][3] ~ N(0, tstep), unif [nsteps][ne][N]     drift-diffusion
 * weights [N] in/out; configs [N][ne][3] out; wsums [nsteps][8] = sum_w w*(ke,ee,ei,ecp,grad2,total),
 * sum_w w, 0; nacc / ntacc [nsteps][ne] accepted diffusion / T-moves. */
int qmcb_dmc_block(qmcb_ctx *ctx, int nsteps, double tstep, double branchcut, double e_trial,
                   double e_est, const double *gauss, const double *unif, const double *ecp_u,
                   const double *ecp_rot, const double *tm_u, const double *tm_rot,
                   const double *tm_sel, const double *tm_acc, double *weights, double *configs,
                   double *wsums, int64_t *nacc, int64_t *ntacc);
/* page-locked host buffers for the per-block variates / results (true async H2D/D2H) */
int qmcb_pinned_alloc(int64_t bytes, void **out);
int qmcb_pinned_free(void *p);

/* ---- Sherman-Morrison kernel on its own (roofline measurement / unit test) --------------
 * sherman_morrison_ms (slater.py:88-94) on device arrays: inv [M][n][n], vec [M][n],
 * mask [M] or NULL, ratio [M].  Launches the same kernel updateinternals uses. */
int qmcb_sm_update_device(int n, int e, int64_t nmat, double *d_inv, const double *d_vec,
                          const uint8_t *d_mask, double *d_ratio, void *stream);
/* host-buffer convenience wrapper (copies in, runs, copies out) */
int qmcb_sm_update(int n, int e, int64_t nmat, double *inv, const double *vec,
                   const uint8_t *mask, double *ratio);

/* ---- host-side random variates of one VMC block, bit-identical to the reference's use of the
 * global legacy numpy RandomState (MT19937): per step and electron normal(scale, (N,3)) then
 * rand(N) (mc.py:119,132); then per electron and ECP atom random(N) and one scipy
 * Rotation.random() (eval_ecp.py:145,263).  key[624]/pos/has_gauss/cached_gauss are the fields
 * of np.random.get_state() and are advanced in place.  ecp_u == NULL skips the ECP draws.
 * The MT19937 stream is walked sequentially; the log/sqrt of the accepted polar pairs runs on
 * nthreads host threads. */
int qmcb_rng_vmc_block(uint32_t *key, int32_t *pos, int32_t *has_gauss, double *cached_gauss,
                       int nsteps, int ne, int64_t N, int necp, double scale, double *gauss,
                       double *unif, double *ecp_u, double *ecp_rot, int nthreads);

/* generic draw program on the same bit-identical generator: op i draws count[i] values into the
 * host buffer at address dst[i]; kind[i] = 0 uniform [0,1) doubles, 1 normals times scale[i], 2 one
 * scipy Rotation.random() matrix (dst[i][9]).  Used for the DMC block (pyqmc/method/dmc.py:150-198). */
int qmcb_rng_program(uint32_t *key, int32_t *pos, int32_t *has_gauss, double *cached_gauss,
                     int64_t nops, const int32_t *kind, const int64_t *count, const uint64_t *dst,
                     const double *scale, int nthreads);

/* two-stage form of the same generator: phase A (sequential walk of the MT19937 stream; fills the
 * uniform outputs, records the accepted polar pairs, advances the state) and phase B (log/sqrt of
 * the pairs and the Gaussian outputs, nthreads host threads) on a plan handle, so phase A of the
 * next block can overlap phase B of the current one. */
void *qmcb_rng_plan_create(void);
void qmcb_rng_plan_destroy(void *plan);
int qmcb_rng_phase_a(void *plan, uint32_t *key, int32_t *pos, int32_t *has_gauss,
                     double *cached_gauss, int nsteps, int ne, int64_t N, int necp, double scale,
                     double *gauss, double *unif, double *ecp_u, double *ecp_rot, int nthreads);
int qmcb_rng_phase_b(void *plan, int nthreads);

/* Host helper (no device): the stochastic comb of the DMC branching step, `branch` (pyqmc/method/dmc.py:358-366) --
 * picked = searchsorted(cumsum(weights), (offset * W + linspace(0, W, n, endpoint=False)) % W), W = sum of the
 * weights -- in the reference's floating-point arithmetic, as one linear pass.  picked [n], *total = W. */
int qmcb_comb_indices(int64_t n, const double *weights, double offset, int64_t *picked, double *total);

/* The dense product of StochasticReconfiguration.avg on its own (stochastic_reconfiguration.py:110-113,
 * einsum "ij,ik->jk" of dp with weights * dp_regularized): C [P][P] = A^T B for host arrays A, B [N][P].
 * variant 0 = FP64-FMA tiles, 1 = DMMA (mma.sync m8n8k4 f64) with a split walker range, -1 = the default
 * qmcb_sr_avg uses.  Runs `reps` timed launches after one warm-up and returns the mean time in *ms
 * (tests / profiles: the tensor-pipe evidence of the path's one GEMM). */
int qmcb_gemm_tn(int device, int64_t N, int P, const double *A, const double *B, double *C,
                 int variant, int reps, double *ms);

/* measured FP64 FMA throughput of the device in TFLOP/s (8 independent DFMA chains per thread): the roof the
 * FP64-bound kernels of this path are reported against in bench.py */
int qmcb_fp64_peak(int device, double *tflops);

/* Orbitals at arbitrary points (open boundaries): out[p][j] = sum_mu chi_mu(pos[p]) coeff[mu][j], the
 * evaluation OBDMAccumulator needs (pyqmc/observables/obdm.py:150-153,231-233 -> orbitals.py:85-96). */
int qmcb_orbitals_at_points(qmcb_ctx *ctx, int64_t npoints, const double *pos, int norb,
                            const double *coeff, double *out);

/* asynchronous form of qmcb_vmc_block_slot for the pipelined driver (pyqmc_b200.mc.vmc): _begin optionally
 * recomputes the factors `recompute_which` from the resident walkers (wf.recompute at the start of vmc_worker,
 * mc.py:110), enqueues the block on the variates of `slot` and the device->host copies of its results (buffers must
 * stay valid, page-locked for true overlap) and returns; _end blocks until they have arrived. */
int qmcb_vmc_block_slot_begin(qmcb_ctx *ctx, int slot, int nsteps, double tstep, int with_energy,
                              int recompute_which, double *configs, double *energy, int64_t *nacc);
int qmcb_vmc_block_slot_end(qmcb_ctx *ctx, int slot);

/* ---- the same bit-identical legacy generator, DEVICE-RESIDENT (csrc/device_rng.cuh): the MT19937 stream of
 * np.random (mc.py:119,132; eval_ecp.py:145,263) is continued on the GPU from the state handed over once, so the
 * variates of a block are produced where they are consumed.  set_state / get_state exchange the fields of
 * np.random.get_state(); programs run asynchronously on the context's copy stream and chain through the state kept
 * on the device.  qmcb_devrng_program: ops as qmcb_rng_program, dst[i] = DEVICE addresses.
 * qmcb_devrng_vmc_block: the draw program of one VMC block into variate slot `slot` (consumed by
 * qmcb_vmc_block_slot).  qmcb_glibc_log_mismatches: how many of nsamples arguments the restated glibc log
 * (csrc/glibc_log.h) rounds differently from this host's libm log(); callers use the device generator only if 0. */
int qmcb_devrng_set_state(qmcb_ctx *ctx, const uint32_t *key, int32_t pos, int32_t has_gauss,
                          double cached_gauss);
int qmcb_devrng_get_state(qmcb_ctx *ctx, uint32_t *key, int32_t *pos, int32_t *has_gauss,
                          double *cached_gauss);
int qmcb_devrng_program(qmcb_ctx *ctx, int64_t nops, const int32_t *kind, const int64_t *count,
                        const uint64_t *dst, const double *scale);
int qmcb_devrng_vmc_block(qmcb_ctx *ctx, int slot, int nsteps, int ne, int64_t N, int necp,
                          double sigma);
/* DMC: the draw program of one dmc_propagate call (+ the branching draw) into DMC variate slot 0/1, and
 * qmcb_dmc_block on such a slot (no variate upload); *branch_draw = the rand() of branch (dmc.py:361) */
int qmcb_devrng_dmc_block(qmcb_ctx *ctx, int slot, int nsteps, int ne, int64_t N, int necp,
                          double sigma, int tmoves, int with_branch);
int qmcb_dmc_block_slot(qmcb_ctx *ctx, int slot, int nsteps, double tstep, double branchcut,
                        double e_trial, double e_est, double *weights, double *configs,
                        double *wsums, int64_t *nacc, int64_t *ntacc, double *branch_draw);
int64_t qmcb_glibc_log_mismatches(int64_t nsamples, uint64_t seed);
/* diagnostics: SM cycles, nanoseconds and state blocks of the last k_mt_generate launch (out3[3]) */
int qmcb_devrng_generator_timing(qmcb_ctx *ctx, int64_t *out3);

#ifdef __cplusplus
}
#endif
#endif
