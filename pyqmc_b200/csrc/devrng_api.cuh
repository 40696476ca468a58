// devrng_api.cuh -- host side of the device-resident legacy generator (device_rng.cuh); included by qmcb200.cu
// inside its extern "C" block.  The generator state lives on the device between programs; the host hands over
// np.random.get_state() once (qmcb_devrng_set_state) and fetches the advanced state when it needs numpy's global
// stream to be where the reference would have left it (qmcb_devrng_get_state).
struct DevRngProgram {
  std::vector<long long> key;  // (kind, count, dst, scale bits) of every op: rebuilt only when it changes
  DBuf<devrng::Op> d_ops;
  DBuf<devrng::OpPlan> d_plan;
  DBuf<devrng::Work> d_work;
  int nops = 0, nwork = 0;
  long long words_bound = 0;
};

struct DevRng {
  DBuf<devrng::State> st;
  DBuf<devrng::GenState> gen;
  // segmented generation (k_mt_jump): per-segment generator states, their streams and events
  static constexpr int NSEG_STREAMS = 8;
  DBuf<devrng::GenState> gen_seg;
  DBuf<uint32_t> d_poly;
  cudaStream_t seg_stream[NSEG_STREAMS] = {};
  std::vector<cudaEvent_t> ev_head, ev_done;
  long long segmented_launches = 0;
  DBuf<uint32_t> W, F0, F1;
  long long cap_blocks = 0;   // capacity of W in 624-word state blocks
  long long gen_blocks = 0;   // blocks generated (block 0 = the state handed in / rebased onto)
  long long flagged = 0;      // attempt groups whose accept flags exist (multiple of 32)
  long long upper = 0;        // upper bound of the stream offset after the programs enqueued so far
  int r0 = 0;                 // alignment class of the stream offset (mod 4)
  cudaStream_t gen_stream = nullptr;
  cudaEvent_t ev_gen = nullptr, ev_plan = nullptr;
  DevRngProgram slot_prog[qmcb_ctx::NSLOT];
  DevRngProgram dmc_prog[2];
  DevRngProgram generic;
  bool have_state = false;
  long long programs_run = 0, rebases = 0;
};

static DevRng* devrng_of(qmcb_ctx* c) {
  if (!c->devrng) {
    DevRng* r = new DevRng();
    // highest priority: the generator is one long-lived CTA that everything downstream waits for; it must not queue
    // behind the hundreds of CTAs of the compute kernels
    int lo = 0, hi = 0;
    cudaDeviceGetStreamPriorityRange(&lo, &hi);
    cudaStreamCreateWithPriority(&r->gen_stream, cudaStreamNonBlocking, hi);
    cudaEventCreateWithFlags(&r->ev_gen, cudaEventDisableTiming);
    cudaEventCreateWithFlags(&r->ev_plan, cudaEventDisableTiming);
    for (int i = 0; i < DevRng::NSEG_STREAMS; ++i) cudaStreamCreateWithPriority(&r->seg_stream[i], cudaStreamNonBlocking, hi);
    c->devrng = r;
  }
  return static_cast<DevRng*>(c->devrng);
}

static void devrng_free(qmcb_ctx* c) {
  if (!c->devrng) return;
  DevRng* r = static_cast<DevRng*>(c->devrng);
  if (r->gen_stream) cudaStreamSynchronize(r->gen_stream);
  if (c->copy_stream) cudaStreamSynchronize(c->copy_stream);
  for (int i = 0; i < DevRng::NSEG_STREAMS; ++i)
    if (r->seg_stream[i]) {
      cudaStreamSynchronize(r->seg_stream[i]);
      cudaStreamDestroy(r->seg_stream[i]);
    }
  for (auto e : r->ev_head) cudaEventDestroy(e);
  for (auto e : r->ev_done) cudaEventDestroy(e);
  r->gen_seg.release();
  r->d_poly.release();
  r->st.release();
  r->gen.release();
  r->W.release();
  r->F0.release();
  r->F1.release();
  DevRngProgram* ps[qmcb_ctx::NSLOT + 3];
  for (int i = 0; i < qmcb_ctx::NSLOT; ++i) ps[i] = &r->slot_prog[i];
  ps[qmcb_ctx::NSLOT] = &r->generic;
  ps[qmcb_ctx::NSLOT + 1] = &r->dmc_prog[0];
  ps[qmcb_ctx::NSLOT + 2] = &r->dmc_prog[1];
  for (auto* p : ps) {
    p->d_ops.release();
    p->d_plan.release();
    p->d_work.release();
  }
  if (r->ev_gen) cudaEventDestroy(r->ev_gen);
  if (r->ev_plan) cudaEventDestroy(r->ev_plan);
  if (r->gen_stream) cudaStreamDestroy(r->gen_stream);
  delete r;
  c->devrng = nullptr;
}

// upper bound of the polar attempts needed for m accepted pairs: mean m/p plus 12 standard deviations
static long long attempts_bound(long long m) {
  const double p = 0.78539816339744830962;
  return (long long)std::ceil(m / p + 12.0 * std::sqrt((double)m * (1.0 - p)) / p) + 64;
}

static int devrng_build(qmcb_ctx* c, DevRngProgram& P, int nops, const int* kind, const long long* count,
                        double* const* dst, const double* scale) {
  std::vector<long long> key;
  key.reserve(4 * (size_t)nops);
  for (int i = 0; i < nops; ++i) {
    long long sb;
    std::memcpy(&sb, &scale[i], 8);
    key.push_back(kind[i]);
    key.push_back(count[i]);
    key.push_back((long long)(uintptr_t)dst[i]);
    key.push_back(sb);
  }
  if (key == P.key && P.d_ops.p) return 0;
  std::vector<devrng::Op> ops(nops);
  std::vector<devrng::Work> work;
  long long words = 0;
  for (int i = 0; i < nops; ++i) {
    devrng::Op& o = ops[i];
    o.dst = dst[i];
    o.kind = kind[i];
    o.count = kind[i] == devrng::KIND_ROTATION ? 4 : count[i];
    o.scale = scale[i];
    if (kind[i] == devrng::KIND_UNIFORM) {
      o.maxchunks = (int)((o.count + devrng::CHUNK_UNIFORM - 1) / devrng::CHUNK_UNIFORM);
      words += 2 * o.count;
    } else if (kind[i] == devrng::KIND_NORMAL || kind[i] == devrng::KIND_ROTATION) {
      const long long m = (o.count + 1) / 2;
      const long long att = attempts_bound(m);
      // a chunk is THREADS bitmap words starting at the word that holds the draw's first attempt
      o.maxchunks = (int)((att + 31 + devrng::CHUNK_GROUPS - 1) / devrng::CHUNK_GROUPS);
      words += 4 * att;
    } else {
      return fail("devrng: unknown op kind");
    }
    for (int k = 0; k < std::max(o.maxchunks, o.count > 0 ? 1 : 0); ++k) work.push_back(devrng::Work{i, k});
  }
  // the previous version of this program may still be in flight
  CK(cudaStreamSynchronize(c->copy_stream));
  if (P.d_ops.ensure(nops) || P.d_plan.ensure(nops) || P.d_work.ensure(work.size() + 1)) return -1;
  CK(cudaMemcpy(P.d_ops.p, ops.data(), nops * sizeof(devrng::Op), cudaMemcpyHostToDevice));
  CK(cudaMemcpy(P.d_work.p, work.data(), work.size() * sizeof(devrng::Work), cudaMemcpyHostToDevice));
  P.nops = nops;
  P.nwork = (int)work.size();
  P.words_bound = words + 1024;
  P.key.swap(key);
  return 0;
}

// (re)allocation of the word buffer and the flag bitmaps; keeps state block 0
static int devrng_reserve(qmcb_ctx* c, DevRng* r, long long blocks) {
  if (blocks <= r->cap_blocks) return 0;
  CK(cudaStreamSynchronize(r->gen_stream));
  CK(cudaStreamSynchronize(c->copy_stream));
  DBuf<uint32_t> W;
  if (W.ensure((size_t)blocks * 624)) return -1;
  if (r->W.p) CK(cudaMemcpy(W.p, r->W.p, 624 * 4, cudaMemcpyDeviceToDevice));
  r->W.release();
  r->W = W;
  const size_t fwords = (size_t)blocks * 156 / 32 + 64;  // 156 attempt groups per state block and class
  r->F0.release();
  r->F1.release();
  if (r->F0.ensure(fwords) || r->F1.ensure(fwords)) return -1;
  r->cap_blocks = blocks;
  r->gen_blocks = std::min<long long>(r->gen_blocks, 1);  // only block 0 was carried over
  r->flagged = 0;
  return 0;
}

// generate + flag so that `blocks` state blocks exist; records ev_gen on the generator stream
static void devrng_launch_generator(DevRng* r, devrng::GenState* g, long long first, int nb, cudaStream_t s) {
  static const int mode = std::getenv("QMCB_MT_MODE") ? std::atoi(std::getenv("QMCB_MT_MODE")) : 1;
  if (nb <= 0) return;
  if (mode == 0)
    devrng::k_mt_generate<0><<<1, 640, 0, s>>>(g, r->W.p, first, nb);
  else
    devrng::k_mt_generate<1><<<1, 640, 0, s>>>(g, r->W.p, first, nb);
}

static int devrng_generate(qmcb_ctx* c, DevRng* r, long long blocks) {
  if (blocks <= r->gen_blocks) return 0;
  const long long nb = blocks - r->gen_blocks, g0 = r->gen_blocks;
  constexpr long long SEG = QMCB_MT_SEG_BLOCKS, HEAD = 32;  // 33 blocks (the start block + 32) feed one jump
  static const bool segmented = std::getenv("QMCB_MT_SEGMENTS") == nullptr || std::atoi(std::getenv("QMCB_MT_SEGMENTS")) != 0;
  if (!segmented || nb < 2 * SEG) {
    devrng_launch_generator(r, r->gen.p, g0, (int)nb, r->gen_stream);
  } else {
    // Segments of SEG blocks generated concurrently, one single-CTA generator each.  The start block of segment k
    // is the jump (k_mt_jump) of the start block of segment k - 1, which needs that segment's first 32 blocks: every
    // segment runs as head (32 blocks) + tail on its own stream, and the jump for the next one sits between them.
    const int S = (int)((nb + SEG - 1) / SEG);
    if (r->gen_seg.ensure((size_t)S)) return -1;
    if (!r->d_poly.p) {
      if (r->d_poly.ensure(QMCB_MT_POLY_WORDS)) return -1;
      CK(cudaMemcpy(r->d_poly.p, qmcb_mt_jump_poly, sizeof(qmcb_mt_jump_poly), cudaMemcpyHostToDevice));
    }
    while ((int)r->ev_head.size() < S) {
      cudaEvent_t a, b;
      CK(cudaEventCreateWithFlags(&a, cudaEventDisableTiming));
      CK(cudaEventCreateWithFlags(&b, cudaEventDisableTiming));
      r->ev_head.push_back(a);
      r->ev_done.push_back(b);
    }
    // the segment streams start behind everything the generator stream has been given so far (rebase, earlier blocks)
    CK(cudaEventRecord(r->ev_done[0], r->gen_stream));
    devrng_launch_generator(r, r->gen.p, g0, (int)HEAD, r->gen_stream);
    CK(cudaEventRecord(r->ev_head[0], r->gen_stream));
    devrng_launch_generator(r, r->gen.p, g0 + HEAD, (int)(SEG - HEAD), r->gen_stream);
    for (int k = 1; k < S; ++k) {
      cudaStream_t s = r->seg_stream[(k - 1) % DevRng::NSEG_STREAMS];
      const long long len = std::min<long long>(SEG, nb - (long long)k * SEG), head = std::min<long long>(HEAD, len);
      if (k == 1) CK(cudaStreamWaitEvent(s, r->ev_done[0], 0));
      CK(cudaStreamWaitEvent(s, r->ev_head[k - 1], 0));
      uint32_t* src = r->W.p + (size_t)(g0 - 1 + (long long)(k - 1) * SEG) * 624;
      devrng::k_mt_jump<<<(624 + devrng::JUMP_OUT - 1) / devrng::JUMP_OUT, devrng::JUMP_THREADS, 0, s>>>(
          src, r->d_poly.p, src + (size_t)SEG * 624, r->gen_seg.p + k);
      devrng_launch_generator(r, r->gen_seg.p + k, g0 + (long long)k * SEG, (int)head, s);
      if (k + 1 < S) CK(cudaEventRecord(r->ev_head[k], s));
      devrng_launch_generator(r, r->gen_seg.p + k, g0 + (long long)k * SEG + head, (int)(len - head), s);
      CK(cudaEventRecord(r->ev_done[k], s));
      c->nlaunch += 3;
    }
    for (int k = 1; k < S; ++k) CK(cudaStreamWaitEvent(r->gen_stream, r->ev_done[k], 0));
    // the generator state continues from the last block of the last segment
    CK(cudaMemcpyAsync(r->gen.p, r->W.p + (size_t)(blocks - 1) * 624, 624 * 4, cudaMemcpyDeviceToDevice, r->gen_stream));
    r->segmented_launches++;
    c->nlaunch += 1;
  }
  r->gen_blocks = blocks;
  const long long avail = blocks * 624;
  const long long g_hi = ((avail - 3) / 4) / 32 * 32;
  if (g_hi > r->flagged) {
    const long long ng = g_hi - r->flagged;
    devrng::k_rng_flags<<<(unsigned)((ng + devrng::THREADS - 1) / devrng::THREADS), devrng::THREADS, 0, r->gen_stream>>>(
        r->W.p, r->F0.p, r->F1.p, r->r0, r->flagged, g_hi);
    r->flagged = g_hi;
  }
  c->nlaunch += 2;
  CK(cudaGetLastError());
  CK(cudaEventRecord(r->ev_gen, r->gen_stream));
  return 0;
}

static int devrng_run(qmcb_ctx* c, DevRngProgram& P) {
  DevRng* r = devrng_of(c);
  if (!r->have_state) return fail("qmcb_devrng_set_state has not been called");
  const long long prog_blocks = (P.words_bound + 623) / 624 + 2;
  long long need_blocks = (r->upper + P.words_bound + 623) / 624 + 2;
  if (need_blocks > r->cap_blocks) {
    if (r->upper > 624) {
      // rebase: move the state block the stream is in to the front of the buffer.  Ordered after every plan so far
      // (copy stream) and after the generator (it may be writing ahead); the generator restarts behind it.
      CK(cudaStreamWaitEvent(c->copy_stream, r->ev_gen, 0));
      devrng::k_rng_rebase<<<1, 320, 0, c->copy_stream>>>(r->st.p, r->gen.p, r->W.p);
      CK(cudaEventRecord(r->ev_plan, c->copy_stream));
      CK(cudaStreamWaitEvent(r->gen_stream, r->ev_plan, 0));
      c->nlaunch++;
      r->gen_blocks = 1;
      r->flagged = 0;
      r->upper = 624;
      r->rebases++;
      need_blocks = (r->upper + P.words_bound + 623) / 624 + 2;
    }
    const long long want = std::max<long long>(need_blocks + prog_blocks, 32 * prog_blocks);
    if (need_blocks > r->cap_blocks && devrng_reserve(c, r, std::min<long long>(want, std::max<long long>(need_blocks + 2, (3LL << 30) / 2496))))
      return -1;
  }
  if (devrng_generate(c, r, need_blocks)) return -1;
  cudaStream_t s = c->copy_stream;
  CK(cudaStreamWaitEvent(s, r->ev_gen, 0));
  const long long nwords = r->gen_blocks * 624, valid = r->flagged;
  devrng::k_rng_plan<<<1, devrng::THREADS, 0, s>>>(P.d_ops.p, P.nops, r->st.p, r->W.p, r->F0.p, r->F1.p, r->r0, valid, nwords,
                                                    P.d_plan.p);
  if (P.nwork > 0)
    devrng::k_rng_fill<<<P.nwork, devrng::THREADS, 0, s>>>(P.d_ops.p, P.d_plan.p, P.d_work.p, r->st.p, r->W.p, r->F0.p, r->F1.p,
                                                           r->r0, valid);
  c->nlaunch += 2;
  CK(cudaGetLastError());
  r->upper += P.words_bound;
  // run the generator ahead by one more program of this size: the next plan then finds its words ready
  const long long ahead = std::min<long long>(r->cap_blocks, (r->upper + P.words_bound + 623) / 624 + 2);
  if (devrng_generate(c, r, ahead)) return -1;
  r->programs_run++;
  return 0;
}

int qmcb_devrng_set_state(qmcb_ctx* c, const uint32_t* key, int32_t pos, int32_t has_gauss, double cached_gauss) {
  Guard g(c);
  if (pos < 0 || pos > 624) return fail("MT19937 position out of range");
  DevRng* r = devrng_of(c);
  CK(cudaStreamSynchronize(r->gen_stream));
  CK(cudaStreamSynchronize(c->copy_stream));
  if (r->st.ensure(1) || r->gen.ensure(1)) return -1;
  if (r->cap_blocks == 0 && devrng_reserve(c, r, 64)) return -1;
  devrng::State h;
  std::memset(&h, 0, sizeof(h));
  h.cur = pos;
  h.has_gauss = has_gauss;
  h.cached = cached_gauss;
  CK(cudaMemcpy(r->st.p, &h, sizeof(h), cudaMemcpyHostToDevice));
  CK(cudaMemcpy(r->gen.p, key, 624 * 4, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(r->W.p, key, 624 * 4, cudaMemcpyHostToDevice));
  r->r0 = pos & 3;
  r->gen_blocks = 1;
  r->flagged = 0;
  r->upper = pos;
  r->have_state = true;
  return 0;
}

int qmcb_devrng_get_state(qmcb_ctx* c, uint32_t* key, int32_t* pos, int32_t* has_gauss, double* cached_gauss) {
  Guard g(c);
  DevRng* r = devrng_of(c);
  if (!r->have_state) return fail("qmcb_devrng_set_state has not been called");
  devrng::k_rng_finalize<<<1, 320, 0, c->copy_stream>>>(r->st.p, r->W.p);
  c->nlaunch++;
  CK(cudaStreamSynchronize(c->copy_stream));
  devrng::State h;
  CK(cudaMemcpy(&h, r->st.p, sizeof(h), cudaMemcpyDeviceToHost));
  if (h.error) return fail("device generator ran out of generated words (draw program bound too small)");
  std::memcpy(key, h.key, sizeof(h.key));
  *pos = h.pos;
  *has_gauss = h.has_gauss;
  *cached_gauss = h.cached;
  return 0;
}

// timing of the last generator launch: SM cycles, nanoseconds, state blocks (diagnostics)
int qmcb_devrng_generator_timing(qmcb_ctx* c, int64_t* out3) {
  Guard g(c);
  DevRng* r = devrng_of(c);
  if (!r->gen.p) return fail("no generator state");
  CK(cudaStreamSynchronize(r->gen_stream));
  devrng::GenState h;
  CK(cudaMemcpy(&h, r->gen.p, sizeof(h), cudaMemcpyDeviceToHost));
  out3[0] = h.cycles;
  out3[1] = h.nanos;
  out3[2] = h.blocks;
  return 0;
}

int qmcb_devrng_program(qmcb_ctx* c, int64_t nops, const int32_t* kind, const int64_t* count, const uint64_t* dst,
                        const double* scale) {
  Guard g(c);
  DevRng* r = devrng_of(c);
  std::vector<int> k(kind, kind + nops);
  std::vector<long long> n(count, count + nops);
  std::vector<double*> d(nops);
  for (int64_t i = 0; i < nops; ++i) d[i] = reinterpret_cast<double*>(static_cast<uintptr_t>(dst[i]));
  if (devrng_build(c, r->generic, (int)nops, k.data(), n.data(), d.data(), scale)) return -1;
  return devrng_run(c, r->generic);
}

// Variates of one VMC block generated into device slot `slot` in the reference's consumption order (the same
// program qmcb_rng_vmc_block runs on the host).  Asynchronous: returns after enqueuing on the copy stream;
// qmcb_vmc_block_slot waits on the slot's event.
int qmcb_devrng_vmc_block(qmcb_ctx* c, int slot, int nsteps, int ne, int64_t N, int necp, double sigma) {
  Guard g(c);
  if (slot < 0 || slot >= qmcb_ctx::NSLOT) return fail("slot out of range");
  DevRng* r = devrng_of(c);
  const size_t nse = (size_t)nsteps * ne;
  if (c->s_gauss[slot].ensure(nse * N * 3) || c->s_unif[slot].ensure(nse * N)) return -1;
  if (necp > 0 && (c->s_u[slot].ensure(nse * necp * N) || c->s_rot[slot].ensure(nse * necp * 9))) return -1;
  std::vector<int> kind;
  std::vector<long long> count;
  std::vector<double*> dst;
  std::vector<double> scale;
  auto push = [&](int k, long long n, double* d, double s) {
    kind.push_back(k);
    count.push_back(n);
    dst.push_back(d);
    scale.push_back(s);
  };
  for (int step = 0; step < nsteps; ++step) {
    for (int e = 0; e < ne; ++e) {
      const size_t se = (size_t)step * ne + e;
      push(devrng::KIND_NORMAL, 3 * N, c->s_gauss[slot].p + se * N * 3, sigma);
      push(devrng::KIND_UNIFORM, N, c->s_unif[slot].p + se * N, 1.0);
    }
    for (int e = 0; e < ne && necp > 0; ++e)
      for (int a = 0; a < necp; ++a) {
        const size_t sea = ((size_t)step * ne + e) * necp + a;
        push(devrng::KIND_UNIFORM, N, c->s_u[slot].p + sea * N, 1.0);
        push(devrng::KIND_ROTATION, 4, c->s_rot[slot].p + sea * 9, 1.0);
      }
  }
  if (devrng_build(c, r->slot_prog[slot], (int)kind.size(), kind.data(), count.data(), dst.data(), scale.data())) return -1;
  // the slot's previous variates may still be read by a block begun with qmcb_vmc_block_slot_begin
  if (c->block_pending[slot]) CK(cudaStreamWaitEvent(c->copy_stream, c->block_done[slot], 0));
  if (devrng_run(c, r->slot_prog[slot])) return -1;
  CK(cudaEventRecord(c->slot_ready[slot], c->copy_stream));
  return 0;
}

// Variates of one DMC block (dmc_propagate, dmc.py:123-221) generated into DMC slot `slot` in the reference's
// consumption order: the energy evaluation before the first step; per step, for every electron the T-move draws
// (per ECP atom random(N) + a rotation; select_walker: N scalar rand(); acceptance rand(N)), for every electron
// normal(N, 3) + rand(N), the energy evaluation; finally (with_branch) the rand() of branch (dmc.py:361).
int qmcb_devrng_dmc_block(qmcb_ctx* c, int slot, int nsteps, int ne, int64_t N, int necp, double sigma, int tmoves,
                          int with_branch) {
  Guard g(c);
  if (slot < 0 || slot > 1) return fail("slot out of range");
  DevRng* r = devrng_of(c);
  qmcb_ctx::DmcSlot& sl = c->dmc_slot[slot];
  const size_t nse = (size_t)nsteps * ne, nu1 = (size_t)ne * necp * N, nr1 = (size_t)ne * necp * 9;
  if (sl.gauss.ensure(nse * N * 3) || sl.unif.ensure(nse * N) || sl.u.ensure((size_t)(nsteps + 1) * nu1) ||
      sl.rot.ensure((size_t)(nsteps + 1) * nr1) || sl.tmu.ensure(nse * necp * N) || sl.tmrot.ensure(nse * necp * 9) ||
      sl.tmsel.ensure(nse * N) || sl.tmacc.ensure(nse * N) || sl.branch.ensure(1))
    return -1;
  if (!sl.ready) CK(cudaEventCreateWithFlags(&sl.ready, cudaEventDisableTiming));
  std::vector<int> kind;
  std::vector<long long> count;
  std::vector<double*> dst;
  std::vector<double> scale;
  auto push = [&](int k, long long n, double* d, double s) {
    kind.push_back(k);
    count.push_back(n);
    dst.push_back(d);
    scale.push_back(s);
  };
  auto energy_draws = [&](int k) {
    for (int e = 0; e < ne; ++e)
      for (int a = 0; a < necp; ++a) {
        const size_t ea = (size_t)e * necp + a;
        push(devrng::KIND_UNIFORM, N, sl.u.p + (size_t)k * nu1 + ea * N, 1.0);
        push(devrng::KIND_ROTATION, 4, sl.rot.p + (size_t)k * nr1 + ea * 9, 1.0);
      }
  };
  energy_draws(0);
  for (int step = 0; step < nsteps; ++step) {
    if (tmoves)
      for (int e = 0; e < ne; ++e) {
        const size_t se = (size_t)step * ne + e;
        for (int a = 0; a < necp; ++a) {
          push(devrng::KIND_UNIFORM, N, sl.tmu.p + (se * necp + a) * N, 1.0);
          push(devrng::KIND_ROTATION, 4, sl.tmrot.p + (se * necp + a) * 9, 1.0);
        }
        push(devrng::KIND_UNIFORM, N, sl.tmsel.p + se * N, 1.0);
        push(devrng::KIND_UNIFORM, N, sl.tmacc.p + se * N, 1.0);
      }
    for (int e = 0; e < ne; ++e) {
      const size_t se = (size_t)step * ne + e;
      push(devrng::KIND_NORMAL, 3 * N, sl.gauss.p + se * N * 3, sigma);
      push(devrng::KIND_UNIFORM, N, sl.unif.p + se * N, 1.0);
    }
    energy_draws(step + 1);
  }
  if (with_branch) push(devrng::KIND_UNIFORM, 1, sl.branch.p, 1.0);
  if (devrng_build(c, r->dmc_prog[slot], (int)kind.size(), kind.data(), count.data(), dst.data(), scale.data())) return -1;
  if (devrng_run(c, r->dmc_prog[slot])) return -1;
  CK(cudaEventRecord(sl.ready, c->copy_stream));
  sl.filled = true;
  return 0;
}

// Host check that glibc_log.h reproduces THIS machine's libm log() (the function numpy's generator calls): the
// device generator is only used when this returns 0 mismatches.
int64_t qmcb_glibc_log_mismatches(int64_t nsamples, uint64_t seed) {
  uint64_t x = seed ? seed : 0x9e3779b97f4a7c15ull;
  int64_t bad = 0;
  for (int64_t i = 0; i < nsamples; ++i) {
    x ^= x << 13;
    x ^= x >> 7;
    x ^= x << 17;
    double v;
    if (i & 1) {
      v = 0.9375 + (double)(x >> 11) * (0.0625 / 9007199254740992.0);
    } else {
      const double a = 2.0 * ((double)(x >> 38) / 67108864.0) - 1.0, b = 2.0 * ((double)(x & 0x3ffffff) / 67108864.0) - 1.0;
      v = a * a + b * b;
      if (!(v < 1.0) || v == 0.0) continue;
    }
    volatile double ref = std::log(v);
    const double mine = qmcb_glibc_log(v), refv = ref;
    if (std::memcmp(&mine, &refv, 8) != 0) ++bad;
  }
  return bad;
}
