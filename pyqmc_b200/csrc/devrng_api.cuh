// devrng_api.cuh -- host side of the device-resident legacy generator (device_rng.cuh); included by qmcb200.cu
// inside its extern "C" block.  The generator state lives on the device between programs; the host hands over
// np.random.get_state() once (qmcb_devrng_set_state) and fetches the advanced state when it needs numpy's global
// stream to be where the reference would have left it (qmcb_devrng_get_state).
struct DevRngProgram {
  std::vector<long long> key;  // (kind, count, dst, scale bits) of every op: rebuilt only when it changes
  DBuf<devrng::Op> d_ops;
  DBuf<devrng::OpPlan> d_plan;
  DBuf<long long> d_prefix;
  DBuf<devrng::Work> d_work;
  int nops = 0, nwork = 0;
  long long words_bound = 0;
};

struct DevRng {
  DBuf<devrng::State> st;
  DBuf<uint32_t> W;
  DevRngProgram slot_prog[qmcb_ctx::NSLOT];
  DevRngProgram generic;
  bool have_state = false;
  long long programs_run = 0;
};

static DevRng* devrng_of(qmcb_ctx* c) {
  if (!c->devrng) c->devrng = new DevRng();
  return static_cast<DevRng*>(c->devrng);
}

static void devrng_free(qmcb_ctx* c) {
  if (!c->devrng) return;
  DevRng* r = static_cast<DevRng*>(c->devrng);
  r->st.release();
  r->W.release();
  DevRngProgram* ps[qmcb_ctx::NSLOT + 1];
  for (int i = 0; i < qmcb_ctx::NSLOT; ++i) ps[i] = &r->slot_prog[i];
  ps[qmcb_ctx::NSLOT] = &r->generic;
  for (auto* p : ps) {
    p->d_ops.release();
    p->d_plan.release();
    p->d_prefix.release();
    p->d_work.release();
  }
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
  int chunk0 = 0;
  for (int i = 0; i < nops; ++i) {
    devrng::Op& o = ops[i];
    o.dst = dst[i];
    o.kind = kind[i];
    o.count = kind[i] == devrng::KIND_ROTATION ? 4 : count[i];
    o.scale = scale[i];
    o.chunk0 = chunk0;
    o.small = kind[i] == devrng::KIND_ROTATION ? 1 : 0;
    if (kind[i] == devrng::KIND_UNIFORM) {
      o.maxchunks = (int)((o.count + devrng::CHUNK_UNIFORM - 1) / devrng::CHUNK_UNIFORM);
      words += 2 * o.count;
    } else if (kind[i] == devrng::KIND_NORMAL || kind[i] == devrng::KIND_ROTATION) {
      const long long m = (o.count + 1) / 2;
      const long long att = attempts_bound(m);
      const long long per = o.small ? 32 : devrng::CHUNK_ATTEMPTS;
      o.maxchunks = (int)((att + per - 1) / per) + 1;
      words += 4 * att;
    } else {
      return fail("devrng: unknown op kind");
    }
    chunk0 += o.maxchunks;
    for (int k = 0; k < std::max(o.maxchunks, o.count > 0 ? 1 : 0); ++k) work.push_back(devrng::Work{i, k});
  }
  if (P.d_ops.ensure(nops) || P.d_plan.ensure(nops) || P.d_prefix.ensure((size_t)chunk0 + 1) || P.d_work.ensure(work.size() + 1))
    return -1;
  // (a program is rebuilt only when its shape changes; the copies are ordered after earlier work of the stream)
  CK(cudaMemcpyAsync(P.d_ops.p, ops.data(), nops * sizeof(devrng::Op), cudaMemcpyHostToDevice, c->copy_stream));
  CK(cudaMemcpyAsync(P.d_work.p, work.data(), work.size() * sizeof(devrng::Work), cudaMemcpyHostToDevice, c->copy_stream));
  CK(cudaStreamSynchronize(c->copy_stream));  // the host vectors die with this call
  P.nops = nops;
  P.nwork = (int)work.size();
  P.words_bound = words + 8LL * devrng::CHUNK_ATTEMPTS;
  P.key.swap(key);
  return 0;
}

static int devrng_run(qmcb_ctx* c, DevRngProgram& P) {
  DevRng* r = devrng_of(c);
  if (!r->have_state) return fail("qmcb_devrng_set_state has not been called");
  const int nblocks = (int)((624 + P.words_bound + 623) / 624) + 2;
  const long long nwords = (long long)nblocks * 624;
  if (r->W.ensure((size_t)nwords)) return -1;
  cudaStream_t s = c->copy_stream;
  devrng::k_mt_generate<<<1, 256, 0, s>>>(r->st.p, r->W.p, nblocks);
  devrng::k_rng_plan<<<1, devrng::PLAN_THREADS, 0, s>>>(P.d_ops.p, P.nops, r->st.p, r->W.p, nwords, P.d_plan.p, P.d_prefix.p);
  if (P.nwork > 0)
    devrng::k_rng_fill<<<P.nwork, devrng::PLAN_THREADS, 0, s>>>(P.d_ops.p, P.d_plan.p, P.d_prefix.p, P.d_work.p, r->st.p, r->W.p,
                                                               nwords);
  devrng::k_rng_finalize<<<1, 256, 0, s>>>(r->st.p, r->W.p);
  c->nlaunch += 4;
  CK(cudaGetLastError());
  r->programs_run++;
  return 0;
}

int qmcb_devrng_set_state(qmcb_ctx* c, const uint32_t* key, int32_t pos, int32_t has_gauss, double cached_gauss) {
  Guard g(c);
  if (pos < 0 || pos > 624) return fail("MT19937 position out of range");
  DevRng* r = devrng_of(c);
  if (r->st.ensure(1)) return -1;
  devrng::State h;
  std::memset(&h, 0, sizeof(h));
  std::memcpy(h.key, key, sizeof(h.key));
  h.pos = pos;
  h.has_gauss = has_gauss;
  h.cached = cached_gauss;
  CK(cudaStreamSynchronize(c->copy_stream));
  CK(cudaMemcpy(r->st.p, &h, sizeof(h), cudaMemcpyHostToDevice));
  r->have_state = true;
  return 0;
}

int qmcb_devrng_get_state(qmcb_ctx* c, uint32_t* key, int32_t* pos, int32_t* has_gauss, double* cached_gauss) {
  Guard g(c);
  DevRng* r = devrng_of(c);
  if (!r->have_state) return fail("qmcb_devrng_set_state has not been called");
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
  if (devrng_run(c, r->slot_prog[slot])) return -1;
  CK(cudaEventRecord(c->slot_ready[slot], c->copy_stream));
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
