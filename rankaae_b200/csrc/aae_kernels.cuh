// Kernels: the persistent per-trial train kernel (all batches of an epoch), the validation/metrics
// kernel (eval block + Shapiro-Wilk + Spearman + ReduceLROnPlateau), and the small state-init kernel.
#pragma once
#include "aae_stages.cuh"

namespace raae {

// functions.py:214-219, float64 like the reference's numpy scalar
__device__ inline float alpha_schedule(const Ctx& c) {
  double pct = (double)c.epoch / c.hp[RAAE_HP_MAX_EPOCH];
  double a = (2.0 / (1.0 + exp(-1.0e4 / c.hp[RAAE_HP_ALPHA_FLAT_STEP] * pct)) - 1.0) * c.hp[RAAE_HP_ALPHA_LIMIT];
  return (float)a;
}

// spec_in += randn_like(spec_in) * spec_noise (trainer.py:112) for the rows idx[0..B) of the training
// split; descriptors of the same rows; z_sample of the MI phase.
__device__ __noinline__ void build_batch(const Ctx& c_ref, const int32_t* __restrict__ idx) {
  const Ctx c = c_ref;                 // register copy of the kernel context (SmemFixed::ctx, shared memory)
  const KParams& p = *c.p;
  const int tid = threadIdx.x, dim = p.cfg.dim_in, xld = p.sl.xld, K = p.cfg.n_aux, ns = p.cfg.nstyle;
  float* xn = c.sc + p.sl.xn;
  float* aux = c.sc + p.sl.aux;
  float* zs = c.sc + p.sl.zs;
  float* xk = c.sc + p.sl.xk;
  float* xm = c.sc + p.sl.xm;
  RAAE_SMEM();
  StageTimer timer_(&sm->prof[kStBatch]);
  __syncthreads();
  const bool dbg = c.a->debug != 0;
  const bool images = (p.cfg.tensor_cores & 4) != 0;
  const float* xsrc = dbg ? c.a->dbg.x_noisy : p.spec_train;
  const float sigma = dbg ? 0.f : (float)c.hp[RAAE_HP_SPEC_NOISE];
  const uint32_t key = stream_key(c.seed, c.step_id, kStreamXNoise);
  const int nch64 = p.sl.nch64, nch128 = p.sl.nch128;
  // Reference row of the operand images: column means over the first (up to 32) rows of the batch.  The tensor core
  // accumulates with truncation, ~1 ulp of the running sum per MMA; spectra are a large common profile plus small
  // variations, and BatchNorm then divides by the small spread, so the images hold (row - reference) and the consumers
  // add reference . W^T (forward) / db (x) reference (weight gradient) back in FP32: exact algebra, 20 x smaller sums.
  // cluster per trial: every CTA builds the rows of ITS tiles; the reference row is the same in all of them (own copy each)
  const int crank = c.crank, csize = c.csize;
  float* xref = c.sc + p.sl.xref + crank * kMaxDim;
  if (images) {
    const int nref = min(32, c.B), c4 = (tid & 63) * 4, g = tid >> 6;
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    if (c4 < dim)
      for (int r = g; r < nref; r += 4) {
        const size_t srow = dbg ? (size_t)r : (size_t)idx[r];
        float4 v = *reinterpret_cast<const float4*>(xsrc + srow * dim + c4);
        if (sigma != 0.f) {
          uint32_t e = (uint32_t)(r * kMaxDim + c4);
          float n0, n1, n2, n3;                    // e is a multiple of 4: two whole Box-Muller pairs
          normal_pair(key, e >> 1, n0, n1);
          normal_pair(key, (e >> 1) + 1u, n2, n3);
          v.x += sigma * n0; v.y += sigma * n1; v.z += sigma * n2; v.w += sigma * n3;
        }
        acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
      }
    float* red = &sm->red[0][0];                 // [4][256]
    *reinterpret_cast<float4*>(red + g * 256 + c4) = acc;
    __syncthreads();
    {
      const float s = (red[tid] + red[256 + tid] + red[512 + tid] + red[768 + tid]) / (float)nref;
      sm->bias[tid] = s;                         // kMaxDim == kThreads == 256
      xref[tid] = s;
    }
    __syncthreads();
  }
  const int q_per_row = images ? nch64 * 16 : (dim >> 2);                 // float4 quads per row (padded to whole 64-col chunks)
  const int rows_all = csize > 1 ? cl::own_tiles(c.B, crank, csize) * kTM
                     : images ? ((c.B + kTM - 1) / kTM) * kTM : c.B;      // whole tiles: rows >= B are zero-filled
  // Four quads per thread and pass: the row indices, then the four gathered loads, are all in flight before the first use
  // (one quad per pass exposed two dependent memory round trips - index, then row - 256 times per thread and batch).
  constexpr int U = 4;
  const int total = rows_all * q_per_row;
  const int qshift = (q_per_row & (q_per_row - 1)) == 0 ? 31 - __clz(q_per_row) : -1;      // power of two: shift instead of divide
  for (int i0 = tid; i0 < total; i0 += kThreads * U) {
    int rr_[U], cc_[U];
    bool ok_[U];
    size_t srow_[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int i = i0 + u * kThreads;
      const int sl = qshift >= 0 ? (i >> qshift) : i / q_per_row;
      cc_[u] = (i - sl * q_per_row) * 4;
      rr_[u] = csize == 1 ? sl : cl::slot_row(sl, crank, csize);
      ok_[u] = i < total && rr_[u] < c.B && cc_[u] < dim;
      srow_[u] = ok_[u] ? (dbg ? (size_t)rr_[u] : (size_t)idx[rr_[u]]) : 0;
    }
    float4 v_[U];
#pragma unroll
    for (int u = 0; u < U; ++u)
      v_[u] = ok_[u] ? *reinterpret_cast<const float4*>(xsrc + srow_[u] * dim + cc_[u]) : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int i = i0 + u * kThreads, r = rr_[u], c4 = cc_[u];
      if (i >= total) break;
      float4 v = v_[u];
      if (ok_[u]) {
        if (sigma != 0.f) {
          uint32_t e = (uint32_t)(r * kMaxDim + c4);
          float n0, n1, n2, n3;
          normal_pair(key, e >> 1, n0, n1);
          normal_pair(key, (e >> 1) + 1u, n2, n3);
          v.x += sigma * n0; v.y += sigma * n1; v.z += sigma * n2; v.w += sigma * n3;
        }
        *reinterpret_cast<float4*>(xn + (size_t)r * xld + c4) = v;
      }
      if (images) {
        const int T = r >> 7, rr = r & 127;
        if (ok_[u]) {
          const float4 xr = *reinterpret_cast<const float4*>(sm->bias + c4);
          v.x -= xr.x; v.y -= xr.y; v.z -= xr.z; v.w -= xr.w;
        }
        // raw fp32 in the operand layout: the consumers split it into rounded hi / lo planes in shared memory, so the
        // batch crosses HBM once per consumer
        float* bk = xk + (size_t)(T * nch64 + (c4 >> 6)) * 8192;
        *reinterpret_cast<float4*>(reinterpret_cast<char*>(bk) + tc::sw128_chunk_off(rr, c4 & 63, tc::kABlockBytes)) = v;
        if (c4 < nch128 * 128) {
          float* bm = xm + (size_t)(T * nch128 + (c4 >> 7)) * 16384;
          *reinterpret_cast<float4*>(reinterpret_cast<char*>(bm) + tc::sw128_32b_chunk_off(rr, c4 & 127, tc::kABlockBytes)) = v;
        }
      }
    }
  }
  const int nslots = csize == 1 ? c.B : cl::own_tiles(c.B, crank, csize) * kTM;
  for (int e = tid; e < nslots * kZ; e += kThreads) {
    const int k = e & 7, r = csize == 1 ? (e >> 3) : cl::slot_row(e >> 3, crank, csize);
    if (r >= c.B) continue;
    const float* arow = dbg ? c.a->dbg.aux + (size_t)r * K : p.aux_train + (size_t)idx[r] * K;
    aux[r * kZ + k] = k < K ? arow[k] : 0.f;
  }
  const float* zsp = dbg ? c.a->dbg.z_sample : nullptr;
  const uint32_t kz = stream_key(c.seed, c.step_id, kStreamZSample);
  for (int e = tid; e < nslots * kZ; e += kThreads) {
    const int k = e & 7, r = csize == 1 ? (e >> 3) : cl::slot_row(e >> 3, crank, csize);
    if (r >= c.B) continue;
    const int i = r * kZ + k;
    zs[i] = k < ns ? (zsp ? zsp[(size_t)r * ns + k] : normal_at(kz, (uint32_t)i)) : 0.f;
  }
  if (images) tc::fence_async_all();     // the images are read by bulk copies (async proxy) in later stages
  __syncthreads();
}

// One iteration of the batch loop body, trainer.py:112-204 (gradient-reversal branch).
__device__ __forceinline__ void train_step(const Ctx& c, int phase_mask, bool run_p0 = true) {
  const KParams& p = *c.p;
  RAAE_SMEM();
  const int tid = threadIdx.x, ns = p.cfg.nstyle;
  const int LE = p.lay.net[kE].n_linear;
  if (tid == 0) {
    sm->alpha = alpha_schedule(c);
    for (int i = 0; i < 8; ++i) sm->loss_acc[i] = 0.0;
  }
  __syncthreads();
  if (!run_p0) phase_mask &= ~(1 << kAdv);       // the adversarial phase back-propagates through the P0 forward
  LayerIn x = wide_in(c.sc + p.sl.xn, p.sl.xld, p.cfg.dim_in, 0);
  x.img = 1;
  const float* zE = c.sc + p.sl.zE;
  const float* meanZ = sm->mean[kE][LE - 1];
  const float* invZ = sm->inv[kE][LE - 1];
  float* dz = c.sc + p.sl.dz;
  const int act = p.cfg.decoder_softplus ? 1 : 2;
  const LayerIn zin = latent_in(zE, ns, LE - 1);

  // P0 (trainer.py:113-114): styles = E(x); spec_out = D(styles) is unused, but the decoder's BatchNorm
  // buffers advance, so its hidden blocks run (the output Linear has no side effect and is skipped).
  if (run_p0) {
    encoder_forward(c, x, 0);
    if (c.a->debug && c.a->dbg.styles && c.crank == 0) {
      for (int i = tid; i < c.B * ns; i += kThreads) {
        int r = i / ns, k = i - r * ns;
        c.a->dbg.styles[i] = (zE[(size_t)r * kZ + k] - meanZ[k]) * invZ[k];
      }
    }
    decoder_forward_hidden(c, zin, 0);
  }

  // P1 adversarial (trainer.py:118-127)
  if (phase_mask & (1 << kAdv)) {
    if (tid == 0) adam_prepare(c, sm, kAdv);
    __syncthreads();
    dis_stage(c, 1, kAdv, c.a->debug ? c.a->dbg.z_real : nullptr, stream_key(c.seed, c.step_id, kStreamZReal));
    encoder_backward(c, x, 0, kAdv, nullptr);
    if (tid == 0) adam_finish(c, kAdv);
    __syncthreads();
  }
  // P2 Kendall constraint (trainer.py:153-161)
  if (phase_mask & (1 << kCorr)) {
    encoder_forward(c, x, 1);
    kendall_stage(c, c.sc + p.sl.aux, 1);
    if (tid == 0) adam_prepare(c, sm, kCorr);
    __syncthreads();
    encoder_backward(c, x, 1, kCorr, nullptr);
    if (tid == 0) adam_finish(c, kCorr);
    __syncthreads();
  }
  // P3 reconstruction (trainer.py:164-172)
  if (phase_mask & (1 << kRecon)) {
    encoder_forward(c, x, 2);
    const LayerIn z = zin;
    decoder_forward_hidden(c, z, 1);
    if (tid == 0) adam_prepare(c, sm, kRecon);
    __syncthreads();
    dec_last(c, kLastRecon, 1, kRecon);
    decoder_backward_hidden(c, z, 1, kRecon, dz);
    encoder_backward(c, x, 2, kRecon, nullptr);
    if (tid == 0) adam_finish(c, kRecon);
    __syncthreads();
  }
  // P4 mutual information (trainer.py:175-186)
  if (phase_mask & (1 << kMI)) {
    encoder_forward(c, x, 3);                       // trainer.py:176: result unused, BN buffers advance
    const LayerIn zs = latent_in(c.sc + p.sl.zs, ns, -1);
    decoder_forward_hidden(c, zs, 2);
    dec_last(c, kLastStoreV, 2, kMI);
    LayerIn y = wide_in(c.sc + p.sl.v, p.sl.vld, p.cfg.dim_out, act);
    y.img = (p.cfg.dim_in == p.cfg.dim_out) ? 2 : 0;     // dec_last(kLastStoreV) wrote the operand image of act(v)
    encoder_forward(c, y, 4);
    mi_mse_stage(c, 1);
    if (tid == 0) adam_prepare(c, sm, kMI);
    __syncthreads();
    encoder_backward(c, y, 4, kMI, c.sc + p.sl.v);
    dec_last(c, kLastFromDv, 2, kMI);
    decoder_backward_hidden(c, zs, 2, kMI, nullptr);
    if (tid == 0) adam_finish(c, kMI);
    __syncthreads();
  }
  // P5 smoothness (trainer.py:189-200): only the decoder is stepped, so the encoder backward is skipped
  if ((phase_mask & (1 << kSmooth)) && (double)c.epoch < c.hp[RAAE_HP_EPOCH_STOP_SMOOTH]) {
    encoder_forward(c, x, 5);
    const LayerIn z = zin;
    decoder_forward_hidden(c, z, 3);
    if (tid == 0) adam_prepare(c, sm, kSmooth);
    __syncthreads();
    dec_last(c, kLastSmooth, 3, kSmooth);
    decoder_backward_hidden(c, z, 3, kSmooth, nullptr);
    if (tid == 0) adam_finish(c, kSmooth);
    __syncthreads();
  }
  if (tid == 0 && c.crank == 0) {
    float* misc = c.st + p.lay.misc_off;
    for (int i = 0; i < RAAE_NUM_PHASES; ++i)
      if (phase_mask & (1 << i)) misc[i] = (float)sm->loss_acc[i];
    if (phase_mask & (1 << kMI)) {
      misc[5] += (float)sm->loss_acc[kMI];
      misc[6] += 1.f;
    }
    if (c.a->debug && c.a->dbg.losses)
      for (int i = 0; i < RAAE_NUM_PHASES; ++i) c.a->dbg.losses[i] = (float)sm->loss_acc[i];
  }
  stage_sync(c);
}

__device__ __forceinline__ void init_ctx(Ctx& c, const KParams& p, const RunArgs& a, int trial) {
  c.p = &p;
  c.a = &a;
  c.st = p.state + (size_t)trial * p.lay.state_floats;
  c.sc = p.scratch + (size_t)trial * p.lay.scratch_floats;
  c.hp = p.hp + (size_t)trial * RAAE_HP_COUNT;
  c.Breal = p.cfg.batch_size;
  c.seed = mix32((uint32_t)(long long)c.hp[RAAE_HP_SEED] * 0x9e3779b9U + (uint32_t)trial * 0x85ebca6bU + 1u);
  c.epoch = a.epoch;
  c.trial = trial;
  c.apply = 1;
  c.train = 1;
  c.x = c.sc + p.sl.xn;
  c.xld = p.sl.xld;
#if RAAE_CLUSTER
  c.crank = (int)cl::ctarank();
  c.csize = (int)cl::nctarank();
#endif
#pragma unroll
  for (int g = 0; g < 2; ++g) {
    const double pd = c.hp[g ? RAAE_HP_DIS_DROPOUT : RAAE_HP_DROPOUT];
    c.drop_scale[g] = pd > 0.0 ? 1.f / (float)(1.0 - pd) : 1.f;
    c.drop_thresh[g] = pd > 0.0 ? (uint32_t)(pd * 65536.0 + 0.5) : 0u;
  }
}

constexpr size_t kSmemBytes = kArenaOffset + (size_t)kArenaFloats * sizeof(float);

// TMEM accumulator + mbarrier for the tcgen05 path (one CTA per SM, so the allocation never contends)
__device__ __forceinline__ void tc_setup(const KParams& p, SmemFixed* sm) {
  // cluster per trial: no CTA may touch a peer's shared memory before that peer runs, nor exit while a peer may still read its own
#if RAAE_CLUSTER
  if (threadIdx.x == 0) sm->xpar = 0u;
  cl::sync();
#endif
  if (!p.cfg.tensor_cores) return;
  if (threadIdx.x < 32) tc::tmem_alloc(&sm->tmem_base, tc::kTmemCols);
  if (threadIdx.x == 0) {
    tc::mbar_init(reinterpret_cast<uint64_t*>(&sm->mbar), 1);
    tc::mbar_init(reinterpret_cast<uint64_t*>(&sm->mbar2), 1);
    sm->tc_phase = 0;
    sm->tc_phase2 = 0;
  }
  tc::fence_before_sync();
  __syncthreads();
  tc::fence_after_sync();
}
__device__ __forceinline__ void tc_teardown(const KParams& p, SmemFixed* sm) {
#if RAAE_CLUSTER
  cl::sync();
#endif
  if (!p.cfg.tensor_cores) return;
  tc::fence_before_sync();
  __syncthreads();
  if (threadIdx.x < 32) tc::tmem_dealloc(sm->tmem_base, tc::kTmemCols);
}

__global__ void __launch_bounds__(kThreads, 1)
raae_train_kernel(const __grid_constant__ KParams p, const __grid_constant__ RunArgs a) {
  RAAE_SMEM();
#if RAAE_CLUSTER
  const int trial = a.trial0 + (int)(blockIdx.x / cl::nctarank());     // one thread-block cluster (ctas_per_trial CTAs) per trial
#else
  const int trial = a.trial0 + blockIdx.x;
#endif
  Ctx& c = sm->ctx;                      // one copy per CTA in shared memory: thread 0 writes, a barrier publishes
  const int bs = p.cfg.batch_size;
  if (threadIdx.x == 0) {
    init_ctx(c, p, a, trial);
    if (a.debug) {
      c.B = a.dbg.rows;
      c.epoch = a.dbg.epoch;
      c.apply = a.dbg.apply_updates;
      c.step_id = 0;
    } else if (a.split) {
      c.B = min(bs, p.n_train - a.step0 * bs);
      c.step_id = (uint32_t)(a.epoch * a.n_steps + a.step0);
      c.apply = 0;
    }
  }
  if (threadIdx.x < 32) sm->prof[threadIdx.x] = 0;
  const long long t_start = clock64();
  __syncthreads();
  tc_setup(p, sm);
  if (a.debug) {
    build_batch(c, nullptr);
    train_step(c, a.dbg.phase_mask);
    tc_teardown(p, sm);
    return;
  }
  const int32_t* perm = a.perm + (size_t)trial * p.n_train;
  if (a.split) {
    // one batch, selected phases, gradients exported instead of applied (data-parallel mode: the host all-reduces
    // them and raae_adam_kernel applies the update).  The batch and the P0 forward belong to the launch that runs P1.
    const bool first = (a.phase_mask & (1 << kAdv)) != 0;
    if (first && a.step0 == 0 && threadIdx.x == 0 && c.crank == 0) { c.st[p.lay.misc_off + 5] = 0.f; c.st[p.lay.misc_off + 6] = 0.f; }
    if (first) build_batch(c, perm + a.step0 * bs);
    train_step(c, a.phase_mask, first);
    tc_teardown(p, sm);
    return;
  }
  if (threadIdx.x == 0 && c.crank == 0) { c.st[p.lay.misc_off + 5] = 0.f; c.st[p.lay.misc_off + 6] = 0.f; }
  for (int s = 0; s < a.n_steps; ++s) {
    if (threadIdx.x == 0) {              // every reader of the previous step's values is behind the barrier that ended it
      c.B = min(bs, p.n_train - s * bs);
      c.step_id = (uint32_t)(a.epoch * a.n_steps + s);
    }
    __syncthreads();
    build_batch(c, perm + s * bs);
    train_step(c, 0x1f);
  }
  if (a.prof && threadIdx.x == 0 && c.crank == 0) {      // [n_trials][32]: per-stage-type SM cycles, slot 15 = whole kernel, 16.. = probes
    for (int i = 0; i < 15; ++i) a.prof[(size_t)trial * 32 + i] += sm->prof[i];
    a.prof[(size_t)trial * 32 + 15] += clock64() - t_start;
    for (int i = 16; i < 32; ++i) a.prof[(size_t)trial * 32 + i] += sm->prof[i];
  }
  tc_teardown(p, sm);
}

// ------------------------------------------------------------------------------------------
// validation block (trainer.py:207-304)
// ------------------------------------------------------------------------------------------
// in-place bitonic sort of (key, idx) pairs in shared memory, npad a power of two
__device__ __forceinline__ void bitonic_sort(float* key, int* idx, int npad) {
  for (int k = 2; k <= npad; k <<= 1) {
    for (int j = k >> 1; j > 0; j >>= 1) {
      __syncthreads();
      for (int i = threadIdx.x; i < npad; i += kThreads) {
        int ixj = i ^ j;
        if (ixj > i) {
          bool up = (i & k) == 0;
          float a = key[i], b = key[ixj];
          if ((a > b) == up) {
            key[i] = b; key[ixj] = a;
            int t = idx[i]; idx[i] = idx[ixj]; idx[ixj] = t;
          }
        }
      }
    }
  }
  __syncthreads();
}

// scipy.stats.shapiro(z_k).statistic per style (Royston weights supplied by the host)
// and max |Spearman| over style pairs (trainer.py:286-293).  Results: sm->zs[2][0] = min W, zs[2][1] = coupling.
__device__ __noinline__ void latent_metrics(const Ctx& c_ref) {
  const Ctx c = c_ref;                 // register copy of the kernel context (SmemFixed::ctx, shared memory)
  RAAE_SMEM();
  StageTimer timer_(&sm->prof[kStBatch]);      // the validation kernel builds no batch: slot 0 is free there
  const KParams& p = *c.p;
  const int n = c.B, ns = p.cfg.nstyle, tid = threadIdx.x;
  const int lE = p.lay.net[kE].n_linear - 1;
  const float* zE = c.sc + p.sl.zE;
  float* ranks = c.sc + p.sl.rank;             // [kZ][max_rows]
  const int rstride = p.cfg.max_rows;
  int npad = 1;
  while (npad < n) npad <<= 1;
  float* key = arena;
  int* idx = reinterpret_cast<int*>(arena + npad);
  float wmin = 1e30f;
  // cluster per trial: style k is sorted by rank k % csize, rank-correlation pair q by rank q % csize; min / max merged at the end
  const int crank = c.crank, csize = c.csize;
  for (int k = 0; k < ns; ++k) {
    if (k % csize != crank) continue;
    __syncthreads();
    const float mu = sm->mean[kE][lE][k], is = sm->inv[kE][lE][k];
    for (int i = tid; i < npad; i += kThreads) {
      key[i] = i < n ? (zE[(size_t)i * kZ + k] - mu) * is : __int_as_float(0x7f800000);
      idx[i] = i;
    }
    bitonic_sort(key, idx, npad);
    double sw = 0.0, sx = 0.0;
    for (int i = tid; i < n; i += kThreads) { sw += (double)p.shapiro_w[i] * (double)key[i]; sx += (double)key[i]; }
    sw = block_sum_d(sw, sm->redd);
    sx = block_sum_d(sx, sm->redd);
    const double xm = sx / (double)n;
    double ss = 0.0;
    for (int i = tid; i < n; i += kThreads) { double d = (double)key[i] - xm; ss += d * d; }
    ss = block_sum_d(ss, sm->redd);
    float W = (float)(sw * sw / ss);
    wmin = fminf(wmin, W);
    // average ranks (scipy.stats.rankdata 'average')
    for (int i = tid; i < n; i += kThreads) {
      int lo = i, hi = i;
      const float v = key[i];
      while (lo > 0 && key[lo - 1] == v) --lo;
      while (hi + 1 < n && key[hi + 1] == v) ++hi;
      ranks[(size_t)k * rstride + idx[i]] = 0.5f * (float)(lo + hi) + 1.f;
    }
  }
  stage_sync(c);                         // the other CTAs' rank columns
  // Pearson correlation of the rank columns
  const double rm = 0.5 * ((double)n + 1.0);
  float cmax = 0.f;
  double var[kZ];
  for (int a = 0; a < ns; ++a) {
    double s = 0.0;
    for (int i = tid; i < n; i += kThreads) { double d = (double)ranks[(size_t)a * rstride + i] - rm; s += d * d; }
    var[a] = block_sum_d(s, sm->redd);
  }
  int pair = 0;
  for (int a = 0; a < ns; ++a)
    for (int b = a + 1; b < ns; ++b) {
      if ((pair++) % csize != crank) continue;
      double s = 0.0;
      for (int i = tid; i < n; i += kThreads)
        s += ((double)ranks[(size_t)a * rstride + i] - rm) * ((double)ranks[(size_t)b * rstride + i] - rm);
      s = block_sum_d(s, sm->redd);
      float r = (float)(s / sqrt(var[a] * var[b]));
      cmax = fmaxf(cmax, fabsf(r));
    }
  if (tid == 0) { sm->zs[2][0] = wmin; sm->zs[2][1] = cmax; }
  __syncthreads();
  if (csize > 1) {
    const uint32_t par = sm->xpar;
    float* x = &sm->xchg[par][0][0];
    if (tid < 2) x[tid] = sm->zs[2][tid];
    cl::sync();
    if (tid == 0) {
      float w = 1e30f, cm = 0.f;
      for (int r = 0; r < csize; ++r) {
        w = fminf(w, cl::ld_f32(cl::map(x, (uint32_t)r)));
        cm = fmaxf(cm, cl::ld_f32(cl::map(x + 1, (uint32_t)r)));
      }
      sm->zs[2][0] = w; sm->zs[2][1] = cm;
      sm->xpar = par ^ 1u;
    }
    __syncthreads();
  }
}

// torch ReduceLROnPlateau(mode="min", threshold_mode="rel", threshold=0.01, cooldown=0, min_lr=0, eps=1e-8)
// on every stepped optimizer with the same metric (trainer.py:303-304, 400-408); thread 0 only.
__device__ inline void plateau_step(const Ctx& c, double metric) {
  const double factor = c.hp[RAAE_HP_SCH_FACTOR], patience = c.hp[RAAE_HP_SCH_PATIENCE];
  for (int o = 0; o < RAAE_NUM_PHASES; ++o) {
    float* s = c.st + c.p->lay.opt[o].scalar_off;
    double lr = (double)s[0], best = (double)s[2], bad = (double)s[3];
    if (metric < best * (1.0 - 0.01)) { best = metric; bad = 0.0; }
    else bad += 1.0;
    if (bad > patience) {
      double nl = lr * factor;
      if (lr - nl > 1e-8) lr = nl;
      bad = 0.0;
    }
    s[0] = (float)lr; s[2] = (float)best; s[3] = (float)bad;
  }
}

__global__ void __launch_bounds__(kThreads, 1)
raae_val_kernel(const __grid_constant__ KParams p, const __grid_constant__ RunArgs a) {
  RAAE_SMEM();
#if RAAE_CLUSTER
  const int trial = a.trial0 + (int)(blockIdx.x / cl::nctarank());
#else
  const int trial = a.trial0 + blockIdx.x;
#endif
  Ctx& c = sm->ctx;                      // one copy per CTA in shared memory: thread 0 writes, a barrier publishes
  const int tid = threadIdx.x, ns = p.cfg.nstyle, K = p.cfg.n_aux;
  const int LE = p.lay.net[kE].n_linear;
#ifdef RAAE_PROFILE_VAL
  const long long t_start_val = clock64();
  if (tid < 32) sm->prof[tid] = 0;
#endif
  if (tid == 0) {
    init_ctx(c, p, a, trial);
    c.train = 0;
    c.apply = 0;
    c.B = p.n_val;
    c.x = p.spec_val;
    c.xld = p.cfg.dim_in;
    c.step_id = 0x40000000u + (uint32_t)a.epoch;
  }
  __syncthreads();
  raae_val_io io = a.val;
  if (io.per_trial) {                    // raae_evaluate_trials: one output block per trial
    const size_t tb = (size_t)(trial - a.trial0);
    if (io.losses) io.losses += tb * RAAE_NUM_PHASES;
    if (io.metrics) io.metrics += tb * 6;
    if (io.z) io.z += tb * (size_t)p.n_val * ns;
  }
  tc_setup(p, sm);
  if (tid == 0) {
    sm->alpha = alpha_schedule(c);
    for (int i = 0; i < 8; ++i) sm->loss_acc[i] = 0.0;
  }
  // descriptors of the validation rows, z_sample
  float* aux = c.sc + p.sl.aux;
  float* zs = c.sc + p.sl.zs;
  const uint32_t kz = stream_key(c.seed, c.step_id, kStreamValZSample);
  {
    const int nslots = c.csize == 1 ? c.B : cl::own_tiles(c.B, c.crank, c.csize) * kTM;      // cluster per trial: own rows
    for (int e = tid; e < nslots * kZ; e += kThreads) {
      const int k = e & 7, r = c.csize == 1 ? (e >> 3) : cl::slot_row(e >> 3, c.crank, c.csize);
      if (r >= c.B) continue;
      const int i = r * kZ + k;
      aux[i] = k < K ? p.aux_val[(size_t)r * K + k] : 0.f;
      zs[i] = k < ns ? (io.z_sample ? io.z_sample[(size_t)r * ns + k] : normal_at(kz, (uint32_t)i)) : 0.f;
    }
  }
  __syncthreads();
  const float* zE = c.sc + p.sl.zE;
  const float* meanZ = sm->mean[kE][LE - 1];
  const float* invZ = sm->inv[kE][LE - 1];
  const int act = p.cfg.decoder_softplus ? 1 : 2;
  // z = E(x_val); spec_out = D(z)   (trainer.py:213-214)
  encoder_forward(c, wide_in(p.spec_val, p.cfg.dim_in, p.cfg.dim_in, 0), 0);
  stage_sync(c);                         // eval-mode stages have no barrier of their own: every CTA's latent rows are in place
  if (io.z && c.crank == 0)
    for (int i = tid; i < c.B * ns; i += kThreads) {
      int r = i / ns, k = i - r * ns;
      io.z[i] = (zE[(size_t)r * kZ + k] - meanZ[k]) * invZ[k];
    }
  latent_metrics(c);
  const float wmin = sm->zs[2][0], coupling = sm->zs[2][1];
  kendall_stage(c, aux, 0);
  dis_stage(c, 0, kAdv, io.z_real, stream_key(c.seed, c.step_id, kStreamValZReal));
  decoder_forward_hidden(c, latent_in(zE, ns, LE - 1), 0);
  dec_last(c, kLastEval, 0, kRecon);
  // mutual information on z_sample (trainer.py:240-246)
  decoder_forward_hidden(c, latent_in(zs, ns, -1), 0);
  dec_last(c, kLastStoreV, 0, kMI);
  {
    LayerIn y = wide_in(c.sc + p.sl.v, p.sl.vld, p.cfg.dim_out, act);
    y.img = (p.cfg.dim_in == p.cfg.dim_out) ? 2 : 0;
    encoder_forward(c, y, 0);
  }
  mi_mse_stage(c, 0);
  if (tid == 0 && c.crank == 0) {
    float* misc = c.st + p.lay.misc_off;
    float avg_mi = io.avg_mutual_info > -1e30f ? io.avg_mutual_info : (misc[6] > 0.f ? misc[5] / misc[6] : 0.f);
    double m[5] = {(double)wmin, sm->loss_acc[kRecon], (double)avg_mi, (double)coupling, sm->loss_acc[kCorr]};
    // combined_metric = -sum(metric_weights * metrics), metric_weights = [1, -1, -0.01, -1, -1]  (trainer.py:35, 297)
    double combined = -(1.0 * m[0] - 1.0 * m[1] - 0.01 * m[2] - 1.0 * m[3] - 1.0 * m[4]);
    if (io.losses)
      for (int i = 0; i < RAAE_NUM_PHASES; ++i) io.losses[i] = (float)sm->loss_acc[i];
    if (io.metrics) {
      for (int i = 0; i < 5; ++i) io.metrics[i] = (float)m[i];
      io.metrics[5] = (float)combined;
    }
    if (!a.debug) {
      // production: losses.csv columns (trainer.py:84-87, 270-279) and the scheduler step
      if (a.out_losses) {
        float* o = a.out_losses + (size_t)trial * 12;
        o[0] = misc[kAdv];    o[1] = (float)sm->loss_acc[kAdv];
        o[2] = 0.f;           o[3] = 0.f;
        o[4] = misc[kCorr];   o[5] = (float)sm->loss_acc[kCorr];
        o[6] = misc[kRecon];  o[7] = (float)sm->loss_acc[kRecon];
        o[8] = (double)a.epoch < c.hp[RAAE_HP_EPOCH_STOP_SMOOTH] ? misc[kSmooth] : 0.f;
        o[9] = (float)sm->loss_acc[kSmooth];
        o[10] = misc[kMI];    o[11] = (float)sm->loss_acc[kMI];
      }
      if (a.out_metrics) {
        float* o = a.out_metrics + (size_t)trial * 6;
        for (int i = 0; i < 5; ++i) o[i] = (float)m[i];
        o[5] = (float)combined;
      }
      plateau_step(c, combined);
    }
  }
#ifdef RAAE_PROFILE_VAL
  // profiling builds only: stage cycles of the validation block in a SECOND [n_trials][32] block behind the train kernel's
  // (tools/stage_profile.py allocates both); slot 0 = latent_metrics, slot 15 = whole kernel
  if (a.prof && tid == 0 && c.crank == 0) {
    long long* vp = a.prof + ((size_t)p.cfg.n_trials + trial) * 32;
    for (int i = 0; i < 15; ++i) vp[i] += sm->prof[i];
    vp[15] += clock64() - t_start_val;
  }
#endif
  tc_teardown(p, sm);
}

// AdamW of optimizer `o` from a gradient vector in global memory (data-parallel mode); grid = (blocks, n_trials)
struct AdamScalars { float decay, w1, b2, w2, ss, bc2s; };
__device__ __forceinline__ AdamScalars adam_scalars(const float* st, const double* hp, const raae_opt_layout& ol, int o) {
  const double lr = (double)st[ol.scalar_off + 0], t = (double)st[ol.scalar_off + 1] + 1.0;
  const double b1 = hp[RAAE_HP_BETA1 + o], b2d = hp[RAAE_HP_BETA2 + o], wd = hp[RAAE_HP_WD + o];
  AdamScalars a;
  a.decay = (float)(1.0 - lr * wd); a.w1 = (float)(1.0 - b1); a.b2 = (float)b2d; a.w2 = (float)(1.0 - b2d);
  a.ss = (float)(lr / (1.0 - pow(b1, t))); a.bc2s = (float)sqrt(1.0 - pow(b2d, t));
  return a;
}
__device__ __forceinline__ void adam_element(const KParams& p, float* st, const raae_opt_layout& ol, const AdamScalars& a, int i, float gi) {
  int net = -1, rel = 0;
  for (int k = 0; k < RAAE_NUM_NETS; ++k)
    if (ol.net_off[k] >= 0 && i >= ol.net_off[k] && i < ol.net_off[k] + p.lay.net[k].n_params) { net = k; rel = i - ol.net_off[k]; }
  float* P = st + p.lay.net[net].param_off + rel;
  float* M = st + ol.m_off + i;
  float* V = st + ol.v_off + i;
  float pp = *P, m = *M, v = *V;
  adamw_update(pp, m, v, gi, a.decay, a.w1, a.b2, a.w2, a.ss, a.bc2s);
  *P = pp;
  *M = m;
  *V = v;
}
__global__ void raae_adam_kernel(const __grid_constant__ KParams p, int o, const float* __restrict__ grads) {
  const int trial = blockIdx.y;
  float* st = p.state + (size_t)trial * p.lay.state_floats;
  const raae_opt_layout& ol = p.lay.opt[o];
  const AdamScalars a = adam_scalars(st, p.hp + (size_t)trial * RAAE_HP_COUNT, ol, o);
  const float* g = grads + (size_t)trial * ol.n;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < ol.n; i += gridDim.x * blockDim.x) adam_element(p, st, ol, a, i, g[i]);
}

// ------------------------------------------------------------------------------------------
// Data-parallel mode, peer-memory exchange: gradient all-reduce (mean) over NVLink P2P loads fused with AdamW.
// Replaces all_reduce + divide + raae_adam_kernel + raae_adam_tick_kernel (trainer.py:125-204 under DDP semantics).
// Protocol, per call `seq` (the same on every rank): block (0,0) release-stores `seq` into word [rank] of every rank's
// flag array (this rank's vector was written by the preceding raae_train_phase launch on the same stream); every block
// acquire-polls its OWN flag array until all `world` words reached `seq`, then each element is the sum of the ranks'
// vectors in rank order (bit-identical on every rank) divided by world.  A vector of phase o is rewritten one step later,
// after >= 3 further exchanges which every peer enters only once its previous launch (the reader) has finished.
// ------------------------------------------------------------------------------------------
struct PeerArgs {
  const float* sums[RAAE_MAX_PEERS];     // every rank's gradient vector of this phase, summed over its replicas (peer-mapped), [opt[o].n]
  unsigned* flags[RAAE_MAX_PEERS];       // every rank's flag words [RAAE_MAX_PEERS]
  const float* lgrads;                   // local: the replicas' vectors [replicas][opt[o].n] as raae_train_phase wrote them
  float* lsum;                           // local: their sum (== lgrads when replicas == 1)
  unsigned* done;                        // local: blocks finished (for the optimizer step counter)
  unsigned* arrive;                      // local: blocks whose slice of lsum is complete
  int world, rank;
  int replicas;                          // trials resident per rank (identical weights, one shard each)
  unsigned seq;
  unsigned long long timeout_ns;         // how long a rank waits for its peers before it gives up (RAAE_PEER_TIMEOUT_S)
  unsigned* error;                       // local: set to `seq` when a peer did not arrive in time (the update is skipped)
};
__device__ __forceinline__ void st_release_sys(unsigned* p, unsigned v) {
  asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ unsigned ld_acquire_sys(const unsigned* p) {
  unsigned v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ float ld_relaxed_sys(const float* p) {
  float v;
  asm volatile("ld.relaxed.sys.global.f32 %0, [%1];" : "=f"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ unsigned long long global_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
  return t;
}
__global__ void raae_adam_peer_kernel(const __grid_constant__ KParams p, int o, const __grid_constant__ PeerArgs pa) {
  const int V = pa.replicas;
  const raae_opt_layout& ol = p.lay.opt[o];
  if (V > 1) {
    // local pre-reduction: the peers read ONE vector per rank, not `replicas` (replica order; every block sums its slice,
    // block 0 publishes once all slices are complete - all blocks of this small grid are resident)
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < ol.n; i += gridDim.x * blockDim.x) {
      float g = 0.f;
#pragma unroll 4
      for (int v = 0; v < V; ++v) g += pa.lgrads[(size_t)v * ol.n + i];
      pa.lsum[i] = g;
    }
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x == 0) atomicAdd(pa.arrive, 1u);
  }
  if (blockIdx.x == 0) {
    if (V > 1) {
      if (threadIdx.x == 0) {
        while (atomicAdd(pa.arrive, 0u) < gridDim.x) __nanosleep(32);
        *pa.arrive = 0u;
        __threadfence_system();
      }
      __syncthreads();
    }
    if (threadIdx.x < pa.world) {
      __threadfence_system();
      st_release_sys(pa.flags[threadIdx.x] + pa.rank, pa.seq);
    }
  }
  __shared__ int s_late;
  if (threadIdx.x == 0) s_late = 0;
  __syncthreads();
  if (threadIdx.x < pa.world) {
    const unsigned* f = pa.flags[pa.rank] + threadIdx.x;
    const unsigned long long t0 = global_ns();
    while ((int)(ld_acquire_sys(f) - pa.seq) < 0) {
      __nanosleep(64);
      // a peer never arrived: give up WITHOUT trapping (a trap poisons the CUDA context of the whole job); the update of this
      // exchange is skipped everywhere the wait failed and the error word tells the host (raae_peer_status)
      if (global_ns() - t0 > pa.timeout_ns) { s_late = 1; break; }
    }
  }
  __syncthreads();
  if (s_late) {
    if (threadIdx.x == 0) { *pa.error = pa.seq; __threadfence_system(); }
    return;
  }
  // `replicas` = the trials resident on every rank: they hold the SAME weights and act as further data-parallel ranks
  // (one CTA each in raae_train_phase), so the mean runs over world x replicas vectors - the ranks' pre-reduced sums added in
  // rank order - and the one update is applied to every local replica's state
  const float wf = (float)(pa.world * V);
  // the replicas share the hyper-parameter row (dp.py) and, as long as their schedulers agree, lr and step count:
  // the float64 AdamW scalars are formed once and only re-formed for a replica whose lr / step differ
  const AdamScalars a0 = adam_scalars(p.state, p.hp, ol, o);
  const float lr0 = p.state[ol.scalar_off + 0], t0 = p.state[ol.scalar_off + 1];
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < ol.n; i += gridDim.x * blockDim.x) {
    float g = 0.f;
#pragma unroll 4
    for (int r = 0; r < pa.world; ++r) g += ld_relaxed_sys(pa.sums[r] + i);
    if (pa.world * V > 1) g = g / wf;
    // the replicas' parameter / moment triples are independent: loads of four replicas are issued before the first store
    // (the compiler cannot hoist them across the stores itself), so the update costs one memory latency per four replicas
    int net = -1, rel = 0;
    for (int k = 0; k < RAAE_NUM_NETS; ++k)
      if (ol.net_off[k] >= 0 && i >= ol.net_off[k] && i < ol.net_off[k] + p.lay.net[k].n_params) { net = k; rel = i - ol.net_off[k]; }
    const size_t po = (size_t)p.lay.net[net].param_off + rel, mo = (size_t)ol.m_off + i, vo = (size_t)ol.v_off + i;
    for (int v0 = 0; v0 < V; v0 += 4) {
      float P4[4], M4[4], V4[4];
      bool same[4];
#pragma unroll
      for (int u = 0; u < 4; ++u)
        if (v0 + u < V) {
          const float* st = p.state + (size_t)(v0 + u) * p.lay.state_floats;
          P4[u] = st[po]; M4[u] = st[mo]; V4[u] = st[vo];
          same[u] = st[ol.scalar_off + 0] == lr0 && st[ol.scalar_off + 1] == t0;
        }
#pragma unroll
      for (int u = 0; u < 4; ++u)
        if (v0 + u < V) {
          float* st = p.state + (size_t)(v0 + u) * p.lay.state_floats;
          const AdamScalars a = same[u] ? a0 : adam_scalars(st, p.hp + (size_t)(v0 + u) * RAAE_HP_COUNT, ol, o);
          float pp = P4[u], m = M4[u], vv = V4[u];
          adamw_update(pp, m, vv, g, a.decay, a.w1, a.b2, a.w2, a.ss, a.bc2s);
          st[po] = pp;
          st[mo] = m;
          st[vo] = vv;
        }
    }
  }
  // optimizer step counters: by the last block to finish (every block has read the scalars by then)
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();
    if (atomicAdd(pa.done, 1u) == gridDim.x - 1) {
      *pa.done = 0;
      for (int v = 0; v < V; ++v) p.state[(size_t)v * p.lay.state_floats + ol.scalar_off + 1] += 1.f;
    }
  }
}
// step counter of optimizer o, after raae_adam_kernel
__global__ void raae_adam_tick_kernel(const __grid_constant__ KParams p, int o, int n_trials) {
  int trial = blockIdx.x * blockDim.x + threadIdx.x;
  if (trial < n_trials) p.state[(size_t)trial * p.lay.state_floats + p.lay.opt[o].scalar_off + 1] += 1.f;
}

// parity hook of the scheduler: feeds metrics[0..n) to plateau_step of `trial` and records (lr, best, num_bad) of every
// optimizer after each step (out [n][RAAE_NUM_PHASES][3]); single thread
__global__ void raae_plateau_debug_kernel(const __grid_constant__ KParams p, int trial, const double* __restrict__ metrics, int n,
                                          float* __restrict__ out) {
  if (blockIdx.x != 0 || threadIdx.x != 0) return;
  Ctx c;
  c.p = &p;
  c.st = p.state + (size_t)trial * p.lay.state_floats;
  c.hp = p.hp + (size_t)trial * RAAE_HP_COUNT;
  for (int e = 0; e < n; ++e) {
    plateau_step(c, metrics[e]);
    for (int o = 0; o < RAAE_NUM_PHASES; ++o) {
      const float* s = c.st + p.lay.opt[o].scalar_off;
      float* d = out + ((size_t)e * RAAE_NUM_PHASES + o) * 3;
      d[0] = s[0]; d[1] = s[2]; d[2] = s[3];
    }
  }
}

// lr <- hp, t <- 0, best <- +inf, bad <- 0; misc zeroed
__global__ void raae_reset_opt_kernel(const __grid_constant__ KParams p, int n_trials) {
  int trial = blockIdx.x * blockDim.x + threadIdx.x;
  if (trial >= n_trials) return;
  float* st = p.state + (size_t)trial * p.lay.state_floats;
  const double* hp = p.hp + (size_t)trial * RAAE_HP_COUNT;
  for (int o = 0; o < RAAE_NUM_PHASES; ++o) {
    float* s = st + p.lay.opt[o].scalar_off;
    s[0] = (float)hp[RAAE_HP_LR0 + o];
    s[1] = 0.f;
    s[2] = __int_as_float(0x7f800000);
    s[3] = 0.f;
  }
  for (int i = 0; i < 16; ++i) st[p.lay.misc_off + i] = 0.f;
}

}  // namespace raae
