// Loss-side stages of the fused AAE step: decoder output layer (+ reconstruction / smoothness losses),
// style discriminator (+ BCE through the gradient-reversal layer), Kendall rank constraint, latent MSE.
#pragma once
#include <type_traits>

#include "aae_step.cuh"

namespace raae {

enum LastMode { kLastRecon = 0, kLastSmooth = 1, kLastStoreV = 2, kLastFromDv = 3, kLastEval = 4 };

// ------------------------------------------------------------------------------------------
// Decoder output layer  v = a @ W^T + b,  y = act(v)   (model.py:558-561), fused per 128-row tile with
//   kLastRecon : recon_loss(scale = use_flex_spec_target)  functions.py:81-107, dL/dv, dW, db, g -> sc.g[0]
//   kLastSmooth: smoothness_loss                            functions.py:194-212, likewise
//   kLastStoreV: v -> sc.v                                  (MI phase forward)
//   kLastFromDv: dL/dv read from sc.v                       (MI phase backward)
//   kLastEval  : plain-MSE reconstruction and smoothness losses only (validation, trainer.py:223-239)
// Losses land in sm->loss_acc[kRecon] / [kSmooth].
// ------------------------------------------------------------------------------------------
// TCB: the backward contractions (g = dv W, dW = dv^T a) run on tcgen05 as well (config tensor_cores bit 5).  The 128 x 256
// gradient tile cannot share the 227 KB of shared memory with its hi / lo operand planes, so it is parked in TENSOR MEMORY
// (the 256 accumulator columns of the forward product, free once v has been copied out) and re-read slab by slab:
//   g  (TMEM columns 256..319) += dv[:, 64 s ..] W[64 s .., :]   per 64-column slab, dv slab staged K-major hi / lo, W^T slab by bulk copy
//   dW (TMEM columns 320..447, kept over the whole batch)        per 128-column half, dv half staged MN-major as ONE plane at a
//                                                                 time (hi: x a_hi, x a_lo; then lo: x a_hi), a staged MN-major hi / lo
// All staging goes through one 64 KB buffer, so the six passes of a tile are serial; they still cost less than half of the
// FP32-FMA contractions they replace.
template <bool TCB>
__device__ __noinline__ void dec_last_t(const Ctx& c_ref, int mode, int inst, int o) {
  const Ctx c = c_ref;                 // register copy of the kernel context (shared memory)
  RAAE_SMEM();
  StageTimer timer_(&sm->prof[mode == kLastStoreV ? kStDecLastStoreV : mode == kLastFromDv ? kStDecLastFromDv : kStDecLastLoss]);
  const raae_net_layout& nl = NL(c, kD);
  const int L = nl.n_linear, l = L - 1, N = nl.out_dim[l];
  const int tid = threadIdx.x, ty = tid >> 4, tx = tid & 15, c4 = tx * 4, lane = tid & 31, warp = tid >> 5;
  const float* Wg = netp(c, kD) + nl.w_off[l];
  const int act = c.p->cfg.decoder_softplus ? 1 : 2;
  float* Y = arena;                         // [kTM][kLDW]
  float* At = Y + kWideTile;                // [kTM][kLD]
  float* Wc = At + kTile;                   // [64][kLD]
  float* Wc2 = Wc + kWTile;                 // second weight-chunk buffer; aliases the row buffers (loss pass only)
  float* rowbuf = Wc + kWTile + warp * 576; // per warp: ypad[272] | ezp[288]
  float* vpanel = c.sc + c.p->sl.v;
  const int vld = c.p->sl.vld;
  const LayerIn in = hidden_out(c, kD, L - 2, inst);
  const bool want_bwd = mode == kLastRecon || mode == kLastSmooth || mode == kLastFromDv;
  const bool flex = c.p->cfg.use_flex_spec_target != 0;
  const bool tc_fwd = (c.p->cfg.tensor_cores & 16) != 0 && mode != kLastFromDv;
  const int nchN = (N + kH - 1) / kH;
  float* wl = c.sc + c.p->sl.wl;
  // tensor-core forward: operand buffers alias the Y tile (they are dead when the accumulator is copied out)
  float* Ahi = Y;                           // [2 K blocks][128][32] swizzled
  float* Alo = Ahi + tc::kATileFloats;
  float* Bb = Alo + tc::kATileFloats;       // 2 x [hi 4096 | lo 4096] weight chunk
  uint64_t* wfull = reinterpret_cast<uint64_t*>(&sm->pipe_bar[0]);     // [2] weight chunk landed
  uint64_t* wdone = reinterpret_cast<uint64_t*>(&sm->pipe_bar[2]);     // [2] MMAs of the chunk completed
  uint64_t* accfull = reinterpret_cast<uint64_t*>(&sm->pipe_bar[4]);   // all MMAs of the tile completed
  const uint32_t d_tmem = sm->tmem_base;
  __syncthreads();
  sm->bias[tid] = tid < N ? netp(c, kD)[nl.b_off[l] + tid] : 0.f;
  if (tc_fwd) {
    if (tid == 0)
      for (int i = 0; i < 5; ++i) tc::mbar_init(reinterpret_cast<uint64_t*>(&sm->pipe_bar[i]), 1);
    // K-major hi / lo image of W [N][64] in global scratch, one [hi 4096 | lo 4096] block per 64 output columns
    for (int i = tid; i < nchN * kH * 16; i += kThreads) {
      const int n = i >> 4, k4 = (i & 15) * 4;
      const float4 w = n < N ? *reinterpret_cast<const float4*>(Wg + (size_t)n * kH + k4) : make_float4(0.f, 0.f, 0.f, 0.f);
      float* blk = wl + (size_t)(n >> 6) * 8192;
      tc::split_store(blk, blk + 4096, tc::sw128_chunk_off(n & 63, k4, tc::kBBlockBytes), w);
    }
    tc::fence_async_all();
  }
  // ---- tcgen05 backward (TCB): buffers, barriers, W^T image ----
  float* const Xb = Y;                      // 16384 floats: K-major hi | lo of a 64-column slab of dv, or ONE MN-major plane of a 128-column half
  float* const Phi = Y + 16384;             // activations a, MN-major (B operand of dW): [2 blocks][128 rows][32] hi ...
  float* const Plo = Phi + tc::kATileFloats; // ... and lo
  float* const Wtb = Wc;                    // [hi 4096 | lo 4096] W^T slab (B operand of g); Wc + the row buffers = 35 KB, 1 KB aligned
  float* const wt = c.sc + c.p->sl.wk;      // its image in global scratch: the encoder's input-weight image is rebuilt by every stage that uses it
  uint64_t* const wtfull = reinterpret_cast<uint64_t*>(&sm->pipe_bar[5]);   // W^T slab landed
  uint64_t* const bdone = reinterpret_cast<uint64_t*>(&sm->pipe_bar[6]);    // MMA batch of the backward completed
  uint32_t n_wt = 0u, n_bd = 0u;            // completed phases seen
  const bool leader = tc::warp_uniform_id() == 0;
  const int erow = 32 * (warp & 3) + lane, eh = warp >> 2;                  // TMEM ownership: lane (row), column half
  const uint32_t trow = (uint32_t)(32 * (warp & 3)) << 16;
  if (TCB) {
    if (tid == 0) { tc::mbar_init(wtfull, 1); tc::mbar_init(bdone, 1); }
    // element (row k, column c) of slab c >> 6 = W[c][k]: K-major hi / lo image of W^T, one [hi 4096 | lo 4096] block per 64 output columns
    for (int i = tid; i < nchN * kH * 16; i += kThreads) {
      const int cc = i >> 4, k4 = (i & 15) * 4;
      const float4 w = cc < N ? *reinterpret_cast<const float4*>(Wg + (size_t)cc * kH + k4) : make_float4(0.f, 0.f, 0.f, 0.f);
      const float wv[4] = {w.x, w.y, w.z, w.w};
      float* blk = wt + (size_t)(cc >> 6) * 8192;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        float h, lo;
        tc::tf32_split(wv[j], h, lo);
        const uint32_t off = tc::sw128_chunk_off(k4 + j, (cc & 63) & ~3, tc::kBBlockBytes) + (uint32_t)((cc & 3) * 4);
        *reinterpret_cast<float*>(reinterpret_cast<char*>(blk) + off) = h;
        *reinterpret_cast<float*>(reinterpret_cast<char*>(blk + 4096) + off) = lo;
      }
    }
    tc::fence_async_all();
  }
  __syncthreads();
  uint32_t n_acc = 0u;                      // completed phases of accfull seen
  float accW[TCB ? 1 : 8][8];               // FP32-FMA path: this thread's 8 x 8 block of dW over the whole batch
#pragma unroll
  for (int i = 0; i < (TCB ? 1 : 8); ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j) accW[i][j] = 0.f;
  float dbp = 0.f;
  float sg4[4] = {0.f, 0.f, 0.f, 0.f}, sgx4[4] = {0.f, 0.f, 0.f, 0.f};
  double loss_a = 0.0, loss_b = 0.0;        // a: reconstruction, b: smoothness
  const float nB = (float)c.B, nN = (float)N, inv_nN = 1.f / (float)N;
  const double inv_B = 1.0 / (double)c.B, inv_BN = 1.0 / ((double)c.B * (double)N);
  // prefix sums of the (symmetric) taps: P[j] = w[0] + ... + w[j]; the replicate-padding overhang of the adjoint folds
  // into the end points with these weights (see the smoothness branch below)
  float tapP[8];
  {
    float acc = 0.f;
#pragma unroll
    for (int j = 0; j < 8; ++j) { acc += kGauss17[j]; tapP[j] = acc; }
  }
  const int ntiles = (c.B + kTM - 1) / kTM;
  float* const row_mae = (mode == kLastEval && c.a->val.row_mae)
                             ? c.a->val.row_mae + (c.a->val.per_trial ? (size_t)(c.trial - c.a->trial0) * c.p->n_val : 0) : nullptr;
  float* const sg_dst = c.csize > 1 ? sm->sgp : sm->sg;         // cluster per trial: partial sums, gathered after the barrier
  float* const sgx_dst = c.csize > 1 ? sm->sgxp : sm->sgx;
  for (int t = c.crank; t < ntiles; t += c.csize) {      // this CTA's tiles; `it` counts them
    const int it = tile_iter(c, t);
    const int row0 = t * kTM, nv = min(kTM, c.B - row0);
    const long long q0 = RAAE_PROFILE ? clock64() : 0ll;
    auto load_w = [&](int ck) {             // elected thread only
      const int b = ck & 1;
      tc::mbar_expect_tx(&wfull[b], 32768u);
      tc::bulk_g2s(Bb + b * 8192, wl + (size_t)ck * 8192, 32768u, &wfull[b]);
    };
    if (tc_fwd && tc::warp_uniform_id() == 0) {
      // the Y tile (which the operand buffers alias) is free since the barrier that ended the previous tile
      if (tc::elect_one()) { load_w(0); if (nchN > 1) load_w(1); }
      __syncwarp();
    }
    build_act_tile(At, in.src, row0, nv, sm->mean[in.snet][in.slayer], sm->inv[in.snet][in.slayer], in.slope, in.mask);
    if (tc_fwd) {
      // v = a W^T on the tensor core: a staged K-major hi / lo from this thread's own elements of At, weight chunks
      // streamed with bulk copies (two buffers), 24 MMAs per 64 output columns into TMEM columns [64 c, 64 c + 64)
      const uint32_t offK = tc::sw128_chunk_off(ty, c4, tc::kABlockBytes);
#pragma unroll
      for (int i = 0; i < kTM / 16; ++i)
        tc::split_store(Ahi, Alo, offK + (uint32_t)(i * 16 * 128), *reinterpret_cast<const float4*>(At + (ty + 16 * i) * kLD + c4));
      tc::fence_async_smem();
      tc::fence_before_sync();
      __syncthreads();
      if (tc::warp_uniform_id() == 0) {
        if (tc::elect_one()) {
          tc::fence_after_sync();
          // buffer b has been used t * uses_b + (ck >> 1) times before chunk ck of tile t: its barrier parities follow
          for (int ck = 0; ck < nchN; ++ck) {
            const int b = ck & 1;
            tc::mbar_wait(&wfull[b], (uint32_t)(((it * ((nchN + 1 - b) >> 1)) + (ck >> 1)) & 1));
            const float* Bh = Bb + b * 8192;
            tc::issue_gemm_3xtf32_acc(d_tmem + (uint32_t)(64 * ck), Ahi, Alo, Bh, Bh + 4096, 0u);
            tc::mma_commit(&wdone[b]);
            if (ck + 2 < nchN) {
              tc::mbar_wait(&wdone[b], (uint32_t)(((it * ((nchN + 1 - b) >> 1)) + (ck >> 1)) & 1));
              load_w(ck + 2);
            }
          }
          tc::mma_commit(accfull);
        }
        __syncwarp();
      }
      tc::mbar_wait(accfull, n_acc & 1u);
      ++n_acc;
      tc::fence_after_sync();
      // accumulator -> Y (+ bias): warp w owns TMEM lanes 32 (w & 3) .. + 31 and output columns 128 (w >> 2) .. + 127
      {
        const int row = 32 * (warp & 3) + lane;
        for (int q4 = 0; q4 < 4; ++q4) {
          const int col0 = 128 * (warp >> 2) + 32 * q4;
          if (col0 < nchN * kH) {
            float v[32];
            tc::tmem_ld32(d_tmem + ((uint32_t)(32 * (warp & 3)) << 16) + (uint32_t)col0, v);
#pragma unroll
            for (int j = 0; j < 32; j += 4) {
              const float4 bb = *reinterpret_cast<const float4*>(sm->bias + col0 + j);
              *reinterpret_cast<float4*>(Y + row * kLDW + col0 + j) = make_float4(v[j] + bb.x, v[j + 1] + bb.y, v[j + 2] + bb.z, v[j + 3] + bb.w);
            }
          }
        }
      }
      tc::fence_before_sync();
    } else if (mode != kLastFromDv) {
      // weight chunks double-buffered with cp.async (second buffer = the per-warp row buffers, unused during the GEMMs):
      // the copy of chunk c+1 overlaps the contraction of chunk c, one barrier per chunk
      prefetch_w_rows64(Wc, kLD, Wg, 0, N);
      cp_async_commit();
      for (int n0 = 0, ci = 0; n0 < N; n0 += kH, ++ci) {
        float* Wcur = (ci & 1) ? Wc2 : Wc;
        cp_async_wait<0>();
        __syncthreads();
        if (n0 + kH < N) { prefetch_w_rows64((ci & 1) ? Wc : Wc2, kLD, Wg, n0 + kH, N); cp_async_commit(); }
        float acc[8][4];
#pragma unroll
        for (int i = 0; i < 8; ++i)
#pragma unroll
          for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
        mma_nt<kH>(At, kLD, Wcur, kLD, acc, ty, tx);
#pragma unroll
        for (int i = 0; i < 8; ++i)
#pragma unroll
          for (int j = 0; j < 4; ++j)
            Y[(ty + 16 * i) * kLDW + n0 + tx + 16 * j] = acc[i][j] + sm->bias[n0 + tx + 16 * j];
      }
    } else {
      build_wide_tile(Y, vpanel, vld, N, row0, nv, 0);
    }
    __syncthreads();
    const long long q1 = RAAE_PROFILE ? clock64() : 0ll;
    if (mode == kLastStoreV) {
      // v -> panel (the MI backward needs act'(v) and overwrites it with dL/dv); y = act(v) - xref -> K-major operand image
      // of the re-encoding forward (same layout and reference row as the batch image, ScratchLayout::yk)
      const int cc = (tid & 63) * 4;
      const bool img = (c.p->cfg.tensor_cores & 4) != 0 && N == c.p->cfg.dim_in;
      const int nch64 = c.p->sl.nch64, nch128 = c.p->sl.nch128;
      float* yk = c.sc + c.p->sl.yk;
      float* ym = c.sc + c.p->sl.ym;
      float* yref = c.sc + c.p->sl.yref + c.crank * kMaxDim;      // per CTA: the consumers of its tiles add its reference back
      if (img && it == 0) {
        // reference row of the image: column means of y over the first (up to 32) rows - y of random latents need not be
        // close to the batch reference, and the tensor core's truncating accumulation wants small centred operands
        const int nref = min(32, nv);
        float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
        if (cc < N)
          for (int r = tid >> 6; r < nref; r += 4) {
            const float4 v4 = *reinterpret_cast<const float4*>(Y + r * kLDW + cc);
            if (act == 1) { acc.x += softplus2_f(v4.x); acc.y += softplus2_f(v4.y); acc.z += softplus2_f(v4.z); acc.w += softplus2_f(v4.w); }
            else { acc.x += fmaxf(v4.x, 0.f); acc.y += fmaxf(v4.y, 0.f); acc.z += fmaxf(v4.z, 0.f); acc.w += fmaxf(v4.w, 0.f); }
          }
        float* red = &sm->red[0][0];                 // [4][256]
        *reinterpret_cast<float4*>(red + (tid >> 6) * 256 + cc) = acc;
        __syncthreads();
        yref[tid] = (red[tid] + red[256 + tid] + red[512 + tid] + red[768 + tid]) / (float)nref;
        __threadfence_block();
        __syncthreads();
      }
      const float4 xr = (img && cc < N) ? *reinterpret_cast<const float4*>(yref + cc) : make_float4(0.f, 0.f, 0.f, 0.f);
      for (int r = tid >> 6; r < kTM; r += 4) {
        float4 y4 = make_float4(0.f, 0.f, 0.f, 0.f);
        if (r < nv && cc < N) {
          const float4 v4 = *reinterpret_cast<const float4*>(Y + r * kLDW + cc);
          *reinterpret_cast<float4*>(vpanel + (size_t)(row0 + r) * vld + cc) = v4;
          if (act == 1) { y4.x = softplus2_f(v4.x); y4.y = softplus2_f(v4.y); y4.z = softplus2_f(v4.z); y4.w = softplus2_f(v4.w); }
          else { y4.x = fmaxf(v4.x, 0.f); y4.y = fmaxf(v4.y, 0.f); y4.z = fmaxf(v4.z, 0.f); y4.w = fmaxf(v4.w, 0.f); }
          y4.x -= xr.x; y4.y -= xr.y; y4.z -= xr.z; y4.w -= xr.w;
        }
        if (img && cc < nch64 * 64) {
          float* bk = yk + (size_t)(t * nch64 + (cc >> 6)) * 8192;
          *reinterpret_cast<float4*>(reinterpret_cast<char*>(bk) + tc::sw128_chunk_off(r, cc & 63, tc::kABlockBytes)) = y4;
          if (c.train && cc < nch128 * 128) {        // MN-major image for the weight gradient of the re-encoding forward
            float* bm = ym + (size_t)(t * nch128 + (cc >> 7)) * 16384;
            *reinterpret_cast<float4*>(reinterpret_cast<char*>(bm) + tc::sw128_32b_chunk_off(r, cc & 127, tc::kABlockBytes)) = y4;
          }
        }
      }
      if (img) tc::fence_async_all();        // read by bulk copies in the next stage
      __syncthreads();
      continue;
    }
    if (mode != kLastFromDv) {
      // ---- per-row losses and dL/dv; one warp per row, lane owns 8 consecutive columns.  The row loop is specialised on
      // (mode, activation, full 256-column rows) so that its unrolled bodies carry no run-time branches ----
      auto row_pass = [&](auto mode_c, auto act_c, auto full_c) {
        constexpr int MODE = decltype(mode_c)::value;
        constexpr int ACT = decltype(act_c)::value;
        constexpr bool FULL = decltype(full_c)::value;
      float* ypad = rowbuf;
        float* ezp = rowbuf + 272;
        const bool need_x = (MODE == kLastRecon) || (MODE == kLastEval);
        float xn[8];                  // target row of the NEXT iteration, prefetched one row ahead
#pragma unroll
        for (int e = 0; e < 8; ++e) xn[e] = 0.f;
        if (need_x && warp < nv) {
          const float* xrow = c.x + (size_t)(row0 + warp) * c.xld;
#pragma unroll
          for (int e = 0; e < 8; ++e) xn[e] = (FULL || lane * 8 + e < N) ? xrow[lane * 8 + e] : 0.f;
        }
        for (int r = warp; r < kTM; r += kThreads / 32) {
          float* yrow = Y + r * kLDW;
          const int col0 = lane * 8;
          if (r >= nv) {
#pragma unroll
            for (int e = 0; e < 8; ++e) yrow[col0 + e] = 0.f;
            continue;
          }
          float x[8];
#pragma unroll
          for (int e = 0; e < 8; ++e) x[e] = xn[e];
          if (need_x && r + kThreads / 32 < nv) {
            const float* xrow = c.x + (size_t)(row0 + r + kThreads / 32) * c.xld;
#pragma unroll
            for (int e = 0; e < 8; ++e) xn[e] = (FULL || col0 + e < N) ? xrow[col0 + e] : 0.f;
          }
          float v[8], y[8], dy[8];
#pragma unroll
          for (int e = 0; e < 8; ++e) {
            v[e] = yrow[col0 + e];
            y[e] = (FULL || col0 + e < N) ? ((ACT == 1) ? softplus2_f(v[e]) : fmaxf(v[e], 0.f)) : 0.f;
            dy[e] = 0.f;
          }
          if (need_x) {
            float sy = 0.f, sx = 0.f;
#pragma unroll
            for (int e = 0; e < 8; ++e) { sy += y[e]; sx += x[e]; }
            float cc = 1.f, dr = 0.f;
            if ((MODE == kLastRecon) && flex) {
#pragma unroll
              for (int of = 16; of > 0; of >>= 1) { sy += __shfl_xor_sync(0xffffffffu, sy, of); sx += __shfl_xor_sync(0xffffffffu, sx, of); }
              // reciprocal-multiply divisions (2 ulp): an IEEE division carries a slow-path branch, and four of them per row
              // cut the unrolled row body into pieces
              float m_out = sy * inv_nN, m_in = sx * inv_nN;
              float rr = __fdividef(fabsf(m_out), fabsf(m_in));
              cc = fminf(fmaxf(rr, 0.7f), 1.3f);
              float sgn = m_out > 0.f ? 1.f : (m_out < 0.f ? -1.f : 0.f);
              dr = __fdividef((0.2f / nB) * (rr - 1.f) * sgn, fabsf(m_in) * nN);
              if (lane == 0) loss_a += 0.1 * (double)((rr - 1.f) * (rr - 1.f)) * inv_B;
            }
            float sq = 0.f;
#pragma unroll
            for (int e = 0; e < 8; ++e) {
              if (FULL || col0 + e < N) {
                float d = y[e] - x[e] * cc;
                sq = fmaf(d, d, sq);
                dy[e] = dr + (2.f / (nB * nN)) * d;
              }
            }
            loss_a += (double)sq * inv_BN;
            if ((MODE == kLastEval) && row_mae != nullptr) {
              // mean absolute error of the row (report tooling: sc/report/analysis.py:425-428)
              float sa = 0.f;
#pragma unroll
              for (int e = 0; e < 8; ++e)
                if (FULL || col0 + e < N) sa += fabsf(y[e] - x[e]);
#pragma unroll
              for (int of = 16; of > 0; of >>= 1) sa += __shfl_xor_sync(0xffffffffu, sa, of);
              if (lane == 0) row_mae[row0 + r] = sa / nN;
            }
          }
          if ((MODE == kLastSmooth) || (MODE == kLastEval)) {
            // replicate-padded copy of the row
#pragma unroll
            for (int e = 0; e < 8; ++e)
              if (FULL || col0 + e < N) ypad[8 + col0 + e] = y[e];
            __syncwarp();
            if (lane < 8) { ypad[lane] = ypad[8]; ypad[8 + N + lane] = ypad[8 + N - 1]; }
            __syncwarp();
            float ee[8], sq = 0.f;
            {
              // sliding window in registers: ypad[col0 .. col0 + 24)
              float win[24];
#pragma unroll
              for (int q4 = 0; q4 < 6; ++q4) {
                float4 t4 = *reinterpret_cast<const float4*>(ypad + col0 + 4 * q4);
                win[4 * q4] = t4.x; win[4 * q4 + 1] = t4.y; win[4 * q4 + 2] = t4.z; win[4 * q4 + 3] = t4.w;
              }
#pragma unroll
              for (int e = 0; e < 8; ++e) {
                ee[e] = 0.f;
                if (FULL || col0 + e < N) {
                  float s = 0.f;
#pragma unroll
                  for (int k = 0; k < 17; ++k) s = fmaf(kGauss17[k], win[e + k], s);
                  ee[e] = y[e] - s;
                  sq = fmaf(ee[e], ee[e], sq);
                }
              }
            }
            loss_b += (double)sq * inv_BN;
            if ((MODE == kLastSmooth)) {
              // adjoint: zero-padded full correlation of e with the taps, overhang folded into the end points
              if (lane < 16) { ezp[lane] = 0.f; ezp[16 + N + lane] = 0.f; }
#pragma unroll
              for (int e = 0; e < 8; ++e)
                if (FULL || col0 + e < N) ezp[16 + col0 + e] = ee[e];
              __syncwarp();
              {
                // (K^T e)[j + 8] = sum_t w[t] e[j + 8 - t] = sum_m w[m] ezp[j + 8 + m]  (taps symmetric, ezp[16 + i] = e[i])
                float win[24];
#pragma unroll
                for (int q4 = 0; q4 < 6; ++q4) {
                  float4 t4 = *reinterpret_cast<const float4*>(ezp + col0 + 8 + 4 * q4);
                  win[4 * q4] = t4.x; win[4 * q4 + 1] = t4.y; win[4 * q4 + 2] = t4.z; win[4 * q4 + 3] = t4.w;
                }
#pragma unroll
                for (int e = 0; e < 8; ++e) {
                  const int j = col0 + e;
                  if (FULL || j < N) {
                    float kt = 0.f;
#pragma unroll
                    for (int k = 0; k < 17; ++k) kt = fmaf(kGauss17[k], win[e + k], kt);
                    // overhang of the replicate padding: sum_{i<8} (K^T e)[i] = sum_{m<8} e[m] P[7-m] folds into y[0],
                    // and by symmetry sum_{q<8} e[N-1-q] P[7-q] into y[N-1]
                    if (e == 0 && col0 == 0) {           // j == 0 (e is a compile-time index: one test per row, not eight)
#pragma unroll
                      for (int m = 0; m < 8; ++m) kt = fmaf(tapP[7 - m], ezp[16 + m], kt);
                    }
                    if (FULL ? (e == 7 && col0 == kMaxDim - 8) : (j == N - 1)) {
#pragma unroll
                      for (int q2 = 0; q2 < 8; ++q2) kt = fmaf(tapP[7 - q2], ezp[16 + N - 1 - q2], kt);
                    }
                    dy[e] = (2.f / (nB * nN)) * (ee[e] - kt);
                  }
                }
              }
            }
            __syncwarp();
          }
#pragma unroll
          for (int e = 0; e < 8; ++e) {
            float g = 0.f;
            if (FULL || col0 + e < N) g = dy[e] * ((ACT == 1) ? softplus2_grad_f(v[e]) : (v[e] > 0.f ? 1.f : 0.f));
            yrow[col0 + e] = g;
          }
        }
      };
      {
        using I0 = std::integral_constant<int, kLastRecon>;
        using I1 = std::integral_constant<int, kLastSmooth>;
        using I4 = std::integral_constant<int, kLastEval>;
        using A1 = std::integral_constant<int, 1>;
        using A2 = std::integral_constant<int, 2>;
        using BT = std::integral_constant<bool, true>;
        using BF = std::integral_constant<bool, false>;
        const bool full = N == kMaxDim;
#define RAAE_ROWPASS(M)                                                     \
        if (act == 1) { if (full) row_pass(M{}, A1{}, BT{}); else row_pass(M{}, A1{}, BF{}); } \
        else          { if (full) row_pass(M{}, A2{}, BT{}); else row_pass(M{}, A2{}, BF{}); }
        if (mode == kLastRecon) { RAAE_ROWPASS(I0) }
        else if (mode == kLastSmooth) { RAAE_ROWPASS(I1) }
        else { RAAE_ROWPASS(I4) }
#undef RAAE_ROWPASS
      }
      __syncthreads();
    }
    const long long q2 = RAAE_PROFILE ? clock64() : 0ll;
    if (RAAE_PROFILE && !RAAE_PROF_FWD && tid == 0 && (mode == kLastRecon || mode == kLastSmooth)) { sm->prof[16] += q1 - q0; sm->prof[17 + (mode == kLastSmooth)] += q2 - q1; }
    if (want_bwd) {
      // db, dW += dv^T a, g = (dv @ W) * dropout
      {
        float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;       // four independent chains instead of 128 dependent additions
#pragma unroll 4
        for (int r = 0; r < kTM; r += 4) {
          s0 += Y[r * kLDW + tid]; s1 += Y[(r + 1) * kLDW + tid]; s2 += Y[(r + 2) * kLDW + tid]; s3 += Y[(r + 3) * kLDW + tid];
        }
        dbp += (s0 + s1) + (s2 + s3);
      }
      float acc[8][4];
      long long q4 = 0ll;
      if constexpr (TCB) {
        // ---- dv: shared memory -> TMEM columns [0, 256): thread = row (TMEM lane), column half eh ----
        if (leader && tc::elect_one()) {                 // first W^T slab: lands behind the copy and the first staging pass
          tc::mbar_expect_tx(wtfull, 32768u);
          tc::bulk_g2s(Wtb, wt, 32768u, wtfull);
        }
#pragma unroll
        for (int q4 = 0; q4 < 4; ++q4) {
          const int col0 = 128 * eh + 32 * q4;
          float v[32];
#pragma unroll
          for (int j = 0; j < 32; j += 4) {
            const float4 t4 = *reinterpret_cast<const float4*>(Y + erow * kLDW + col0 + j);
            v[j] = t4.x; v[j + 1] = t4.y; v[j + 2] = t4.z; v[j + 3] = t4.w;
          }
          tc::tmem_st32(d_tmem + trow + (uint32_t)col0, v);
        }
        tc::tmem_wait_st();
        tc::fence_before_sync();
        __syncthreads();                                 // every row of dv is in TMEM: the Y area becomes the operand buffers
        tc::fence_after_sync();
        // a -> MN-major hi / lo (this thread's (ty, c4) chunks); rows >= nv of the activation tile are zero
        {
          const uint32_t offM = tc::sw128_32b_chunk_off(ty, c4, tc::kABlockBytes);
#pragma unroll
          for (int i = 0; i < kTM / 16; ++i)
            tc::split_store(Phi, Plo, offM + (uint32_t)(i * 16 * 128), *reinterpret_cast<const float4*>(At + (ty + 16 * i) * kLD + c4));
        }
        // ---- g += dv[:, 64 s ..] W[64 s .., :]: slab staged K-major hi / lo from TMEM (row erow, columns 64 s + 32 eh ..) ----
        for (int s2 = 0; s2 < nchN; ++s2) {
          {
            float v[32];
            tc::tmem_ld32(d_tmem + trow + (uint32_t)(64 * s2 + 32 * eh), v);
#pragma unroll
            for (int j = 0; j < 32; j += 4)
              tc::split_store(Xb, Xb + tc::kATileFloats, tc::sw128_chunk_off(erow, 32 * eh + j, tc::kABlockBytes),
                              make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]));
          }
          tc::fence_async_smem();
          tc::fence_before_sync();
          __syncthreads();
          if (leader && tc::elect_one()) {
            tc::mbar_wait(wtfull, n_wt & 1u);
            tc::fence_after_sync();
            tc::issue_gemm_3xtf32_acc(d_tmem + 256u, Xb, Xb + tc::kATileFloats, Wtb, Wtb + 4096, s2 > 0 ? 1u : 0u);
            tc::mma_commit(bdone);
          }
          ++n_wt;
          tc::mbar_wait(bdone, n_bd & 1u);
          ++n_bd;
          if (s2 + 1 < nchN && leader && tc::elect_one()) {   // the slab buffer is free again
            tc::mbar_expect_tx(wtfull, 32768u);
            tc::bulk_g2s(Wtb, wt + (size_t)(s2 + 1) * 8192, 32768u, wtfull);
          }
        }
        // ---- dW[128 h ..][:] += dv[:, 128 h ..]^T a: the half staged MN-major, its rounded hi plane first (x a_hi, x a_lo),
        //      then the residual plane (x a_hi); thread (erow, eh) owns columns 128 h + 64 eh .. + 63 (blocks 2 eh, 2 eh + 1).
        //      The half is read from TMEM once; the residual plane is formed in registers while the MMAs of the hi plane run ----
        for (int h2 = 0; h2 < (nchN + 1) / 2; ++h2) {
          float v0[32], v1[32];
          tc::tmem_ld32(d_tmem + trow + (uint32_t)(128 * h2 + 64 * eh), v0);
          tc::tmem_ld32(d_tmem + trow + (uint32_t)(128 * h2 + 64 * eh + 32), v1);
          const uint32_t dW_t = d_tmem + 320u + (uint32_t)(64 * h2);
          char* const xb0 = reinterpret_cast<char*>(Xb);
#pragma unroll
          for (int j = 0; j < 32; j += 4) {
            float4 h0, h1;
            h0.x = tc::tf32_rna(v0[j]); h0.y = tc::tf32_rna(v0[j + 1]); h0.z = tc::tf32_rna(v0[j + 2]); h0.w = tc::tf32_rna(v0[j + 3]);
            h1.x = tc::tf32_rna(v1[j]); h1.y = tc::tf32_rna(v1[j + 1]); h1.z = tc::tf32_rna(v1[j + 2]); h1.w = tc::tf32_rna(v1[j + 3]);
            *reinterpret_cast<float4*>(xb0 + tc::sw128_32b_chunk_off(erow, 64 * eh + j, tc::kABlockBytes)) = h0;
            *reinterpret_cast<float4*>(xb0 + tc::sw128_32b_chunk_off(erow, 64 * eh + 32 + j, tc::kABlockBytes)) = h1;
            v0[j] -= h0.x; v0[j + 1] -= h0.y; v0[j + 2] -= h0.z; v0[j + 3] -= h0.w;       // exact residuals
            v1[j] -= h1.x; v1[j + 1] -= h1.y; v1[j + 2] -= h1.z; v1[j + 3] -= h1.w;
          }
          tc::fence_async_smem();
          tc::fence_before_sync();
          __syncthreads();
          if (leader && tc::elect_one()) {
            tc::fence_after_sync();
            tc::issue_gemm_tn128_pass(dW_t, Xb, tc::kABlockBytes, Phi, tc::kABlockBytes, it > 0 ? 1u : 0u);
            tc::issue_gemm_tn128_pass(dW_t, Xb, tc::kABlockBytes, Plo, tc::kABlockBytes, 1u);
            tc::mma_commit(bdone);
          }
          tc::mbar_wait(bdone, n_bd & 1u);
          ++n_bd;
#pragma unroll
          for (int j = 0; j < 32; j += 4) {
            *reinterpret_cast<float4*>(xb0 + tc::sw128_32b_chunk_off(erow, 64 * eh + j, tc::kABlockBytes)) = make_float4(v0[j], v0[j + 1], v0[j + 2], v0[j + 3]);
            *reinterpret_cast<float4*>(xb0 + tc::sw128_32b_chunk_off(erow, 64 * eh + 32 + j, tc::kABlockBytes)) = make_float4(v1[j], v1[j + 1], v1[j + 2], v1[j + 3]);
          }
          tc::fence_async_smem();
          tc::fence_before_sync();
          __syncthreads();
          if (leader && tc::elect_one()) {
            tc::fence_after_sync();
            tc::issue_gemm_tn128_pass(dW_t, Xb, tc::kABlockBytes, Phi, tc::kABlockBytes, 1u);
            tc::mma_commit(bdone);
          }
          tc::mbar_wait(bdone, n_bd & 1u);
          ++n_bd;
        }
        // ---- g: accumulator -> shared memory (the staging buffer is free) in the (ty, c4) ownership of the epilogue ----
        {
          float v[32];
          tc::fence_after_sync();
          tc::tmem_ld32(d_tmem + trow + (uint32_t)(256 + 32 * eh), v);
          tc::fence_before_sync();
#pragma unroll
          for (int j = 0; j < 32; j += 4) *reinterpret_cast<float4*>(Xb + erow * kLD + 32 * eh + j) = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
        }
        __syncthreads();
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const float4 g4 = *reinterpret_cast<const float4*>(Xb + (ty + 16 * i) * kLD + c4);
          acc[i][0] = g4.x; acc[i][1] = g4.y; acc[i][2] = g4.z; acc[i][3] = g4.w;
        }
      } else {
      const long long q3 = RAAE_PROFILE ? clock64() : 0ll;
      mma_tn8(Y, kLDW, 8 * (tid >> 3), At, kLD, 8 * (tid & 7), 0, kTM, accW);
      q4 = RAAE_PROFILE ? clock64() : 0ll;
      if (RAAE_PROFILE && !RAAE_PROF_FWD && tid == 0 && (mode == kLastRecon || mode == kLastSmooth)) { sm->prof[19] += q3 - q2; sm->prof[20] += q4 - q3; }
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
      prefetch_w_rows64(Wc, kLD, Wg, 0, N);
      cp_async_commit();
      for (int n0 = 0, ci = 0; n0 < N; n0 += kH, ++ci) {
        float* Wcur = (ci & 1) ? Wc2 : Wc;
        cp_async_wait<0>();
        __syncthreads();
        if (n0 + kH < N) { prefetch_w_rows64((ci & 1) ? Wc : Wc2, kLD, Wg, n0 + kH, N); cp_async_commit(); }
        mma_nn<kH>(Y + n0, kLDW, Wcur, kLD, acc, ty, tx);
      }
      }
      // branch-free body (rows >= nv of dv and of the activation tile are zero, their keep bits clear); guarded store
      const uint32_t kbits = mask_keep4_rows(in.mask, row0, ty, c4, nv);
      float* const g_dst = c.sc + c.p->sl.g[0] + (size_t)row0 * kH + c4;
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const int r = ty + 16 * i;
        const uint32_t kb = kbits >> (4 * i);
        const float4 a = *reinterpret_cast<const float4*>(At + r * kLD + c4);
        float4 gm;
        gm.x = (kb & 1u) ? acc[i][0] * in.mask.scale : 0.f;
        gm.y = (kb & 2u) ? acc[i][1] * in.mask.scale : 0.f;
        gm.z = (kb & 4u) ? acc[i][2] * in.mask.scale : 0.f;
        gm.w = (kb & 8u) ? acc[i][3] * in.mask.scale : 0.f;
        sg4[0] += gm.x; sg4[1] += gm.y; sg4[2] += gm.z; sg4[3] += gm.w;
        sgx4[0] = fmaf(acc[i][0], a.x, sgx4[0]); sgx4[1] = fmaf(acc[i][1], a.y, sgx4[1]);
        sgx4[2] = fmaf(acc[i][2], a.z, sgx4[2]); sgx4[3] = fmaf(acc[i][3], a.w, sgx4[3]);
        if (r < nv) *reinterpret_cast<float4*>(g_dst + (size_t)r * kH) = gm;
      }
      if (RAAE_PROFILE && !RAAE_PROF_FWD && tid == 0 && (mode == kLastRecon || mode == kLastSmooth)) sm->prof[21] += clock64() - q4;
    }
    __syncthreads();
  }
  if (mode == kLastRecon || mode == kLastEval) {
    double s = block_sum_d(loss_a, sm->redd);
    if (tid == 0) sm->loss_acc[kRecon] = s;
    cluster_allreduce_d(c, sm, &sm->loss_acc[kRecon], 1);
  }
  if (mode == kLastSmooth || mode == kLastEval) {
    double s = block_sum_d(loss_b, sm->redd);
    if (tid == 0) sm->loss_acc[kSmooth] = s;
    cluster_allreduce_d(c, sm, &sm->loss_acc[kSmooth], 1);
  }
  if (want_bwd) {
    sm->red[ty][c4 + 0] = sg4[0]; sm->red[ty][c4 + 1] = sg4[1]; sm->red[ty][c4 + 2] = sg4[2]; sm->red[ty][c4 + 3] = sg4[3];
    __syncthreads();
    if (tid < kH) { float s = 0.f; for (int i = 0; i < 16; ++i) s += sm->red[i][tid]; sg_dst[tid] = s; }
    __syncthreads();
    sm->red[ty][c4 + 0] = sgx4[0]; sm->red[ty][c4 + 1] = sgx4[1]; sm->red[ty][c4 + 2] = sgx4[2]; sm->red[ty][c4 + 3] = sgx4[3];
    __syncthreads();
    if (tid < kH) { float s = 0.f; for (int i = 0; i < 16; ++i) s += sm->red[i][tid]; sgx_dst[tid] = s; }
    float* gradW = Y;            // dense [N][64]
    float* gradb = Y + N * kH;   // [N], right behind dW: [W | b] is one run of the parameter vector
    if constexpr (TCB) {
      // TMEM columns 320 + 64 h ..: D[lane = output column 128 h + lane][k] (identity row map); a CTA without rows contributes zeros
      const bool any_tile = has_tiles(c, ntiles);
      tc::fence_after_sync();
      for (int h2 = 0; h2 < (nchN + 1) / 2; ++h2) {
        float v[32];
        tc::tmem_ld32(d_tmem + trow + (uint32_t)(320 + 64 * h2 + 32 * eh), v);
        const int cc = 128 * h2 + erow;
        if (cc < N) {
#pragma unroll
          for (int j = 0; j < 32; j += 4)
            *reinterpret_cast<float4*>(gradW + cc * kH + 32 * eh + j) =
                any_tile ? make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]) : make_float4(0.f, 0.f, 0.f, 0.f);
        }
      }
      tc::fence_before_sync();
    } else {
      const int m0 = 8 * (tid >> 3), n0 = 8 * (tid & 7);
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j)
          if (m0 + i < N) gradW[(m0 + i) * kH + n0 + j] = accW[i][j];
    }
    if (tid < N) gradb[tid] = dbp;
    __syncthreads();
    if (c.csize > 1) { cl::sync(); cluster_gather_sg(c, sm); }
    adam_apply(c, sm, o, kD, nl.w_off[l], N * kH + N, gradW);
    stage_sync(c);
  }
  __syncthreads();
}

__device__ __forceinline__ void dec_last(const Ctx& c, int mode, int inst, int o) {
  const bool bwd = mode == kLastRecon || mode == kLastSmooth || mode == kLastFromDv;
  if (bwd && (c.p->cfg.tensor_cores & 32)) dec_last_t<true>(c, mode, inst, o);
  else dec_last_t<false>(c, mode, inst, o);
}

// ------------------------------------------------------------------------------------------
// Style discriminator on z_real and on the encoder output, BCE-with-logits, backward through the
// gradient-reversal layer.  adversarial_loss functions.py:109-132, DiscriminatorFC model.py:631-663,
// GradientReversalLayer model.py:8-22.  dis_layers == 3:  Linear(ns,64) PReLU Drop Linear(64,64) PReLU Drop Linear(64,1)
// Output: sm->loss_acc[kAdv]; with backward: discriminator gradients (AdamW / export) and
// sc.dz = -alpha * dL/dstyles for the fake rows.
// ------------------------------------------------------------------------------------------
__device__ __noinline__ void dis_stage(const Ctx& c_ref, int backward, int o, const float* z_real_ptr, uint32_t key_zreal) {
  const Ctx c = c_ref;                 // register copy of the kernel context (SmemFixed::ctx, shared memory)
  RAAE_SMEM();
  StageTimer timer_(&sm->prof[kStDis]);
  const raae_net_layout& nl = NL(c, kS);
  const raae_net_layout& el = NL(c, kE);
  const int ns = nl.in_dim[0];
  const int tid = threadIdx.x, ty = tid >> 4, tx = tid & 15, c4 = tx * 4, ch = tid & 63, q = tid >> 6, lane = tid & 31,
            warp = tid >> 5;
  float* U1 = arena;
  float* H1 = U1 + kTile;
  float* U2 = H1 + kTile;
  float* H2 = U2 + kTile;
  float* W1s = H2 + kTile;            // [64][kLD]
  float* Zt = W1s + kWTile;           // [kTM][kZ]
  float* W0s = Zt + kTM * kZ;         // [64][9]
  float* w2s = W0s + kH * 9;          // [64]
  float* a0s = w2s + kH;              // [64] slopes of layer 0
  float* a1s = a0s + kH;              // [64] slopes of layer 1
  float* b0s = a1s + kH;
  float* b1s = b0s + kH;
  const float* P = netp(c, kS);
  const float* zE = c.sc + c.p->sl.zE;
  const int lE = el.n_linear - 1;
  const float noise = c.train ? (float)c.hp[RAAE_HP_DIS_NOISE] : 0.f;
  const float* eps_real = c.a->debug ? c.a->dbg.dis_eps_real : nullptr;
  const float* eps_fake = c.a->debug ? c.a->dbg.dis_eps_fake : nullptr;
  const uint32_t key_er = stream_key(c.seed, c.step_id, kStreamDisEpsReal);
  const uint32_t key_ef = stream_key(c.seed, c.step_id, kStreamDisEpsFake);
  const MaskSrc mk_real0 = make_mask(c, kS, 0, 0), mk_real1 = make_mask(c, kS, 0, 1);
  const MaskSrc mk_fake0 = make_mask(c, kS, 1, 0), mk_fake1 = make_mask(c, kS, 1, 1);
  __syncthreads();
  load_w_rows(W1s, kLD, P + nl.w_off[1], kH, 0, kH);
  for (int i = tid; i < kH * 9; i += kThreads) {
    int n = i / 9, k = i - n * 9;
    W0s[i] = k < ns ? P[nl.w_off[0] + n * ns + k] : 0.f;
  }
  if (tid < kH) {
    w2s[tid] = P[nl.w_off[2] + tid];
    a0s[tid] = P[nl.a_off[0] + tid];
    a1s[tid] = P[nl.a_off[1] + tid];
    b0s[tid] = P[nl.b_off[0] + tid];
    b1s[tid] = P[nl.b_off[1] + tid];
  }
  __syncthreads();
  const float b2 = P[nl.b_off[2]];
  const float alpha = sm->alpha;
  float accW1[8][8];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j) accW1[i][j] = 0.f;
  float accW0[kZ] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};   // (ch, q) partials of dW0[ch][:]
  float dW2p[4] = {0.f, 0.f, 0.f, 0.f}, da1p[4] = {0.f, 0.f, 0.f, 0.f}, db1p[4] = {0.f, 0.f, 0.f, 0.f};   // (ty, c4) partials
  float da0p[4] = {0.f, 0.f, 0.f, 0.f}, db0p[4] = {0.f, 0.f, 0.f, 0.f};   // (ty, c4) partials
  float db2p = 0.f;                                   // threads 0..127: one row of every tile each
  double lossp = 0.0;
  const int tiles_real = (c.Breal + kTM - 1) / kTM, tiles_fake = (c.B + kTM - 1) / kTM;
  for (int t = 0; t < tiles_real + tiles_fake; ++t) {
    const bool fake = t >= tiles_real;
    // cluster per trial: tile tl of either half belongs to rank tl % csize (the fake rows follow the row ownership of zE / dz)
    if (((fake ? t - tiles_real : t) % c.csize) != c.crank) continue;
    const int nrows = fake ? c.B : c.Breal;
    const int row0 = (fake ? t - tiles_real : t) * kTM, nv = min(kTM, nrows - row0);
    RAAE_PROBE_INIT();
    const MaskSrc mk0 = fake ? mk_fake0 : mk_real0;        // register copies
    const MaskSrc mk1 = fake ? mk_fake1 : mk_real1;
    const float label = fake ? 0.f : 1.f;
    const double inv_rows = 1.0 / (double)nrows;
    // dropout keep bits of this thread's (row group, 4 channels) blocks of both hidden layers: drawn once per tile
    // (straight-line mixer chains), used by the forward and by the backward passes
    uint32_t kbits0 = mask_keep4_rows(mk0, row0, ty, c4, nv);
    uint32_t kbits1 = mask_keep4_rows(mk1, row0, ty, c4, nv);
    asm volatile("" : "+r"(kbits0), "+r"(kbits1));     // opaque: the bit tests are re-derived where used, not kept in registers
    // ---- input rows (+ input noise, model.py:659-660) ----
    for (int i = tid; i < kTM * kZ; i += kThreads) {
      int r = i >> 3, k = i & 7;
      float v = 0.f;
      if (r < nv && k < ns) {
        int row = row0 + r;
        if (fake) v = (zE[(size_t)row * kZ + k] - sm->mean[kE][lE][k]) * sm->inv[kE][lE][k];
        else v = z_real_ptr ? z_real_ptr[(size_t)row * ns + k] : normal_at(key_zreal, (uint32_t)(row * kZ + k));
        if (noise != 0.f) {
          const float* ep = fake ? eps_fake : eps_real;
          float e = ep ? ep[(size_t)row * ns + k] : normal_at(fake ? key_ef : key_er, (uint32_t)(row * kZ + k));
          v = v + noise * e;
        }
      }
      Zt[i] = v;
    }
    __syncthreads();
    RAAE_PROBE(23);     // input rows + noise
    // ---- layer 0: thread (ty, c4) owns channels c4..c4+3 of rows ty, ty + 16, ...: its four weight rows stay in registers,
    //      the input row is two broadcast loads, one mask draw covers the four channels, float4 stores ----
    {
      float w0r[4][kZ];
#pragma unroll
      for (int j = 0; j < 4; ++j)
#pragma unroll
        for (int k = 0; k < kZ; ++k) w0r[j][k] = W0s[(c4 + j) * 9 + k];
      const float4 b0v = *reinterpret_cast<const float4*>(b0s + c4), a0v = *reinterpret_cast<const float4*>(a0s + c4);
      const float bb[4] = {b0v.x, b0v.y, b0v.z, b0v.w}, aa[4] = {a0v.x, a0v.y, a0v.z, a0v.w};
#pragma unroll 2
      for (int i = 0; i < kTM / 16; ++i) {
        const int r = ty + 16 * i;
        const float4 z0 = *reinterpret_cast<const float4*>(Zt + r * kZ);
        const float4 z1 = *reinterpret_cast<const float4*>(Zt + r * kZ + 4);
        const uint32_t kb = kbits0 >> (4 * i);
        float u[4], h[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          float a = bb[j];
          a = fmaf(z0.x, w0r[j][0], a); a = fmaf(z0.y, w0r[j][1], a); a = fmaf(z0.z, w0r[j][2], a); a = fmaf(z0.w, w0r[j][3], a);
          a = fmaf(z1.x, w0r[j][4], a); a = fmaf(z1.y, w0r[j][5], a); a = fmaf(z1.z, w0r[j][6], a); a = fmaf(z1.w, w0r[j][7], a);
          u[j] = r < nv ? a : 0.f;
          h[j] = (kb >> j & 1u) ? prelu_f(a, aa[j]) * mk0.scale : 0.f;
        }
        *reinterpret_cast<float4*>(U1 + r * kLD + c4) = make_float4(u[0], u[1], u[2], u[3]);
        *reinterpret_cast<float4*>(H1 + r * kLD + c4) = make_float4(h[0], h[1], h[2], h[3]);
      }
    }
    __syncthreads();
    RAAE_PROBE(24);     // layer 0
    // ---- layer 1 ----
    {
      float acc[8][4];
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
      mma_nt<kH>(H1, kLD, W1s, kLD, acc, ty, tx);
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) U2[(ty + 16 * i) * kLD + tx + 16 * j] = acc[i][j] + b1s[tx + 16 * j];
    }
    __syncthreads();
    {
      const float4 a1v = *reinterpret_cast<const float4*>(a1s + c4);
#pragma unroll 4
      for (int i = 0; i < kTM / 16; ++i) {
        const int r = ty + 16 * i;
        const uint32_t kb = kbits1 >> (4 * i);
        const float4 u = *reinterpret_cast<const float4*>(U2 + r * kLD + c4);
        float4 h;
        h.x = (kb & 1u) ? prelu_f(u.x, a1v.x) * mk1.scale : 0.f;
        h.y = (kb & 2u) ? prelu_f(u.y, a1v.y) * mk1.scale : 0.f;
        h.z = (kb & 4u) ? prelu_f(u.z, a1v.z) * mk1.scale : 0.f;
        h.w = (kb & 8u) ? prelu_f(u.w, a1v.w) * mk1.scale : 0.f;
        *reinterpret_cast<float4*>(H2 + r * kLD + c4) = h;
      }
    }
    __syncthreads();
    RAAE_PROBE(25);     // layer 1 GEMM + H2
    // ---- logits, BCE-with-logits (mean over the rows of this half), dL/dlogit ----
    // the dot products by warps (shuffle reduction), then the scalar loss arithmetic of the 128 rows by 128 threads at once
    // (it was a serial chain of 16 rows on lane 0 of every warp)
    {
      const float w2a = w2s[lane], w2b = w2s[lane + 32];
#pragma unroll 4
      for (int r = warp; r < kTM; r += kThreads / 32) {
        float s = H2[r * kLD + lane] * w2a + H2[r * kLD + lane + 32] * w2b;
#pragma unroll
        for (int of = 16; of > 0; of >>= 1) s += __shfl_xor_sync(0xffffffffu, s, of);
        if (lane == 0) sm->logit[r] = s + b2;
      }
    }
    __syncthreads();
    if (tid < kTM) {
      float dl = 0.f;
      if (tid < nv) {
        const float x = sm->logit[tid];
        const float li = fmaxf(x, 0.f) - x * label + log1pf(expf(-fabsf(x)));
        lossp += (double)li * inv_rows;                // inv_rows: one float64 division per tile, not per row
        dl = (sigmoid_f(x) - label) / (float)nrows;
      }
      sm->dlogit[tid] = dl;
      db2p += dl;                                      // per-thread partial sum of dL/dlogit (reduced at the end)
    }
    __syncthreads();
    RAAE_PROBE(26);     // logits + BCE
    if (backward) {
      // ---- layer 2 and the PReLU/dropout of layer 1 ----
      {
        const float4 a1v = *reinterpret_cast<const float4*>(a1s + c4), w2v = *reinterpret_cast<const float4*>(w2s + c4);
#pragma unroll 2
        for (int i = 0; i < kTM / 16; ++i) {
          const int r = ty + 16 * i;
          float4 du = make_float4(0.f, 0.f, 0.f, 0.f);
          {                                    // branch-free: rows >= nv carry dlogit = 0, h = u = 0 and clear keep bits
            const float dl = sm->dlogit[r];
            const float4 h = *reinterpret_cast<const float4*>(H2 + r * kLD + c4);
            const float4 u = *reinterpret_cast<const float4*>(U2 + r * kLD + c4);
            uint32_t kb = kbits1 >> (4 * i);
            asm volatile("" : "+r"(kb));
#define RAAE_DU2(comp, idx, bit)                                                         \
            {                                                                            \
              dW2p[idx] = fmaf(dl, h.comp, dW2p[idx]);                                   \
              const float g = (kb & bit) ? dl * w2v.comp * mk1.scale : 0.f;              \
              const bool pos = u.comp > 0.f;                                             \
              du.comp = pos ? g : a1v.comp * g;                                          \
              da1p[idx] += pos ? 0.f : u.comp * g;                                       \
              db1p[idx] += du.comp;                                                      \
            }
            RAAE_DU2(x, 0, 1u) RAAE_DU2(y, 1, 2u) RAAE_DU2(z, 2, 4u) RAAE_DU2(w, 3, 8u)
#undef RAAE_DU2
          }
          *reinterpret_cast<float4*>(U2 + r * kLD + c4) = du;
        }
      }
      __syncthreads();
      RAAE_PROBE(16);   // du2 pass
      // ---- layer 1: dW1 += du2^T h1, dh1 = du2 @ W1 ----
      {
        const int qq = tid >> 6, tt = tid & 63;
        mma_tn8(U2, kLD, 8 * (tt >> 3), H1, kLD, 8 * (tt & 7), 32 * qq, 32 * qq + 32, accW1);
        float acc[8][4];
#pragma unroll
        for (int i = 0; i < 8; ++i)
#pragma unroll
          for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
        mma_nn<kH>(U2, kLD, W1s, kLD, acc, ty, tx);
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          int r = ty + 16 * i;
          float4 du = make_float4(0.f, 0.f, 0.f, 0.f);
          {                                    // branch-free: rows >= nv have clear keep bits and u = 0
            uint32_t kb = kbits0 >> (4 * i);
            asm volatile("" : "+r"(kb));
            float4 u = *reinterpret_cast<const float4*>(U1 + r * kLD + c4);
#define RAAE_DIS(comp, idx, bit)                                             \
            {                                                                \
              float g = (kb & bit) ? acc[i][idx] * mk0.scale : 0.f;          \
              bool pos = u.comp > 0.f;                                       \
              du.comp = pos ? g : a0s[c4 + idx] * g;                         \
              da0p[idx] += pos ? 0.f : u.comp * g;                           \
              db0p[idx] += du.comp;                                          \
            }
            RAAE_DIS(x, 0, 1u) RAAE_DIS(y, 1, 2u) RAAE_DIS(z, 2, 4u) RAAE_DIS(w, 3, 8u)
#undef RAAE_DIS
          }
          *reinterpret_cast<float4*>(U1 + r * kLD + c4) = du;
        }
      }
      __syncthreads();
      RAAE_PROBE(17);   // dW1 + dh1 GEMMs + epilogue
      // ---- layer 0: dW0 += du1^T z;  fake rows: dz = -alpha * du1 @ W0 (gradient reversal) ----
      // dW0[ch][k] over rows q, q + 4, ...: one du1 load and one broadcast input row per 8 FMAs
#pragma unroll 4
      for (int i = 0; i < kTM / 4; ++i) {
        const int r = q + 4 * i;
        const float d = U1[r * kLD + ch];
        const float4 z0 = *reinterpret_cast<const float4*>(Zt + r * kZ);
        const float4 z1 = *reinterpret_cast<const float4*>(Zt + r * kZ + 4);
        accW0[0] = fmaf(d, z0.x, accW0[0]); accW0[1] = fmaf(d, z0.y, accW0[1]); accW0[2] = fmaf(d, z0.z, accW0[2]); accW0[3] = fmaf(d, z0.w, accW0[3]);
        accW0[4] = fmaf(d, z1.x, accW0[4]); accW0[5] = fmaf(d, z1.y, accW0[5]); accW0[6] = fmaf(d, z1.z, accW0[6]); accW0[7] = fmaf(d, z1.w, accW0[7]);
      }
      if (fake) {
        // dz[r][k] = -alpha sum_n du1[r][n] W0[n][k]: two threads per row (32 channels each), partial sums combined by shuffle
        float* dz = c.sc + c.p->sl.dz;
        const int r = tid >> 1, half = tid & 1;
        float s8[kZ];
#pragma unroll
        for (int k = 0; k < kZ; ++k) s8[k] = 0.f;
#pragma unroll 2
        for (int n4 = 0; n4 < 32; n4 += 4) {
          const float4 d4 = *reinterpret_cast<const float4*>(U1 + r * kLD + 32 * half + n4);
          const float dv[4] = {d4.x, d4.y, d4.z, d4.w};
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const float* wrow = W0s + (32 * half + n4 + e) * 9;
#pragma unroll
            for (int k = 0; k < kZ; ++k) s8[k] = fmaf(dv[e], wrow[k], s8[k]);
          }
        }
#pragma unroll
        for (int k = 0; k < kZ; ++k) s8[k] += __shfl_xor_sync(0xffffffffu, s8[k], 1);
        if (half == 0 && r < nv) {
          *reinterpret_cast<float4*>(dz + (size_t)(row0 + r) * kZ) = make_float4(-alpha * s8[0], -alpha * s8[1], -alpha * s8[2], -alpha * s8[3]);
          *reinterpret_cast<float4*>(dz + (size_t)(row0 + r) * kZ + 4) = make_float4(-alpha * s8[4], -alpha * s8[5], -alpha * s8[6], -alpha * s8[7]);
        }
      }
      __syncthreads();
      RAAE_PROBE(22);   // dW0 + dz
    }
  }
  {
    double s = block_sum_d(lossp, sm->redd);
    if (tid == 0) sm->loss_acc[kAdv] = s;
    cluster_allreduce_d(c, sm, &sm->loss_acc[kAdv], 1);
  }
  if (backward) {
    // the whole discriminator gradient in parameter order (one AdamW pass): W0 [64 ns] | b0 | a0 | W1 [64][64] | b1 | a1 | W2 | b2
    float* gW0 = arena;
    float* gb0 = gW0 + kH * ns;
    float* ga0 = gb0 + kH;
    float* gW1 = ga0 + kH;
    float* gb1 = gW1 + kH * kH;
    float* ga1 = gb1 + kH;
    float* gW2 = ga1 + kH;
    float* gb2 = gW2 + kH;
    float* gsm = arena + 8192;         // scratch for the dW0 row groups, behind the gradient vector
    __syncthreads();
    {
      const int qq = tid >> 6, tt = tid & 63, m0 = 8 * (tt >> 3), n0 = 8 * (tt & 7);
      for (int pass = 0; pass < 4; ++pass) {
        if (qq == pass) {
#pragma unroll
          for (int i = 0; i < 8; ++i)
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              float* dst = gW1 + (m0 + i) * kH + n0 + j;
              *dst = pass == 0 ? accW1[i][j] : *dst + accW1[i][j];
            }
        }
        __syncthreads();
      }
    }
    {
      // reduce the four row groups of dW0 through the (free) tile area behind the gradient vectors
      float* part = gsm;                     // [4][64][8]
#pragma unroll
      for (int k = 0; k < kZ; ++k) part[(q * kH + ch) * kZ + k] = accW0[k];
    }
    __syncthreads();
#pragma unroll
    for (int e = 0; e < 2; ++e) {
      const float* part = gsm;
      int oo = tid + kThreads * e, n = oo >> 3, k = oo & 7;
      if (k < ns) gW0[n * ns + k] = part[(0 * kH + n) * kZ + k] + part[(1 * kH + n) * kZ + k] + part[(2 * kH + n) * kZ + k] + part[(3 * kH + n) * kZ + k];
    }
    {
      const float tot = block_sum(db2p, sm->redw);
      if (tid == 0) gb2[0] = tot;
    }
    __syncthreads();
    // (ty, c4) partials of dW2 / da1 / db1 -> column sums over the 16 row groups
    {
      float* const dst3[3] = {gW2, ga1, gb1};
      const float* const src3[3] = {dW2p, da1p, db1p};
#pragma unroll
      for (int v = 0; v < 3; ++v) {
        sm->red[ty][c4 + 0] = src3[v][0]; sm->red[ty][c4 + 1] = src3[v][1]; sm->red[ty][c4 + 2] = src3[v][2]; sm->red[ty][c4 + 3] = src3[v][3];
        __syncthreads();
        if (tid < kH) { float sacc = 0.f; for (int i = 0; i < 16; ++i) sacc += sm->red[i][tid]; dst3[v][tid] = sacc; }
        __syncthreads();
      }
    }
    sm->red[ty][c4 + 0] = da0p[0]; sm->red[ty][c4 + 1] = da0p[1]; sm->red[ty][c4 + 2] = da0p[2]; sm->red[ty][c4 + 3] = da0p[3];
    __syncthreads();
    if (tid < kH) { float s = 0.f; for (int i = 0; i < 16; ++i) s += sm->red[i][tid]; ga0[tid] = s; }
    __syncthreads();
    sm->red[ty][c4 + 0] = db0p[0]; sm->red[ty][c4 + 1] = db0p[1]; sm->red[ty][c4 + 2] = db0p[2]; sm->red[ty][c4 + 3] = db0p[3];
    __syncthreads();
    if (tid < kH) { float s = 0.f; for (int i = 0; i < 16; ++i) s += sm->red[i][tid]; gb0[tid] = s; }
    __syncthreads();
    if (c.csize > 1) cl::sync();
    adam_apply(c, sm, o, kS, nl.w_off[0], nl.n_params, gW0);
    stage_sync(c);
  }
  __syncthreads();
}

// ------------------------------------------------------------------------------------------
// Kendall rank constraint, kendall_constraint functions.py:37-79, over all ordered pairs of rows.
//   t = sign(d_i - d_j), p = (s_i - s_j) t;  per descriptor k: n_same = #(p > 0), n_opp = #(p < 0) (>= 1),
//   w_k = n_opp / max(n_same, n_opp) rescales the p > 0 entries (kendall_activation);
//   loss = -sum(p) / ((B^2 - B) n_aux);   dloss/ds_ik = -2/norm * sum_j t_ijk (w_k if p_ijk > 0 else 1)
// The pair tensor is never materialised: one pass accumulates, per row, A = sum_j [p>0] t and
// Bn = sum_j [p<=0] t, and per descriptor the counts and partial sums.
// Output: sm->loss_acc[kCorr]; with want_grad: sc.dz (zero beyond n_aux).
// ------------------------------------------------------------------------------------------
constexpr int kKendallChunk = 2048;

// pair loop of one row i against the staged rows j, K descriptors (compile-time so that the loop is branch-free).
// Per (pair, k): t = sign(d_i - d_j), p = (s_i - s_j) t.  With dd = d_i - d_j, ds = s_i - s_j: sign(p) = sign(ds dd),
// |p| = |ds| wherever p != 0, and t = copysign(1, dd) wherever dd != 0 - so the body is 3 float ops, 3 compares and 6
// PREDICATED accumulations (13 instructions instead of 17: ptxas turns the C++ selects into FSEL + FADD pairs, the
// inline PTX keeps them as one predicated instruction each; the stage is issue-bound).  Accumulated: counts of p > 0 /
// p < 0, the sums of |p| over either set, the row's A = sum [p > 0] t and T = sum t.
// (ds dd cannot underflow to zero for distinct BatchNorm-ed latents / descriptors: that needs |ds|, |dd| < 1e-19.)
__device__ __forceinline__ void kendall_pair(float dd, float ds, float& fp, float& fn, float& A, float& T, int& cs, int& co) {
  asm("{\n\t.reg .pred pp, pn, pz;\n\t.reg .f32 q, a, sg;\n\t"
      "mul.rn.f32 q, %6, %7;\n\t"
      "setp.gt.f32 pp, q, 0f00000000;\n\t"
      "setp.lt.f32 pn, q, 0f00000000;\n\t"
      "setp.neu.f32 pz, %6, 0f00000000;\n\t"
      "abs.f32 a, %7;\n\t"
      "copysign.f32 sg, %6, 0f3F800000;\n\t"
      "@pp add.f32 %0, %0, a;\n\t"
      "@pn add.f32 %1, %1, a;\n\t"
      "@pp add.f32 %2, %2, sg;\n\t"
      "@pz add.f32 %3, %3, sg;\n\t"
      "@pp add.s32 %4, %4, 1;\n\t"
      "@pn add.s32 %5, %5, 1;\n\t}"
      : "+f"(fp), "+f"(fn), "+f"(A), "+f"(T), "+r"(cs), "+r"(co)
      : "f"(dd), "f"(ds));
}
template <int K>
__device__ __forceinline__ void kendall_row(const float* __restrict__ Ss, const float* __restrict__ Ds, int nj, const float (&si)[kZ],
                                            const float (&di)[kZ], float (&A)[kZ], float (&T)[kZ], float (&fp)[kZ], float (&fn)[kZ],
                                            int (&cs)[kZ], int (&co)[kZ]) {
#pragma unroll 2
  for (int j = 0; j < nj; ++j) {
    const float4 s0 = *reinterpret_cast<const float4*>(Ss + j * kZ);
    const float4 d0 = *reinterpret_cast<const float4*>(Ds + j * kZ);
    float4 s1 = make_float4(0.f, 0.f, 0.f, 0.f), d1 = make_float4(0.f, 0.f, 0.f, 0.f);
    if (K > 4) {
      s1 = *reinterpret_cast<const float4*>(Ss + j * kZ + 4);
      d1 = *reinterpret_cast<const float4*>(Ds + j * kZ + 4);
    }
    const float sj[kZ] = {s0.x, s0.y, s0.z, s0.w, s1.x, s1.y, s1.z, s1.w};
    const float dj[kZ] = {d0.x, d0.y, d0.z, d0.w, d1.x, d1.y, d1.z, d1.w};
#pragma unroll
    for (int k = 0; k < K; ++k) kendall_pair(di[k] - dj[k], si[k] - sj[k], fp[k], fn[k], A[k], T[k], cs[k], co[k]);
  }
}

__device__ __noinline__ void kendall_stage(const Ctx& c_ref, const float* __restrict__ aux, int want_grad) {
  const Ctx c = c_ref;                 // register copy of the kernel context (SmemFixed::ctx, shared memory)
  RAAE_SMEM();
  StageTimer timer_(&sm->prof[kStKendall]);
  const raae_net_layout& el = NL(c, kE);
  const int lE = el.n_linear - 1;
  const int K = c.p->cfg.n_aux, B = c.B, tid = threadIdx.x;
  float* Ss = arena;                          // [chunk][kZ] styles
  float* Ds = Ss + kKendallChunk * kZ;        // [chunk][kZ] descriptors
  const float* zE = c.sc + c.p->sl.zE;
  float* kacc = c.sc + c.p->sl.g[0];          // [rows][32] per-row A | Bn per partner half (spill for multi-chunk batches)
  float* dz = c.sc + c.p->sl.dz;
  int cs[kZ], co[kZ];
  double sp[kZ], sn[kZ];
#pragma unroll
  for (int k = 0; k < kZ; ++k) { cs[k] = 0; co[k] = 0; sp[k] = 0.0; sn[k] = 0.0; }
  const int nchunks = (B + kKendallChunk - 1) / kKendallChunk;
  // cluster per trial: a CTA owns few rows (128 at 8 CTAs and 1024 rows), so two threads share a row and split its partners
  constexpr int JS = RAAE_CLUSTER ? 2 : 1;
  constexpr int rpp = kThreads / JS;                              // rows per pass
  // cluster per trial: a CTA pairs ITS rows i (slot -> row) with all rows j; the other CTAs' latents / descriptors must be
  // visible (in eval mode the producing stage has no barrier of its own)
  const int crank = c.crank, csize = c.csize;
  const int nslots = csize == 1 ? B : cl::own_tiles(B, crank, csize) * kTM;
  if (csize > 1) cl::sync();
  for (int ck = 0; ck < nchunks; ++ck) {
    const int j0 = ck * kKendallChunk, nj = min(kKendallChunk, B - j0);
    __syncthreads();
    for (int i = tid; i < nj * kZ; i += kThreads) {
      int r = i >> 3, k = i & 7;
      float s = 0.f, d = 0.f;
      if (k < K) {
        s = (zE[(size_t)(j0 + r) * kZ + k] - sm->mean[kE][lE][k]) * sm->inv[kE][lE][k];
        d = aux[(size_t)(j0 + r) * kZ + k];
      }
      Ss[i] = s; Ds[i] = d;
    }
    __syncthreads();
    const int lrow = tid % rpp, half = tid / rpp;
    for (int ib = 0; ib < nslots; ib += rpp) {
      // without the gradient only the totals are needed and p_ij = p_ji: each unordered pair is visited once (j > i) and
      // the totals are doubled at the end; alternate passes run the rows in reverse so that every thread gets long and
      // short rows (row i has B - 1 - i partners)
      const int sl = (want_grad || !((ib / rpp) & 1)) ? ib + lrow : ib + rpp - 1 - lrow;
      if (sl >= nslots) continue;
      const int i = csize == 1 ? sl : cl::slot_row(sl, crank, csize);
      if (i >= B) continue;
      int jb = want_grad ? 0 : max(0, i + 1 - j0);               // first staged row of this chunk that row i pairs with
      if (jb >= nj) continue;
      int je = nj;
      if (JS > 1) { const int len = nj - jb; je = jb + (len * (half + 1)) / JS; jb = jb + (len * half) / JS; }
      float si[kZ], di[kZ], A[kZ], T[kZ], fp[kZ], fn[kZ];
#pragma unroll
      for (int k = 0; k < kZ; ++k) {
        si[k] = k < K ? (zE[(size_t)i * kZ + k] - sm->mean[kE][lE][k]) * sm->inv[kE][lE][k] : 0.f;
        di[k] = k < K ? aux[(size_t)i * kZ + k] : 0.f;
        A[k] = 0.f; T[k] = 0.f; fp[k] = 0.f; fn[k] = 0.f;
      }
      switch (K) {
        case 1: kendall_row<1>(Ss + jb * kZ, Ds + jb * kZ, je - jb, si, di, A, T, fp, fn, cs, co); break;
        case 2: kendall_row<2>(Ss + jb * kZ, Ds + jb * kZ, je - jb, si, di, A, T, fp, fn, cs, co); break;
        case 3: kendall_row<3>(Ss + jb * kZ, Ds + jb * kZ, je - jb, si, di, A, T, fp, fn, cs, co); break;
        case 4: kendall_row<4>(Ss + jb * kZ, Ds + jb * kZ, je - jb, si, di, A, T, fp, fn, cs, co); break;
        case 5: kendall_row<5>(Ss + jb * kZ, Ds + jb * kZ, je - jb, si, di, A, T, fp, fn, cs, co); break;
        case 6: kendall_row<6>(Ss + jb * kZ, Ds + jb * kZ, je - jb, si, di, A, T, fp, fn, cs, co); break;
        case 7: kendall_row<7>(Ss + jb * kZ, Ds + jb * kZ, je - jb, si, di, A, T, fp, fn, cs, co); break;
        default: kendall_row<8>(Ss + jb * kZ, Ds + jb * kZ, je - jb, si, di, A, T, fp, fn, cs, co); break;
      }
#pragma unroll
      for (int k = 0; k < kZ; ++k) { sp[k] += (double)fp[k]; sn[k] -= (double)fn[k]; }
      if (want_grad) {
        // A = sum [p > 0] t, Bn = sum [p <= 0] t = T - A, accumulated over the chunks of a large batch
#pragma unroll
        for (int k = 0; k < kZ; ++k) {            // row i: [half 0: A | Bn][half 1: A | Bn]
          float a = A[k], bn = T[k] - A[k];
          float* ka = kacc + (size_t)i * 32 + 16 * half;
          if (ck > 0) { a += ka[k]; bn += ka[8 + k]; }
          ka[k] = a; ka[8 + k] = bn;
        }
      }
    }
  }
  // block totals per descriptor (cluster per trial: summed over the CTAs)
  for (int k = 0; k < kZ; ++k) {
    if (k >= K) break;
    const double b0 = block_sum_d(sp[k], sm->redd), b1 = block_sum_d(sn[k], sm->redd);
    const double b2 = block_sum_d((double)cs[k], sm->redd), b3 = block_sum_d((double)co[k], sm->redd);
    if (tid == 0) { sm->ktot[4 * k] = b0; sm->ktot[4 * k + 1] = b1; sm->ktot[4 * k + 2] = b2; sm->ktot[4 * k + 3] = b3; }
  }
  cluster_allreduce_d(c, sm, sm->ktot, 4 * K);
  __syncthreads();
  double loss = 0.0;
  for (int k = 0; k < kZ; ++k) {
    if (k >= K) break;
    const double sym = want_grad ? 1.0 : 2.0;             // ordered pairs = 2 x unordered pairs
    double tsp = sym * sm->ktot[4 * k];
    double tsn = sym * sm->ktot[4 * k + 1];
    double tcs = sym * sm->ktot[4 * k + 2];
    double tco = sym * sm->ktot[4 * k + 3];
    double w = 1.0;
    if (c.p->cfg.kendall_activation) {
      double n_same = tcs > 1.0 ? tcs : 1.0, n_opp = tco > 1.0 ? tco : 1.0;
      w = n_opp / (n_same > n_opp ? n_same : n_opp);
    }
    loss += (double)(float)w * tsp + tsn;
    if (tid == 0) sm->kw[k] = (float)w;
  }
  const double norm = ((double)B * (double)B - (double)B) * (double)K;
  if (tid == 0) sm->loss_acc[kCorr] = -loss / norm;
  __syncthreads();
  if (want_grad) {
    const float scale = (float)(-2.0 / norm);
    for (int e = tid; e < nslots * kZ; e += kThreads) {
      const int sl = e >> 3, k = e & 7, r = csize == 1 ? sl : cl::slot_row(sl, crank, csize);
      if (r >= B) continue;
      float g = 0.f;
      if (k < K) {
        float a = kacc[(size_t)r * 32 + k], bn = kacc[(size_t)r * 32 + 8 + k];
        if (JS > 1) { a += kacc[(size_t)r * 32 + 16 + k]; bn += kacc[(size_t)r * 32 + 24 + k]; }
        g = scale * (sm->kw[k] * a + bn);
      }
      dz[(size_t)r * kZ + k] = g;
    }
  }
  __syncthreads();
}

// MSE between the re-encoded latent and z_sample (mutual_info_loss functions.py:174-192)
__device__ __noinline__ void mi_mse_stage(const Ctx& c_ref, int want_grad) {
  const Ctx c = c_ref;                 // register copy of the kernel context (SmemFixed::ctx, shared memory)
  RAAE_SMEM();
  StageTimer timer_(&sm->prof[kStMiMse]);
  const raae_net_layout& el = NL(c, kE);
  const int lE = el.n_linear - 1, ns = c.p->cfg.nstyle, tid = threadIdx.x;
  const float* zE = c.sc + c.p->sl.zE;
  const float* zs = c.sc + c.p->sl.zs;
  float* dz = c.sc + c.p->sl.dz;
  const float cnt = (float)c.B * (float)ns;
  double lp = 0.0;
  const int crank = c.crank, csize = c.csize;
  const int nslots = csize == 1 ? c.B : cl::own_tiles(c.B, crank, csize) * kTM;      // cluster per trial: this CTA's rows
  __syncthreads();
  for (int e = tid; e < nslots * kZ; e += kThreads) {
    const int k = e & 7, r = csize == 1 ? (e >> 3) : cl::slot_row(e >> 3, crank, csize);
    if (r >= c.B) continue;
    const int i = r * kZ + k;
    float d = 0.f;
    if (k < ns) d = (zE[i] - sm->mean[kE][lE][k]) * sm->inv[kE][lE][k] - zs[i];
    lp += (double)(d * d);
    if (want_grad) dz[i] = 2.f * d / cnt;
  }
  double s = block_sum_d(lp, sm->redd);
  if (tid == 0) sm->loss_acc[kMI] = s;
  cluster_allreduce_d(c, sm, &sm->loss_acc[kMI], 1);
  __syncthreads();
  if (tid == 0) sm->loss_acc[kMI] = sm->loss_acc[kMI] / (double)cnt;
  __syncthreads();
}

}  // namespace raae
