// The fused AAE train step: stage functions executed by ONE CTA per trial (256 threads).
//
// Data flow: per-row activations live in the trial's scratch block in HBM/L2 as
// row-major [rows][64] panels; a stage streams 128-row tiles of them through shared memory, runs the
// dense contractions as register-tiled FP32 FMA on the smem tiles, and fuses bias / PReLU / BatchNorm
// statistics / dropout / loss gradients / AdamW around them.  BatchNorm couples the whole batch, so a
// stage == one layer (forward) or one layer's backward; stages are separated by __syncthreads().
#pragma once
#include "aae_common.cuh"
#include "aae_tc.cuh"

namespace raae {

// Per-trial context of the running kernel.  ONE copy per CTA in shared memory (SmemFixed::ctx): thread 0 writes it, a
// barrier publishes it, and every stage function starts with a register copy of it (`const Ctx c = c_ref`).  It used to
// live in the kernel's local-memory frame, where the copy at every stage entry missed the (28 KB, streamed-through) L1.
struct Ctx {
  const KParams* p;
  const RunArgs* a;
  float* st;          // trial state block
  float* sc;          // trial scratch block
  const double* hp;   // trial hyper-parameters
  int B;              // rows of the current batch
  int Breal;          // rows of z_real (cfg batch_size; trainer.py:121)
  const float* x;     // input spectra rows of the current batch / validation set
  int xld;
  uint32_t seed, step_id;
  int train;          // BN batch statistics + dropout + noise
  int apply;          // apply optimizer updates
  int epoch;
  int trial;
  float drop_scale[2];     // [0] encoder / decoder, [1] discriminator: 1 / (1 - p), read once per kernel from the hp row
  uint32_t drop_thresh[2]; // round(p * 65536); 0 = no dropout
#if RAAE_CLUSTER
  int crank, csize;        // rank of this CTA in the trial's thread-block cluster / cluster size (ctas_per_trial)
#else
  // one CTA per trial (this translation unit): every cluster branch folds away at compile time
  static constexpr int crank = 0, csize = 1;
#endif
};

struct SmemFixed {
  float mean[2][RAAE_MAX_LAYERS][kH];   // per net (E, D) / layer: BN mean used by the current forward
  float inv[2][RAAE_MAX_LAYERS][kH];    // 1/sqrt(var + eps)
  float bias[kMaxDim];
  float slope[kH];
  float shift[kH];
  float cg[kH], cgx[kH];                // mean(g), mean(g * xhat) of the layer being back-propagated
  float sg[kH], sgx[kH];                // the same sums being produced for the next lower layer
  float red[16][kH];
  float redw[32];
  double redd[8];
  float ad[8];                          // AdamW scalars of the running phase
  float zs[4][kZ];                      // latent-channel scalars: [0] mean(dz) [1] mean(dz*zhat) ...
  float logit[kTM], dlogit[kTM];
  double loss_acc[8];
  int kcount[2][kZ];
  float kw[kZ];
  float alpha;
  long long prof[32];                   // per-stage-type cycle counters (thread 0), see StageId; 16.. = sub-stage probes
  unsigned long long mbar;              // mbarrier the tcgen05 commits arrive on
  unsigned long long pipe_bar[16];       // mbarriers of the bulk-copy / MMA pipelines (re-initialised by every stage that uses them)
  unsigned long long mbar2;             // second mbarrier (double-buffered accumulators of the pipelined forward)
  uint32_t tc_phase2;                   // parity of the next completion of mbar2
  uint32_t tmem_base;                   // TMEM address returned by tcgen05.alloc
  uint32_t tc_phase;                    // parity of the next mbarrier completion
  // thread-block cluster per trial (ctas_per_trial > 1): partial vectors the other CTAs of the trial read through
  // distributed shared memory.  Double-buffered, so an all-reduce costs ONE cluster barrier: buffer b is rewritten by
  // exchange k + 2, which a CTA can only reach through the barrier of exchange k + 1, i.e. after every reader of k is done.
  uint32_t xpar;                        // buffer of the next exchange
  float xchg[2][4][kH];
  double xchgd[2][40];
  float sgp[kH], sgxp[kH];              // this CTA's partial sums behind sg / sgx
  double ktot[4 * kZ];                  // Kendall totals per descriptor: sum |p| over p > 0, -(sum |p| over p < 0), counts
  Ctx ctx;                              // the kernel's context (see Ctx)
};

enum StageId { kStBatch = 0, kStFwdWide, kStFwdHidden, kStFwdLatent, kStFwdEncLast, kStBwdWide, kStBwdHidden, kStBwdLatent,
               kStBwdEncLast, kStDecLastLoss, kStDecLastStoreV, kStDecLastFromDv, kStDis, kStKendall, kStMiMse, kStCount };

// scoped cycle counter: thread 0 adds the elapsed SM cycles of the enclosing stage to sm->prof[id]
struct StageTimer {
  long long t0;
  long long* slot;
  __device__ __forceinline__ StageTimer(long long* s) : slot(s) { if (threadIdx.x == 0) t0 = clock64(); }
  __device__ __forceinline__ ~StageTimer() { if (threadIdx.x == 0) *slot += clock64() - t0; }
};

// sub-stage probe: thread 0 adds the cycles since the previous probe to sm->prof[slot]
#ifndef RAAE_PROF_FWD
#define RAAE_PROF_FWD 0                 // 1: sub-stage probes of fwd_hidden64_tc in slots 16..22 (tools/stage_profile.py raw)
#endif
#define RAAE_FPROBE(slot) do { if (RAAE_PROF_FWD) RAAE_PROBE(slot); } while (0)
// sub-stage probes cost a clock read + a shared-memory update per tile: compiled in only for profiling builds
// (RAAE_NVCC_EXTRA="-DRAAE_PROFILE=1"); the per-stage StageTimer (one read per stage) is always on
#ifndef RAAE_PROFILE
#define RAAE_PROFILE RAAE_PROF_FWD
#endif
#if RAAE_PROFILE
#define RAAE_PROBE_INIT() long long probe_t_ = clock64()
#define RAAE_PROBE(slot) do { if (threadIdx.x == 0) { long long now_ = clock64(); sm->prof[slot] += now_ - probe_t_; probe_t_ = now_; } } while (0)
#else
#define RAAE_PROBE_INIT() do { } while (0)
#define RAAE_PROBE(slot) do { } while (0)
#endif

// Shared memory is always reached through the `extern __shared__` symbol (never through a pointer stored
// in a struct) so that the compiler emits LDS/STS with 32-bit addresses instead of generic LD/ST.
constexpr size_t kArenaOffset = ((sizeof(SmemFixed) + 1023) / 1024) * 1024;   // tcgen05 SWIZZLE_128B tiles need 1 KB alignment
#define RAAE_SMEM()                                                            \
  extern __shared__ __align__(1024) unsigned char raae_smem_raw[];             \
  SmemFixed* const sm = reinterpret_cast<SmemFixed*>(raae_smem_raw);           \
  float* const arena = reinterpret_cast<float*>(raae_smem_raw + kArenaOffset); \
  (void)arena

constexpr int kArenaFloats = 50944;     // 199 KB; see the per-stage carve-ups below
constexpr int kTile = kTM * kLD;        // 8704 floats
constexpr int kWideTile = kTM * kLDW;   // 33280 floats
constexpr int kWTile = kH * kLD;        // 4352 floats

__device__ __forceinline__ const raae_net_layout& NL(const Ctx& c, int net) { return c.p->lay.net[net]; }
__device__ __forceinline__ float* netp(const Ctx& c, int net) { return c.st + c.p->lay.net[net].param_off; }

// ------------------------------------------------------------------------------------------
// cluster collectives (no-ops for csize == 1).  Called by ALL threads of ALL CTAs of the trial, the same number of times.
// ------------------------------------------------------------------------------------------
// index of tile t among this CTA's tiles (t = crank, crank + csize, ...); with one CTA per trial it IS the tile index
__device__ __forceinline__ int tile_iter(const Ctx& c, int t) {
#if RAAE_CLUSTER
  return (t - c.crank) / c.csize;
#else
  return t;
#endif
}
// does this CTA own any tile of a batch of ntiles tiles?  (always, with one CTA per trial: ntiles >= 1)
__device__ __forceinline__ bool has_tiles(const Ctx& c, int ntiles) {
#if RAAE_CLUSTER
  return c.crank < ntiles;
#else
  return true;
#endif
}
// barrier between stages: orders this trial's global / shared writes before the reads of the next stage in every CTA
__device__ __forceinline__ void stage_sync(const Ctx& c) {
  if (c.csize > 1) cl::sync();
  else __syncthreads();
}
// in-place sum over the cluster of vec[0..n) (shared memory, n <= 256); every CTA ends with the same bits (rank order)
__device__ __forceinline__ void cluster_allreduce_f(const Ctx& c, SmemFixed* sm, float* vec, int n) {
  if (c.csize == 1) return;
  const int tid = threadIdx.x;
  __syncthreads();
  const uint32_t par = sm->xpar;
  float* x = &sm->xchg[par][0][0];
  if (tid < n) x[tid] = vec[tid];
  cl::sync();
  if (tid < n) {
    float v[RAAE_MAX_CTAS];
#pragma unroll
    for (int r = 0; r < RAAE_MAX_CTAS; ++r) v[r] = r < c.csize ? cl::ld_f32(cl::map(x + tid, (uint32_t)r)) : 0.f;   // all loads in flight
    float s = 0.f;
#pragma unroll
    for (int r = 0; r < RAAE_MAX_CTAS; ++r) s += v[r];
    vec[tid] = s;
  }
  if (tid == 0) sm->xpar = par ^ 1u;     // every thread read `par` before the cluster barrier; published by the barrier below
  __syncthreads();
}
__device__ __forceinline__ void cluster_allreduce_d(const Ctx& c, SmemFixed* sm, double* vec, int n) {   // n <= 40
  if (c.csize == 1) return;
  const int tid = threadIdx.x;
  __syncthreads();
  const uint32_t par = sm->xpar;
  double* x = &sm->xchgd[par][0];
  if (tid < n) x[tid] = vec[tid];
  cl::sync();
  if (tid < n) {
    double v[RAAE_MAX_CTAS];
#pragma unroll
    for (int r = 0; r < RAAE_MAX_CTAS; ++r) v[r] = r < c.csize ? cl::ld_f64(cl::map(x + tid, (uint32_t)r)) : 0.0;
    double s = 0.0;
#pragma unroll
    for (int r = 0; r < RAAE_MAX_CTAS; ++r) s += v[r];
    vec[tid] = s;
  }
  if (tid == 0) sm->xpar = par ^ 1u;
  __syncthreads();
}
// totals of the BN-backward sums: the stage left this CTA's partials in sm->sgp / sgxp; call after the stage's first cluster
// barrier (the partials stay untouched until the barrier that ends the stage)
__device__ __forceinline__ void cluster_gather_sg(const Ctx& c, SmemFixed* sm) {
  if (c.csize == 1) return;
  const int tid = threadIdx.x;
  if (tid < 2 * kH) {
    const float* src = tid < kH ? &sm->sgp[tid] : &sm->sgxp[tid - kH];
    float v[RAAE_MAX_CTAS];
#pragma unroll
    for (int r = 0; r < RAAE_MAX_CTAS; ++r) v[r] = r < c.csize ? cl::ld_f32(cl::map(src, (uint32_t)r)) : 0.f;
    float s = 0.f;
#pragma unroll
    for (int r = 0; r < RAAE_MAX_CTAS; ++r) s += v[r];
    if (tid < kH) sm->sg[tid] = s; else sm->sgx[tid - kH] = s;
  }
}

__device__ inline MaskSrc make_mask(const Ctx& c, int net, int inst, int layer) {
  MaskSrc m;
  m.ptr = nullptr; m.key = 0u; m.thresh = 0u; m.scale = 1.f;
  if (!c.train) return m;
  const int g = net == kS ? 1 : 0;
  if (c.drop_thresh[g] == 0u) return m;
  m.scale = c.drop_scale[g];
  if (c.a->debug) {
    const uint8_t* ptr = net == kE ? c.a->dbg.mask_enc[inst][layer]
                       : net == kD ? c.a->dbg.mask_dec[inst][layer] : c.a->dbg.mask_dis[inst][layer];
    if (ptr) { m.ptr = ptr; return m; }
  }
  uint32_t kind = (net == kE ? kStreamEncMask : net == kD ? kStreamDecMask : kStreamDisMask) + inst * 8 + layer;
  m.key = stream_key(c.seed, c.step_id, kind);
  m.thresh = c.drop_thresh[g];
  return m;
}

// column sums over the 32 lanes of a warp: on return lane j holds sum_lanes v[j] (31 shuffles, recursive halving)
__device__ __forceinline__ float warp_colsum32(float (&v)[32]) {
  const int lane = threadIdx.x & 31;
#pragma unroll
  for (int s = 16; s > 0; s >>= 1) {
    const bool up = (lane & s) != 0;
#pragma unroll
    for (int i = 0; i < s; ++i) {
      const float send = up ? v[i] : v[i + s];
      const float keep = up ? v[i + s] : v[i];
      v[i] = keep + __shfl_xor_sync(0xffffffffu, send, s);
    }
  }
  return v[0];
}

// ------------------------------------------------------------------------------------------
// tile builders (global -> transformed smem tile, zero-filled beyond `nv` rows)
// ------------------------------------------------------------------------------------------
// a = dropout(BN(PReLU(u)))  for a [kTM][64] panel of pre-activations
__device__ __forceinline__ void build_act_tile(float* __restrict__ At, const float* __restrict__ u, int row0, int nv,
                                               const float* mean, const float* inv, const float* __restrict__ slope_g,
                                               const MaskSrc& mk) {
  const int c4 = (threadIdx.x & 15) * 4, r0 = threadIdx.x >> 4;
  const float4 mu = *reinterpret_cast<const float4*>(mean + c4);
  const float4 is = *reinterpret_cast<const float4*>(inv + c4);
  const float4 sl = *reinterpret_cast<const float4*>(slope_g + c4);
  float4 uu[kTM / 16];
#pragma unroll
  for (int i = 0; i < kTM / 16; ++i) {       // all loads in flight before the first use
    const int r = r0 + 16 * i;
    uu[i] = r < nv ? *reinterpret_cast<const float4*>(u + (size_t)(row0 + r) * kH + c4) : make_float4(0.f, 0.f, 0.f, 0.f);
  }
  const uint32_t kbits = mask_keep4_rows(mk, row0, r0, c4, nv);     // rows >= nv: all dropped -> zero-filled
#pragma unroll
  for (int i = 0; i < kTM / 16; ++i) {
    const int r = r0 + 16 * i;
    const uint32_t kb = kbits >> (4 * i);
    float4 o;
    o.x = (kb & 1u) ? (prelu_f(uu[i].x, sl.x) - mu.x) * is.x * mk.scale : 0.f;
    o.y = (kb & 2u) ? (prelu_f(uu[i].y, sl.y) - mu.y) * is.y * mk.scale : 0.f;
    o.z = (kb & 4u) ? (prelu_f(uu[i].z, sl.z) - mu.z) * is.z * mk.scale : 0.f;
    o.w = (kb & 8u) ? (prelu_f(uu[i].w, sl.w) - mu.w) * is.w * mk.scale : 0.f;
    *reinterpret_cast<float4*>(At + r * kLD + c4) = o;
  }
}

// same transform applied IN PLACE to a raw tile prefetched by prefetch_panel_tile (own elements only)
__device__ __forceinline__ void transform_act_tile(float* __restrict__ At, int row0, int nv, const float* mean,
                                                   const float* inv, const float* __restrict__ slope_s, const MaskSrc& mk) {
  const int c4 = (threadIdx.x & 15) * 4, r0 = threadIdx.x >> 4;
  const float4 mu = *reinterpret_cast<const float4*>(mean + c4);
  const float4 is = *reinterpret_cast<const float4*>(inv + c4);
  const float4 sl = *reinterpret_cast<const float4*>(slope_s + c4);
  const uint32_t kbits = mask_keep4_rows(mk, row0, r0, c4, nv);     // rows >= nv: all dropped -> zero-filled
#pragma unroll
  for (int i = 0; i < kTM / 16; ++i) {
    const int r = r0 + 16 * i;
    const float4 uu = *reinterpret_cast<const float4*>(At + r * kLD + c4);
    const uint32_t kb = kbits >> (4 * i);
    float4 o;
    o.x = (kb & 1u) ? (prelu_f(uu.x, sl.x) - mu.x) * is.x * mk.scale : 0.f;
    o.y = (kb & 2u) ? (prelu_f(uu.y, sl.y) - mu.y) * is.y * mk.scale : 0.f;
    o.z = (kb & 4u) ? (prelu_f(uu.z, sl.z) - mu.z) * is.z * mk.scale : 0.f;
    o.w = (kb & 8u) ? (prelu_f(uu.w, sl.w) - mu.w) * is.w * mk.scale : 0.f;
    *reinterpret_cast<float4*>(At + r * kLD + c4) = o;
  }
}

// columns [k0, k0 + 64) of wide rows -> [kTM][kLD]; optional decoder activation on load
__device__ __forceinline__ void build_wide_chunk(float* __restrict__ At, const float* __restrict__ src, int ld, int dim,
                                                 int k0, int row0, int nv, int act /*0 none 1 softplus 2 relu*/) {
  const int c4 = (threadIdx.x & 15) * 4, r0 = threadIdx.x >> 4;
  float4 v[kTM / 16];
#pragma unroll
  for (int i = 0; i < kTM / 16; ++i) {
    const int r = r0 + 16 * i;
    v[i] = (r < nv && k0 + c4 < dim) ? *reinterpret_cast<const float4*>(src + (size_t)(row0 + r) * ld + k0 + c4)
                                     : make_float4(0.f, 0.f, 0.f, 0.f);
  }
#pragma unroll
  for (int i = 0; i < kTM / 16; ++i) {
    const int r = r0 + 16 * i;
    float4 o = v[i];
    if (r < nv && k0 + c4 < dim) {
      if (act == 1) { o.x = softplus2_f(o.x); o.y = softplus2_f(o.y); o.z = softplus2_f(o.z); o.w = softplus2_f(o.w); }
      else if (act == 2) { o.x = fmaxf(o.x, 0.f); o.y = fmaxf(o.y, 0.f); o.z = fmaxf(o.z, 0.f); o.w = fmaxf(o.w, 0.f); }
    }
    *reinterpret_cast<float4*>(At + r * kLD + c4) = o;
  }
}

// whole wide rows -> [kTM][kLDW]
__device__ __forceinline__ void build_wide_tile(float* __restrict__ Xt, const float* __restrict__ src, int ld, int dim,
                                                int row0, int nv, int act) {
  const int c4 = (threadIdx.x & 63) * 4, r0 = threadIdx.x >> 6;
  for (int b = 0; b < kTM / 4; b += 8) {
    float4 v[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int r = r0 + 4 * (b + i);
      v[i] = (r < nv && c4 < dim) ? *reinterpret_cast<const float4*>(src + (size_t)(row0 + r) * ld + c4)
                                  : make_float4(0.f, 0.f, 0.f, 0.f);
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int r = r0 + 4 * (b + i);
      float4 o = v[i];
      if (r < nv && c4 < dim) {
        if (act == 1) { o.x = softplus2_f(o.x); o.y = softplus2_f(o.y); o.z = softplus2_f(o.z); o.w = softplus2_f(o.w); }
        else if (act == 2) { o.x = fmaxf(o.x, 0.f); o.y = fmaxf(o.y, 0.f); o.z = fmaxf(o.z, 0.f); o.w = fmaxf(o.w, 0.f); }
      }
      *reinterpret_cast<float4*>(Xt + r * kLDW + c4) = o;
    }
  }
}

// latent rows [rows][kZ] -> [kTM][kZ]; optional BN transform (encoder output)
__device__ __forceinline__ void build_latent_tile(float* __restrict__ Zt, const float* __restrict__ src, int row0, int nv,
                                                  int nstyle, const float* mean, const float* inv) {
  // one float4 per thread: row tid / 2, columns 4 (tid & 1) .. (kTM * kZ / 4 == kThreads)
  const int r = threadIdx.x >> 1, k0 = (threadIdx.x & 1) * 4;
  float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
  if (r < nv) v = *reinterpret_cast<const float4*>(src + (size_t)(row0 + r) * kZ + k0);
  float o[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
  for (int e = 0; e < 4; ++e) {
    const int k = k0 + e;
    if (r < nv && k < nstyle) { if (mean) o[e] = (o[e] - mean[k]) * inv[k]; }
    else o[e] = 0.f;
  }
  *reinterpret_cast<float4*>(Zt + r * kZ + k0) = make_float4(o[0], o[1], o[2], o[3]);
}

// weights [n_rows][K] (nn.Linear layout) rows [n0, n0+64) -> smem [64][ld], zero-filled
__device__ __forceinline__ void load_w_rows(float* __restrict__ Ws, int ld, const float* __restrict__ W, int K, int n0,
                                            int n_rows) {
  const int k4n = (K + 3) >> 2;  // K is a multiple of 4 for every wide/hidden layer
  const int per_row = ld >> 2, total = kH * per_row;
  // batches of 4 float4 per thread: all loads of a batch are in flight before the first store
  for (int o0 = threadIdx.x; o0 < total; o0 += kThreads * 4) {
    float4 v[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int o = o0 + kThreads * u;
      const int n = o / per_row, k4 = o - n * per_row;
      v[u] = (o < total && n0 + n < n_rows && k4 < k4n) ? *reinterpret_cast<const float4*>(W + (size_t)(n0 + n) * K + 4 * k4)
                                                        : make_float4(0.f, 0.f, 0.f, 0.f);
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int o = o0 + kThreads * u;
      const int n = o / per_row, k4 = o - n * per_row;
      if (o < total) *reinterpret_cast<float4*>(Ws + n * ld + 4 * k4) = v[u];
    }
  }
}

// asynchronous variant for 64-column weight rows: rows [n0, n0+64) x [0, 64) -> smem [64][ld] with cp.async (rows beyond
// n_rows are zero-filled with plain stores); the caller commits / waits
__device__ __forceinline__ void prefetch_w_rows64(float* __restrict__ Ws, int ld, const float* __restrict__ W, int n0, int n_rows) {
#pragma unroll
  for (int u = 0; u < 4; ++u) {
    const int o = threadIdx.x + kThreads * u, n = o >> 4, k4 = (o & 15) * 4;
    if (n0 + n < n_rows) cp_async16(Ws + n * ld + k4, W + (size_t)(n0 + n) * kH + k4);
    else *reinterpret_cast<float4*>(Ws + n * ld + k4) = make_float4(0.f, 0.f, 0.f, 0.f);
  }
}

// ------------------------------------------------------------------------------------------
// AdamW (torch/optim/adam.py single-tensor path)
// ------------------------------------------------------------------------------------------
// thread 0 only; followed by __syncthreads() at the call site
__device__ __forceinline__ void adam_prepare(const Ctx& c, SmemFixed* sm, int o) {
  const raae_opt_layout& ol = c.p->lay.opt[o];
  double lr = (double)c.st[ol.scalar_off + 0];
  double t = (double)c.st[ol.scalar_off + 1] + 1.0;
  double b1 = c.hp[RAAE_HP_BETA1 + o], b2 = c.hp[RAAE_HP_BETA2 + o], wd = c.hp[RAAE_HP_WD + o];
  double bc1 = 1.0 - pow(b1, t), bc2 = 1.0 - pow(b2, t);
  sm->ad[0] = (float)(1.0 - lr * wd);
  sm->ad[1] = (float)(1.0 - b1);
  sm->ad[2] = (float)b2;
  sm->ad[3] = (float)(1.0 - b2);
  sm->ad[4] = (float)(lr / bc1);
  sm->ad[5] = (float)sqrt(bc2);
}

// thread 0 only, after every parameter of the phase has been updated
__device__ inline void adam_finish(const Ctx& c, int o) {
  if (c.apply && c.crank == 0) c.st[c.p->lay.opt[o].scalar_off + 1] += 1.f;
}

// Cluster variant of adam_apply: every CTA of the trial holds its PARTIAL gradient at the same shared-memory address `g`
// (complete since the stage's first cluster barrier); rank r owns elements [r per, (r + 1) per), sums the partials of all
// ranks for them through distributed shared memory (rank order) and applies / exports them.  The caller ends the stage
// with another cluster barrier before `g` is reused and before any CTA reads the updated parameters.
__device__ __noinline__ void adam_apply_cluster(const Ctx& c_ref, const SmemFixed* sm, int o, int net, int poff, int n,
                                                const float* g) {
  const Ctx c = c_ref;
  const raae_opt_layout& ol = c.p->lay.opt[o];
  if (ol.net_off[net] < 0) return;
  float* dbg = c.a->debug ? c.a->dbg.grads[o]
             : (c.a->grads_out[o] ? c.a->grads_out[o] + (size_t)c.trial * ol.n : nullptr);
  if (!dbg && !c.apply) return;
  const int C = c.csize;
  const int per = (((n + C - 1) / C) + 3) & ~3;
  const int i0 = min(n, c.crank * per), i1 = min(n, i0 + per);
  uint32_t gbase[RAAE_MAX_CTAS];
#pragma unroll
  for (int r = 0; r < RAAE_MAX_CTAS; ++r) gbase[r] = cl::map(g, (uint32_t)min(r, C - 1));
  float* P = c.st + c.p->lay.net[net].param_off + poff;
  float* M = c.st + ol.m_off + ol.net_off[net] + poff;
  float* V = c.st + ol.v_off + ol.net_off[net] + poff;
  const float decay = sm->ad[0], w1 = sm->ad[1], b2 = sm->ad[2], w2 = sm->ad[3], ss = sm->ad[4], bc2s = sm->ad[5];
  const bool apply = c.apply != 0;
  for (int base = i0 + threadIdx.x; base < i1; base += kThreads * 4) {
    // branch-free: every DSMEM load of the four elements is in flight before the first sum (out-of-range elements read
    // element i0 and are never stored), the parameter / moment loads are predicated, the stores guarded
    float pv[4], mv[4], vv[4], gv[4], part[4][RAAE_MAX_CTAS];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const int i = base + k * kThreads;
      const int is = i < i1 ? i : i0;
#pragma unroll
      for (int r = 0; r < RAAE_MAX_CTAS; ++r) part[k][r] = cl::ld_f32(gbase[r] + 4u * (uint32_t)is);   // gbase[r >= C] repeats rank C - 1
      const bool ok = apply && i < i1;
      pv[k] = ok ? P[i] : 0.f; mv[k] = ok ? M[i] : 0.f; vv[k] = ok ? V[i] : 0.f;
    }
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      float sgrad = 0.f;
#pragma unroll
      for (int r = 0; r < RAAE_MAX_CTAS; ++r) sgrad += r < C ? part[k][r] : 0.f;       // rank order
      gv[k] = sgrad;
    }
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const int i = base + k * kThreads;
      const float gi = gv[k];
      float pp = pv[k], m = mv[k], v = vv[k];
      adamw_update(pp, m, v, gi, decay, w1, b2, w2, ss, bc2s);
      if (i < i1) {
        if (dbg) dbg[ol.net_off[net] + poff + i] = gi;
        if (apply) { P[i] = pp; M[i] = m; V[i] = v; }
      }
    }
  }
}

// all threads: update `n` parameters at offset `poff` of net `net` with gradient g (shared memory),
// and/or export the gradient to the debug buffer.
__device__ __forceinline__ void adam_apply(const Ctx& c, const SmemFixed* sm, int o, int net, int poff, int n,
                                           const float* __restrict__ g) {
  if (c.csize > 1) { adam_apply_cluster(c, sm, o, net, poff, n, g); return; }
  const raae_opt_layout& ol = c.p->lay.opt[o];
  float* dbg = c.a->debug ? c.a->dbg.grads[o]
             : (c.a->grads_out[o] ? c.a->grads_out[o] + (size_t)c.trial * ol.n : nullptr);
  if (dbg && ol.net_off[net] >= 0)
    for (int i = threadIdx.x; i < n; i += kThreads) dbg[ol.net_off[net] + poff + i] = g[i];
  if (!c.apply || ol.net_off[net] < 0) return;
  float* P = c.st + c.p->lay.net[net].param_off + poff;
  float* M = c.st + ol.m_off + ol.net_off[net] + poff;
  float* V = c.st + ol.v_off + ol.net_off[net] + poff;
  const float decay = sm->ad[0], w1 = sm->ad[1], b2 = sm->ad[2], w2 = sm->ad[3], ss = sm->ad[4], bc2s = sm->ad[5];
  // batches of 8 elements per thread: all 24 loads of a batch are in flight before the first use
  for (int base = threadIdx.x; base < n; base += kThreads * 8) {
    float pv[8], mv[8], vv[8], gv[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      const int i = base + k * kThreads;
      const bool ok = i < n;                  // predicated loads, no branch per element
      pv[k] = ok ? P[i] : 0.f; mv[k] = ok ? M[i] : 0.f; vv[k] = ok ? V[i] : 0.f; gv[k] = ok ? g[i] : 0.f;
    }
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      const int i = base + k * kThreads;
      float pp = pv[k], m = mv[k], v = vv[k];          // branch-free arithmetic, guarded stores: the eight elements interleave
      adamw_update(pp, m, v, gv[k], decay, w1, b2, w2, ss, bc2s);
      if (i < n) { P[i] = pp; M[i] = m; V[i] = v; }
    }
  }
}

// ------------------------------------------------------------------------------------------
// forward of one hidden block:  u = in @ W^T + b  (stored),  batch statistics of PReLU(u)
// ------------------------------------------------------------------------------------------
enum InKind { kInHidden = 0, kInWide = 1, kInLatent = 2 };

struct LayerIn {
  int kind;
  const float* src;    // hidden: u_prev [rows][64]; wide: [rows][ld]; latent: [rows][kZ]
  int ld, dim, act;    // wide only
  int img;             // wide only: 1 = the rows are the noised batch, whose operand images exist (ScratchLayout::xk / xm);
                       // 2 = the rows are y = act(v) of the MI phase (ScratchLayout::yk, forward only)
  int snet, slayer;    // hidden: BN statistics sm->mean/inv[snet][slayer] of the producing layer;
                       // latent: BN of the encoder output, or slayer < 0 for raw rows
  const float* slope;  // hidden: PReLU slopes (global) of the producing layer
  MaskSrc mask;        // hidden: dropout of the producing layer
};

// finalize BN statistics of channel c from the shifted sums; updates running buffers in train mode
// rm_old / rv_old: the running buffers of the channel as loaded at the START of the stage (bn_running_load) - loaded here,
// at its end, the read-modify-write exposed a DRAM round trip in front of every stage's closing barrier
__device__ __forceinline__ void bn_finalize(const Ctx& c, SmemFixed* sm, int net, int l, int ch, float shift, float s1,
                                            float s2, int nrows, bool running, float rm_old, float rv_old) {
  const raae_net_layout& nl = NL(c, net);
  float n = (float)nrows;
  float d = s1 / n;
  float mean = shift + d;
  float var = fmaxf(s2 / n - d * d, 0.f);
  sm->mean[net][l][ch] = mean;
  sm->inv[net][l][ch] = 1.f / sqrtf(var + kBnEps);
#ifdef RAAE_DEBUG_STATS
  if (net == kE && l == 0) { float* dd = c.sc + c.p->sl.rank; dd[ch] = mean; dd[64 + ch] = var; }
#endif
  if (!running) return;
  float* rm = c.st + nl.rm_off[l];
  float* rv = c.st + nl.rv_off[l];
  float unb = nrows > 1 ? var * (n / (n - 1.f)) : var;
  rm[ch] = (1.f - kBnMomentum) * rm_old + kBnMomentum * mean;
  rv[ch] = (1.f - kBnMomentum) * rv_old + kBnMomentum * unb;
}
// the running mean / variance of channel threadIdx.x of layer l (threads < 64 of the CTA that advances them, train mode)
__device__ __forceinline__ void bn_running_load(const Ctx& c, int net, int l, float& rm_old, float& rv_old) {
  rm_old = 0.f; rv_old = 0.f;
  if (c.train && threadIdx.x < kH && c.crank == 0) {
    const raae_net_layout& nl = NL(c, net);
    rm_old = c.st[nl.rm_off[l] + threadIdx.x];
    rv_old = c.st[nl.rv_off[l] + threadIdx.x];
  }
}

// Cluster merge of BatchNorm batch statistics (all threads; the values of threads tid < kH count): every CTA contributes the
// mean and the centred sum of squares (M2) of ITS rows; exact parallel-variance merge in rank order, identical in every CTA;
// rank 0 advances the running buffers.
__device__ __forceinline__ void bn_merge_cluster(const Ctx& c, SmemFixed* sm, int net, int l, float mean_l, float m2_l, float rm_old,
                                                 float rv_old) {
  const int tid = threadIdx.x;
  __syncthreads();
  const uint32_t par = sm->xpar;
  float* x = &sm->xchg[par][0][0];
  if (tid < kH) { x[tid] = mean_l; x[kH + tid] = m2_l; }
  cl::sync();
  if (tid < kH) {
    float mr[RAAE_MAX_CTAS], qr[RAAE_MAX_CTAS], nr[RAAE_MAX_CTAS];
#pragma unroll
    for (int r = 0; r < RAAE_MAX_CTAS; ++r) {            // all loads in flight before the first use
      const bool on = r < c.csize;
      mr[r] = on ? cl::ld_f32(cl::map(x + tid, (uint32_t)r)) : 0.f;
      qr[r] = on ? cl::ld_f32(cl::map(x + kH + tid, (uint32_t)r)) : 0.f;
      nr[r] = on ? (float)cl::own_rows(c.B, r, c.csize) : 0.f;
    }
    float mean = 0.f;
#pragma unroll
    for (int r = 0; r < RAAE_MAX_CTAS; ++r) mean += nr[r] * mr[r];
    mean /= (float)c.B;
    float m2 = 0.f;
#pragma unroll
    for (int r = 0; r < RAAE_MAX_CTAS; ++r) { const float d = mr[r] - mean; m2 += qr[r] + nr[r] * d * d; }
    bn_finalize(c, sm, net, l, tid, mean, 0.f, m2, c.B, c.crank == 0, rm_old, rv_old);
  }
  if (tid == 0) sm->xpar = par ^ 1u;     // every thread read `par` before the cluster barrier; published by the caller's barrier
  __syncthreads();
}
// end of a forward stage in train mode (all threads): shifted single-pass sums of this CTA's rows (threads tid < kH) ->
// sm->mean / inv of the layer (+ running buffers)
__device__ __forceinline__ void bn_stats_finish(const Ctx& c, SmemFixed* sm, int net, int l, float shift, float s1, float s2,
                                                float rm_old, float rv_old) {
  if (c.csize == 1) {
    if (threadIdx.x < kH) bn_finalize(c, sm, net, l, threadIdx.x, shift, s1, s2, c.B, true, rm_old, rv_old);
    return;
  }
  const float nl = (float)cl::own_rows(c.B, c.crank, c.csize);
  float mean_l = 0.f, m2_l = 0.f;
  if (nl > 0.f) { const float d = s1 / nl; mean_l = shift + d; m2_l = fmaxf(s2 - s1 * d, 0.f); }
  bn_merge_cluster(c, sm, net, l, mean_l, m2_l, rm_old, rv_old);
}

__device__ __noinline__ void fwd_hidden_edge(const Ctx& c_ref, int net, int l, const LayerIn& in_ref, float* __restrict__ u_out) {
  const Ctx c = c_ref;                 // register copy of the kernel context (SmemFixed::ctx, shared memory)
  const LayerIn in = in_ref;
  RAAE_SMEM();
  StageTimer timer_(&sm->prof[in.kind == kInWide ? kStFwdWide : in.kind == kInHidden ? kStFwdHidden : kStFwdLatent]);
  float rm_old, rv_old;                // running BatchNorm buffers of this layer, loaded a whole stage ahead of their update
  bn_running_load(c, net, l, rm_old, rv_old);
  const raae_net_layout& nl = NL(c, net);
  const int tid = threadIdx.x, ty = tid >> 4, tx = tid & 15, ch = tid & 63, q = tid >> 6;
  const int K = nl.in_dim[l];
  const float* Wg = netp(c, net) + nl.w_off[l];
  float* Ws = arena;                          // hidden: [64][kLD]; wide: [64][kLDW]; latent: [64][9]
  float* At = arena + kH * kLDW;              // [kTM][kLD] or [kTM][kZ]
  float* Ot = At + kTile;                     // [kTM][kLD]
  __syncthreads();
  if (in.kind == kInLatent) {
    for (int o = tid; o < kH * 9; o += kThreads) {
      int n = o / 9, k = o - n * 9;
      Ws[o] = k < K ? Wg[n * K + k] : 0.f;
    }
  } else {
    load_w_rows(Ws, in.kind == kInWide ? kLDW : kLD, Wg, K, 0, kH);
  }
  if (tid < kH) {
    sm->bias[tid] = netp(c, net)[nl.b_off[l] + tid];
    sm->slope[tid] = netp(c, net)[nl.a_off[l] + tid];
    if (!c.train) {
      sm->mean[net][l][tid] = c.st[nl.rm_off[l] + tid];
      sm->inv[net][l][tid] = 1.f / sqrtf(c.st[nl.rv_off[l] + tid] + kBnEps);
    }
  }
  __syncthreads();
  float4 s1v = make_float4(0.f, 0.f, 0.f, 0.f), s2v = make_float4(0.f, 0.f, 0.f, 0.f);
  const int ntiles = (c.B + kTM - 1) / kTM;
  for (int t = c.crank; t < ntiles; t += c.csize) {
    const int row0 = t * kTM, nv = min(kTM, c.B - row0);
    if (in.kind == kInLatent) {
      build_latent_tile(At, in.src, row0, nv, K, in.slayer >= 0 ? sm->mean[in.snet][in.slayer] : nullptr,
                        in.slayer >= 0 ? sm->inv[in.snet][in.slayer] : nullptr);
      __syncthreads();
      {
        // this thread's weight row in registers, the latent row as two broadcast loads
        float wr[kZ];
#pragma unroll
        for (int k = 0; k < kZ; ++k) wr[k] = Ws[ch * 9 + k];
        const float bc = sm->bias[ch];
#pragma unroll 4
        for (int i = 0; i < kTM / 4; ++i) {
          const int r = q + 4 * i;
          const float4 z0 = *reinterpret_cast<const float4*>(At + r * kZ);
          const float4 z1 = *reinterpret_cast<const float4*>(At + r * kZ + 4);
          float acc = bc;
          acc = fmaf(z0.x, wr[0], acc); acc = fmaf(z0.y, wr[1], acc); acc = fmaf(z0.z, wr[2], acc); acc = fmaf(z0.w, wr[3], acc);
          acc = fmaf(z1.x, wr[4], acc); acc = fmaf(z1.y, wr[5], acc); acc = fmaf(z1.z, wr[6], acc); acc = fmaf(z1.w, wr[7], acc);
          Ot[r * kLD + ch] = acc;
        }
      }
    } else {
      float acc[8][4];
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
      if (in.kind == kInHidden) {
        build_act_tile(At, in.src, row0, nv, sm->mean[in.snet][in.slayer], sm->inv[in.snet][in.slayer], in.slope, in.mask);
        __syncthreads();
        mma_nt<kH>(At, kLD, Ws, kLD, acc, ty, tx);
      } else {
        for (int k0 = 0; k0 < K; k0 += kH) {
          build_wide_chunk(At, in.src, in.ld, in.dim, k0, row0, nv, in.act);
          __syncthreads();
          mma_nt<kH>(At, kLD, Ws + k0, kLDW, acc, ty, tx);
          __syncthreads();
        }
      }
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) Ot[(ty + 16 * i) * kLD + tx + 16 * j] = acc[i][j] + sm->bias[tx + 16 * j];
    }
    __syncthreads();
    // elementwise epilogue: thread (ty, tx) owns channels 4tx..4tx+3 of rows ty, ty+16, ...
    const int c4 = tx * 4;
    const float4 a_sl = *reinterpret_cast<const float4*>(sm->slope + c4);
    float4 uo[kTM / 16];
#pragma unroll
    for (int i = 0; i < kTM / 16; ++i) uo[i] = *reinterpret_cast<const float4*>(Ot + (ty + 16 * i) * kLD + c4);
    if (c.train && t == c.crank) {
      float4 sp = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
      for (int i = 0; i < kTM / 16; ++i)
        if (ty + 16 * i < nv) {
          sp.x += prelu_f(uo[i].x, a_sl.x); sp.y += prelu_f(uo[i].y, a_sl.y);
          sp.z += prelu_f(uo[i].z, a_sl.z); sp.w += prelu_f(uo[i].w, a_sl.w);
        }
      *reinterpret_cast<float4*>(&sm->red[ty][c4]) = sp;
      __syncthreads();
      if (tid < kH) {
        float sacc = 0.f;
#pragma unroll
        for (int i = 0; i < 16; ++i) sacc += sm->red[i][tid];
        sm->shift[tid] = sacc / (float)nv;
      }
      __syncthreads();
    }
    const float4 sh = c.train ? *reinterpret_cast<const float4*>(sm->shift + c4) : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
    for (int i = 0; i < kTM / 16; ++i) {
      const int r = ty + 16 * i;
      if (r < nv) {
        *reinterpret_cast<float4*>(u_out + (size_t)(row0 + r) * kH + c4) = uo[i];
        float d;
        d = prelu_f(uo[i].x, a_sl.x) - sh.x; s1v.x += d; s2v.x = fmaf(d, d, s2v.x);
        d = prelu_f(uo[i].y, a_sl.y) - sh.y; s1v.y += d; s2v.y = fmaf(d, d, s2v.y);
        d = prelu_f(uo[i].z, a_sl.z) - sh.z; s1v.z += d; s2v.z = fmaf(d, d, s2v.z);
        d = prelu_f(uo[i].w, a_sl.w) - sh.w; s1v.w += d; s2v.w = fmaf(d, d, s2v.w);
      }
    }
    __syncthreads();
  }
  if (c.train) {
    const int c4 = tx * 4;
    *reinterpret_cast<float4*>(&sm->red[ty][c4]) = s1v;
    __syncthreads();
    float a1 = 0.f, a2 = 0.f;
    if (tid < kH) {
#pragma unroll
      for (int i = 0; i < 16; ++i) a1 += sm->red[i][tid];
    }
    __syncthreads();
    *reinterpret_cast<float4*>(&sm->red[ty][c4]) = s2v;
    __syncthreads();
    if (tid < kH) {
#pragma unroll
      for (int i = 0; i < 16; ++i) a2 += sm->red[i][tid];
    }
    bn_stats_finish(c, sm, net, l, tid < kH ? sm->shift[tid] : 0.f, a1, a2, rm_old, rv_old);
  }
  __syncthreads();
}

// forward of a hidden block whose input is the latent (K = nstyle <= 8: decoder layer 0).  Thread (ty, c4) owns channels
// c4..c4+3 of rows ty, ty + 16, ...: its four weight rows stay in registers, a latent row is two broadcast loads, the outputs
// go straight into the epilogue's registers (no bounce through a [128][64] tile, one barrier less per tile), and the next
// tile's latent rows (one float4 per thread) are loaded while this one is contracted.
__device__ __noinline__ void fwd_latent(const Ctx& c_ref, int net, int l, const LayerIn& in_ref, float* __restrict__ u_out) {
  const Ctx c = c_ref;                 // register copy of the kernel context (SmemFixed::ctx, shared memory)
  const LayerIn in = in_ref;
  RAAE_SMEM();
  StageTimer timer_(&sm->prof[kStFwdLatent]);
  float rm_old, rv_old;                // running BatchNorm buffers of this layer, loaded a whole stage ahead of their update
  bn_running_load(c, net, l, rm_old, rv_old);
  const raae_net_layout& nl = NL(c, net);
  const int tid = threadIdx.x, ty = tid >> 4, c4 = (tid & 15) * 4;
  const int K = nl.in_dim[l];
  const float* Wg = netp(c, net) + nl.w_off[l];
  float* Ws = arena;                          // [64][9]
  float* Zt = arena + 1024;                   // [kTM][kZ]
  const float* zmean = in.slayer >= 0 ? sm->mean[in.snet][in.slayer] : nullptr;
  const float* zinv = in.slayer >= 0 ? sm->inv[in.snet][in.slayer] : nullptr;
  const int ntiles = (c.B + kTM - 1) / kTM;
  const int zr = tid >> 1, zk0 = (tid & 1) * 4;      // this thread's float4 of a latent tile: row tid / 2, columns 4 (tid & 1) ..
  auto load_z = [&](int t) {
    const int row0 = t * kTM;
    return (t < ntiles && row0 + zr < c.B) ? *reinterpret_cast<const float4*>(in.src + (size_t)(row0 + zr) * kZ + zk0)
                                           : make_float4(0.f, 0.f, 0.f, 0.f);
  };
  float4 zraw = load_z(c.crank);
  __syncthreads();
  for (int o = tid; o < kH * 9; o += kThreads) {
    int n = o / 9, k = o - n * 9;
    Ws[o] = k < K ? Wg[n * K + k] : 0.f;
  }
  if (tid < kH) {
    sm->bias[tid] = netp(c, net)[nl.b_off[l] + tid];
    sm->slope[tid] = netp(c, net)[nl.a_off[l] + tid];
    if (!c.train) {
      sm->mean[net][l][tid] = c.st[nl.rm_off[l] + tid];
      sm->inv[net][l][tid] = 1.f / sqrtf(c.st[nl.rv_off[l] + tid] + kBnEps);
    }
  }
  __syncthreads();
  float wr[4][kZ];
#pragma unroll
  for (int j = 0; j < 4; ++j)
#pragma unroll
    for (int k = 0; k < kZ; ++k) wr[j][k] = Ws[(c4 + j) * 9 + k];
  const float4 b4 = *reinterpret_cast<const float4*>(sm->bias + c4);
  const float bb[4] = {b4.x, b4.y, b4.z, b4.w};
  const float4 a_sl = *reinterpret_cast<const float4*>(sm->slope + c4);
  float4 s1v = make_float4(0.f, 0.f, 0.f, 0.f), s2v = make_float4(0.f, 0.f, 0.f, 0.f);
  for (int t = c.crank; t < ntiles; t += c.csize) {
    const int row0 = t * kTM, nv = min(kTM, c.B - row0);
    {
      // latent tile (normalised with the BatchNorm of the encoder output where it applies; zero beyond nstyle / nv)
      float o[4] = {zraw.x, zraw.y, zraw.z, zraw.w};
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const int k = zk0 + e;
        if (zr < nv && k < K) { if (zmean) o[e] = (o[e] - zmean[k]) * zinv[k]; }
        else o[e] = 0.f;
      }
      *reinterpret_cast<float4*>(Zt + zr * kZ + zk0) = make_float4(o[0], o[1], o[2], o[3]);
    }
    __syncthreads();
    zraw = load_z(t + c.csize);                  // in flight during the contraction and the epilogue
    float4 uo[kTM / 16];
#pragma unroll
    for (int i = 0; i < kTM / 16; ++i) {
      const int r = ty + 16 * i;
      const float4 z0 = *reinterpret_cast<const float4*>(Zt + r * kZ);
      const float4 z1 = *reinterpret_cast<const float4*>(Zt + r * kZ + 4);
      float a[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        float acc = bb[j];
        acc = fmaf(z0.x, wr[j][0], acc); acc = fmaf(z0.y, wr[j][1], acc); acc = fmaf(z0.z, wr[j][2], acc); acc = fmaf(z0.w, wr[j][3], acc);
        acc = fmaf(z1.x, wr[j][4], acc); acc = fmaf(z1.y, wr[j][5], acc); acc = fmaf(z1.z, wr[j][6], acc); acc = fmaf(z1.w, wr[j][7], acc);
        a[j] = acc;
      }
      uo[i] = make_float4(a[0], a[1], a[2], a[3]);
    }
    if (c.train && t == c.crank) {
      float4 sp = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
      for (int i = 0; i < kTM / 16; ++i)
        if (ty + 16 * i < nv) {
          sp.x += prelu_f(uo[i].x, a_sl.x); sp.y += prelu_f(uo[i].y, a_sl.y);
          sp.z += prelu_f(uo[i].z, a_sl.z); sp.w += prelu_f(uo[i].w, a_sl.w);
        }
      *reinterpret_cast<float4*>(&sm->red[ty][c4]) = sp;
      __syncthreads();
      if (tid < kH) {
        float sacc = 0.f;
#pragma unroll
        for (int i = 0; i < 16; ++i) sacc += sm->red[i][tid];
        sm->shift[tid] = sacc / (float)nv;
      }
      __syncthreads();
    }
    const float4 sh = c.train ? *reinterpret_cast<const float4*>(sm->shift + c4) : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
    for (int i = 0; i < kTM / 16; ++i) {
      const int r = ty + 16 * i;
      if (r < nv) {
        *reinterpret_cast<float4*>(u_out + (size_t)(row0 + r) * kH + c4) = uo[i];
        float d;
        d = prelu_f(uo[i].x, a_sl.x) - sh.x; s1v.x += d; s2v.x = fmaf(d, d, s2v.x);
        d = prelu_f(uo[i].y, a_sl.y) - sh.y; s1v.y += d; s2v.y = fmaf(d, d, s2v.y);
        d = prelu_f(uo[i].z, a_sl.z) - sh.z; s1v.z += d; s2v.z = fmaf(d, d, s2v.z);
        d = prelu_f(uo[i].w, a_sl.w) - sh.w; s1v.w += d; s2v.w = fmaf(d, d, s2v.w);
      }
    }
    __syncthreads();                             // the latent tile is rewritten by the next iteration
  }
  if (c.train) {
    *reinterpret_cast<float4*>(&sm->red[ty][c4]) = s1v;
    __syncthreads();
    float a1 = 0.f, a2 = 0.f;
    if (tid < kH) {
#pragma unroll
      for (int i = 0; i < 16; ++i) a1 += sm->red[i][tid];
    }
    __syncthreads();
    *reinterpret_cast<float4*>(&sm->red[ty][c4]) = s2v;
    __syncthreads();
    if (tid < kH) {
#pragma unroll
      for (int i = 0; i < 16; ++i) a2 += sm->red[i][tid];
    }
    bn_stats_finish(c, sm, net, l, tid < kH ? sm->shift[tid] : 0.f, a1, a2, rm_old, rv_old);
  }
  __syncthreads();
}

// forward of a hidden block whose input is another hidden block's panel (K = 64): the raw pre-activation tile of
// tile t+1 is prefetched with cp.async while tile t runs its contraction and epilogue.
__device__ __noinline__ void fwd_hidden64(const Ctx& c_ref, int net, int l, const LayerIn& in_ref, float* __restrict__ u_out) {
  const Ctx c = c_ref;                 // register copy of the kernel context (SmemFixed::ctx, shared memory)
  const LayerIn in = in_ref;
  RAAE_SMEM();
  StageTimer timer_(&sm->prof[kStFwdHidden]);
  float rm_old, rv_old;                // running BatchNorm buffers of this layer, loaded a whole stage ahead of their update
  bn_running_load(c, net, l, rm_old, rv_old);
  const raae_net_layout& nl = NL(c, net);
  const int tid = threadIdx.x, ty = tid >> 4, tx = tid & 15, c4 = tx * 4;
  const float* Wg = netp(c, net) + nl.w_off[l];
  float* Ws = arena;                           // [64][kLD]
  float* Rb[2] = {arena + kWTile, arena + kWTile + kTile};
  float* Ot = arena + kWTile + 2 * kTile;      // [kTM][kLD]
  float* slope_in = sm->cg;                    // PReLU slopes of the producing layer (cg is free during forwards)
  const float* mean_in = sm->mean[in.snet][in.slayer];
  const float* inv_in = sm->inv[in.snet][in.slayer];
  const int ntiles = (c.B + kTM - 1) / kTM;
  __syncthreads();
  const int t_first = c.crank, tstep = c.csize;
  if (has_tiles(c, ntiles)) prefetch_panel_tile(Rb[0], in.src, t_first * kTM, min(kTM, c.B - t_first * kTM));
  cp_async_commit();
  load_w_rows(Ws, kLD, Wg, kH, 0, kH);
  if (tid < kH) {
    sm->bias[tid] = netp(c, net)[nl.b_off[l] + tid];
    sm->slope[tid] = netp(c, net)[nl.a_off[l] + tid];
    slope_in[tid] = in.slope[tid];
    if (!c.train) {
      sm->mean[net][l][tid] = c.st[nl.rm_off[l] + tid];
      sm->inv[net][l][tid] = 1.f / sqrtf(c.st[nl.rv_off[l] + tid] + kBnEps);
    }
  }
  __syncthreads();
  float4 s1v = make_float4(0.f, 0.f, 0.f, 0.f), s2v = make_float4(0.f, 0.f, 0.f, 0.f);
  for (int t = t_first; t < ntiles; t += tstep) {
    const int it = tile_iter(c, t);
    const int row0 = t * kTM, nv = min(kTM, c.B - row0);
    float* At = Rb[it & 1];
    if (t + tstep < ntiles) prefetch_panel_tile(Rb[(it + 1) & 1], in.src, row0 + tstep * kTM, min(kTM, c.B - row0 - tstep * kTM));
    cp_async_commit();
    cp_async_wait<1>();
    transform_act_tile(At, row0, nv, mean_in, inv_in, slope_in, in.mask);
    __syncthreads();
    float acc[8][4];
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
    mma_nt<kH>(At, kLD, Ws, kLD, acc, ty, tx);
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) Ot[(ty + 16 * i) * kLD + tx + 16 * j] = acc[i][j] + sm->bias[tx + 16 * j];
    __syncthreads();
    const float4 a_sl = *reinterpret_cast<const float4*>(sm->slope + c4);
    float4 uo[kTM / 16];
#pragma unroll
    for (int i = 0; i < kTM / 16; ++i) uo[i] = *reinterpret_cast<const float4*>(Ot + (ty + 16 * i) * kLD + c4);
    if (c.train && it == 0) {
      float4 sp = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
      for (int i = 0; i < kTM / 16; ++i)
        if (ty + 16 * i < nv) {
          sp.x += prelu_f(uo[i].x, a_sl.x); sp.y += prelu_f(uo[i].y, a_sl.y);
          sp.z += prelu_f(uo[i].z, a_sl.z); sp.w += prelu_f(uo[i].w, a_sl.w);
        }
      *reinterpret_cast<float4*>(&sm->red[ty][c4]) = sp;
      __syncthreads();
      if (tid < kH) {
        float sacc = 0.f;
#pragma unroll
        for (int i = 0; i < 16; ++i) sacc += sm->red[i][tid];
        sm->shift[tid] = sacc / (float)nv;
      }
      __syncthreads();
    }
    const float4 sh = c.train ? *reinterpret_cast<const float4*>(sm->shift + c4) : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
    for (int i = 0; i < kTM / 16; ++i) {
      const int r = ty + 16 * i;
      if (r < nv) {
        *reinterpret_cast<float4*>(u_out + (size_t)(row0 + r) * kH + c4) = uo[i];
        float d;
        d = prelu_f(uo[i].x, a_sl.x) - sh.x; s1v.x += d; s2v.x = fmaf(d, d, s2v.x);
        d = prelu_f(uo[i].y, a_sl.y) - sh.y; s1v.y += d; s2v.y = fmaf(d, d, s2v.y);
        d = prelu_f(uo[i].z, a_sl.z) - sh.z; s1v.z += d; s2v.z = fmaf(d, d, s2v.z);
        d = prelu_f(uo[i].w, a_sl.w) - sh.w; s1v.w += d; s2v.w = fmaf(d, d, s2v.w);
      }
    }
    // no barrier here: the next iteration's barrier (after its transform) orders these Ot reads before the next
    // Ot writes, and the buffer prefetched next was last read by the contraction two barriers ago
  }
  cp_async_wait<0>();
  if (c.train) {
    __syncthreads();
    *reinterpret_cast<float4*>(&sm->red[ty][c4]) = s1v;
    __syncthreads();
    float a1 = 0.f, a2 = 0.f;
    if (tid < kH) {
#pragma unroll
      for (int i = 0; i < 16; ++i) a1 += sm->red[i][tid];
    }
    __syncthreads();
    *reinterpret_cast<float4*>(&sm->red[ty][c4]) = s2v;
    __syncthreads();
    if (tid < kH) {
#pragma unroll
      for (int i = 0; i < 16; ++i) a2 += sm->red[i][tid];
    }
    bn_stats_finish(c, sm, net, l, tid < kH ? sm->shift[tid] : 0.f, a1, a2, rm_old, rv_old);
  }
  __syncthreads();
}

// fwd_hidden64 with the contraction on the tensor core (tcgen05.mma kind::tf32, 3 x TF32 split), software-pipelined over
// the 128-row tiles: two operand buffers and two TMEM accumulators (columns [0,64) and [64,128)), so the MMAs of tile
// t+1 run while tile t is read back; the raw pre-activation tile of tile t+2 is in flight (cp.async) meanwhile.
// The epilogue works in the TMEM layout (thread = row, 32 consecutive columns): bias, store, and the shifted
// single-pass BatchNorm sums as per-thread partials that are reduced over the rows once per layer.
__device__ __noinline__ void fwd_hidden64_tc(const Ctx& c_ref, int net, int l, const LayerIn& in_ref, float* __restrict__ u_out) {
  const Ctx c = c_ref;                 // register copy of the kernel context (SmemFixed::ctx, shared memory)
  const LayerIn in = in_ref;
  RAAE_SMEM();
  StageTimer timer_(&sm->prof[kStFwdHidden]);
  float rm_old, rv_old;                // running BatchNorm buffers of this layer, loaded a whole stage ahead of their update
  bn_running_load(c, net, l, rm_old, rv_old);
  const raae_net_layout& nl = NL(c, net);
  const int tid = threadIdx.x, ty = tid >> 4, tx = tid & 15, c4 = tx * 4, warp = tid >> 5, lane = tid & 31;
  const float* Wg = netp(c, net) + nl.w_off[l];
  float* A0 = arena;                               // operand buffer b: hi at A0 + b * 2 * kATileFloats, lo right after
  float* Whi = A0 + 4 * tc::kATileFloats;          // [2 K blocks][64 rows][32] swizzled
  float* Wlo = Whi + tc::kBTileFloats;
  float* Raw = Wlo + tc::kBTileFloats;             // raw prefetch tile [kTM][64] (every thread re-reads only what it copied)
  float* slope_in = sm->cg;
  const float* src = in.src;
  const MaskSrc mk = in.mask;
  const int B = c.B, train = c.train, ntiles = (B + kTM - 1) / kTM;
  const uint32_t d_tmem = sm->tmem_base;
  uint64_t* mbar0 = reinterpret_cast<uint64_t*>(&sm->mbar);
  uint64_t* mbar1 = reinterpret_cast<uint64_t*>(&sm->mbar2);
  __syncthreads();
  auto prefetch_raw = [&](int t) {
    const int row0 = t * kTM, nv = min(kTM, B - row0);
#pragma unroll
    for (int i = 0; i < kTM / 16; ++i) {
      const int r = ty + 16 * i;
      if (r < nv) cp_async16(Raw + r * kH + c4, src + (size_t)(row0 + r) * kH + c4);
    }
    cp_async_commit();
  };
  const int t_first = c.crank, tstep = c.csize;      // this CTA's tiles: t_first, t_first + tstep, ... (cluster per trial)
  RAAE_PROBE_INIT();
  if (has_tiles(c, ntiles)) prefetch_raw(t_first);
  {
    // W_l [64 n][64 k] -> hi / lo, K-major SWIZZLE_128B
    float4 w[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) w[i] = *reinterpret_cast<const float4*>(Wg + (size_t)(ty + 16 * i) * kH + c4);
#pragma unroll
    for (int i = 0; i < 4; ++i) tc::split_store(Whi, Wlo, tc::sw128_chunk_off(ty + 16 * i, c4, tc::kBBlockBytes), w[i]);
  }
  if (tid < kH) {
    sm->bias[tid] = netp(c, net)[nl.b_off[l] + tid];
    sm->slope[tid] = netp(c, net)[nl.a_off[l] + tid];
    slope_in[tid] = in.slope[tid];
    if (!train) {
      sm->mean[net][l][tid] = c.st[nl.rm_off[l] + tid];
      sm->inv[net][l][tid] = 1.f / sqrtf(c.st[nl.rv_off[l] + tid] + kBnEps);
    }
  }
  __syncthreads();
  const float4 mu = *reinterpret_cast<const float4*>(sm->mean[in.snet][in.slayer] + c4);
  const float4 is = *reinterpret_cast<const float4*>(sm->inv[in.snet][in.slayer] + c4);
  const float4 sl = *reinterpret_cast<const float4*>(slope_in + c4);
  const uint32_t offK = tc::sw128_chunk_off(ty, c4, tc::kABlockBytes);
  uint32_t ph0 = sm->tc_phase, ph1 = sm->tc_phase2;
  // transform the raw tile t (own elements) into the hi / lo operands of buffer it & 1 (it = index among this CTA's tiles),
  // then start its MMAs
  auto stage = [&](int t, int it) {
    const int row0 = t * kTM, nv = min(kTM, B - row0);
    float* Ahi = A0 + (it & 1) * 2 * tc::kATileFloats;
    float* Alo = Ahi + tc::kATileFloats;
    cp_async_wait<0>();
    float4 uu[kTM / 16];
#pragma unroll
    for (int i = 0; i < kTM / 16; ++i) uu[i] = *reinterpret_cast<const float4*>(Raw + (ty + 16 * i) * kH + c4);
    if (t + tstep < ntiles) prefetch_raw(t + tstep);   // overwrites only this thread's own (already read) elements
    const uint32_t kbits = mask_keep4_rows(mk, row0, ty, c4, nv);
#pragma unroll
    for (int i = 0; i < kTM / 16; ++i) {
      const uint32_t kb = kbits >> (4 * i);
      float4 o;
      o.x = (kb & 1u) ? (prelu_f(uu[i].x, sl.x) - mu.x) * is.x * mk.scale : 0.f;
      o.y = (kb & 2u) ? (prelu_f(uu[i].y, sl.y) - mu.y) * is.y * mk.scale : 0.f;
      o.z = (kb & 4u) ? (prelu_f(uu[i].z, sl.z) - mu.z) * is.z * mk.scale : 0.f;
      o.w = (kb & 8u) ? (prelu_f(uu[i].w, sl.w) - mu.w) * is.w * mk.scale : 0.f;
      tc::split_store(Ahi, Alo, offK + (uint32_t)(i * 16 * 128), o);
    }
  };
  auto issue = [&](int it) {             // after a barrier that follows stage(.., it) and every TMEM read of item it-2
    if (tc::warp_uniform_id() == 0 && tc::elect_one()) {
      float* Ahi = A0 + (it & 1) * 2 * tc::kATileFloats;
      tc::fence_after_sync();
      tc::issue_gemm_3xtf32(d_tmem + (uint32_t)(64 * (it & 1)), Ahi, Ahi + tc::kATileFloats, Whi, Wlo);
      tc::mma_commit((it & 1) ? mbar1 : mbar0);
    }
  };
  float s1[32], s2[32];
#pragma unroll
  for (int j = 0; j < 32; ++j) { s1[j] = 0.f; s2[j] = 0.f; }
  const int erow = 32 * (warp & 3) + lane, ecol0 = 32 * (warp >> 2);
  RAAE_FPROBE(16);                       // weights staged, constants loaded
  if (has_tiles(c, ntiles)) stage(t_first, 0);
  tc::fence_async_smem();                // generic-proxy writes -> visible to the tensor core (async proxy)
  __syncthreads();
  RAAE_FPROBE(17);                       // first tile landed and staged
  if (has_tiles(c, ntiles)) issue(0);
  for (int t = t_first; t < ntiles; t += tstep) {
    const int it = tile_iter(c, t);
    const int row0 = t * kTM, nv = min(kTM, B - row0);
    // operands of the next tile are staged while the MMAs of tile t run; the accumulator of tile t is read back BEFORE the
    // MMAs of the next tile are queued (a tcgen05.ld issued behind a queued MMA batch waits for it), and the epilogue
    // arithmetic then overlaps them
    if (t + tstep < ntiles) stage(t + tstep, it + 1);
    if (it & 1) { tc::mbar_wait(mbar1, ph1); ph1 ^= 1u; }
    else        { tc::mbar_wait(mbar0, ph0); ph0 ^= 1u; }
    tc::fence_after_sync();
    float v[32];
    tc::tmem_ld32(d_tmem + ((uint32_t)(32 * (warp & 3)) << 16) + (uint32_t)(64 * (it & 1) + ecol0), v);
    tc::fence_before_sync();
    tc::fence_async_smem();              // late: the staging stores have drained behind the mbarrier wait and the TMEM load
    __syncthreads();
    RAAE_FPROBE(18);                     // MMAs + accumulator read-back (+ staging of the next tile)
    if (t + tstep < ntiles) issue(it + 1);
#pragma unroll
    for (int j = 0; j < 32; j += 4) {
      const float4 bb = *reinterpret_cast<const float4*>(sm->bias + ecol0 + j);
      v[j] += bb.x; v[j + 1] += bb.y; v[j + 2] += bb.z; v[j + 3] += bb.w;
    }
    if (erow < nv) {
      float* urow = u_out + (size_t)(row0 + erow) * kH + ecol0;
#pragma unroll
      for (int j = 0; j < 32; j += 4) *reinterpret_cast<float4*>(urow + j) = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
    }
    if (train) {
      if (it == 0) {
        // shift of the single-pass variance: column means of PReLU(u) over the first tile (operand buffer 0 is free:
        // its MMAs have completed; buffer 1 may be in use by tile 1)
        float pv[32];
#pragma unroll
        for (int j = 0; j < 32; ++j) pv[j] = erow < nv ? prelu_f(v[j], sm->slope[ecol0 + j]) : 0.f;
        sm->red[warp & 3][ecol0 + lane] = warp_colsum32(pv);
        __syncthreads();
        if (tid < kH) sm->shift[tid] = (sm->red[0][tid] + sm->red[1][tid] + sm->red[2][tid] + sm->red[3][tid]) / (float)nv;
        __syncthreads();
      }
      if (erow < nv) {
#pragma unroll
        for (int j = 0; j < 32; j += 4) {
          const float4 a4 = *reinterpret_cast<const float4*>(sm->slope + ecol0 + j);
          const float4 sh = *reinterpret_cast<const float4*>(sm->shift + ecol0 + j);
          float d;
          d = prelu_f(v[j], a4.x) - sh.x;     s1[j] += d;     s2[j] = fmaf(d, d, s2[j]);
          d = prelu_f(v[j + 1], a4.y) - sh.y; s1[j + 1] += d; s2[j + 1] = fmaf(d, d, s2[j + 1]);
          d = prelu_f(v[j + 2], a4.z) - sh.z; s1[j + 2] += d; s2[j + 2] = fmaf(d, d, s2[j + 2]);
          d = prelu_f(v[j + 3], a4.w) - sh.w; s1[j + 3] += d; s2[j + 3] = fmaf(d, d, s2[j + 3]);
        }
      }
    }
  }
  RAAE_FPROBE(19);                       // epilogues (stores, shift, partial sums)
  tc::fence_before_sync();
  __syncthreads();
  if (tid == 0) { sm->tc_phase = ph0; sm->tc_phase2 = ph1; }
  if (train) {
    // column sums of the per-thread (row, 32-column) partials: warp-level transposing reductions, then the four row
    // quarters of every column block are added
    const float r1 = warp_colsum32(s1);
    const float r2 = warp_colsum32(s2);
    sm->red[warp & 3][ecol0 + lane] = r1;
    sm->red[4 + (warp & 3)][ecol0 + lane] = r2;
    __syncthreads();
    float a1 = 0.f, a2 = 0.f;
    if (tid < kH) {
      a1 = sm->red[0][tid] + sm->red[1][tid] + sm->red[2][tid] + sm->red[3][tid];
      a2 = sm->red[4][tid] + sm->red[5][tid] + sm->red[6][tid] + sm->red[7][tid];
    }
    RAAE_FPROBE(20);                     // column reductions
    bn_stats_finish(c, sm, net, l, tid < kH ? sm->shift[tid] : 0.f, a1, a2, rm_old, rv_old);
    RAAE_FPROBE(21);                     // statistics (cluster: exchange + merge)
  }
  __syncthreads();
}

// Forward of the input block of the encoder on the noised batch, fed from the raw K-major operand image
// (ScratchLayout::xk) that build_batch wrote.  Warp 0 drives a two-deep pipeline of bulk asynchronous copies (raw A chunk
// 32 KB + weight chunk hi/lo 32 KB per 64 input columns) and the 24 MMAs of every chunk; warps 1..3 split every raw
// A chunk into its rounded hi (in place) and lo planes in shared memory;
// warps 4..7 read the finished 128 x 64 accumulators back (two TMEM accumulators, so the MMAs of tile t+1 overlap the
// epilogue of tile t), add the bias, store u and keep the BatchNorm sums per warp with warp-local shifts that are
// merged exactly at the end.  HBM traffic: the batch once (the hi/lo split never leaves the SM).
__device__ __noinline__ void fwd_wide_img(const Ctx& c_ref, int net, int l, const float* __restrict__ xk, const float* __restrict__ xref,
                                          float* __restrict__ u_out) {
  const Ctx c = c_ref;                 // register copy of the kernel context (SmemFixed::ctx, shared memory)
  RAAE_SMEM();
  StageTimer timer_(&sm->prof[kStFwdWide]);
  float rm_old, rv_old;                // running BatchNorm buffers of this layer, loaded a whole stage ahead of their update
  bn_running_load(c, net, l, rm_old, rv_old);
  const raae_net_layout& nl = NL(c, net);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int K = nl.in_dim[l], nch = c.p->sl.nch64;
  const float* Wg = netp(c, net) + nl.w_off[l];
  float* wk = c.sc + c.p->sl.wk;
  // Loop order: groups of up to 8 row tiles; inside a group the 64-column chunks are the OUTER loop and the tiles the
  // inner one, with the accumulators of all tiles of the group resident in TMEM (8 x 64 = all 512 columns).  The weight
  // chunk is then fetched once per chunk instead of once per (tile, chunk) - a bulk copy costs ~850 cycles of serialised
  // service per SM whatever its size up to 32 KB (tools/bulk_probe.cu), so halving their number matters - and the A
  // stream is the only per-item copy.
  float* Araw = arena;                       // 2 x [8192] raw chunk, rounded in place (= hi operand)
  float* Alo = arena + 2 * 8192;             // 2 x [8192] lo plane
  float* Bbuf = arena + 4 * 8192;            // 2 x [hi 4096 | lo 4096] weight chunk (double-buffered over the chunks)
  // cluster per trial: this CTA walks ITS tiles (local index lt -> tile crank + lt * csize)
  const int B = c.B, train = c.train, ntiles = cl::own_tiles(B, c.crank, c.csize), ngroups = (ntiles + 7) / 8;
  const int crank = c.crank, csize = c.csize;
  const uint32_t d_tmem = sm->tmem_base;
  uint64_t* full = reinterpret_cast<uint64_t*>(&sm->pipe_bar[0]);      // [2] A chunk landed
  uint64_t* empty = reinterpret_cast<uint64_t*>(&sm->pipe_bar[2]);     // [2] MMAs of the A buffer completed
  uint64_t* conv = reinterpret_cast<uint64_t*>(&sm->pipe_bar[4]);      // [2] hi / lo planes of the A buffer written
  uint64_t* wfull = reinterpret_cast<uint64_t*>(&sm->pipe_bar[6]);     // [2] weight chunk landed
  uint64_t* accfull = reinterpret_cast<uint64_t*>(&sm->pipe_bar[8]);   // accumulators of the group complete
  uint64_t* accfree = reinterpret_cast<uint64_t*>(&sm->pipe_bar[9]);   // accumulators of the group read back (4 warps arrive)
  __syncthreads();
  if (tid == 0) {
    for (int i = 0; i < 9; ++i) tc::mbar_init(reinterpret_cast<uint64_t*>(&sm->pipe_bar[i]), 1);
    tc::mbar_init(accfree, 8);
  }
  // K-major hi / lo image of W_l [64][K] in global scratch, one [hi 4096 | lo 4096] block per 64 input columns
  for (int i = tid; i < kH * nch * 16; i += kThreads) {
    const int n = i / (nch * 16), k4 = (i - n * (nch * 16)) * 4;
    const float4 w = k4 < K ? *reinterpret_cast<const float4*>(Wg + (size_t)n * K + k4) : make_float4(0.f, 0.f, 0.f, 0.f);
    float* blk = wk + (size_t)(k4 >> 6) * 8192;
    tc::split_store(blk, blk + 4096, tc::sw128_chunk_off(n, k4 & 63, tc::kBBlockBytes), w);
  }
  {
    // effective bias: b + W xref (the images are centred on xref); 4 threads per output channel, float64 partial sums
    // (a float32 dot product here costs the deep-stack parity case its margin: the constant feeds PReLU before BatchNorm)
    const int n = tid >> 2, part = tid & 3;
    double acc = 0.0;
    for (int k = part * 4; k < K; k += 16) {
      const float4 w = *reinterpret_cast<const float4*>(Wg + (size_t)n * K + k);
      const float4 xr = *reinterpret_cast<const float4*>(xref + k);
      acc += (double)w.x * xr.x + (double)w.y * xr.y + (double)w.z * xr.z + (double)w.w * xr.w;
    }
    acc += __shfl_xor_sync(0xffffffffu, acc, 1);
    acc += __shfl_xor_sync(0xffffffffu, acc, 2);
    if (part == 0) sm->bias[n] = (float)((double)netp(c, net)[nl.b_off[l] + n] + acc);
  }
  if (tid < kH) {
    sm->slope[tid] = netp(c, net)[nl.a_off[l] + tid];
    if (!train) {
      sm->mean[net][l][tid] = c.st[nl.rm_off[l] + tid];
      sm->inv[net][l][tid] = 1.f / sqrtf(c.st[nl.rv_off[l] + tid] + kBnEps);
    }
  }
  tc::fence_async_all();
  __syncthreads();
  // Roles per group of tiles: warp 0 = producer / issuer (elected lane), warps 1..7 = converters (224 threads); afterwards
  // all 8 warps read the group's accumulators back.  Items are counted in execution order: for g, for ck, for tl < T_g.
  const int warp_u = tc::warp_uniform_id();
  const int erow = 32 * (warp & 3) + lane, eh = warp >> 2;                 // epilogue ownership: TMEM lane, column half
  float shv[32], s1v[32], s2v[32];                                          // this thread's row x 32 columns: shift, partial sums
#pragma unroll
  for (int j = 0; j < 32; ++j) { shv[j] = 0.f; s1v[j] = 0.f; s2v[j] = 0.f; }
  float sh = 0.f;                                                           // lane j: shift of column 32 eh + j
  int nrows_w = 0;                                                          // rows this warp has accumulated
  // driver state (meaningful in the elected lane of warp 0 only; elect.sync picks the same lane every time)
  int it = 0, wit = 0, wloaded = 0;
  int lg = 0, lck = 0, ltl = 0, lit = 0;                                    // next A item to load
  int cit = 0;                                                              // converters: next item
  for (int g = 0; g < ngroups; ++g) {
    const int Tg = min(8, ntiles - 8 * g);
    if (warp_u == 0) {
      if (tc::elect_one()) {
        auto load_next = [&]() {
          if (lg >= ngroups) return;
          const int b = lit & 1, tile = crank + csize * (8 * lg + ltl);
          tc::mbar_expect_tx(&full[b], 32768u);
          tc::bulk_g2s(Araw + b * 8192, xk + ((size_t)tile * nch + lck) * 8192, 32768u, &full[b]);
          ++lit;
          const int Tl = min(8, ntiles - 8 * lg);
          if (++ltl == Tl) { ltl = 0; if (++lck == nch) { lck = 0; ++lg; } }
        };
        auto load_w = [&](int ck, int n) {                    // n-th weight chunk load overall -> buffer n & 1
          tc::mbar_expect_tx(&wfull[n & 1], 32768u);
          tc::bulk_g2s(Bbuf + (n & 1) * 8192, wk + (size_t)ck * 8192, 32768u, &wfull[n & 1]);
        };
        if (g == 0) {
          load_next();
          load_next();
          load_w(0, wloaded++);
        } else {
          tc::mbar_wait(accfree, (uint32_t)((g - 1) & 1));                       // the previous group has been read back
        }
        for (int ck = 0; ck < nch; ++ck, ++wit) {
          // prefetch the next weight chunk into the other buffer, whose previous user (chunk wit - 1) is complete when
          // the MMAs of its last item (item it - 1) are
          const bool more_w = (ck + 1 < nch) || (g + 1 < ngroups);
          if (more_w) {
            if (it >= 1) tc::mbar_wait(&empty[(it - 1) & 1], (uint32_t)(((it - 1) >> 1) & 1));
            load_w(ck + 1 < nch ? ck + 1 : 0, wloaded++);
          }
          tc::mbar_wait(&wfull[wit & 1], (uint32_t)((wit >> 1) & 1));
          const float* Bh = Bbuf + (wit & 1) * 8192;
          for (int tl = 0; tl < Tg; ++tl, ++it) {
            const int b = it & 1;
            tc::mbar_wait(&conv[b], (uint32_t)((it >> 1) & 1));                    // raw landed and split
            tc::fence_after_sync();
            tc::issue_gemm_3xtf32_acc(d_tmem + (uint32_t)(64 * tl), Araw + b * 8192, Alo + b * 8192, Bh, Bh + 4096,
                                      ck > 0 ? 1u : 0u);
            tc::mma_commit(&empty[b]);
            if (lit < it + 3) {                                                     // keep two items in flight
              tc::mbar_wait(&empty[b], (uint32_t)((it >> 1) & 1));
              load_next();
            }
          }
        }
        tc::mma_commit(accfull);
      }
      __syncwarp();
    } else {
      // converters (224 threads): round-to-nearest hi / lo split, element for element in the swizzled layout
      const int ctid = tid - 32;
      for (int k0 = 0; k0 < Tg * nch; ++k0, ++cit) {
        const int b = cit & 1;
        tc::mbar_wait(&full[b], (uint32_t)((cit >> 1) & 1));
        float4* src = reinterpret_cast<float4*>(Araw + b * 8192);   // rounded in place: becomes the hi plane
        float4* dst = reinterpret_cast<float4*>(Alo + b * 8192);
        // 2048 quads over 224 threads: nine unguarded quads per thread (all nine loads in flight, straight-line split and
        // stores - a guard around the two stores of every quad compiled into a branch diamond per quad), the last 32 quads
        // by the first converter warp
        {
          float4 x[9];
#pragma unroll
          for (int k = 0; k < 9; ++k) x[k] = src[ctid + 224 * k];
#pragma unroll
          for (int k = 0; k < 9; ++k) {
            float4 hi, lo;
            tc::tf32_split(x[k].x, hi.x, lo.x);
            tc::tf32_split(x[k].y, hi.y, lo.y);
            tc::tf32_split(x[k].z, hi.z, lo.z);
            tc::tf32_split(x[k].w, hi.w, lo.w);
            src[ctid + 224 * k] = hi;
            dst[ctid + 224 * k] = lo;
          }
          if (ctid < 2048 - 9 * 224) {            // warp-uniform: the first converter warp
            const int q = 9 * 224 + ctid;
            const float4 xv = src[q];
            float4 hi, lo;
            tc::tf32_split(xv.x, hi.x, lo.x);
            tc::tf32_split(xv.y, hi.y, lo.y);
            tc::tf32_split(xv.z, hi.z, lo.z);
            tc::tf32_split(xv.w, hi.w, lo.w);
            src[q] = hi;
            dst[q] = lo;
          }
        }
        tc::fence_async_smem();
        asm volatile("bar.sync 2, 224;" ::: "memory");
        if (ctid == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(tc::smem_u32(&conv[b])) : "memory");
      }
    }
    // ---- read-back of the group's accumulators by all 8 warps: TMEM lane (row) erow, columns 32 eh .. + 31 of every tile ----
    tc::mbar_wait(accfull, (uint32_t)(g & 1));
    tc::fence_after_sync();
    for (int tl = 0; tl < Tg; ++tl) {
      const int t = crank + csize * (8 * g + tl);
      const int row0 = t * kTM, nv = min(kTM, B - row0);
      const int nvw = max(0, min(32, nv - 32 * (warp & 3)));                // valid rows of this warp in the tile
      float v[32];
      tc::tmem_ld32(d_tmem + ((uint32_t)(32 * (warp & 3)) << 16) + (uint32_t)(64 * tl + 32 * eh), v);
#pragma unroll
      for (int j = 0; j < 32; j += 4) {
        const float4 bb = *reinterpret_cast<const float4*>(sm->bias + 32 * eh + j);
        v[j] += bb.x; v[j + 1] += bb.y; v[j + 2] += bb.z; v[j + 3] += bb.w;
      }
      if (erow < nv) {
        float* urow = u_out + (size_t)(row0 + erow) * kH + 32 * eh;
#pragma unroll
        for (int j = 0; j < 32; j += 4) *reinterpret_cast<float4*>(urow + j) = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
      }
      if (train && nvw > 0) {
        float pv[32];
#pragma unroll
        for (int j = 0; j < 32; ++j) pv[j] = erow < nv ? prelu_f(v[j], sm->slope[32 * eh + j]) : 0.f;
        if (nrows_w == 0) {
          // warp-local shift: mean of the warp's first rows, broadcast into every thread's registers
          float qv[32];
#pragma unroll
          for (int j = 0; j < 32; ++j) qv[j] = pv[j];
          sh = warp_colsum32(qv) / (float)nvw;
#pragma unroll
          for (int j = 0; j < 32; ++j) shv[j] = __shfl_sync(0xffffffffu, sh, j);
        }
        if (erow < nv) {
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            const float d = pv[j] - shv[j];
            s1v[j] += d;
            s2v[j] = fmaf(d, d, s2v[j]);
          }
        }
      }
      if (train) nrows_w += nvw;
    }
    tc::fence_before_sync();
    __syncwarp();
    if (lane == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(tc::smem_u32(accfree)) : "memory");
  }
  if (train) {
    // per-warp (count, mean, M2) of its columns -> sm->red rows 0..7; the column reductions run once per stage
    const float s1 = warp_colsum32(s1v), s2 = warp_colsum32(s2v);
    const float nw = (float)nrows_w;
    const float mean_w = nrows_w > 0 ? sh + s1 / nw : 0.f;
    const float m2_w = nrows_w > 0 ? fmaxf(s2 - s1 * s1 / nw, 0.f) : 0.f;
    sm->red[warp & 3][32 * eh + lane] = mean_w;
    sm->red[4 + (warp & 3)][32 * eh + lane] = m2_w;
    if (lane == 0 && eh == 0) sm->redw[warp & 3] = nw;
  }
  __syncthreads();
  if (train) {
    // exact merge of the four row groups (parallel variance): mean = sum n_w mean_w / n, M2 = sum M2_w + n_w (mean_w - mean)^2
    float mean = 0.f, m2 = 0.f;
    if (tid < kH) {
      const float n = csize == 1 ? (float)B : sm->redw[0] + sm->redw[1] + sm->redw[2] + sm->redw[3];   // rows of this CTA
      if (n > 0.f) {
#pragma unroll
        for (int w = 0; w < 4; ++w) mean += sm->redw[w] * sm->red[w][tid];
        mean /= n;
#pragma unroll
        for (int w = 0; w < 4; ++w) { const float d = sm->red[w][tid] - mean; m2 += sm->red[4 + w][tid] + sm->redw[w] * d * d; }
      }
    }
    if (csize == 1) { if (tid < kH) bn_finalize(c, sm, net, l, tid, mean, 0.f, m2, B, true, rm_old, rv_old); }
    else bn_merge_cluster(c, sm, net, l, mean, m2, rm_old, rv_old);
  }
  __syncthreads();
}

// statistics of nstyle columns of a [rows][kZ] panel (two-pass); results in sm->zs[0] (mean), zs[1] (biased var).
// Cluster per trial: every CTA sums over ITS rows (slot -> row, cl::slot_row) and the sums are all-reduced.
__device__ __forceinline__ void latent_colstats(const Ctx& c, SmemFixed* sm, const float* __restrict__ z, int nrows) {
  const int tid = threadIdx.x, k = tid & 7, g = tid >> 3;
  const int crank = c.crank, csize = c.csize;
  const int nslots = csize == 1 ? nrows : cl::own_tiles(nrows, crank, csize) * kTM;
  float s = 0.f;
  for (int r0 = g; r0 < nslots; r0 += 32 * 8) {            // 8 loads in flight per thread
    float zv[8];
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      const int sl = r0 + 32 * u, r = csize == 1 ? sl : cl::slot_row(sl, crank, csize);
      zv[u] = (sl < nslots && r < nrows) ? z[(size_t)r * kZ + k] : 0.f;
    }
#pragma unroll
    for (int u = 0; u < 8; ++u) s += zv[u];
  }
  float* red = &sm->red[0][0];
  __syncthreads();
  red[tid] = s;
  __syncthreads();
  if (tid < kZ) {
    float t = 0.f;
    for (int i = 0; i < 32; ++i) t += red[i * 8 + tid];
    sm->zs[0][tid] = t;
  }
  cluster_allreduce_f(c, sm, sm->zs[0], kZ);
  __syncthreads();
  if (tid < kZ) sm->zs[0][tid] = sm->zs[0][tid] / (float)nrows;
  __syncthreads();
  const float mu = sm->zs[0][k];
  s = 0.f;
  for (int r0 = g; r0 < nslots; r0 += 32 * 8) {
    float zv[8];
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      const int sl = r0 + 32 * u, r = csize == 1 ? sl : cl::slot_row(sl, crank, csize);
      zv[u] = (sl < nslots && r < nrows) ? z[(size_t)r * kZ + k] : mu;
    }
#pragma unroll
    for (int u = 0; u < 8; ++u) { const float d = zv[u] - mu; s = fmaf(d, d, s); }
  }
  __syncthreads();
  red[tid] = s;
  __syncthreads();
  if (tid < kZ) {
    float t = 0.f;
    for (int i = 0; i < 32; ++i) t += red[i * 8 + tid];
    sm->zs[1][tid] = t;
  }
  cluster_allreduce_f(c, sm, sm->zs[1], kZ);
  __syncthreads();
  if (tid < kZ) sm->zs[1][tid] = sm->zs[1][tid] / (float)nrows;
  __syncthreads();
}

// last encoder Linear (64 -> nstyle) + BatchNorm1d(nstyle).  zE receives the PRE-BN output; the BN
// statistics go to sm->mean/inv[kE][L-1][0..nstyle).
__device__ __noinline__ void fwd_enc_last(const Ctx& c_ref, const LayerIn& in_ref) {
  const Ctx c = c_ref;                 // register copy of the kernel context (SmemFixed::ctx, shared memory)
  const LayerIn in = in_ref;
  RAAE_SMEM();
  StageTimer timer_(&sm->prof[kStFwdEncLast]);
  const raae_net_layout& nl = NL(c, kE);
  const int L = nl.n_linear, l = L - 1, ns = nl.out_dim[l];
  const int tid = threadIdx.x;
  float* Ws = arena;                   // [kZ][kLD]
  float* At = arena + kH * kLDW;       // [kTM][kLD]
  float* zE = c.sc + c.p->sl.zE;
  __syncthreads();
  for (int o = tid; o < kZ * kH; o += kThreads) {
    int n = o >> 6, k = o & 63;
    Ws[n * kLD + k] = n < ns ? netp(c, kE)[nl.w_off[l] + n * kH + k] : 0.f;
  }
  if (tid < kZ) sm->bias[tid] = tid < ns ? netp(c, kE)[nl.b_off[l] + tid] : 0.f;
  __syncthreads();
  const int ntiles = (c.B + kTM - 1) / kTM;
  for (int t = c.crank; t < ntiles; t += c.csize) {
    const int row0 = t * kTM, nv = min(kTM, c.B - row0);
    build_act_tile(At, in.src, row0, nv, sm->mean[in.snet][in.slayer], sm->inv[in.snet][in.slayer], in.slope, in.mask);
    __syncthreads();
    {
      // thread (row r = tid / 2, latents n0 = 4 (tid & 1) .. + 3): one pass over the activation row feeds four independent
      // accumulation chains (the chain order over k is the one of a single accumulator per output, as before); float4 store
      const int r = tid >> 1, n0 = (tid & 1) * 4;
      float acc[4] = {sm->bias[n0], sm->bias[n0 + 1], sm->bias[n0 + 2], sm->bias[n0 + 3]};
#pragma unroll 4
      for (int k = 0; k < kH; k += 4) {
        const float4 a = *reinterpret_cast<const float4*>(At + r * kLD + k);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const float4 w = *reinterpret_cast<const float4*>(Ws + (n0 + j) * kLD + k);
          acc[j] = fmaf(a.x, w.x, acc[j]); acc[j] = fmaf(a.y, w.y, acc[j]); acc[j] = fmaf(a.z, w.z, acc[j]); acc[j] = fmaf(a.w, w.w, acc[j]);
        }
      }
      if (r < nv)
        *reinterpret_cast<float4*>(zE + (size_t)(row0 + r) * kZ + n0) =
            make_float4(n0 < ns ? acc[0] : 0.f, n0 + 1 < ns ? acc[1] : 0.f, n0 + 2 < ns ? acc[2] : 0.f, n0 + 3 < ns ? acc[3] : 0.f);
    }
    __syncthreads();
  }
  if (c.train) {
    __threadfence_block();
    latent_colstats(c, sm, zE, c.B);
    if (tid < ns) {
      float mean = sm->zs[0][tid], var = sm->zs[1][tid], n = (float)c.B;
      sm->mean[kE][l][tid] = mean;
      sm->inv[kE][l][tid] = 1.f / sqrtf(var + kBnEps);
      if (c.crank == 0) {
        float* rm = c.st + nl.rm_off[l];
        float* rv = c.st + nl.rv_off[l];
        float unb = c.B > 1 ? var * (n / (n - 1.f)) : var;
        rm[tid] = (1.f - kBnMomentum) * rm[tid] + kBnMomentum * mean;
        rv[tid] = (1.f - kBnMomentum) * rv[tid] + kBnMomentum * unb;
      }
    }
  } else if (tid < ns) {
    sm->mean[kE][l][tid] = c.st[nl.rm_off[l] + tid];
    sm->inv[kE][l][tid] = 1.f / sqrtf(c.st[nl.rv_off[l] + tid] + kBnEps);
  }
  if (tid >= ns && tid < kZ) { sm->mean[kE][l][tid] = 0.f; sm->inv[kE][l][tid] = 0.f; }
  __syncthreads();
}

__device__ __forceinline__ void fwd_hidden(const Ctx& c, int net, int l, const LayerIn& in, float* __restrict__ u_out) {
  if (in.kind == kInHidden) {
    if (c.p->cfg.tensor_cores & 1) fwd_hidden64_tc(c, net, l, in, u_out);
    else fwd_hidden64(c, net, l, in, u_out);
  } else if (in.kind == kInWide && (c.p->cfg.tensor_cores & 4) && in.img) {
    fwd_wide_img(c, net, l, c.sc + (in.img == 2 ? c.p->sl.yk : c.p->sl.xk),
                 c.sc + (in.img == 2 ? c.p->sl.yref : c.p->sl.xref) + c.crank * kMaxDim, u_out);
  } else if (in.kind == kInLatent) {
    fwd_latent(c, net, l, in, u_out);
  } else {
    fwd_hidden_edge(c, net, l, in, u_out);
  }
}

// LayerIn describing "the activations coming out of hidden layer l of `net`, forward instance inst"
__device__ __forceinline__ LayerIn hidden_out(const Ctx& c, int net, int l, int inst) {
  LayerIn in;
  in.kind = kInHidden;
  in.src = c.sc + (net == kE ? c.p->sl.uE[l] : c.p->sl.uD[l]);
  in.ld = kH; in.dim = kH; in.act = 0; in.img = 0;
  in.snet = net; in.slayer = l;
  in.slope = netp(c, net) + NL(c, net).a_off[l];
  in.mask = make_mask(c, net, inst, l);
  return in;
}

__device__ __forceinline__ LayerIn wide_in(const float* src, int ld, int dim, int act) {
  LayerIn in;
  in.kind = kInWide; in.src = src; in.ld = ld; in.dim = dim; in.act = act; in.img = 0;
  in.snet = 0; in.slayer = -1; in.slope = nullptr;
  in.mask.ptr = nullptr; in.mask.key = 0; in.mask.thresh = 0; in.mask.scale = 1.f;
  return in;
}

// bn_layer >= 0: rows are the pre-BN encoder output, normalised with sm->mean/inv[kE][bn_layer]
__device__ __forceinline__ LayerIn latent_in(const float* src, int nstyle, int bn_layer) {
  LayerIn in;
  in.kind = kInLatent; in.src = src; in.ld = kZ; in.dim = nstyle; in.act = 0; in.img = 0;
  in.snet = kE; in.slayer = bn_layer; in.slope = nullptr;
  in.mask.ptr = nullptr; in.mask.key = 0; in.mask.thresh = 0; in.mask.scale = 1.f;
  return in;
}

// FCEncoder.forward (model.py:330-378): x -> zE (pre-BN) + BN statistics of every layer
__device__ __forceinline__ void encoder_forward(const Ctx& c, const LayerIn& x, int inst) {
  const raae_net_layout& nl = NL(c, kE);
  const int L = nl.n_linear;
  fwd_hidden(c, kE, 0, x, c.sc + c.p->sl.uE[0]);
  for (int l = 1; l < L - 1; ++l) fwd_hidden(c, kE, l, hidden_out(c, kE, l - 1, inst), c.sc + c.p->sl.uE[l]);
  fwd_enc_last(c, hidden_out(c, kE, L - 2, inst));
  if (c.train && threadIdx.x == 0 && c.crank == 0) c.st[nl.nbt_off] += 1.f;
}

// hidden blocks of FCDecoder.forward (model.py:518-570); the output Linear is a separate stage
__device__ __forceinline__ void decoder_forward_hidden(const Ctx& c, const LayerIn& z, int inst) {
  const raae_net_layout& nl = NL(c, kD);
  const int L = nl.n_linear;
  fwd_hidden(c, kD, 0, z, c.sc + c.p->sl.uD[0]);
  for (int l = 1; l < L - 1; ++l) fwd_hidden(c, kD, l, hidden_out(c, kD, l - 1, inst), c.sc + c.p->sl.uD[l]);
  if (c.train && threadIdx.x == 0 && c.crank == 0) c.st[nl.nbt_off] += 1.f;
}

// ------------------------------------------------------------------------------------------
// backward of one hidden block (PReLU -> BN -> dropout already folded into g_in), fused AdamW
// ------------------------------------------------------------------------------------------
// On entry sm->sg / sm->sgx hold sum_r g and sum_r g * xhat of THIS layer (g = dL/dxhat).
// Produces dW, db, dslope (-> AdamW / debug export) and, if g_out != null, the gradient w.r.t. the
// producing layer's xhat plus its sums in sm->sg / sm->sgx.
//   in.kind == kInHidden : input activations rebuilt from in.src;    g_out [rows][64]
//   in.kind == kInWide   : input rows in.src (K = dim);              dx_out (optional) [rows][ld]: receives
//                           dL/dx * act'(v) IN PLACE of the pre-activation stored there (MI phase)
//   in.kind == kInLatent : input latent rows;                        dz_out (optional) [rows][kZ]
__device__ __noinline__ void bwd_hidden_edge(const Ctx& c_ref, int net, int l, const LayerIn& in_ref, const float* __restrict__ u_l,
                                             const float* __restrict__ g_in, float* g_out, int o) {
  const Ctx c = c_ref;                 // register copy of the kernel context (SmemFixed::ctx, shared memory)
  const LayerIn in = in_ref;
  RAAE_SMEM();
  StageTimer timer_(&sm->prof[in.kind == kInWide ? kStBwdWide : in.kind == kInHidden ? kStBwdHidden : kStBwdLatent]);
  const raae_net_layout& nl = NL(c, net);
  const int tid = threadIdx.x, ty = tid >> 4, tx = tid & 15;
  const int K = nl.in_dim[l];
  const float* Wg = netp(c, net) + nl.w_off[l];
  // arena: [Dt kTile][At: kTile | wide kWideTile][Ws kWTile (hidden NN operand / wide chunk / latent [64][9])]
  float* Dt = arena;
  float* At = arena + kTile;
  float* Ws = At + (in.kind == kInWide ? kWideTile : kTile);
  float* gradW = At;                          // reused after the tile loop: dense [64][K]
  const bool want_out = g_out != nullptr;
  __syncthreads();
  if (tid < kH) {
    float nB = (float)c.B;
    sm->cg[tid] = sm->sg[tid] / nB;
    sm->cgx[tid] = sm->sgx[tid] / nB;
    sm->slope[tid] = netp(c, net)[nl.a_off[l] + tid];
  }
  if (in.kind == kInHidden && want_out) load_w_rows(Ws, kLD, Wg, K, 0, kH);
  if (in.kind == kInLatent)
    for (int i = tid; i < kH * 9; i += kThreads) {
      int n = i / 9, k = i - n * 9;
      Ws[i] = k < K ? Wg[n * K + k] : 0.f;
    }
  __syncthreads();
  const int c4 = tx * 4;
  const float4 mu = *reinterpret_cast<const float4*>(sm->mean[net][l] + c4);
  const float4 is = *reinterpret_cast<const float4*>(sm->inv[net][l] + c4);
  const float4 sl = *reinterpret_cast<const float4*>(sm->slope + c4);
  const float4 cg = *reinterpret_cast<const float4*>(sm->cg + c4);
  const float4 cgx = *reinterpret_cast<const float4*>(sm->cgx + c4);
  float db4[4] = {0.f, 0.f, 0.f, 0.f}, ds4[4] = {0.f, 0.f, 0.f, 0.f};
  float sg4[4] = {0.f, 0.f, 0.f, 0.f}, sgx4[4] = {0.f, 0.f, 0.f, 0.f};
  float accW[8][8];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j) accW[i][j] = 0.f;
  float accS[kZ] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};   // latent dW: (ch, row group) partials of dW[ch][:]
  const int ntiles = (c.B + kTM - 1) / kTM;
  float* const sg_dst = c.csize > 1 ? sm->sgp : sm->sg;         // cluster: partial sums, gathered after the barrier
  float* const sgx_dst = c.csize > 1 ? sm->sgxp : sm->sgx;
  for (int t = c.crank; t < ntiles; t += c.csize) {
    const int row0 = t * kTM, nv = min(kTM, c.B - row0);
    // 1. du = PReLU'(u) * BN'(g); loads batched 4 rows (8 float4) at a time
#pragma unroll
    for (int b = 0; b < kTM / 16; b += 4) {
      float4 gg[4], uu[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int r = ty + 16 * (b + i);
        if (r < nv) {
          gg[i] = *reinterpret_cast<const float4*>(g_in + (size_t)(row0 + r) * kH + c4);
          uu[i] = *reinterpret_cast<const float4*>(u_l + (size_t)(row0 + r) * kH + c4);
        } else {
          gg[i] = make_float4(0.f, 0.f, 0.f, 0.f);
          uu[i] = make_float4(0.f, 0.f, 0.f, 0.f);
        }
      }
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int r = ty + 16 * (b + i);
        float4 du = make_float4(0.f, 0.f, 0.f, 0.f);
        if (r < nv) {
          const float4 g = gg[i], u = uu[i];
#define RAAE_DU(comp, idx)                                                        \
          {                                                                       \
            float xh = (prelu_f(u.comp, sl.comp) - mu.comp) * is.comp;            \
            float dh = (g.comp - cg.comp - xh * cgx.comp) * is.comp;              \
            bool pos = u.comp > 0.f;                                              \
            du.comp = pos ? dh : sl.comp * dh;                                    \
            ds4[idx] += pos ? 0.f : u.comp * dh;                                  \
            db4[idx] += du.comp;                                                  \
          }
          RAAE_DU(x, 0) RAAE_DU(y, 1) RAAE_DU(z, 2) RAAE_DU(w, 3)
#undef RAAE_DU
        }
        *reinterpret_cast<float4*>(Dt + r * kLD + c4) = du;
      }
    }
    // 2. the layer's input activations
    if (in.kind == kInHidden) build_act_tile(At, in.src, row0, nv, sm->mean[in.snet][in.slayer], sm->inv[in.snet][in.slayer], in.slope, in.mask);
    else if (in.kind == kInWide) build_wide_tile(At, in.src, in.ld, in.dim, row0, nv, in.act);
    else build_latent_tile(At, in.src, row0, nv, K, in.slayer >= 0 ? sm->mean[in.snet][in.slayer] : nullptr,
                        in.slayer >= 0 ? sm->inv[in.snet][in.slayer] : nullptr);
    __syncthreads();
    // 3. dW += du^T a
    if (in.kind == kInHidden) {
      const int qq = tid >> 6, tt = tid & 63;
      mma_tn8(Dt, kLD, 8 * (tt >> 3), At, kLD, 8 * (tt & 7), 32 * qq, 32 * qq + 32, accW);
    } else if (in.kind == kInWide) {
      mma_tn8(Dt, kLD, 8 * (tid >> 5), At, kLDW, 8 * (tid & 31), 0, kTM, accW);
    } else {
      // dW[ch][k] over rows q, q + 4, ...: one du load and one broadcast latent row per 8 FMAs
      const int ch = tid & 63, q = tid >> 6;
#pragma unroll 4
      for (int i = 0; i < kTM / 4; ++i) {
        const int r = q + 4 * i;
        const float d = Dt[r * kLD + ch];
        const float4 z0 = *reinterpret_cast<const float4*>(At + r * kZ);
        const float4 z1 = *reinterpret_cast<const float4*>(At + r * kZ + 4);
        accS[0] = fmaf(d, z0.x, accS[0]); accS[1] = fmaf(d, z0.y, accS[1]); accS[2] = fmaf(d, z0.z, accS[2]); accS[3] = fmaf(d, z0.w, accS[3]);
        accS[4] = fmaf(d, z1.x, accS[4]); accS[5] = fmaf(d, z1.y, accS[5]); accS[6] = fmaf(d, z1.z, accS[6]); accS[7] = fmaf(d, z1.w, accS[7]);
      }
    }
    // 4. gradient w.r.t. the input
    if (want_out && in.kind == kInHidden) {
      float acc[8][4];
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
      mma_nn<kH>(Dt, kLD, Ws, kLD, acc, ty, tx);
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        int r = ty + 16 * i;
        if (r < nv) {
          uint32_t kb = mask_keep4(in.mask, row0 + r, c4);
          float4 a = *reinterpret_cast<const float4*>(At + r * kLD + c4);
          float4 gm;
          gm.x = (kb & 1u) ? acc[i][0] * in.mask.scale : 0.f;
          gm.y = (kb & 2u) ? acc[i][1] * in.mask.scale : 0.f;
          gm.z = (kb & 4u) ? acc[i][2] * in.mask.scale : 0.f;
          gm.w = (kb & 8u) ? acc[i][3] * in.mask.scale : 0.f;
          sg4[0] += gm.x; sg4[1] += gm.y; sg4[2] += gm.z; sg4[3] += gm.w;
          sgx4[0] = fmaf(acc[i][0], a.x, sgx4[0]); sgx4[1] = fmaf(acc[i][1], a.y, sgx4[1]);
          sgx4[2] = fmaf(acc[i][2], a.z, sgx4[2]); sgx4[3] = fmaf(acc[i][3], a.w, sgx4[3]);
          *reinterpret_cast<float4*>(g_out + (size_t)(row0 + r) * kH + c4) = gm;
        }
      }
    } else if (want_out && in.kind == kInWide) {
      // dL/dx chunk by chunk, multiplied by act'(v) and written over v (g_out == the v panel, stride in.ld).  The weight
      // chunks are double-buffered with cp.async; act'(v) comes from y = act(v), which the input tile holds:
      // softplus_2'(v) = sigmoid(2 v) = 1 - exp(-2 y), relu'(v) = [y > 0] - so v is not read back.
      auto prefetch_wchunk = [&](float* dst, int k0) {
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const int e = tid + kThreads * u, n = e >> 4, k4 = (e & 15) * 4;
          if (k0 + k4 < K) cp_async16(dst + n * kLD + k4, Wg + (size_t)n * K + k0 + k4);
          else *reinterpret_cast<float4*>(dst + n * kLD + k4) = make_float4(0.f, 0.f, 0.f, 0.f);
        }
        cp_async_commit();
      };
      float* Ws2 = Ws + kWTile;
      __syncthreads();                       // every warp is done with the weight buffers of the previous tile
      prefetch_wchunk(Ws, 0);
      for (int k0 = 0, ci = 0; k0 < in.dim; k0 += kH, ++ci) {
        float* Wcur = (ci & 1) ? Ws2 : Ws;
        cp_async_wait<0>();
        __syncthreads();
        if (k0 + kH < in.dim) prefetch_wchunk((ci & 1) ? Ws : Ws2, k0 + kH);
        float acc[8][4];
#pragma unroll
        for (int i = 0; i < 8; ++i)
#pragma unroll
          for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
        mma_nn<kH>(Dt, kLD, Wcur, kLD, acc, ty, tx);
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          int r = ty + 16 * i;
          if (r < nv && k0 + c4 < in.dim) {
            const float4 y = *reinterpret_cast<const float4*>(At + r * kLDW + k0 + c4);
            float4 v;
            if (in.act == 1) {
              v.x = acc[i][0] * -expm1f(-2.f * y.x); v.y = acc[i][1] * -expm1f(-2.f * y.y);
              v.z = acc[i][2] * -expm1f(-2.f * y.z); v.w = acc[i][3] * -expm1f(-2.f * y.w);
            } else {
              v.x = y.x > 0.f ? acc[i][0] : 0.f; v.y = y.y > 0.f ? acc[i][1] : 0.f;
              v.z = y.z > 0.f ? acc[i][2] : 0.f; v.w = y.w > 0.f ? acc[i][3] : 0.f;
            }
            *reinterpret_cast<float4*>(g_out + (size_t)(row0 + r) * in.ld + k0 + c4) = v;
          }
        }
      }
    } else if (want_out && in.kind == kInLatent) {
      // dz[r][k] = sum_n du[r][n] W[n][k]: two threads per row (32 channels each), partial sums combined by shuffle
      const int r = tid >> 1, half = tid & 1;
      float s8[kZ];
#pragma unroll
      for (int k = 0; k < kZ; ++k) s8[k] = 0.f;
#pragma unroll 2
      for (int n4 = 0; n4 < 32; n4 += 4) {
        const float4 d4 = *reinterpret_cast<const float4*>(Dt + r * kLD + 32 * half + n4);
        const float dv[4] = {d4.x, d4.y, d4.z, d4.w};
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const float* wrow = Ws + (32 * half + n4 + e) * 9;
#pragma unroll
          for (int k = 0; k < kZ; ++k) s8[k] = fmaf(dv[e], wrow[k], s8[k]);
        }
      }
#pragma unroll
      for (int k = 0; k < kZ; ++k) s8[k] += __shfl_xor_sync(0xffffffffu, s8[k], 1);
      if (half == 0 && r < nv) {
        *reinterpret_cast<float4*>(g_out + (size_t)(row0 + r) * kZ) = make_float4(s8[0], s8[1], s8[2], s8[3]);
        *reinterpret_cast<float4*>(g_out + (size_t)(row0 + r) * kZ + 4) = make_float4(s8[4], s8[5], s8[6], s8[7]);
      }
    }
    __syncthreads();
  }
  // ---- reductions, gradient export, AdamW ----
  float* gb = gradW + kH * K;    // [64] db | [64] dslope, right behind dW: [W | b | a] is one run of the parameter vector
  sm->red[ty][c4 + 0] = db4[0]; sm->red[ty][c4 + 1] = db4[1]; sm->red[ty][c4 + 2] = db4[2]; sm->red[ty][c4 + 3] = db4[3];
  __syncthreads();
  if (tid < kH) { float s = 0.f; for (int i = 0; i < 16; ++i) s += sm->red[i][tid]; gb[tid] = s; }
  __syncthreads();
  sm->red[ty][c4 + 0] = ds4[0]; sm->red[ty][c4 + 1] = ds4[1]; sm->red[ty][c4 + 2] = ds4[2]; sm->red[ty][c4 + 3] = ds4[3];
  __syncthreads();
  if (tid < kH) { float s = 0.f; for (int i = 0; i < 16; ++i) s += sm->red[i][tid]; gb[kH + tid] = s; }
  __syncthreads();
  if (want_out && in.kind == kInHidden) {
    sm->red[ty][c4 + 0] = sg4[0]; sm->red[ty][c4 + 1] = sg4[1]; sm->red[ty][c4 + 2] = sg4[2]; sm->red[ty][c4 + 3] = sg4[3];
    __syncthreads();
    if (tid < kH) { float s = 0.f; for (int i = 0; i < 16; ++i) s += sm->red[i][tid]; sg_dst[tid] = s; }
    __syncthreads();
    sm->red[ty][c4 + 0] = sgx4[0]; sm->red[ty][c4 + 1] = sgx4[1]; sm->red[ty][c4 + 2] = sgx4[2]; sm->red[ty][c4 + 3] = sgx4[3];
    __syncthreads();
    if (tid < kH) { float s = 0.f; for (int i = 0; i < 16; ++i) s += sm->red[i][tid]; sgx_dst[tid] = s; }
    __syncthreads();
  }
  if (in.kind == kInHidden) {
    const int qq = tid >> 6, tt = tid & 63, m0 = 8 * (tt >> 3), n0 = 8 * (tt & 7);
    for (int pass = 0; pass < 4; ++pass) {
      if (qq == pass) {
#pragma unroll
        for (int i = 0; i < 8; ++i)
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            float* dst = gradW + (m0 + i) * kH + n0 + j;
            *dst = pass == 0 ? accW[i][j] : *dst + accW[i][j];
          }
      }
      __syncthreads();
    }
  } else if (in.kind == kInWide) {
    const int m0 = 8 * (tid >> 5), n0 = 8 * (tid & 31);
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
      for (int j = 0; j < 8; ++j)
        if (n0 + j < K) gradW[(m0 + i) * K + n0 + j] = accW[i][j];
    __syncthreads();
  } else {
    {
      // reduce the four row groups through shared memory (the area behind gradW / gb is free now)
      float* part = gradW + 1024;            // [4][64][8]
      const int ch = tid & 63, q = tid >> 6;
#pragma unroll
      for (int k = 0; k < kZ; ++k) part[(q * kH + ch) * kZ + k] = accS[k];
      __syncthreads();
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        int oo = tid + kThreads * e, n = oo >> 3, k = oo & 7;
        if (k < K) gradW[n * K + k] = part[(0 * kH + n) * kZ + k] + part[(1 * kH + n) * kZ + k] + part[(2 * kH + n) * kZ + k] + part[(3 * kH + n) * kZ + k];
      }
    }
    __syncthreads();
  }
  if (c.csize > 1) {
    cl::sync();                                // every CTA's partial gradients / sums are in place
    if (want_out && in.kind == kInHidden) cluster_gather_sg(c, sm);
  }
  adam_apply(c, sm, o, net, nl.w_off[l], kH * K + 2 * kH, gradW);
  stage_sync(c);
}

// backward of a hidden block whose input is the latent (decoder layer 0, K = nstyle <= 8).  All 16 global loads of a tile
// (g, u) are in flight before the first use; du stays in registers in the (ty, c4) ownership for the weight gradient
// (dW[c4 + j][k] += du z, the latent row as two broadcast loads: no shared-memory round trip, 8 x 4 accumulators per thread)
// and goes to shared memory only for dz = du W, which needs whole rows (two threads per row, combined by a shuffle).
__device__ __noinline__ void bwd_latent(const Ctx& c_ref, int net, int l, const LayerIn& in_ref, const float* __restrict__ u_l,
                                        const float* __restrict__ g_in, float* g_out, int o) {
  const Ctx c = c_ref;                 // register copy of the kernel context (SmemFixed::ctx, shared memory)
  const LayerIn in = in_ref;
  RAAE_SMEM();
  StageTimer timer_(&sm->prof[kStBwdLatent]);
  const raae_net_layout& nl = NL(c, net);
  const int tid = threadIdx.x, ty = tid >> 4, tx = tid & 15, c4 = tx * 4;
  const int K = nl.in_dim[l];
  const float* Wg = netp(c, net) + nl.w_off[l];
  float* Dt = arena;                          // [kTM][kLD] du
  float* Zt = arena + kTile;                  // [kTM][kZ] latent rows
  float* Ws = Zt + kTM * kZ;                  // [64][9]
  float* gradW = Ws + kH * 9 + 16;            // dense [64][K] + [64] db + [64] dslope (reused after the tile loop)
  float* part = gradW + kH * kZ + 2 * kH;     // [16 row groups][64][8] partials of dW
  const bool want_out = g_out != nullptr;
  const float* zmean = in.slayer >= 0 ? sm->mean[in.snet][in.slayer] : nullptr;
  const float* zinv = in.slayer >= 0 ? sm->inv[in.snet][in.slayer] : nullptr;
  const int ntiles = (c.B + kTM - 1) / kTM;
  __syncthreads();
  if (tid < kH) {
    float nB = (float)c.B;
    sm->cg[tid] = sm->sg[tid] / nB;
    sm->cgx[tid] = sm->sgx[tid] / nB;
    sm->slope[tid] = netp(c, net)[nl.a_off[l] + tid];
  }
  for (int i = tid; i < kH * 9; i += kThreads) {
    int n = i / 9, k = i - n * 9;
    Ws[i] = k < K ? Wg[n * K + k] : 0.f;
  }
  __syncthreads();
  const float4 mu = *reinterpret_cast<const float4*>(sm->mean[net][l] + c4);
  const float4 is = *reinterpret_cast<const float4*>(sm->inv[net][l] + c4);
  const float4 sl = *reinterpret_cast<const float4*>(sm->slope + c4);
  const float4 cg = *reinterpret_cast<const float4*>(sm->cg + c4);
  const float4 cgx = *reinterpret_cast<const float4*>(sm->cgx + c4);
  float db4[4] = {0.f, 0.f, 0.f, 0.f}, ds4[4] = {0.f, 0.f, 0.f, 0.f};
  float accW[4][kZ];
#pragma unroll
  for (int j = 0; j < 4; ++j)
#pragma unroll
    for (int k = 0; k < kZ; ++k) accW[j][k] = 0.f;
  const int zr = tid >> 1, zk0 = (tid & 1) * 4;      // this thread's float4 of a latent tile
  for (int t = c.crank; t < ntiles; t += c.csize) {
    const int row0 = t * kTM, nv = min(kTM, c.B - row0);
    float4 gg[kTM / 16], uu[kTM / 16];
#pragma unroll
    for (int i = 0; i < kTM / 16; ++i) {
      const int r = ty + 16 * i;
      const bool ok = r < nv;
      const size_t go = (size_t)(row0 + (ok ? r : 0)) * kH + c4;
      gg[i] = ok ? *reinterpret_cast<const float4*>(g_in + go) : make_float4(0.f, 0.f, 0.f, 0.f);
      uu[i] = ok ? *reinterpret_cast<const float4*>(u_l + go) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
    {
      // latent tile (normalised with the BatchNorm of the encoder output where it applies; zero beyond nstyle / nv)
      const float4 v = zr < nv ? *reinterpret_cast<const float4*>(in.src + (size_t)(row0 + zr) * kZ + zk0) : make_float4(0.f, 0.f, 0.f, 0.f);
      float ov[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const int k = zk0 + e;
        if (zr < nv && k < K) { if (zmean) ov[e] = (ov[e] - zmean[k]) * zinv[k]; }
        else ov[e] = 0.f;
      }
      *reinterpret_cast<float4*>(Zt + zr * kZ + zk0) = make_float4(ov[0], ov[1], ov[2], ov[3]);
    }
    // du = PReLU'(u) BN'(g) (zero for rows >= nv), kept in registers and written to shared memory for dz
    float4 dur[kTM / 16];
#pragma unroll
    for (int i = 0; i < kTM / 16; ++i) {
      const bool valid = ty + 16 * i < nv;
      const float4 g = gg[i], u = uu[i];
      float4 du;
#define RAAE_DU(comp, idx)                                                          \
      {                                                                             \
        float xh = (prelu_f(u.comp, sl.comp) - mu.comp) * is.comp;                  \
        float dh = (g.comp - cg.comp - xh * cgx.comp) * is.comp;                    \
        bool pos = u.comp > 0.f;                                                    \
        float d = pos ? dh : sl.comp * dh;                                          \
        d = valid ? d : 0.f;                                                        \
        du.comp = d;                                                                \
        ds4[idx] += (pos || !valid) ? 0.f : u.comp * dh;                            \
        db4[idx] += d;                                                              \
      }
      RAAE_DU(x, 0) RAAE_DU(y, 1) RAAE_DU(z, 2) RAAE_DU(w, 3)
#undef RAAE_DU
      dur[i] = du;
      if (want_out) *reinterpret_cast<float4*>(Dt + (ty + 16 * i) * kLD + c4) = du;
    }
    __syncthreads();
    // dW[c4 + j][k] += du[r][c4 + j] z[r][k]
#pragma unroll
    for (int i = 0; i < kTM / 16; ++i) {
      const int r = ty + 16 * i;
      const float4 z0 = *reinterpret_cast<const float4*>(Zt + r * kZ);
      const float4 z1 = *reinterpret_cast<const float4*>(Zt + r * kZ + 4);
      const float zz[kZ] = {z0.x, z0.y, z0.z, z0.w, z1.x, z1.y, z1.z, z1.w};
      const float dd[4] = {dur[i].x, dur[i].y, dur[i].z, dur[i].w};
#pragma unroll
      for (int j = 0; j < 4; ++j)
#pragma unroll
        for (int k = 0; k < kZ; ++k) accW[j][k] = fmaf(dd[j], zz[k], accW[j][k]);
    }
    if (want_out) {
      // dz[r][k] = sum_n du[r][n] W[n][k]: two threads per row (32 channels each), partial sums combined by shuffle
      const int r = tid >> 1, half = tid & 1;
      float s8[kZ];
#pragma unroll
      for (int k = 0; k < kZ; ++k) s8[k] = 0.f;
#pragma unroll 2
      for (int n4 = 0; n4 < 32; n4 += 4) {
        const float4 d4 = *reinterpret_cast<const float4*>(Dt + r * kLD + 32 * half + n4);
        const float dv[4] = {d4.x, d4.y, d4.z, d4.w};
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const float* wrow = Ws + (32 * half + n4 + e) * 9;
#pragma unroll
          for (int k = 0; k < kZ; ++k) s8[k] = fmaf(dv[e], wrow[k], s8[k]);
        }
      }
#pragma unroll
      for (int k = 0; k < kZ; ++k) s8[k] += __shfl_xor_sync(0xffffffffu, s8[k], 1);
      if (half == 0 && r < nv) {
        *reinterpret_cast<float4*>(g_out + (size_t)(row0 + r) * kZ) = make_float4(s8[0], s8[1], s8[2], s8[3]);
        *reinterpret_cast<float4*>(g_out + (size_t)(row0 + r) * kZ + 4) = make_float4(s8[4], s8[5], s8[6], s8[7]);
      }
    }
    __syncthreads();
  }
  // ---- reductions over the 16 row groups (fixed order), gradient export, AdamW ----
  float* gb = gradW + kH * K;    // [64] db | [64] dslope, right behind dW: [W | b | a] is one run of the parameter vector
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    *reinterpret_cast<float4*>(part + ((ty * kH) + c4 + j) * kZ) = make_float4(accW[j][0], accW[j][1], accW[j][2], accW[j][3]);
    *reinterpret_cast<float4*>(part + ((ty * kH) + c4 + j) * kZ + 4) = make_float4(accW[j][4], accW[j][5], accW[j][6], accW[j][7]);
  }
  sm->red[ty][c4 + 0] = db4[0]; sm->red[ty][c4 + 1] = db4[1]; sm->red[ty][c4 + 2] = db4[2]; sm->red[ty][c4 + 3] = db4[3];
  __syncthreads();
#pragma unroll
  for (int e = 0; e < 2; ++e) {
    const int oo = tid + kThreads * e, n = oo >> 3, k = oo & 7;
    if (k < K) {
      float sacc = 0.f;
#pragma unroll
      for (int g = 0; g < 16; ++g) sacc += part[((g * kH) + n) * kZ + k];
      gradW[n * K + k] = sacc;
    }
  }
  if (tid < kH) { float sacc = 0.f; for (int i = 0; i < 16; ++i) sacc += sm->red[i][tid]; gb[tid] = sacc; }
  __syncthreads();
  sm->red[ty][c4 + 0] = ds4[0]; sm->red[ty][c4 + 1] = ds4[1]; sm->red[ty][c4 + 2] = ds4[2]; sm->red[ty][c4 + 3] = ds4[3];
  __syncthreads();
  if (tid < kH) { float sacc = 0.f; for (int i = 0; i < 16; ++i) sacc += sm->red[i][tid]; gb[kH + tid] = sacc; }
  __syncthreads();
  if (c.csize > 1) cl::sync();                 // every CTA's partial gradients are in place
  adam_apply(c, sm, o, net, nl.w_off[l], kH * K + 2 * kH, gradW);
  stage_sync(c);
}

// backward of a hidden block whose input is another hidden block's panel (K = 64), software-pipelined: the raw
// g / u / u_prev tiles of tile t+1 are prefetched with cp.async while tile t runs its two contractions.
__device__ __noinline__ void bwd_hidden64(const Ctx& c_ref, int net, int l, const LayerIn& in_ref, const float* __restrict__ u_l,
                                          const float* __restrict__ g_in, float* __restrict__ g_out, int o) {
  const Ctx c = c_ref;                 // register copy of the kernel context (SmemFixed::ctx, shared memory)
  const LayerIn in = in_ref;
  RAAE_SMEM();
  StageTimer timer_(&sm->prof[kStBwdHidden]);
  const raae_net_layout& nl = NL(c, net);
  const int tid = threadIdx.x, ty = tid >> 4, tx = tid & 15, c4 = tx * 4;
  const float* Wg = netp(c, net) + nl.w_off[l];
  float* Gb[2] = {arena, arena + kTile};                     // g tile -> du in place
  float* Pb[2] = {arena + 2 * kTile, arena + 3 * kTile};     // u_prev tile -> input activations in place
  float* Ub = arena + 4 * kTile;                             // u_l tile (free again after the du pass)
  float* Ws = arena + 5 * kTile;                             // [64][kLD] W_l, natural layout (NN operand)
  float* slope_in = sm->shift;                               // slopes of the producing layer (shift is free in backward)
  const float* mean_in = sm->mean[in.snet][in.slayer];
  const float* inv_in = sm->inv[in.snet][in.slayer];
  const int ntiles = (c.B + kTM - 1) / kTM;
  const int t_first = c.crank, tstep = c.csize;
  float* const sg_dst = c.csize > 1 ? sm->sgp : sm->sg;
  float* const sgx_dst = c.csize > 1 ? sm->sgxp : sm->sgx;
  __syncthreads();
  {
    if (has_tiles(c, ntiles)) {
      const int nv0 = min(kTM, c.B - t_first * kTM);
      prefetch_panel_tile(Gb[0], g_in, t_first * kTM, nv0);
      prefetch_panel_tile(Ub, u_l, t_first * kTM, nv0);
      prefetch_panel_tile(Pb[0], in.src, t_first * kTM, nv0);
    }
    cp_async_commit();
  }
  if (tid < kH) {
    float nB = (float)c.B;
    sm->cg[tid] = sm->sg[tid] / nB;
    sm->cgx[tid] = sm->sgx[tid] / nB;
    sm->slope[tid] = netp(c, net)[nl.a_off[l] + tid];
    slope_in[tid] = in.slope[tid];
  }
  load_w_rows(Ws, kLD, Wg, kH, 0, kH);
  __syncthreads();
  const float4 mu = *reinterpret_cast<const float4*>(sm->mean[net][l] + c4);
  const float4 is = *reinterpret_cast<const float4*>(sm->inv[net][l] + c4);
  const float4 sl = *reinterpret_cast<const float4*>(sm->slope + c4);
  const float4 cg = *reinterpret_cast<const float4*>(sm->cg + c4);
  const float4 cgx = *reinterpret_cast<const float4*>(sm->cgx + c4);
  float db4[4] = {0.f, 0.f, 0.f, 0.f}, ds4[4] = {0.f, 0.f, 0.f, 0.f};
  float sg4[4] = {0.f, 0.f, 0.f, 0.f}, sgx4[4] = {0.f, 0.f, 0.f, 0.f};
  float accW[8][8];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j) accW[i][j] = 0.f;
  for (int t = t_first; t < ntiles; t += tstep) {
    const int it = tile_iter(c, t);
    const int row0 = t * kTM, nv = min(kTM, c.B - row0);
    float* Dt = Gb[it & 1];
    float* At = Pb[it & 1];
    cp_async_wait<0>();
    // 1. du = PReLU'(u) * BN'(g), in place over the g tile (own elements)
#pragma unroll
    for (int i = 0; i < kTM / 16; ++i) {
      const int r = ty + 16 * i;
      float4 du = make_float4(0.f, 0.f, 0.f, 0.f);
      if (r < nv) {
        const float4 g = *reinterpret_cast<const float4*>(Dt + r * kLD + c4);
        const float4 u = *reinterpret_cast<const float4*>(Ub + r * kLD + c4);
#define RAAE_DU(comp, idx)                                                      \
        {                                                                       \
          float xh = (prelu_f(u.comp, sl.comp) - mu.comp) * is.comp;            \
          float dh = (g.comp - cg.comp - xh * cgx.comp) * is.comp;              \
          bool pos = u.comp > 0.f;                                              \
          du.comp = pos ? dh : sl.comp * dh;                                    \
          ds4[idx] += pos ? 0.f : u.comp * dh;                                  \
          db4[idx] += du.comp;                                                  \
        }
        RAAE_DU(x, 0) RAAE_DU(y, 1) RAAE_DU(z, 2) RAAE_DU(w, 3)
#undef RAAE_DU
      }
      *reinterpret_cast<float4*>(Dt + r * kLD + c4) = du;
    }
    // 2. the layer's input activations, in place over the u_prev tile
    transform_act_tile(At, row0, nv, mean_in, inv_in, slope_in, in.mask);
    __syncthreads();
    // prefetch tile t+1 (Ub is free; the other two go to the alternate buffers)
    if (t + tstep < ntiles) {
      const int rown = row0 + tstep * kTM, nvn = min(kTM, c.B - rown);
      prefetch_panel_tile(Gb[(it + 1) & 1], g_in, rown, nvn);
      prefetch_panel_tile(Ub, u_l, rown, nvn);
      prefetch_panel_tile(Pb[(it + 1) & 1], in.src, rown, nvn);
    }
    cp_async_commit();
    // 3. dW += du^T a
    {
      const int qq = tid >> 6, tt = tid & 63;
      mma_tn8(Dt, kLD, 8 * (tt >> 3), At, kLD, 8 * (tt & 7), 32 * qq, 32 * qq + 32, accW);
    }
    // 4. g_prev = (du @ W) * dropout mask of the producing layer, and its BN-backward sums
    {
      float acc[8][4];
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
      mma_nn<kH>(Dt, kLD, Ws, kLD, acc, ty, tx);
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        int r = ty + 16 * i;
        if (r < nv) {
          uint32_t kb = mask_keep4(in.mask, row0 + r, c4);
          float4 a = *reinterpret_cast<const float4*>(At + r * kLD + c4);
          float4 gm;
          gm.x = (kb & 1u) ? acc[i][0] * in.mask.scale : 0.f;
          gm.y = (kb & 2u) ? acc[i][1] * in.mask.scale : 0.f;
          gm.z = (kb & 4u) ? acc[i][2] * in.mask.scale : 0.f;
          gm.w = (kb & 8u) ? acc[i][3] * in.mask.scale : 0.f;
          sg4[0] += gm.x; sg4[1] += gm.y; sg4[2] += gm.z; sg4[3] += gm.w;
          sgx4[0] = fmaf(acc[i][0], a.x, sgx4[0]); sgx4[1] = fmaf(acc[i][1], a.y, sgx4[1]);
          sgx4[2] = fmaf(acc[i][2], a.z, sgx4[2]); sgx4[3] = fmaf(acc[i][3], a.w, sgx4[3]);
          *reinterpret_cast<float4*>(g_out + (size_t)(row0 + r) * kH + c4) = gm;
        }
      }
    }
    // no barrier: the next iteration only touches its own elements of the alternate buffers before its barrier
  }
  cp_async_wait<0>();
  __syncthreads();
  // ---- reductions, gradient export, AdamW ----
  float* gradW = Pb[0];            // dense [64][64]
  float* gb = gradW + kH * kH;     // [64] db | [64] dslope, right behind dW (one AdamW pass over [W | b | a])
  sm->red[ty][c4 + 0] = db4[0]; sm->red[ty][c4 + 1] = db4[1]; sm->red[ty][c4 + 2] = db4[2]; sm->red[ty][c4 + 3] = db4[3];
  __syncthreads();
  if (tid < kH) { float s = 0.f; for (int i = 0; i < 16; ++i) s += sm->red[i][tid]; gb[tid] = s; }
  __syncthreads();
  sm->red[ty][c4 + 0] = ds4[0]; sm->red[ty][c4 + 1] = ds4[1]; sm->red[ty][c4 + 2] = ds4[2]; sm->red[ty][c4 + 3] = ds4[3];
  __syncthreads();
  if (tid < kH) { float s = 0.f; for (int i = 0; i < 16; ++i) s += sm->red[i][tid]; gb[kH + tid] = s; }
  __syncthreads();
  sm->red[ty][c4 + 0] = sg4[0]; sm->red[ty][c4 + 1] = sg4[1]; sm->red[ty][c4 + 2] = sg4[2]; sm->red[ty][c4 + 3] = sg4[3];
  __syncthreads();
  if (tid < kH) { float s = 0.f; for (int i = 0; i < 16; ++i) s += sm->red[i][tid]; sg_dst[tid] = s; }
  __syncthreads();
  sm->red[ty][c4 + 0] = sgx4[0]; sm->red[ty][c4 + 1] = sgx4[1]; sm->red[ty][c4 + 2] = sgx4[2]; sm->red[ty][c4 + 3] = sgx4[3];
  __syncthreads();
  if (tid < kH) { float s = 0.f; for (int i = 0; i < 16; ++i) s += sm->red[i][tid]; sgx_dst[tid] = s; }
  {
    const int qq = tid >> 6, tt = tid & 63, m0 = 8 * (tt >> 3), n0 = 8 * (tt & 7);
    for (int pass = 0; pass < 4; ++pass) {
      if (qq == pass) {
#pragma unroll
        for (int i = 0; i < 8; ++i)
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            float* dst = gradW + (m0 + i) * kH + n0 + j;
            *dst = pass == 0 ? accW[i][j] : *dst + accW[i][j];
          }
      }
      __syncthreads();
    }
  }
  if (c.csize > 1) { cl::sync(); cluster_gather_sg(c, sm); }
  adam_apply(c, sm, o, net, nl.w_off[l], kH * kH + 2 * kH, gradW);
  stage_sync(c);
}

// bwd_hidden64 with both contractions on the tensor core: g_prev = du W (128 x 64 x 64, read back every tile) and
// dW += du^T a (64 x 64, K = batch rows, accumulated in TMEM over the whole batch and read back once).
// du is staged K-major (SWIZZLE_128B) for the first product and, once that product has completed, re-staged from
// registers MN-major (SWIZZLE_128B_BASE32B, the only MN-major layout of 32-bit operands; a K-major operand in that
// swizzle faults, tools/tc_probe.cu variant 5) into the same buffer for the second; the input activations are staged
// MN-major only.  The g_prev epilogue overlaps the dW MMAs.  All 24 global loads of a tile (g, u_l, u_prev) are issued
// before the first use, so a tile exposes one memory latency.
__device__ __noinline__ void bwd_hidden64_tc(const Ctx& c_ref, int net, int l, const LayerIn& in_ref, const float* __restrict__ u_l,
                                             const float* __restrict__ g_in, float* __restrict__ g_out, int o) {
  const Ctx c = c_ref;                 // register copy of the kernel context (SmemFixed::ctx, shared memory)
  const LayerIn in = in_ref;
  RAAE_SMEM();
  StageTimer timer_(&sm->prof[kStBwdHidden]);
  const raae_net_layout& nl = NL(c, net);
  const int tid = threadIdx.x, ty = tid >> 4, tx = tid & 15, c4 = tx * 4, warp = tid >> 5, lane = tid & 31;
  const float* Wg = netp(c, net) + nl.w_off[l];
  float* Dhi = arena;                                  // du tile (K-major, then MN-major)
  float* Dlo = Dhi + tc::kATileFloats;
  float* Phi = Dlo + tc::kATileFloats;                 // input-activation tile (MN-major)
  float* Plo = Phi + tc::kATileFloats;
  float* Wthi = Plo + tc::kATileFloats;                // W^T (K-major): rows = input channel k, K index = output channel n
  float* Wtlo = Wthi + tc::kBTileFloats;
  float* Ot = Wtlo + tc::kBTileFloats;                 // [kTM][kLD] g_prev tile bounced from the TMEM layout to (ty, c4) ownership
  float* slope_in = sm->shift;
  const float* u_prev = in.src;
  const MaskSrc mk = in.mask;
  const int B = c.B, ntiles = (B + kTM - 1) / kTM;
  const int t_first = c.crank, tstep = c.csize;        // this CTA's tiles (cluster per trial)
  float* const sg_dst = c.csize > 1 ? sm->sgp : sm->sg;
  float* const sgx_dst = c.csize > 1 ? sm->sgxp : sm->sgx;
  const uint32_t d_tmem = sm->tmem_base;
  uint64_t* mbar = reinterpret_cast<uint64_t*>(&sm->mbar);
  __syncthreads();
  if (tid < kH) {
    float nB = (float)B;
    sm->cg[tid] = sm->sg[tid] / nB;
    sm->cgx[tid] = sm->sgx[tid] / nB;
    sm->slope[tid] = netp(c, net)[nl.a_off[l] + tid];
    slope_in[tid] = in.slope[tid];
  }
  {
    // W[n][k] -> W^T operand element (row k, column n): a warp covers 32 consecutive n of one 4-column group, so the
    // four transposed scalar stores of a thread land in 32 distinct banks across the warp
    float4 w[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int e = tid + kThreads * i, n = e & 63, k4 = (e >> 6) * 4;
      w[i] = *reinterpret_cast<const float4*>(Wg + (size_t)n * kH + k4);
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int e = tid + kThreads * i, n = e & 63, k4 = (e >> 6) * 4;
      const float wv[4] = {w[i].x, w[i].y, w[i].z, w[i].w};
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        float h, lo;
        tc::tf32_split(wv[j], h, lo);
        const uint32_t off = tc::sw128_chunk_off(k4 + j, n & ~3, tc::kBBlockBytes) + (uint32_t)((n & 3) * 4);
        *reinterpret_cast<float*>(reinterpret_cast<char*>(Wthi) + off) = h;
        *reinterpret_cast<float*>(reinterpret_cast<char*>(Wtlo) + off) = lo;
      }
    }
  }
  __syncthreads();
  const float4 mu = *reinterpret_cast<const float4*>(sm->mean[net][l] + c4);
  const float4 is = *reinterpret_cast<const float4*>(sm->inv[net][l] + c4);
  const float4 sl = *reinterpret_cast<const float4*>(sm->slope + c4);
  const float4 cg = *reinterpret_cast<const float4*>(sm->cg + c4);
  const float4 cgx = *reinterpret_cast<const float4*>(sm->cgx + c4);
  const float4 mu_in = *reinterpret_cast<const float4*>(sm->mean[in.snet][in.slayer] + c4);
  const float4 is_in = *reinterpret_cast<const float4*>(sm->inv[in.snet][in.slayer] + c4);
  const float4 sl_in = *reinterpret_cast<const float4*>(slope_in + c4);
  float db4[4] = {0.f, 0.f, 0.f, 0.f}, ds4[4] = {0.f, 0.f, 0.f, 0.f};
  float sg4[4] = {0.f, 0.f, 0.f, 0.f}, sgx4[4] = {0.f, 0.f, 0.f, 0.f};
  uint32_t phase = sm->tc_phase;
  const int erow = 32 * (warp & 3) + lane, ecol0 = 32 * (warp >> 2);     // epilogue ownership (TMEM lane, column block)
  // swizzled staging offsets of this thread's chunks: row = ty + 16 i => (row & 7) and (row & 3) do not depend on i
  const uint32_t offK = tc::sw128_chunk_off(ty, c4, tc::kABlockBytes);          // + i * 16 rows * 128 B
  const uint32_t offM = tc::sw128_32b_chunk_off(ty, c4, tc::kABlockBytes);
  RAAE_PROBE_INIT();
  RAAE_PROBE(27);
  // g / u_l of tile t+1 are loaded into registers before the epilogue of tile t (their latency hides behind it);
  // u_prev of tile t is loaded at the top of the tile and consumed after the du pass
  float4 gg[kTM / 16], uu[kTM / 16];
  auto load_gu = [&](int t) {
    const int row0 = t * kTM, nv = min(kTM, B - row0);
#pragma unroll
    for (int i = 0; i < kTM / 16; ++i) {
      const int r = ty + 16 * i;
      const size_t go = (size_t)(row0 + r) * kH + c4;
      if (r < nv) {
        gg[i] = *reinterpret_cast<const float4*>(g_in + go);
        uu[i] = *reinterpret_cast<const float4*>(u_l + go);
      } else {
        gg[i] = make_float4(0.f, 0.f, 0.f, 0.f);
        uu[i] = make_float4(0.f, 0.f, 0.f, 0.f);
      }
    }
  };
  if (has_tiles(c, ntiles)) load_gu(t_first);
  for (int t = t_first; t < ntiles; t += tstep) {
    const int it = tile_iter(c, t);
    const int row0 = t * kTM, nv = min(kTM, B - row0);
    float4 up[kTM / 16];
#pragma unroll
    for (int i = 0; i < kTM / 16; ++i) {
      const int r = ty + 16 * i;
      up[i] = r < nv ? *reinterpret_cast<const float4*>(u_prev + (size_t)(row0 + r) * kH + c4) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
    // the dW MMAs of the previous tile still read both operand buffers
    if (it > 0) { tc::mbar_wait(mbar, phase); phase ^= 1u; }
    RAAE_PROBE(28);
    // ---- 1. du = PReLU'(u) BN'(g): kept in registers, staged K-major ----
    float4 dur[kTM / 16];
#pragma unroll
    for (int i = 0; i < kTM / 16; ++i) {
      const int r = ty + 16 * i;
      float4 du = make_float4(0.f, 0.f, 0.f, 0.f);
      const float4 g = gg[i], u = uu[i];
      const bool valid = r < nv;
#define RAAE_DU(comp, idx)                                                          \
      {                                                                             \
        float xh = (prelu_f(u.comp, sl.comp) - mu.comp) * is.comp;                  \
        float dh = (g.comp - cg.comp - xh * cgx.comp) * is.comp;                    \
        bool pos = u.comp > 0.f;                                                    \
        float d = pos ? dh : sl.comp * dh;                                          \
        d = valid ? d : 0.f;                                                        \
        du.comp = d;                                                                \
        ds4[idx] += (pos || !valid) ? 0.f : u.comp * dh;                            \
        db4[idx] += d;                                                              \
      }
      RAAE_DU(x, 0) RAAE_DU(y, 1) RAAE_DU(z, 2) RAAE_DU(w, 3)
#undef RAAE_DU
      dur[i] = du;
      tc::split_store(Dhi, Dlo, offK + (uint32_t)(i * 16 * 128), du);
    }
    tc::fence_async_smem();
    tc::fence_before_sync();           // TMEM reads of the previous tile's epilogue precede the next MMA
    __syncthreads();
    if (tc::warp_uniform_id() == 0 && tc::elect_one()) {
      tc::fence_after_sync();
      tc::issue_gemm_3xtf32(d_tmem, Dhi, Dlo, Wthi, Wtlo);                              // g_prev = du W
      tc::mma_commit(mbar);
    }
    // ---- 2. the layer's input activations, staged MN-major WHILE the g_prev MMAs run (their operand buffers are free: the
    //         dW MMAs of the previous tile were waited for above; measured 3 % of the stage against staging them before the
    //         MMAs are issued); the keep bits are reused by the epilogue ----
    const uint32_t kbits = mask_keep4_rows(mk, row0, ty, c4, nv);
#pragma unroll
    for (int i = 0; i < kTM / 16; ++i) {
      float4 a;
      const uint32_t kb = kbits >> (4 * i);
      a.x = (kb & 1u) ? (prelu_f(up[i].x, sl_in.x) - mu_in.x) * is_in.x * mk.scale : 0.f;
      a.y = (kb & 2u) ? (prelu_f(up[i].y, sl_in.y) - mu_in.y) * is_in.y * mk.scale : 0.f;
      a.z = (kb & 4u) ? (prelu_f(up[i].z, sl_in.z) - mu_in.z) * is_in.z * mk.scale : 0.f;
      a.w = (kb & 8u) ? (prelu_f(up[i].w, sl_in.w) - mu_in.w) * is_in.w * mk.scale : 0.f;
      tc::split_store(Phi, Plo, offM + (uint32_t)(i * 16 * 128), a);
    }
    RAAE_PROBE(29);
    tc::mbar_wait(mbar, phase);
    phase ^= 1u;
    RAAE_PROBE(30);
    // ---- 3. read the g_prev accumulator back (before the dW MMAs are queued) and bounce it through shared memory to the
    //         (ty, c4) ownership of the staging passes; re-stage du MN-major (its K-major copy has been consumed) ----
    {
      float v[32];
      tc::fence_after_sync();
      tc::tmem_ld32(d_tmem + ((uint32_t)(32 * (warp & 3)) << 16) + (uint32_t)ecol0, v);
      tc::fence_before_sync();
#pragma unroll
      for (int j = 0; j < 32; j += 4) *reinterpret_cast<float4*>(Ot + erow * kLD + ecol0 + j) = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
    }
#pragma unroll
    for (int i = 0; i < kTM / 16; ++i) tc::split_store(Dhi, Dlo, offM + (uint32_t)(i * 16 * 128), dur[i]);
    tc::fence_async_smem();
    __syncthreads();
    if (tc::warp_uniform_id() == 0 && tc::elect_one()) {
      tc::fence_after_sync();
      tc::issue_gemm_tn_3xtf32(d_tmem + 64, Dhi, Dlo, Phi, Plo, it > 0 ? 1u : 0u);
      tc::mma_commit(mbar);
    }
    if (t + tstep < ntiles) load_gu(t + tstep);
    // ---- 4. g_prev epilogue (overlaps the dW MMAs): dropout mask of the producing layer, BN-backward partial sums,
    //         coalesced store; a = hi + lo is read back from this thread's own staged chunks ----
    uint32_t kbits_e = kbits;
    asm volatile("" : "+r"(kbits_e));    // opaque copy: otherwise the 32 bit tests of the staging pass are kept alive in 32 registers (spills)
#pragma unroll
    for (int i = 0; i < kTM / 16; ++i) {
      // branch-free body (rows >= nv carry zeros: du = 0 there, and their keep bits are clear); only the store is guarded,
      // so the eight row groups interleave
      const int r = ty + 16 * i;
      const float4 g4 = *reinterpret_cast<const float4*>(Ot + r * kLD + c4);
      const uint32_t off = offM + (uint32_t)(i * 16 * 128);
      const float4 ah = *reinterpret_cast<const float4*>(reinterpret_cast<const char*>(Phi) + off);
      const float4 al = *reinterpret_cast<const float4*>(reinterpret_cast<const char*>(Plo) + off);
      const uint32_t kb = kbits_e >> (4 * i);
      float4 gm;
      gm.x = (kb & 1u) ? g4.x * mk.scale : 0.f;
      gm.y = (kb & 2u) ? g4.y * mk.scale : 0.f;
      gm.z = (kb & 4u) ? g4.z * mk.scale : 0.f;
      gm.w = (kb & 8u) ? g4.w * mk.scale : 0.f;
      sg4[0] += gm.x; sg4[1] += gm.y; sg4[2] += gm.z; sg4[3] += gm.w;
      sgx4[0] = fmaf(g4.x, ah.x + al.x, sgx4[0]); sgx4[1] = fmaf(g4.y, ah.y + al.y, sgx4[1]);
      sgx4[2] = fmaf(g4.z, ah.z + al.z, sgx4[2]); sgx4[3] = fmaf(g4.w, ah.w + al.w, sgx4[3]);
      if (r < nv) *reinterpret_cast<float4*>(g_out + (size_t)(row0 + r) * kH + c4) = gm;
    }
    RAAE_PROBE(31);
    // no barrier here: the next tile overwrites the operand buffers only after it has waited for this tile's dW MMAs,
    // but its P stores must not overtake the P reads of this epilogue in other warps -> barrier
    __syncthreads();
  }
  const bool any_tile = has_tiles(c, ntiles);      // a CTA without rows (short batch, large cluster) contributes zeros
  if (any_tile) { tc::mbar_wait(mbar, phase); phase ^= 1u; }      // last dW MMAs
  if (tid == 0) sm->tc_phase = phase;
  // ---- weight gradient: TMEM columns [64,128), M = 64 layout (row n -> lane 32 (n / 16) + n % 16) ----
  float* gradW = Phi;                 // dense [64][64]
  float* gb = gradW + kH * kH;        // [64] db | [64] dslope, right behind dW (one AdamW pass over [W | b | a])
  tc::fence_after_sync();
  __syncthreads();
  if (warp < 4) {
    float v[32];
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      tc::tmem_ld32(d_tmem + ((uint32_t)(32 * warp) << 16) + (uint32_t)(64 + 32 * h), v);
      if (!any_tile) {
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] = 0.f;
      }
      if (lane < 16) {
        const int n = 16 * warp + lane;
#pragma unroll
        for (int j = 0; j < 32; j += 4) *reinterpret_cast<float4*>(gradW + n * kH + 32 * h + j) = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
      }
    }
  }
  tc::fence_before_sync();
  sm->red[ty][c4 + 0] = db4[0]; sm->red[ty][c4 + 1] = db4[1]; sm->red[ty][c4 + 2] = db4[2]; sm->red[ty][c4 + 3] = db4[3];
  __syncthreads();
  if (tid < kH) { float s = 0.f; for (int i = 0; i < 16; ++i) s += sm->red[i][tid]; gb[tid] = s; }
  __syncthreads();
  sm->red[ty][c4 + 0] = ds4[0]; sm->red[ty][c4 + 1] = ds4[1]; sm->red[ty][c4 + 2] = ds4[2]; sm->red[ty][c4 + 3] = ds4[3];
  __syncthreads();
  if (tid < kH) { float s = 0.f; for (int i = 0; i < 16; ++i) s += sm->red[i][tid]; gb[kH + tid] = s; }
  __syncthreads();
  sm->red[ty][c4 + 0] = sg4[0]; sm->red[ty][c4 + 1] = sg4[1]; sm->red[ty][c4 + 2] = sg4[2]; sm->red[ty][c4 + 3] = sg4[3];
  __syncthreads();
  if (tid < kH) { float s = 0.f; for (int i = 0; i < 16; ++i) s += sm->red[i][tid]; sg_dst[tid] = s; }
  __syncthreads();
  sm->red[ty][c4 + 0] = sgx4[0]; sm->red[ty][c4 + 1] = sgx4[1]; sm->red[ty][c4 + 2] = sgx4[2]; sm->red[ty][c4 + 3] = sgx4[3];
  __syncthreads();
  if (tid < kH) { float s = 0.f; for (int i = 0; i < 16; ++i) s += sm->red[i][tid]; sgx_dst[tid] = s; }
  __syncthreads();
  RAAE_PROBE(27);
  if (c.csize > 1) { cl::sync(); cluster_gather_sg(c, sm); }
  adam_apply(c, sm, o, net, nl.w_off[l], kH * kH + 2 * kH, gradW);
  stage_sync(c);
  RAAE_PROBE(27);
}

// Backward of the input block of the encoder on the noised batch (no input gradient): du = PReLU'(u) BN'(g) per tile,
// staged MN-major as the B operand (N = 64 output channels); the batch comes from the centred MN-major operand image
// (ScratchLayout::xm) in 128-column chunks (bulk copy 64 KB, rounded hi / lo split in shared memory) as the A operand
// (M = 128 input columns), and dW^T[chunk] += x^T du accumulates in TMEM over the whole batch (M = 128: full-rate MMAs,
// half as many as with M = 64).  dW = dWc + db (x) xref undoes the centring.
// With dx_out != null (MI phase: the input is y = act(v), imaged by the decoder output stage) the input gradient
// dL/dy act'(v) = (du W) act'(v) is also produced, tile by tile after the weight-gradient MMAs (FP32 FMA on the du tile
// rebuilt from its hi + lo planes, weight chunks double-buffered with cp.async in the then idle chunk buffers), and
// written over v.
__device__ __noinline__ void bwd_wide_img(const Ctx& c_ref, int net, int l, const float* __restrict__ u_l,
                                          const float* __restrict__ g_in, int o, const float* __restrict__ xm,
                                          const float* __restrict__ xref, float* dx_out, int dx_ld, int act) {
  const Ctx c = c_ref;                 // register copy of the kernel context (SmemFixed::ctx, shared memory)
  RAAE_SMEM();
  StageTimer timer_(&sm->prof[kStBwdWide]);
  const raae_net_layout& nl = NL(c, net);
  const int tid = threadIdx.x, ty = tid >> 4, tx = tid & 15, c4 = tx * 4, warp = tid >> 5, lane = tid & 31;
  const int K = nl.in_dim[l], nch = c.p->sl.nch128;
  float* Dhi = arena;                        // du tile, MN-major [2 blocks][128 rows][32]
  float* Dlo = Dhi + 8192;
  float* Xhi = Dlo + 8192;                   // batch chunk, MN-major [4 blocks][128 rows][32]: raw, rounded in place
  float* Xlo = Xhi + 16384;
  const int B = c.B, ntiles = (B + kTM - 1) / kTM;
  const uint32_t d_tmem = sm->tmem_base;
  uint64_t* full = reinterpret_cast<uint64_t*>(&sm->pipe_bar[0]);      // chunk landed
  uint64_t* done = reinterpret_cast<uint64_t*>(&sm->pipe_bar[1]);      // MMAs of the chunk completed
  __syncthreads();
  if (tid == 0) {
    tc::mbar_init(full, 1);
    tc::mbar_init(done, 1);
  }
  if (tid < kH) {
    float nB = (float)B;
    sm->cg[tid] = sm->sg[tid] / nB;
    sm->cgx[tid] = sm->sgx[tid] / nB;
    sm->slope[tid] = netp(c, net)[nl.a_off[l] + tid];
  }
  __syncthreads();
  const float4 mu = *reinterpret_cast<const float4*>(sm->mean[net][l] + c4);
  const float4 is = *reinterpret_cast<const float4*>(sm->inv[net][l] + c4);
  const float4 sl = *reinterpret_cast<const float4*>(sm->slope + c4);
  const float4 cg = *reinterpret_cast<const float4*>(sm->cg + c4);
  const float4 cgx = *reinterpret_cast<const float4*>(sm->cgx + c4);
  float db4[4] = {0.f, 0.f, 0.f, 0.f}, ds4[4] = {0.f, 0.f, 0.f, 0.f};
  const uint32_t offM = tc::sw128_32b_chunk_off(ty, c4, tc::kABlockBytes);
  const bool leader = tc::warp_uniform_id() == 0;
  uint32_t nfull = 0u, ndone = 0u;                       // completed phases seen (parity = count & 1)
  auto load_chunk = [&](int t, int ck) {                   // elected thread only
    // eight concurrent 8 KB copies: one bulk copy keeps only a few KB in flight and is latency-limited (~7 B/cycle)
    tc::mbar_expect_tx(full, 65536u);
    const float* src = xm + ((size_t)t * nch + ck) * 16384;
#pragma unroll
    for (int part = 0; part < 8; ++part) tc::bulk_g2s(Xhi + part * 2048, src + part * 2048, 8192u, full);
  };
  for (int t = c.crank; t < ntiles; t += c.csize) {      // this CTA's tiles (cluster per trial)
    const int it = tile_iter(c, t);
    const int row0 = t * kTM, nv = min(kTM, B - row0);
    if (leader) {
      if (tc::elect_one()) load_chunk(t, 0);               // overlaps the du pass
      __syncwarp();
    }
    // ---- du = PReLU'(u) BN'(g) -> MN-major hi / lo ----
    float4 gg[kTM / 16], uu[kTM / 16];
#pragma unroll
    for (int i = 0; i < kTM / 16; ++i) {
      const int r = ty + 16 * i;
      const size_t go = (size_t)(row0 + r) * kH + c4;
      if (r < nv) {
        gg[i] = *reinterpret_cast<const float4*>(g_in + go);
        uu[i] = *reinterpret_cast<const float4*>(u_l + go);
      } else {
        gg[i] = make_float4(0.f, 0.f, 0.f, 0.f);
        uu[i] = make_float4(0.f, 0.f, 0.f, 0.f);
      }
    }
#pragma unroll
    for (int i = 0; i < kTM / 16; ++i) {
      const int r = ty + 16 * i;
      float4 du = make_float4(0.f, 0.f, 0.f, 0.f);
      const float4 g = gg[i], u = uu[i];
      const bool valid = r < nv;
#define RAAE_DU(comp, idx)                                                          \
      {                                                                             \
        float xh = (prelu_f(u.comp, sl.comp) - mu.comp) * is.comp;                  \
        float dh = (g.comp - cg.comp - xh * cgx.comp) * is.comp;                    \
        bool pos = u.comp > 0.f;                                                    \
        float d = pos ? dh : sl.comp * dh;                                          \
        d = valid ? d : 0.f;                                                        \
        du.comp = d;                                                                \
        ds4[idx] += (pos || !valid) ? 0.f : u.comp * dh;                            \
        db4[idx] += d;                                                              \
      }
      RAAE_DU(x, 0) RAAE_DU(y, 1) RAAE_DU(z, 2) RAAE_DU(w, 3)
#undef RAAE_DU
      tc::split_store(Dhi, Dlo, offM + (uint32_t)(i * 16 * 128), du);
    }
    // ---- chunks of 128 input columns ----
    for (int ck = 0; ck < nch; ++ck) {
      tc::mbar_wait(full, nfull & 1u);
      ++nfull;
      {
        float4* X = reinterpret_cast<float4*>(Xhi);
        float4* XL = reinterpret_cast<float4*>(Xlo);
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          float4 x[8];
#pragma unroll
          for (int k = 0; k < 8; ++k) x[k] = X[tid + kThreads * (8 * h + k)];
#pragma unroll
          for (int k = 0; k < 8; ++k) {
            float4 hi, lo;
            tc::tf32_split(x[k].x, hi.x, lo.x);
            tc::tf32_split(x[k].y, hi.y, lo.y);
            tc::tf32_split(x[k].z, hi.z, lo.z);
            tc::tf32_split(x[k].w, hi.w, lo.w);
            X[tid + kThreads * (8 * h + k)] = hi;
            XL[tid + kThreads * (8 * h + k)] = lo;
          }
        }
      }
      tc::fence_async_smem();
      __syncthreads();
      if (leader) {
        if (tc::elect_one()) {
          tc::fence_after_sync();
          const uint32_t acc = d_tmem + (uint32_t)(64 * ck);
          tc::issue_gemm_tn128_pass(acc, Xlo, tc::kABlockBytes, Dhi, tc::kABlockBytes, it > 0 ? 1u : 0u);
          tc::issue_gemm_tn128_pass(acc, Xhi, tc::kABlockBytes, Dlo, tc::kABlockBytes, 1u);
          tc::issue_gemm_tn128_pass(acc, Xhi, tc::kABlockBytes, Dhi, tc::kABlockBytes, 1u);
          tc::mma_commit(done);
        }
        __syncwarp();
      }
      // the chunk buffers (and, after the last chunk, the du tile) are reused: wait for these MMAs
      tc::mbar_wait(done, ndone & 1u);
      ++ndone;
      if (ck + 1 < nch && leader) {
        if (tc::elect_one()) load_chunk(t, ck + 1);
        __syncwarp();
      }
    }
    if (dx_out != nullptr) {
      // ---- input gradient of the tile (all MMAs of the tile have completed: the chunk buffers are idle) ----
      const float* Wg = netp(c, net) + nl.w_off[l];
      float* Dt = Xhi;                         // [kTM][kLD] du = hi + lo, row-major
      float* Ws0 = Xlo;                        // two [64][kLD] weight chunks
      float* Ws1 = Xlo + kWTile;
      auto prefetch_wchunk = [&](float* dst, int k0) {
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const int e = tid + kThreads * u, n = e >> 4, k4 = (e & 15) * 4;
          if (k0 + k4 < K) cp_async16(dst + n * kLD + k4, Wg + (size_t)n * K + k0 + k4);
          else *reinterpret_cast<float4*>(dst + n * kLD + k4) = make_float4(0.f, 0.f, 0.f, 0.f);
        }
        cp_async_commit();
      };
      prefetch_wchunk(Ws0, 0);
#pragma unroll
      for (int i = 0; i < kTM / 16; ++i) {      // this thread's own staged chunks
        const uint32_t off = offM + (uint32_t)(i * 16 * 128);
        const float4 h = *reinterpret_cast<const float4*>(reinterpret_cast<const char*>(Dhi) + off);
        const float4 lo = *reinterpret_cast<const float4*>(reinterpret_cast<const char*>(Dlo) + off);
        *reinterpret_cast<float4*>(Dt + (ty + 16 * i) * kLD + c4) = make_float4(h.x + lo.x, h.y + lo.y, h.z + lo.z, h.w + lo.w);
      }
      for (int k0 = 0, ci = 0; k0 < K; k0 += kH, ++ci) {
        float* Wcur = (ci & 1) ? Ws1 : Ws0;
        // act'(v) of this chunk: loads in flight during the contraction
        float4 vv[kTM / 16];
#pragma unroll
        for (int i = 0; i < kTM / 16; ++i) {
          const int r = ty + 16 * i;
          vv[i] = (r < nv && k0 + c4 < K) ? *reinterpret_cast<const float4*>(dx_out + (size_t)(row0 + r) * dx_ld + k0 + c4)
                                          : make_float4(0.f, 0.f, 0.f, 0.f);
        }
        cp_async_wait<0>();
        __syncthreads();
        if (k0 + kH < K) prefetch_wchunk((ci & 1) ? Ws0 : Ws1, k0 + kH);
        float acc[8][4];
#pragma unroll
        for (int i = 0; i < 8; ++i)
#pragma unroll
          for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
        mma_nn<kH>(Dt, kLD, Wcur, kLD, acc, ty, tx);
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const int r = ty + 16 * i;
          float4 v = vv[i];                    // branch-free arithmetic (the activation is warp-uniform), guarded store
          if (act == 1) {
            v.x = acc[i][0] * softplus2_grad_f(v.x); v.y = acc[i][1] * softplus2_grad_f(v.y);
            v.z = acc[i][2] * softplus2_grad_f(v.z); v.w = acc[i][3] * softplus2_grad_f(v.w);
          } else {
            v.x = v.x > 0.f ? acc[i][0] : 0.f; v.y = v.y > 0.f ? acc[i][1] : 0.f;
            v.z = v.z > 0.f ? acc[i][2] : 0.f; v.w = v.w > 0.f ? acc[i][3] : 0.f;
          }
          if (r < nv && k0 + c4 < K) *reinterpret_cast<float4*>(dx_out + (size_t)(row0 + r) * dx_ld + k0 + c4) = v;
        }
      }
      __syncthreads();                         // the chunk buffers are reloaded by the next tile
    }
  }
  // ---- weight gradient from the TMEM accumulators: lane = input column of the chunk, column = output channel ----
  float* gradW = Xhi;                 // dense [64][K]
  float* gb = gradW + kH * K;         // [64] db | [64] dslope, right behind dW (K = 256: the first floats of the idle Xlo)
  tc::fence_after_sync();
  __syncthreads();
  sm->red[ty][c4 + 0] = db4[0]; sm->red[ty][c4 + 1] = db4[1]; sm->red[ty][c4 + 2] = db4[2]; sm->red[ty][c4 + 3] = db4[3];
  __syncthreads();
  if (tid < kH) { float s = 0.f; for (int i = 0; i < 16; ++i) s += sm->red[i][tid]; gb[tid] = s; }
  __syncthreads();
  sm->red[ty][c4 + 0] = ds4[0]; sm->red[ty][c4 + 1] = ds4[1]; sm->red[ty][c4 + 2] = ds4[2]; sm->red[ty][c4 + 3] = ds4[3];
  __syncthreads();
  if (tid < kH) { float s = 0.f; for (int i = 0; i < 16; ++i) s += sm->red[i][tid]; gb[kH + tid] = s; }
  __syncthreads();
  {
    // warp w reads TMEM lanes 32 (w & 3) .. + 31 (input columns of the chunk) and output channels 32 (w >> 2) .. + 31
    const int n0 = 32 * (warp >> 2);
    const bool any_tile = has_tiles(c, ntiles);            // a CTA without rows contributes zeros
    for (int ck = 0; ck < nch; ++ck) {
      float v[32];
      tc::tmem_ld32(d_tmem + ((uint32_t)(32 * (warp & 3)) << 16) + (uint32_t)(64 * ck + n0), v);
      if (!any_tile) {
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] = 0.f;
      }
      const int k = 128 * ck + 32 * (warp & 3) + lane;
      if (k < K) {
        const float xr = xref[k];
#pragma unroll
        for (int j = 0; j < 32; ++j) gradW[(n0 + j) * K + k] = fmaf(gb[n0 + j], xr, v[j]);
      }
    }
  }
  tc::fence_before_sync();
  __syncthreads();
  if (c.csize > 1) cl::sync();
  adam_apply(c, sm, o, net, nl.w_off[l], kH * K + 2 * kH, gradW);
  stage_sync(c);
}

// ------------------------------------------------------------------------------------------
// last encoder layer backward: BN(nstyle) -> Linear(64, nstyle); input gradient for hidden layer L-2
// ------------------------------------------------------------------------------------------
__device__ __noinline__ void bwd_enc_last(const Ctx& c_ref, const LayerIn& in_ref, float* __restrict__ g_out, int o) {
  const Ctx c = c_ref;                 // register copy of the kernel context (SmemFixed::ctx, shared memory)
  const LayerIn in = in_ref;
  RAAE_SMEM();
  StageTimer timer_(&sm->prof[kStBwdEncLast]);
  const raae_net_layout& nl = NL(c, kE);
  const int L = nl.n_linear, l = L - 1, ns = nl.out_dim[l];
  const int tid = threadIdx.x;
  const float* zE = c.sc + c.p->sl.zE;
  const float* dz = c.sc + c.p->sl.dz;
  float* Ws = arena;                   // [kZ][kLD]
  float* At = Ws + kZ * kLD;           // [kTM][kLD]
  float* D5 = At + kTile;              // [kTM][kZ]
  float* gradW = D5 + kTM * kZ;        // [kZ][64] + [kZ]
  float* const sg_dst = c.csize > 1 ? sm->sgp : sm->sg;
  float* const sgx_dst = c.csize > 1 ? sm->sgxp : sm->sgx;
  __syncthreads();
  // batch means of dz and dz * zhat (cluster per trial: sums over this CTA's rows, all-reduced)
  {
    const int k = tid & 7, g = tid >> 3;
    const int crank = c.crank, csize = c.csize;
    const int nslots = csize == 1 ? c.B : cl::own_tiles(c.B, crank, csize) * kTM;
    const float mu = sm->mean[kE][l][k], is = sm->inv[kE][l][k];
    float s0 = 0.f, s1 = 0.f;
    for (int r0 = g; r0 < nslots; r0 += 32 * 8) {          // 16 loads in flight per thread
      float dv[8], zv[8];
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        const int sl = r0 + 32 * u, r = csize == 1 ? sl : cl::slot_row(sl, crank, csize);
        const bool ok = sl < nslots && r < c.B;
        dv[u] = ok ? dz[(size_t)r * kZ + k] : 0.f;
        zv[u] = ok ? zE[(size_t)r * kZ + k] : mu;
      }
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        s0 += dv[u];
        s1 = fmaf(dv[u], (zv[u] - mu) * is, s1);
      }
    }
    float* red = &sm->red[0][0];
    red[tid] = s0;
    red[256 + tid] = s1;
    __syncthreads();
    if (tid < kZ) {
      float t0 = 0.f, t1 = 0.f;
      for (int i = 0; i < 32; ++i) { t0 += red[i * 8 + tid]; t1 += red[256 + i * 8 + tid]; }
      sm->zs[2][tid] = t0;
      sm->zs[3][tid] = t1;
    }
    cluster_allreduce_f(c, sm, &sm->zs[2][0], 2 * kZ);        // zs[2] and zs[3] are contiguous
    __syncthreads();
    if (tid < 2 * kZ) (&sm->zs[2][0])[tid] = (&sm->zs[2][0])[tid] / (float)c.B;
  }
  for (int i = tid; i < kZ * kH; i += kThreads) {
    int n = i >> 6, k = i & 63;
    Ws[n * kLD + k] = n < ns ? netp(c, kE)[nl.w_off[l] + n * kH + k] : 0.f;
  }
  __syncthreads();
  // thread (ty, c4) owns channels c4..c4+3 of rows ty, ty + 16, ...: its 8 x 4 block of W stays in registers, a row of the
  // latent gradient is two broadcast loads, one mask draw covers the four channels, float4 loads / stores
  const int ty = tid >> 4, c4 = (tid & 15) * 4;
  float accW[kZ][4], wcol[kZ][4];
#pragma unroll
  for (int n = 0; n < kZ; ++n) {
    const float4 w = *reinterpret_cast<const float4*>(Ws + n * kLD + c4);
    wcol[n][0] = w.x; wcol[n][1] = w.y; wcol[n][2] = w.z; wcol[n][3] = w.w;
    accW[n][0] = accW[n][1] = accW[n][2] = accW[n][3] = 0.f;
  }
  float accB = 0.f;                                   // threads < 64: latent n = tid & 7, rows (tid >> 3) + 8 j of every tile
  float sgp4[4] = {0.f, 0.f, 0.f, 0.f}, sgxp4[4] = {0.f, 0.f, 0.f, 0.f};
  const int ntiles = (c.B + kTM - 1) / kTM;
  for (int t = c.crank; t < ntiles; t += c.csize) {
    const int row0 = t * kTM, nv = min(kTM, c.B - row0);
    build_act_tile(At, in.src, row0, nv, sm->mean[in.snet][in.slayer], sm->inv[in.snet][in.slayer], in.slope, in.mask);
    {
      // one float4 of zE and of dz per thread (row tid / 2, columns 4 (tid & 1) ..): a single memory latency per tile
      const int r = tid >> 1, k0 = (tid & 1) * 4;
      float4 zz = make_float4(0.f, 0.f, 0.f, 0.f), dd = make_float4(0.f, 0.f, 0.f, 0.f);
      if (r < nv) {
        zz = *reinterpret_cast<const float4*>(zE + (size_t)(row0 + r) * kZ + k0);
        dd = *reinterpret_cast<const float4*>(dz + (size_t)(row0 + r) * kZ + k0);
      }
      const float zv[4] = {zz.x, zz.y, zz.z, zz.w}, dv[4] = {dd.x, dd.y, dd.z, dd.w};
      float o[4];
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const int k = k0 + e;
        const float is = sm->inv[kE][l][k];
        const float zh = (zv[e] - sm->mean[kE][l][k]) * is;
        o[e] = (r < nv && k < ns) ? (dv[e] - sm->zs[2][k] - zh * sm->zs[3][k]) * is : 0.f;
      }
      *reinterpret_cast<float4*>(D5 + r * kZ + k0) = make_float4(o[0], o[1], o[2], o[3]);
    }
    __syncthreads();
    // dW[n][c4 + j] += D5[r][n] a[r][c4 + j];  g_prev[r][c4 + j] = sum_n D5[r][n] W[n][c4 + j]
    const uint32_t kbits = mask_keep4_rows(in.mask, row0, ty, c4, nv);
#pragma unroll 2
    for (int i = 0; i < kTM / 16; ++i) {
      const int r = ty + 16 * i;
      const float4 d0 = *reinterpret_cast<const float4*>(D5 + r * kZ);
      const float4 d1 = *reinterpret_cast<const float4*>(D5 + r * kZ + 4);
      const float dr[kZ] = {d0.x, d0.y, d0.z, d0.w, d1.x, d1.y, d1.z, d1.w};
      const float4 a4 = *reinterpret_cast<const float4*>(At + r * kLD + c4);
      const float av[4] = {a4.x, a4.y, a4.z, a4.w};
      float gg[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
      for (int n = 0; n < kZ; ++n)
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          accW[n][j] = fmaf(dr[n], av[j], accW[n][j]);
          gg[j] = fmaf(dr[n], wcol[n][j], gg[j]);
        }
      {
        // rows >= nv: D5 and the activations are zero there and the keep bits clear; only the store is guarded
        const uint32_t kb = kbits >> (4 * i);
        float gm[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          gm[j] = (kb >> j & 1u) ? gg[j] * in.mask.scale : 0.f;
          sgp4[j] += gm[j];
          sgxp4[j] = fmaf(gg[j], av[j], sgxp4[j]);
        }
        if (r < nv) *reinterpret_cast<float4*>(g_out + (size_t)(row0 + r) * kH + c4) = make_float4(gm[0], gm[1], gm[2], gm[3]);
      }
    }
    if (tid < 64) {
      const int n = tid & 7, g = tid >> 3;
#pragma unroll 4
      for (int j = 0; j < kTM / 8; ++j) accB += D5[(g + 8 * j) * kZ + n];
    }
    __syncthreads();
  }
  // reduce the 16 row groups: dW partials [16][8][64] in the At area (free now), the BN-backward sums and the bias partials
  float* part = At;                      // [16][kZ][64] = 8192 floats <= kTile
#pragma unroll
  for (int n = 0; n < kZ; ++n)
    *reinterpret_cast<float4*>(part + (ty * kZ + n) * kH + c4) = make_float4(accW[n][0], accW[n][1], accW[n][2], accW[n][3]);
  __syncthreads();
#pragma unroll
  for (int e = 0; e < 2; ++e) {
    const int oo = tid + kThreads * e, n = oo >> 6, k = oo & 63;
    if (n < ns) {
      float sacc = 0.f;
#pragma unroll
      for (int g = 0; g < 16; ++g) sacc += part[(g * kZ + n) * kH + k];
      gradW[n * kH + k] = sacc;
    }
  }
  sm->red[ty][c4 + 0] = sgp4[0]; sm->red[ty][c4 + 1] = sgp4[1]; sm->red[ty][c4 + 2] = sgp4[2]; sm->red[ty][c4 + 3] = sgp4[3];
  __syncthreads();
  if (tid < kH) { float sacc = 0.f; for (int i = 0; i < 16; ++i) sacc += sm->red[i][tid]; sg_dst[tid] = sacc; }
  __syncthreads();
  sm->red[ty][c4 + 0] = sgxp4[0]; sm->red[ty][c4 + 1] = sgxp4[1]; sm->red[ty][c4 + 2] = sgxp4[2]; sm->red[ty][c4 + 3] = sgxp4[3];
  __syncthreads();
  if (tid < kH) { float sacc = 0.f; for (int i = 0; i < 16; ++i) sacc += sm->red[i][tid]; sgx_dst[tid] = sacc; }
  __syncthreads();
  if (tid < 64) sm->red[0][tid] = accB;               // [8 row groups][8 latents]
  __syncthreads();
  if (tid < ns) {
    float sacc = 0.f;
#pragma unroll
    for (int g = 0; g < 8; ++g) sacc += sm->red[0][g * 8 + tid];
    gradW[ns * kH + tid] = sacc;                      // [W | b] contiguous
  }
  __syncthreads();
  if (c.csize > 1) { cl::sync(); cluster_gather_sg(c, sm); }
  adam_apply(c, sm, o, kE, nl.w_off[l], ns * kH + ns, gradW);
  stage_sync(c);
}

__device__ __forceinline__ void bwd_hidden(const Ctx& c, int net, int l, const LayerIn& in, const float* __restrict__ u_l,
                                           const float* __restrict__ g_in, float* g_out, int o) {
  if (in.kind == kInHidden && g_out != nullptr) {
    if (c.p->cfg.tensor_cores & 2) bwd_hidden64_tc(c, net, l, in, u_l, g_in, g_out, o);
    else bwd_hidden64(c, net, l, in, u_l, g_in, g_out, o);
  } else if (in.kind == kInWide && in.img == 1 && g_out == nullptr && (c.p->cfg.tensor_cores & 4)) {
    bwd_wide_img(c, net, l, u_l, g_in, o, c.sc + c.p->sl.xm, c.sc + c.p->sl.xref + c.crank * kMaxDim, nullptr, 0, 0);
  } else if (in.kind == kInWide && in.img == 2 && g_out != nullptr && (c.p->cfg.tensor_cores & 4)) {
    bwd_wide_img(c, net, l, u_l, g_in, o, c.sc + c.p->sl.ym, c.sc + c.p->sl.yref + c.crank * kMaxDim, g_out, in.ld, in.act);
  } else if (in.kind == kInLatent) {
    bwd_latent(c, net, l, in, u_l, g_in, g_out, o);
  } else {
    bwd_hidden_edge(c, net, l, in, u_l, g_in, g_out, o);
  }
}

// FCEncoder backward from sc.dz; x = the encoder's input rows.  dx_out != null (MI phase): the decoder
// pre-activation panel that receives dL/dv in place.
__device__ __forceinline__ void encoder_backward(const Ctx& c, const LayerIn& x, int inst, int o, float* dx_out) {
  const raae_net_layout& nl = NL(c, kE);
  const int L = nl.n_linear;
  float* g0 = c.sc + c.p->sl.g[0];
  float* g1 = c.sc + c.p->sl.g[1];
  bwd_enc_last(c, hidden_out(c, kE, L - 2, inst), g0, o);
  float* gin = g0;
  float* gout = g1;
  for (int l = L - 2; l >= 1; --l) {
    bwd_hidden(c, kE, l, hidden_out(c, kE, l - 1, inst), c.sc + c.p->sl.uE[l], gin, gout, o);
    float* t = gin; gin = gout; gout = t;
  }
  bwd_hidden(c, kE, 0, x, c.sc + c.p->sl.uE[0], gin, dx_out, o);
}

// hidden blocks of the decoder backward; on entry g[0] / sm->sg,sgx hold the gradient w.r.t. the last
// hidden block's output.  dz_out != null: gradient w.r.t. the latent input -> sc.dz
__device__ __forceinline__ void decoder_backward_hidden(const Ctx& c, const LayerIn& z, int inst, int o, float* dz_out) {
  const raae_net_layout& nl = NL(c, kD);
  const int L = nl.n_linear;
  float* gin = c.sc + c.p->sl.g[0];
  float* gout = c.sc + c.p->sl.g[1];
  for (int l = L - 2; l >= 1; --l) {
    bwd_hidden(c, kD, l, hidden_out(c, kD, l - 1, inst), c.sc + c.p->sl.uD[l], gin, gout, o);
    float* t = gin; gin = gout; gout = t;
  }
  bwd_hidden(c, kD, 0, z, c.sc + c.p->sl.uD[0], gin, dz_out, o);
}

}  // namespace raae
