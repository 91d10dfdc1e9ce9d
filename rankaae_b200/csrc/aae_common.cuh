// Shared definitions of the fused sm_100a AAE train step (device side).
//
// Reference semantics: /root/reference/sc/clustering/trainer.py:103-304 (loop body + eval block),
// model.py:330-378 (FCEncoder), 518-570 (FCDecoder), 631-663 (DiscriminatorFC), functions.py:37-219.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "rankaae_b200.h"

// The device code is compiled twice: RAAE_CLUSTER == 0 (rankaae_b200.cu: one CTA per trial, the ensemble path - cluster rank 0
// and size 1 are compile-time constants, so none of the cluster code exists there) and RAAE_CLUSTER == 1
// (kernels_cluster.cu: ctas_per_trial 2 / 4 / 8, namespace raae_cn).
#ifndef RAAE_CLUSTER
#define RAAE_CLUSTER 0
#endif
#if RAAE_CLUSTER
#define raae raae_cn
#endif

namespace raae {

constexpr int kH = RAAE_HIDDEN;   // hidden width
constexpr int kTM = 128;          // batch rows per tile
constexpr int kLD = 68;           // smem leading dimension of a [rows][64] tile (16 B aligned, conflict-free)
constexpr int kLDW = 260;         // smem leading dimension of a [rows][256] tile
constexpr int kThreads = 256;
constexpr int kZ = RAAE_ZPAD;     // padded latent width
constexpr int kMaxDim = 256;      // max dim_in / dim_out
constexpr float kBnEps = 1e-5f;
constexpr float kBnMomentum = 0.1f;
constexpr float kAdamEps = 1e-8f;

enum Phase { kAdv = 0, kCorr = 1, kRecon = 2, kMI = 3, kSmooth = 4 };
enum NetId { kE = 0, kD = 1, kS = 2 };

// scratch block sub-buffers (offsets in floats from the trial's scratch base)
struct ScratchLayout {
  int xn;                       // [rows][xld]   noised spectra of the batch (trainer.py:112)
  int aux;                      // [rows][kZ]    descriptors of the batch
  int uE[RAAE_MAX_LAYERS];      // [rows][64]    pre-activation of encoder hidden layer l
  int zE;                       // [rows][kZ]    pre-BN output of the last encoder Linear
  int dz;                       // [rows][kZ]    gradient w.r.t. the encoder output
  int uD[RAAE_MAX_LAYERS];      // [rows][64]    pre-activation of decoder hidden layer l
  int v;                        // [rows][vld]   pre-activation of the decoder output (MI phase) / its gradient
  int g[2];                     // [rows][64]    gradient w.r.t. a hidden block's BN output (ping-pong)
  int zs;                       // [rows][kZ]    z_sample (MI phase)
  int rank;                     // [kZ][rows]    validation: per-style ranks (Spearman)
  int xld, vld;                 // padded row strides
  // tensor-core operand images of the noised batch, written once per step by build_batch and streamed into shared
  // memory with bulk asynchronous copies (no staging work in the consuming stages); see DESIGN.md
  int xk;                       // [tiles][nch64][8192] fp32 (batch - xref), K-major SWIZZLE_128B blocks (forward of the input layer)
  int xm;                       // [tiles][nch128][16384] fp32 (batch - xref), MN-major SW128_32B blocks (weight gradient of the input layer)
  int wk;                       // [nch64][hi 4096 | lo 4096] K-major image of the input layer's weights (rebuilt by each forward)
  int yk;                       // like xk, for y = act(v) of the MI phase (written by the decoder output stage, mode kLastStoreV)
  int ym;                       // like xm, for y (reserved for the MI-phase weight gradient)
  int wl;                       // [nch64 of dim_out][hi 4096 | lo 4096] K-major image of the decoder output weights (rebuilt per call)
  int yref;                     // [kMaxDim] reference row of the y image (column means of the first rows of act(v))
  int xref;                     // [kMaxDim] reference row the images are centred on (mean of the first rows of the batch)
  int nch64, nch128;
  int total;
};

struct KParams {
  raae_layout lay;
  raae_config cfg;
  ScratchLayout sl;
  float* state;
  float* scratch;
  const double* hp;
  const float* spec_train;
  const float* aux_train;
  int n_train;
  const float* spec_val;
  const float* aux_val;
  int n_val;
  const float* shapiro_w;
};

struct RunArgs {
  int trial0;                   // first trial handled by blockIdx.x == 0
  int epoch;
  int n_steps;                  // batches in this launch (production) — ignored in debug mode
  int debug;                    // 1 = teacher-forced single step driven by `dbg`
  const int32_t* perm;          // [n_trials][n_train] shuffled row order of this epoch (production)
  float* out_losses;            // [n_trials][12] (production, may be null)
  float* out_metrics;           // [n_trials][6]
  long long* prof;              // optional [n_trials][32] stage cycle counters (raae_set_profile_buffer)
  // split-phase (data-parallel) launches: one batch `step0`, phases `phase_mask`, gradients exported, no update
  int split;                    // 1 = split-phase launch
  int step0;                    // batch index inside the epoch
  int phase_mask;
  float* grads_out[RAAE_NUM_PHASES];   // [n_trials][opt[o].n] per phase, optimizer parameter order
  raae_debug_io dbg;
  raae_val_io val;
};

// ------------------------------------------------------------------------------------------
// counter-based RNG (production mode).  Parity tests inject every draw explicitly, so the generator
// only has to be statistically sound: two rounds of a 32-bit avalanche mixer keyed per stream.
// ------------------------------------------------------------------------------------------
__host__ __device__ __forceinline__ uint32_t mix32(uint32_t x) {
  x ^= x >> 16; x *= 0x7feb352dU; x ^= x >> 15; x *= 0x846ca68bU; x ^= x >> 16;
  return x;
}
__device__ __forceinline__ uint32_t hash2(uint32_t key, uint32_t idx) { return mix32(mix32(idx ^ key) + key * 0x9e3779b9U); }

enum StreamKind { kStreamXNoise = 1, kStreamEncMask = 16, kStreamDecMask = 80, kStreamDisMask = 120,
                  kStreamZReal = 140, kStreamDisEpsReal = 141, kStreamDisEpsFake = 142, kStreamZSample = 143,
                  kStreamValZSample = 150, kStreamValZReal = 151 };

__device__ __forceinline__ uint32_t stream_key(uint32_t seed, uint32_t step_id, uint32_t kind) {
  return mix32(seed ^ mix32(step_id * 256u + kind));
}

// hardware square root (<= 1 ulp, no slow path - and so no branch - around it)
__device__ __forceinline__ float sqrt_approx_f(float x) {
  float r;
  asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
}

// standard normal for element idx of a stream (Box-Muller on two hashed uniforms per pair)
__device__ __forceinline__ float normal_at(uint32_t key, uint32_t idx) {
  uint32_t pair = idx >> 1;
  uint32_t h1 = hash2(key, 2u * pair), h2 = hash2(key, 2u * pair + 1u);
  float u1 = ((h1 >> 8) + 0.5f) * (1.0f / 16777216.0f);
  float u2 = ((h2 >> 8) + 0.5f) * (1.0f / 16777216.0f);
  float r = sqrt_approx_f(-2.0f * __logf(u1));
  float s, c;
  __sincosf(6.28318530717958647692f * u2, &s, &c);
  return (idx & 1u) ? r * s : r * c;
}

// both normals of pair `pair` of a stream: elements 2 pair and 2 pair + 1 of normal_at
__device__ __forceinline__ void normal_pair(uint32_t key, uint32_t pair, float& n0, float& n1) {
  uint32_t h1 = hash2(key, 2u * pair), h2 = hash2(key, 2u * pair + 1u);
  float u1 = ((h1 >> 8) + 0.5f) * (1.0f / 16777216.0f);
  float u2 = ((h2 >> 8) + 0.5f) * (1.0f / 16777216.0f);
  float r = sqrt_approx_f(-2.0f * __logf(u1));
  float s, c;
  __sincosf(6.28318530717958647692f * u2, &s, &c);
  n0 = r * c;
  n1 = r * s;
}

struct MaskSrc {
  const uint8_t* ptr;   // explicit keep-mask [rows][64] or null
  uint32_t key;
  uint32_t thresh;      // drop when the 16-bit draw < thresh = round(p * 65536); 0 = keep everything
  float scale;          // 1 / (1 - p)
};

// Branch-free Bernoulli draws: one mixer round per PAIR of channels, 16 bits each; a channel is dropped when its
// draw < thresh = round(p * 65536).
__device__ __forceinline__ uint32_t mask_word(uint32_t key, uint32_t pair) { return mix32(pair * 0x9e3779b1U + key); }

__device__ __forceinline__ bool mask_keep(const MaskSrc& m, int row, int c) {
  if (m.ptr) return m.ptr[row * kH + c] != 0;
  uint32_t idx = (uint32_t)(row * kH + c);
  uint32_t w = mask_word(m.key, idx >> 1);
  uint32_t d = (idx & 1u) ? (w >> 16) : (w & 0xffffu);
  return d >= m.thresh;          // thresh == 0 keeps everything
}

// four consecutive channels c..c+3 (c % 4 == 0): returns keep bits
__device__ __forceinline__ uint32_t mask_keep4(const MaskSrc& m, int row, int c) {
  if (m.ptr) {
    uint32_t w = *reinterpret_cast<const uint32_t*>(m.ptr + row * kH + c);
    return ((w & 0xffu) ? 1u : 0u) | ((w & 0xff00u) ? 2u : 0u) | ((w & 0xff0000u) ? 4u : 0u) | ((w & 0xff000000u) ? 8u : 0u);
  }
  if (m.thresh == 0u) return 0xfu;
  uint32_t pair = (uint32_t)(row * kH + c) >> 1;
  uint32_t w0 = mask_word(m.key, pair), w1 = mask_word(m.key, pair + 1u);
  return ((w0 & 0xffffu) >= m.thresh ? 1u : 0u) | ((w0 >> 16) >= m.thresh ? 2u : 0u) |
         ((w1 & 0xffffu) >= m.thresh ? 4u : 0u) | ((w1 >> 16) >= m.thresh ? 8u : 0u);
}

// Keep bits of the kTM / 16 row groups a staging thread owns (rows r0 + 16 i of the tile that starts at row0, channels
// c..c+3): bits [4 i, 4 i + 4).  Rows >= nv read as dropped.  Same draws as mask_keep4; the source of the mask is decided
// ONCE (one uniform branch) and the hashed path is straight-line code - sixteen independent mixer chains the scheduler can
// interleave - whereas a mask_keep4 call per row group left a branch diamond per group in the unrolled staging loops and
// serialised them (two warps per scheduler cannot hide a dependent chain of ~25 integer operations per group).
__device__ __forceinline__ uint32_t mask_keep4_rows(const MaskSrc& m, int row0, int r0, int c, int nv) {
  uint32_t bits = 0u;
  if (m.ptr) {
    uint32_t w[kTM / 16];
#pragma unroll
    for (int i = 0; i < kTM / 16; ++i) {
      const int r = r0 + 16 * i;
      w[i] = r < nv ? *reinterpret_cast<const uint32_t*>(m.ptr + (size_t)(row0 + r) * kH + c) : 0u;
    }
#pragma unroll
    for (int i = 0; i < kTM / 16; ++i)
      bits |= (((w[i] & 0xffu) ? 1u : 0u) | ((w[i] & 0xff00u) ? 2u : 0u) | ((w[i] & 0xff0000u) ? 4u : 0u) |
               ((w[i] & 0xff000000u) ? 8u : 0u)) << (4 * i);
    return bits;
  }
  if (m.thresh == 0u) {
    bits = 0xffffffffu;
  } else {
    const uint32_t pair0 = (uint32_t)((row0 + r0) * kH + c) >> 1;
#pragma unroll
    for (int i = 0; i < kTM / 16; ++i) {
      const uint32_t pair = pair0 + (uint32_t)(i * 16 * kH / 2);
      const uint32_t w0 = mask_word(m.key, pair), w1 = mask_word(m.key, pair + 1u);
      bits |= (((w0 & 0xffffu) >= m.thresh ? 1u : 0u) | ((w0 >> 16) >= m.thresh ? 2u : 0u) |
               ((w1 & 0xffffu) >= m.thresh ? 4u : 0u) | ((w1 >> 16) >= m.thresh ? 8u : 0u)) << (4 * i);
    }
  }
  // rows r0 + 16 i >= nv: groups i >= ceil((nv - r0) / 16)
  const int nvalid = nv > r0 ? (nv - r0 + 15) >> 4 : 0;
  return nvalid >= kTM / 16 ? bits : (bits & ((1u << (4 * nvalid)) - 1u));
}

// ------------------------------------------------------------------------------------------
// scalar math restated from SURVEY.md Appendix A
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ float prelu_f(float u, float a) { return u > 0.f ? u : a * u; }

// nn.Softplus(beta=2, threshold=20): model.py:535.  softplus(x) = max(x, 0) + log1p(exp(-|x|)) keeps the
// argument of the logarithm in (1, 2], where the hardware exp/log pair is accurate to ~1e-7 absolute.
// Branch-free: beyond the threshold (2 v > 20) exp(-2 v) < 2.1e-9 vanishes against 1 in float32, so the same expression
// returns 0.5 * (2 v + 0) = v exactly, which is what the thresholded module returns.  (Written as `bv > 20 ? v : ...` it
// compiled into a branch diamond per element that serialised the eight elements of a row-pass iteration.)
__device__ __forceinline__ float softplus2_f(float v) {
  float bv = 2.f * v;
  return 0.5f * (fmaxf(bv, 0.f) + __logf(1.f + __expf(-fabsf(bv))));
}
__device__ __forceinline__ float sigmoid_f(float x) {
  float e = __expf(-fabsf(x));
  float s = __fdividef(1.f, 1.f + e);
  return x >= 0.f ? s : e * s;
}
__device__ __forceinline__ float softplus2_grad_f(float v) {       // sigmoid(2 v); == 1.f beyond the threshold (1 + 2.1e-9 == 1)
  return sigmoid_f(2.f * v);
}

// One AdamW element update (torch.optim.AdamW: decoupled decay, exp_avg, exp_avg_sq, bias-corrected step), shared by the fused
// stages and the stand-alone optimizer kernels so that every path produces the same bits.  Square root and the two
// divisions are the hardware approximations (<= 2 ulp each, i.e. <= 1e-9 of the parameter per step at the usual
// lr / |p| ratios; flush-to-zero only matters where eps = 1e-8 dominates the denominator anyway): the IEEE versions carry
// slow-path calls, which put a branch diamond around every element of the unrolled update loop and serialised it.
__device__ __forceinline__ void adamw_update(float& p, float& m, float& v, float g, float decay, float w1, float b2, float w2, float ss,
                                             float bc2s) {
  // explicit roundings (fmaf / __fmul_rn): left to the compiler, `v b2 + (w2 g) g` contracts one way in one call site and the
  // other way in another, and the split-phase path would drift from the fused one by an ulp per step
  const float pp = __fmul_rn(p, decay);
  m = fmaf(g - m, w1, m);
  v = fmaf(__fmul_rn(w2, g), g, __fmul_rn(v, b2));
  const float denom = __fdividef(sqrt_approx_f(v), bc2s) + kAdamEps;
  p = fmaf(-ss, __fdividef(m, denom), pp);
}

// 17-tap Gaussian (sigma 3) exactly as torch builds it in float32 (model.py:186-206)
__device__ __constant__ float kGauss17[17] = {
    0.0038155282381922007f, 0.008779441937804222f, 0.018076900392770767f, 0.03330628201365471f,
    0.05491277202963829f,   0.08101504296064377f,  0.10695548355579376f,  0.1263529658317566f,
    0.13357123732566833f,   0.1263529658317566f,   0.10695548355579376f,  0.08101504296064377f,
    0.05491277202963829f,   0.03330628201365471f,  0.018076900392770767f, 0.008779441937804222f,
    0.0038155282381922007f};

// ------------------------------------------------------------------------------------------
// register-tile SIMT contractions on smem tiles (FP32 FMA: the parity mode)
// ------------------------------------------------------------------------------------------
// C[m][n] += sum_k A[m*lda + k] * B[n*ldb + k]     m = ty + 16 i (i < 8), n = tx + 16 j (j < 4)
template <int K>
__device__ __forceinline__ void mma_nt(const float* __restrict__ A, int lda, const float* __restrict__ B, int ldb,
                                       float (&acc)[8][4], int ty, int tx) {
#pragma unroll 2
  for (int k = 0; k < K; k += 4) {
    float4 a[8], b[4];
#pragma unroll
    for (int i = 0; i < 8; ++i) a[i] = *reinterpret_cast<const float4*>(A + (ty + 16 * i) * lda + k);
#pragma unroll
    for (int j = 0; j < 4; ++j) b[j] = *reinterpret_cast<const float4*>(B + (tx + 16 * j) * ldb + k);
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        acc[i][j] = fmaf(a[i].x, b[j].x, acc[i][j]);
        acc[i][j] = fmaf(a[i].y, b[j].y, acc[i][j]);
        acc[i][j] = fmaf(a[i].z, b[j].z, acc[i][j]);
        acc[i][j] = fmaf(a[i].w, b[j].w, acc[i][j]);
      }
  }
}

// C[m][n] += sum_k A[m*lda + k] * B[k*ldb + n]     m = ty + 16 i (i < 8), n = 4 tx + j (j < 4)
template <int K>
__device__ __forceinline__ void mma_nn(const float* __restrict__ A, int lda, const float* __restrict__ B, int ldb,
                                       float (&acc)[8][4], int ty, int tx) {
#pragma unroll 2
  for (int k = 0; k < K; k += 4) {
    float4 a[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) a[i] = *reinterpret_cast<const float4*>(A + (ty + 16 * i) * lda + k);
#pragma unroll
    for (int kk = 0; kk < 4; ++kk) {
      float4 b = *reinterpret_cast<const float4*>(B + (k + kk) * ldb + 4 * tx);
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        float av = kk == 0 ? a[i].x : kk == 1 ? a[i].y : kk == 2 ? a[i].z : a[i].w;
        acc[i][0] = fmaf(av, b.x, acc[i][0]);
        acc[i][1] = fmaf(av, b.y, acc[i][1]);
        acc[i][2] = fmaf(av, b.z, acc[i][2]);
        acc[i][3] = fmaf(av, b.w, acc[i][3]);
      }
    }
  }
}

// C[m0+i][n0+j] += sum_{r in [r0, r1)} A[r*lda + m0 + i] * B[r*ldb + n0 + j]      i, j < 8
__device__ __forceinline__ void mma_tn8(const float* __restrict__ A, int lda, int m0, const float* __restrict__ B,
                                        int ldb, int n0, int r0, int r1, float (&acc)[8][8]) {
#pragma unroll 2
  for (int r = r0; r < r1; ++r) {
    float4 a0 = *reinterpret_cast<const float4*>(A + r * lda + m0);
    float4 a1 = *reinterpret_cast<const float4*>(A + r * lda + m0 + 4);
    float4 b0 = *reinterpret_cast<const float4*>(B + r * ldb + n0);
    float4 b1 = *reinterpret_cast<const float4*>(B + r * ldb + n0 + 4);
    float av[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
    float bv[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
      for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
  }
}

// ------------------------------------------------------------------------------------------
// asynchronous global -> shared copies (LDGSTS): tile prefetch that needs no registers
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ void cp_async16(float* smem_dst, const float* gmem_src) {
  uint32_t d = (uint32_t)__cvta_generic_to_shared(smem_dst);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(d), "l"(gmem_src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;\n" ::"n"(N) : "memory"); }

// rows [row0, row0 + nv) of a row-major [rows][64] panel -> raw smem tile [kTM][kLD]; every thread copies the SAME
// elements it later transforms in place (c4 = 4 (tid & 15), rows (tid >> 4) + 16 i), so no barrier is needed in between
__device__ __forceinline__ void prefetch_panel_tile(float* __restrict__ dst, const float* __restrict__ panel, int row0, int nv) {
  const int c4 = (threadIdx.x & 15) * 4, r0 = threadIdx.x >> 4;
#pragma unroll
  for (int i = 0; i < kTM / 16; ++i) {
    const int r = r0 + 16 * i;
    if (r < nv) cp_async16(dst + r * kLD + c4, panel + (size_t)(row0 + r) * kH + c4);
  }
}

// block-wide sum, result broadcast to every thread; `red` holds >= 8 floats of smem.
__device__ __forceinline__ float block_sum(float v, float* red) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  __syncthreads();
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
  __syncthreads();
  float s = 0.f;
#pragma unroll
  for (int w = 0; w < kThreads / 32; ++w) s += red[w];
  return s;
}

// ------------------------------------------------------------------------------------------
// thread-block cluster per trial (raae_config::ctas_per_trial > 1): cluster rank / barrier / distributed shared memory
// ------------------------------------------------------------------------------------------
namespace cl {
__device__ __forceinline__ uint32_t ctarank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ uint32_t nctarank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_nctarank;" : "=r"(r)); return r; }
// every thread of every CTA of the cluster; release / acquire at cluster scope: global and shared-memory writes made before
// the barrier by any CTA are visible to every CTA after it (ptxas invalidates L1 on the acquire side)
__device__ __forceinline__ void sync() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// address of this CTA's shared-memory location `p` in the shared memory of cluster rank `rank`
__device__ __forceinline__ uint32_t map(const void* p, uint32_t rank) {
  uint32_t a = (uint32_t)__cvta_generic_to_shared(p), r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(rank));
  return r;
}
__device__ __forceinline__ float ld_f32(uint32_t a) { float v; asm volatile("ld.shared::cluster.f32 %0, [%1];" : "=f"(v) : "r"(a) : "memory"); return v; }
__device__ __forceinline__ double ld_f64(uint32_t a) { double v; asm volatile("ld.shared::cluster.f64 %0, [%1];" : "=d"(v) : "r"(a) : "memory"); return v; }
// rows are owned tile-wise: 128-row tile t belongs to cluster rank t % csize in EVERY stage, so per-row panels never cross CTAs
__device__ __forceinline__ int own_tiles(int B, int crank, int csize) {
  const int nt = (B + kTM - 1) / kTM;
  return nt > crank ? (nt - crank + csize - 1) / csize : 0;
}
__device__ __forceinline__ int own_rows(int B, int crank, int csize) {
  const int nt = (B + kTM - 1) / kTM;
  if (crank >= nt) return 0;
  int rows = ((nt - 1 - crank) / csize + 1) * kTM;
  if ((nt - 1) % csize == crank) rows -= nt * kTM - B;          // the short last tile is this rank's
  return rows;
}
// row of own-row slot s (slots of a CTA: own_tiles x 128; the row may be >= B in the last tile)
__device__ __forceinline__ int slot_row(int s, int crank, int csize) { return ((s >> 7) * csize + crank) * kTM + (s & (kTM - 1)); }
}  // namespace cl

__device__ __forceinline__ double block_sum_d(double v, double* red) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  __syncthreads();
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
  __syncthreads();
  double s = 0.0;
#pragma unroll
  for (int w = 0; w < kThreads / 32; ++w) s += red[w];
  return s;
}

}  // namespace raae
