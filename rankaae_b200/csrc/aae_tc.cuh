// tcgen05 (5th-generation tensor core) path for the 128 x 64 x 64 contractions of the hidden blocks.
//
// Operands are staged in shared memory in the canonical K-major SWIZZLE_128B layout and described to the tensor
// core by 64-bit matrix descriptors; the FP32 accumulator (128 lanes x 64 columns) lives in tensor memory (TMEM)
// and is read back with tcgen05.ld.  FP32 accuracy is kept with the 3 x TF32 error-compensated split
//   a b ~= a_hi b_hi + a_lo b_hi + a_hi b_lo,     x_hi = x rounded to TF32 (nearest), x_lo = x - x_hi
// (single-pass TF32 fails parity on these networks: SURVEY.md Appendix D-6).
//
// Bit layouts follow cute/arch/mma_sm100_desc.hpp (UMMA::SmemDescriptor / UMMA::InstrDescriptor) of the CUTLASS
// headers vendored in this image; nothing is included from them.
#pragma once
#include "aae_common.cuh"

namespace raae {
namespace tc {

constexpr int kTmemCols = 512;                // the whole TMEM of the SM (one CTA per SM); every stage carves its own accumulators
constexpr int kABlockBytes = kTM * 128;       // one 32-float K block of a 128-row operand
constexpr int kBBlockBytes = kH * 128;        // one 32-float K block of a 64-row operand
constexpr int kATileFloats = kTM * kH;        // 8192 floats: two K blocks
constexpr int kBTileFloats = kH * kH;         // 4096 floats

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// ---- TMEM allocation (one warp) ----
__device__ __forceinline__ void tmem_alloc(uint32_t* dst_smem, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)), "r"(ncols) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void fence_before_sync() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void fence_after_sync() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// One lane of a converged warp.  Issuing the MMAs from `warp_uniform_id() == 0 && elect_one()` instead of `tid == 0`
// lets the compiler keep the descriptors in uniform registers (no per-instruction ELECT loop around UTCHMMA).
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred));
  return pred != 0;
}
__device__ __forceinline__ int warp_uniform_id() { return __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0); }

// ---- mbarrier ----
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  uint32_t done = 0, addr = smem_u32(bar), spins = 0;
  while (!done) {
    if (++spins > (1u << 22)) asm volatile("trap;");      // a lost commit must fail loudly, never hang the GPU
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.b32 %0, 1, 0, p;\n\t}"
        : "=r"(done) : "r"(addr), "r"(parity) : "memory");
  }
}
// ---- bulk asynchronous copies global -> shared (UBLKCP), completion counted in bytes on an mbarrier ----
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* smem_dst, const void* gmem_src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(smem_u32(smem_dst)), "l"(gmem_src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
// generic-proxy writes (to global or shared memory) -> visible to later async-proxy reads (bulk copies, tensor core)
__device__ __forceinline__ void fence_async_all() { asm volatile("fence.proxy.async;" ::: "memory"); }

// arrives on `bar` when every previously issued tcgen05.mma of this thread has completed
__device__ __forceinline__ void mma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}

// ---- descriptors ----
// K-major, SWIZZLE_128B: rows of 128 B, 8-row groups 1024 B apart (SBO), LBO unused (1), descriptor version 1
__device__ __forceinline__ uint64_t make_desc_k_sw128(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr >> 4) & 0x3fffu);      // start address, bits [0,14)
  d |= (uint64_t)1u << 16;                          // leading byte offset (unused for swizzled K-major), bits [16,30)
  d |= (uint64_t)(1024u >> 4) << 32;                // stride byte offset, bits [32,46)
  d |= (uint64_t)1u << 46;                          // version = 1 (Blackwell), bits [46,48)
  d |= (uint64_t)2u << 61;                          // layout type SWIZZLE_128B, bits [61,64)
  return d;
}
// MN-major operands of 32-bit types must use the SWIZZLE_128B_BASE32B layout (CUTLASS sm100_smem_selector: "for
// mn-major tf32 operands, SW128_32B is the only available smem layout"): 32 contiguous MN elements per 128 B row, the
// 32-byte chunk index XOR-ed with (row & 3), 4-row K groups `sbo_bytes` apart, MN blocks `lbo_bytes` apart.
__device__ __forceinline__ uint64_t make_desc_mn_sw128_32b(uint32_t smem_addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr >> 4) & 0x3fffu);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3fffu) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3fffu) << 32;
  d |= (uint64_t)1u << 46;
  d |= (uint64_t)1u << 61;                          // layout type SWIZZLE_128B_BASE32B
  return d;
}
// byte offset of the 16-byte chunk holding elements (row, n .. n+3), n % 4 == 0, of an MN-major SW128_32B tile
// [rows (K)][64 (MN)] stored as two 32-column blocks of [rows][128 B]
__device__ __forceinline__ uint32_t sw128_32b_chunk_off(int row, int n, int block_bytes) {
  return (uint32_t)((n >> 5) * block_bytes + row * 128 + (((((n & 31) >> 3) ^ (row & 3))) << 5) + ((n & 7) << 2));
}
// kind::tf32, FP32 accumulate, A and B MN-major, M = 64, N = 64 (weight gradient du^T a, K = batch rows)
constexpr uint32_t kIdescTf32_TN_64x64 = (1u << 4) | (2u << 7) | (2u << 10) | (1u << 15) | (1u << 16) | ((64u >> 3) << 17) | ((64u >> 4) << 24);
// kind::tf32, FP32 accumulate, A and B K-major, M = 128, N = 64
constexpr uint32_t kIdescTf32_128x64 = (1u << 4) | (2u << 7) | (2u << 10) | ((64u >> 3) << 17) | ((128u >> 4) << 24);

__device__ __forceinline__ void mma_tf32(uint32_t d_tmem, uint64_t da, uint64_t db, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(d_tmem), "l"(da), "l"(db), "r"(idesc), "r"(accumulate) : "memory");
}

// ---- operand staging ----
// byte offset of the 16-byte chunk holding elements (row, k .. k+3), k % 4 == 0, of a K-major SW128 tile
__device__ __forceinline__ uint32_t sw128_chunk_off(int row, int k, int block_bytes) {
  return (uint32_t)((k >> 5) * block_bytes + row * 128 + ((((k & 31) >> 2) ^ (row & 7)) << 4));
}
// Round-to-nearest TF32 split x = hi + lo (+ <= 2^-24 |x|): hi = rna_tf32(x), lo = rna_tf32(x - hi).  The tensor core
// TRUNCATES the 13 low mantissa bits of its operands (tools/tc_probe.cu variant 6); splitting by truncation leaves a
// one-sided error of ~2^-21 |x| per operand that adds up coherently over K (measured: 2e-6 absolute on the K = 256
// input layer, 50 x the FP32-FMA error, enough to fail gradient parity on its low-variance channels), whereas both
// rounded halves are exact TF32 numbers, so nothing is truncated and the residual has a random sign.
// round to nearest (ties away from zero) to 10 mantissa bits, i.e. cvt.rna.tf32.f32, done with two full-rate integer
// instructions: the conversion instruction issues at a fraction of the ALU rate and the staging passes need two per
// element (ncu: math-pipe throttle on that line).  Adding half a TF32 ulp to the sign-magnitude bit pattern and clearing
// the 13 low bits is exact for finite inputs (a carry into the exponent is the correct round-up to the next binade).
__device__ __forceinline__ float tf32_rna(float x) {
  return __uint_as_float((__float_as_uint(x) + 0x1000u) & 0xffffe000u);
}
// The low half is handed to the tensor core as the exact FP32 residual x - hi (so hi + lo == x bit for bit) and the hardware
// truncates it to TF32 itself: |x - hi| <= 2^-11 |x|, the truncation drops at most 2^-10 of that, and because hi was ROUNDED
// the residual - and with it the truncation error - has a random sign (truncating x itself is what adds up coherently).
// Two integer instructions per element less than rounding the residual as well; the dropped term stays at the level of the
// lo x lo product the 3 x TF32 scheme omits anyway (2^-22 |x|).
__device__ __forceinline__ void tf32_split(float x, float& hi, float& lo) {
  hi = tf32_rna(x);
  lo = x - hi;
}
__device__ __forceinline__ void split_store(float* hi_base, float* lo_base, uint32_t off_bytes, float4 v) {
  float4 h, l;
  tf32_split(v.x, h.x, l.x);
  tf32_split(v.y, h.y, l.y);
  tf32_split(v.z, h.z, l.z);
  tf32_split(v.w, h.w, l.w);
  *reinterpret_cast<float4*>(reinterpret_cast<char*>(hi_base) + off_bytes) = h;
  *reinterpret_cast<float4*>(reinterpret_cast<char*>(lo_base) + off_bytes) = l;
}

// D[128 x 64] (TMEM) (+)= A[128 x 64] * B[64 x 64]^T with the 3 x TF32 split; single thread
__device__ __forceinline__ void issue_gemm_3xtf32_acc(uint32_t d_tmem, const float* a_hi, const float* a_lo, const float* b_hi,
                                                      const float* b_lo, uint32_t accumulate_first) {
  const uint32_t ah = smem_u32(a_hi), al = smem_u32(a_lo), bh = smem_u32(b_hi), bl = smem_u32(b_lo);
  uint32_t first = accumulate_first;
#pragma unroll
  for (int pass = 0; pass < 3; ++pass) {
    const uint32_t abase = pass == 0 ? al : ah;           // a_lo b_hi, a_hi b_lo, a_hi b_hi (small terms first)
    const uint32_t bbase = pass == 1 ? bl : bh;
#pragma unroll
    for (int s = 0; s < kH / 8; ++s) {                    // K = 8 per instruction (32 bytes)
      const uint32_t koff = (uint32_t)((s & 3) * 32);
      uint64_t da = make_desc_k_sw128(abase + (s >> 2) * kABlockBytes + koff);
      uint64_t db = make_desc_k_sw128(bbase + (s >> 2) * kBBlockBytes + koff);
      mma_tf32(d_tmem, da, db, kIdescTf32_128x64, first);
      first = 1;
    }
  }
}

__device__ __forceinline__ void issue_gemm_3xtf32(uint32_t d_tmem, const float* a_hi, const float* a_lo, const float* b_hi,
                                                  const float* b_lo) {
  issue_gemm_3xtf32_acc(d_tmem, a_hi, a_lo, b_hi, b_lo, 0u);
}

// D[64 x 64] (TMEM, M = 64 layout) (+)= A^T B over the 128 rows of two [128][64] tiles staged MN-major (SW128_32B),
// 3 x TF32 split; single thread
__device__ __forceinline__ void issue_gemm_tn_3xtf32(uint32_t d_tmem, const float* a_hi, const float* a_lo, const float* b_hi,
                                                     const float* b_lo, uint32_t accumulate_first) {
  const uint32_t ah = smem_u32(a_hi), al = smem_u32(a_lo), bh = smem_u32(b_hi), bl = smem_u32(b_lo);
  uint32_t acc = accumulate_first;
#pragma unroll
  for (int pass = 0; pass < 3; ++pass) {
    const uint32_t abase = pass == 0 ? al : ah;
    const uint32_t bbase = pass == 1 ? bl : bh;
#pragma unroll
    for (int s = 0; s < kTM / 8; ++s) {                   // K = 8 batch rows per instruction = one 1 KB row group
      uint64_t da = make_desc_mn_sw128_32b(abase + s * 1024, kABlockBytes, 512);
      uint64_t db = make_desc_mn_sw128_32b(bbase + s * 1024, kABlockBytes, 512);
      mma_tf32(d_tmem, da, db, kIdescTf32_TN_64x64, acc);
      acc = 1;
    }
  }
}

// D[128 x 64] (TMEM, identity row map) (+)= A^T B with A = [128 rows (K)][128 (M)] as four 32-column blocks `a_block_bytes`
// apart and B = [128 rows (K)][64 (N)] as two blocks `b_block_bytes` apart, both MN-major SW128_32B; ONE product of the
// 3 x TF32 split per call (the caller streams the hi / lo halves of A separately); single thread
constexpr uint32_t kIdescTf32_TN_128x64 = (1u << 4) | (2u << 7) | (2u << 10) | (1u << 15) | (1u << 16) | ((64u >> 3) << 17) | ((128u >> 4) << 24);
__device__ __forceinline__ void issue_gemm_tn128_pass(uint32_t d_tmem, const float* a, uint32_t a_block_bytes, const float* b,
                                                      uint32_t b_block_bytes, uint32_t accumulate_first) {
  const uint32_t a0 = smem_u32(a), b0 = smem_u32(b);
  uint32_t acc = accumulate_first;
#pragma unroll
  for (int s = 0; s < kTM / 8; ++s) {
    uint64_t da = make_desc_mn_sw128_32b(a0 + s * 1024, a_block_bytes, 512);
    uint64_t db = make_desc_mn_sw128_32b(b0 + s * 1024, b_block_bytes, 512);
    mma_tf32(d_tmem, da, db, kIdescTf32_TN_128x64, acc);
    acc = 1;
  }
}

// this thread's row (TMEM lane) of 32 consecutive accumulator columns
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, float (&v)[32]) {
  uint32_t r[32];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr) : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
}

// this thread's row (TMEM lane): 32 consecutive columns written from registers (TMEM used as storage: the decoder output stage
// parks its 128 x 256 gradient tile there while shared memory holds the operand planes)
__device__ __forceinline__ void tmem_st32(uint32_t taddr, const float (&v)[32]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
      "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};"
      ::"r"(taddr), "r"(__float_as_uint(v[0])), "r"(__float_as_uint(v[1])), "r"(__float_as_uint(v[2])), "r"(__float_as_uint(v[3])),
        "r"(__float_as_uint(v[4])), "r"(__float_as_uint(v[5])), "r"(__float_as_uint(v[6])), "r"(__float_as_uint(v[7])),
        "r"(__float_as_uint(v[8])), "r"(__float_as_uint(v[9])), "r"(__float_as_uint(v[10])), "r"(__float_as_uint(v[11])),
        "r"(__float_as_uint(v[12])), "r"(__float_as_uint(v[13])), "r"(__float_as_uint(v[14])), "r"(__float_as_uint(v[15])),
        "r"(__float_as_uint(v[16])), "r"(__float_as_uint(v[17])), "r"(__float_as_uint(v[18])), "r"(__float_as_uint(v[19])),
        "r"(__float_as_uint(v[20])), "r"(__float_as_uint(v[21])), "r"(__float_as_uint(v[22])), "r"(__float_as_uint(v[23])),
        "r"(__float_as_uint(v[24])), "r"(__float_as_uint(v[25])), "r"(__float_as_uint(v[26])), "r"(__float_as_uint(v[27])),
        "r"(__float_as_uint(v[28])), "r"(__float_as_uint(v[29])), "r"(__float_as_uint(v[30])), "r"(__float_as_uint(v[31]))
      : "memory");
}
__device__ __forceinline__ void tmem_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

}  // namespace tc
}  // namespace raae
