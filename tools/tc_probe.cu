// Standalone probe of the tcgen05 descriptor conventions used in rankaae_b200/csrc/aae_tc.cuh.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O2 -Iinclude -Irankaae_b200/csrc tools/tc_probe.cu -o /tmp/tc_probe
// Computes C = A^T B for A, B = [128 rows][64] staged K-major SW128 and read as MN-major operands, for several
// (LBO, SBO, M) variants, and prints the max error against a host reference.
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <vector>
#include "aae_tc.cuh"
using namespace raae;

__global__ void probe(const float* A, const float* B, float* C, int variant) {
  extern __shared__ __align__(1024) unsigned char smem[];
  float* As = reinterpret_cast<float*>(smem);            // 32 KB
  float* Bs = As + 8192;                                  // 32 KB
  __shared__ uint32_t tmem_base;
  __shared__ unsigned long long mbar;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  if (warp == 0) tc::tmem_alloc(&tmem_base, 128);
  if (tid == 0) tc::mbar_init(reinterpret_cast<uint64_t*>(&mbar), 1);
  tc::fence_before_sync(); __syncthreads(); tc::fence_after_sync();
  for (int i = tid; i < 128 * 16; i += blockDim.x) {
    int r = i >> 4, c4 = (i & 15) * 4;
    uint32_t off = (variant == 4 || variant == 6) ? tc::sw128_chunk_off(r, c4, tc::kABlockBytes) : tc::sw128_32b_chunk_off(r, c4, tc::kABlockBytes);
    uint32_t offb = (variant == 4 || variant == 5 || variant == 6) ? tc::sw128_chunk_off(r, c4, tc::kABlockBytes) : off;
    *reinterpret_cast<float4*>(reinterpret_cast<char*>(As) + off) = *reinterpret_cast<const float4*>(A + r * 64 + c4);
    *reinterpret_cast<float4*>(reinterpret_cast<char*>(Bs) + offb) = *reinterpret_cast<const float4*>(B + r * 64 + c4);
  }
  tc::fence_async_smem(); __syncthreads();
  const uint32_t d = tmem_base;
  if (tid == 0) {
    tc::fence_after_sync();
    uint32_t a0 = tc::smem_u32(As), b0 = tc::smem_u32(Bs);
    if (variant == 4 || variant == 5 || variant == 6) {       // K-major product C[128][64] = A[128][64] B[0:64][64]^T; 5: A staged SW128_32B
      for (int s = 0; s < 8; ++s) {
        uint64_t da = tc::make_desc_k_sw128(a0 + (s >> 2) * tc::kABlockBytes + (s & 3) * 32);
        if (variant == 5) da = (da & ~((uint64_t)7u << 61)) | ((uint64_t)1u << 61);   // layout type SWIZZLE_128B_BASE32B
        uint64_t db = tc::make_desc_k_sw128(b0 + (s >> 2) * tc::kABlockBytes + (s & 3) * 32);
        tc::mma_tf32(d, da, db, tc::kIdescTf32_128x64, s > 0);
      }
      tc::mma_commit(reinterpret_cast<uint64_t*>(&mbar));
    } else {
    const int M = (variant & 1) ? 128 : 64;
    uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | (1u << 15) | (1u << 16) | ((64u >> 3) << 17) | (((uint32_t)M >> 4) << 24);
    for (int s = 0; s < 16; ++s) {
      uint64_t da, db;
      if (variant & 2) {    // LBO / SBO swapped
        da = tc::make_desc_mn_sw128_32b(a0 + s * 1024, 512, 16384); db = tc::make_desc_mn_sw128_32b(b0 + s * 1024, 512, 16384);
      } else {
        da = tc::make_desc_mn_sw128_32b(a0 + s * 1024, 16384, 512); db = tc::make_desc_mn_sw128_32b(b0 + s * 1024, 16384, 512);
      }
      tc::mma_tf32(d, da, db, idesc, s > 0);
    }
    tc::mma_commit(reinterpret_cast<uint64_t*>(&mbar));
    }
  }
  tc::mbar_wait(reinterpret_cast<uint64_t*>(&mbar), 0);
  tc::fence_after_sync();
  if (warp < 4) {
    float v[32];
    for (int h = 0; h < 2; ++h) {
      tc::tmem_ld32(d + ((uint32_t)(32 * warp) << 16) + 32 * h, v);
      for (int j = 0; j < 32; ++j) C[(32 * warp + lane) * 64 + 32 * h + j] = v[j];   // raw dump: [lane 0..127][col 0..63]
    }
  }
  tc::fence_before_sync(); __syncthreads();
  if (warp == 0) tc::tmem_dealloc(tmem_base, 128);
}

int main() {
  std::vector<float> A(128 * 64), B(128 * 64), C(128 * 64);
  srand(1);
  auto tf = [](float x) { uint32_t u; memcpy(&u, &x, 4); u &= 0xffffe000u; memcpy(&x, &u, 4); return x; };
  for (auto& x : A) x = tf((rand() % 2001 - 1000) / 1000.f);
  for (auto& x : B) x = tf((rand() % 2001 - 1000) / 1000.f);
  std::vector<double> ref(64 * 64, 0.0);
  for (int m = 0; m < 64; ++m) for (int n = 0; n < 64; ++n) { double s = 0; for (int r = 0; r < 128; ++r) s += (double)A[r * 64 + m] * B[r * 64 + n]; ref[m * 64 + n] = s; }
  float *dA, *dB, *dC;
  cudaMalloc(&dA, A.size() * 4); cudaMalloc(&dB, B.size() * 4); cudaMalloc(&dC, C.size() * 4);
  cudaMemcpy(dA, A.data(), A.size() * 4, cudaMemcpyHostToDevice);
  cudaMemcpy(dB, B.data(), B.size() * 4, cudaMemcpyHostToDevice);
  cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 66 * 1024);
  std::vector<float> A2(A), B2(B);
  for (auto& x : A2) { uint32_t u; memcpy(&u, &x, 4); u |= (uint32_t)(rand() & 0x1fff); memcpy(&x, &u, 4); }
  for (auto& x : B2) { uint32_t u; memcpy(&u, &x, 4); u |= (uint32_t)(rand() & 0x1fff); memcpy(&x, &u, 4); }
  for (int variant = 0; variant < 7; ++variant) {
    if (variant == 5) continue;     // K-major + SW128_32B faults (misaligned address): recorded in aae_step.cuh
    if (variant == 6) {             // operands with non-zero low mantissa bits: does the tensor core truncate them?
      cudaMemcpy(dA, A2.data(), A.size() * 4, cudaMemcpyHostToDevice);
      cudaMemcpy(dB, B2.data(), B.size() * 4, cudaMemcpyHostToDevice);
    }
    cudaMemset(dC, 0, C.size() * 4);
    probe<<<1, 256, 66 * 1024>>>(dA, dB, dC, variant);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("variant %d: CUDA error %s\n", variant, cudaGetErrorString(e)); return 1; }
    cudaMemcpy(C.data(), dC, C.size() * 4, cudaMemcpyDeviceToHost);
    int nz = 0; for (float x : C) nz += x != 0.f;
    printf("variant %d: nonzeros %d  C[0..3] = %g %g %g %g  ref[0..3] = %g %g %g %g\n", variant, nz, C[0], C[1], C[2], C[3], ref[0], ref[1], ref[2], ref[3]);
    if (variant == 4 || variant == 5 || variant == 6) {
      double err = 0;
      for (int m = 0; m < 128; ++m) for (int n = 0; n < 64; ++n) { double s = 0; for (int k = 0; k < 64; ++k) s += (double)A[m * 64 + k] * B[n * 64 + k]; err = fmax(err, fabs(C[m * 64 + n] - s)); }
      printf("variant %d (K-major%s): max err vs the product of the TRUNCATED operands %.3e\n", variant, variant == 6 ? ", low mantissa bits set" : " reference case", err);
      continue;
    }
    const int M = (variant & 1) ? 128 : 64;
    // try row->lane maps: identity and the M=64 "16 rows per 32-lane quarter" map
    for (int map = 0; map < 2; ++map) {
      double err = 0, mag = 0;
      for (int m = 0; m < 64; ++m) for (int n = 0; n < 64; ++n) {
        int lane_ = map == 0 ? m : 32 * (m / 16) + (m % 16);
        err = fmax(err, fabs(C[lane_ * 64 + n] - ref[m * 64 + n])); mag = fmax(mag, fabs(ref[m * 64 + n]));
      }
      printf("variant %d (M=%d, %s) rowmap %s: max err %.3e (ref max %.2f)\n", variant, M, (variant & 2) ? "LBO=512,SBO=16384" : "LBO=16384,SBO=512",
             map == 0 ? "identity" : "16-per-quarter", err, mag);
    }
  }
  return 0;
}
