// Throughput probe: legacy mma.sync.m16n8k8 TF32 (HMMA path) vs FFMA on sm_100a.  One CTA of 256 threads per SM; every warp
// keeps NACC independent accumulators in flight.  Prints warp-instructions per cycle per SM and the FMA-equivalent rate.
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o gpurun_out/mma_probe tools/mma_probe.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

template <int NACC>
__global__ void __launch_bounds__(256, 1) k_mma(float* out, int iters, long long* cyc) {
  float c[NACC][4];
  uint32_t a[4], b[2];
  for (int i = 0; i < 4; ++i) a[i] = __float_as_uint(1.0f + threadIdx.x * 1e-3f + i);
  for (int i = 0; i < 2; ++i) b[i] = __float_as_uint(0.5f + threadIdx.x * 1e-3f + i);
#pragma unroll
  for (int n = 0; n < NACC; ++n) c[n][0] = c[n][1] = c[n][2] = c[n][3] = 0.f;
  __syncthreads();
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int n = 0; n < NACC; ++n)
      asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                   : "+f"(c[n][0]), "+f"(c[n][1]), "+f"(c[n][2]), "+f"(c[n][3])
                   : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
  }
  long long t1 = clock64();
  float s = 0.f;
#pragma unroll
  for (int n = 0; n < NACC; ++n) s += c[n][0] + c[n][1] + c[n][2] + c[n][3];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

__global__ void __launch_bounds__(256, 1) k_ffma(float* out, int iters, long long* cyc) {
  float c[32];
  float a = 1.0f + threadIdx.x * 1e-3f, b = 0.5f + threadIdx.x * 1e-4f;
#pragma unroll
  for (int n = 0; n < 32; ++n) c[n] = (float)n;
  __syncthreads();
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int n = 0; n < 32; ++n) c[n] = fmaf(a, c[n], b);
  }
  long long t1 = clock64();
  float s = 0.f;
#pragma unroll
  for (int n = 0; n < 32; ++n) s += c[n];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

int main() {
  float* out; long long* cyc;
  cudaMalloc(&out, 148 * 256 * 4); cudaMalloc(&cyc, 148 * 8);
  long long h[148];
  const int iters = 20000;
  auto report = [&](const char* name, double instr_per_warp_iter, double fma_per_instr) {
    cudaDeviceSynchronize();
    cudaMemcpy(h, cyc, 148 * 8, cudaMemcpyDeviceToHost);
    double c = 0; for (int i = 0; i < 148; ++i) c += h[i]; c /= 148;
    double wi = 8.0 * iters * instr_per_warp_iter;      // warp-instructions per SM
    printf("%-28s %10.0f cycles  %.3f warp-instr/clk/SM  = %.0f FMA/clk/SM\n", name, c, wi / c, wi * fma_per_instr / c);
  };
  for (int rep = 0; rep < 2; ++rep) {
    k_mma<4><<<148, 256>>>(out, iters, cyc);  report("mma.sync tf32 m16n8k8 x4", 4, 1024);
    k_mma<8><<<148, 256>>>(out, iters, cyc);  report("mma.sync tf32 m16n8k8 x8", 8, 1024);
    k_mma<16><<<148, 256>>>(out, iters, cyc); report("mma.sync tf32 m16n8k8 x16", 16, 1024);
    k_ffma<<<148, 256>>>(out, iters, cyc);    report("ffma x32", 32, 32);
  }
  printf("%s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
