// Per-SM throughput of cp.async.bulk (global -> shared) streams: every CTA (one per SM) streams its own region of
// `mb_per_cta` MB through a ring of `stages` buffers of `chunk_kb` KB; prints bytes / cycle / SM.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O2 -Iinclude -Irankaae_b200/csrc tools/bulk_probe.cu -o tools/bulk_probe.bin
#include <cstdio>
#include <cstdlib>
#include "aae_tc.cuh"
using namespace raae;

__global__ void stream(const float* src, size_t floats_per_cta, int chunk_floats, int stages, int parts, long long* cycles, float* sink) {
  extern __shared__ __align__(1024) unsigned char smem[];
  float* buf = reinterpret_cast<float*>(smem);
  __shared__ unsigned long long bars[8];
  const float* base = src + (size_t)blockIdx.x * floats_per_cta;
  const int nchunks = (int)(floats_per_cta / chunk_floats);
  if (threadIdx.x == 0) for (int i = 0; i < stages; ++i) tc::mbar_init(reinterpret_cast<uint64_t*>(&bars[i]), 1);
  __syncthreads();
  float acc = 0.f;
  long long t0 = clock64();
  if (threadIdx.x == 0) {
    auto load = [&](int c) {
      const int st = c % stages;
      tc::mbar_expect_tx(reinterpret_cast<uint64_t*>(&bars[st]), (uint32_t)chunk_floats * 4u);
      const int pf = chunk_floats / parts;
      for (int p = 0; p < parts; ++p)
        tc::bulk_g2s(buf + (size_t)st * chunk_floats + p * pf, base + (size_t)c * chunk_floats + p * pf, (uint32_t)pf * 4u,
                     reinterpret_cast<uint64_t*>(&bars[st]));
    };
    for (int c = 0; c < stages && c < nchunks; ++c) load(c);
    for (int c = 0; c < nchunks; ++c) {
      const int st = c % stages;
      tc::mbar_wait(reinterpret_cast<uint64_t*>(&bars[st]), (uint32_t)((c / stages) & 1));
      acc += buf[(size_t)st * chunk_floats + (c & 1023)];
      if (c + stages < nchunks) load(c + stages);
    }
  }
  long long t1 = clock64();
  if (threadIdx.x == 0) { cycles[blockIdx.x] = t1 - t0; sink[blockIdx.x] = acc; }
}

int main(int argc, char** argv) {
  const int ctas = 148;
  const size_t mb = argc > 1 ? atoi(argv[1]) : 8;
  const size_t floats_per_cta = mb * 1024 * 1024 / 4;
  float* src; long long* cyc; float* sink;
  cudaMalloc(&src, ctas * floats_per_cta * 4); cudaMemset(src, 0, ctas * floats_per_cta * 4);
  cudaMalloc(&cyc, ctas * 8); cudaMalloc(&sink, ctas * 4);
  cudaFuncSetAttribute(stream, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  const int cfgs[][3] = {{32, 2, 1}, {32, 4, 1}, {32, 4, 4}, {64, 2, 1}, {64, 2, 8}, {16, 8, 1}, {8, 16, 1}};
  for (auto& cf : cfgs) {
    const int chunk_floats = cf[0] * 256, stages = cf[1], parts = cf[2];
    if ((size_t)chunk_floats * 4 * stages > 192 * 1024) continue;
    for (int n : {148, 16, 1}) {
      stream<<<n, 32, (size_t)chunk_floats * 4 * stages>>>(src, floats_per_cta, chunk_floats, stages, parts, cyc, sink);
      cudaError_t e = cudaDeviceSynchronize();
      if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return 1; }
      long long h[148]; cudaMemcpy(h, cyc, n * 8, cudaMemcpyDeviceToHost);
      double mean = 0; for (int i = 0; i < n; ++i) mean += (double)h[i]; mean /= n;
      printf("chunk %3d KB x %2d stages, %d copies per chunk, %3d CTAs: %.1f B/cycle/SM (%.2f TB/s aggregate at 1.9 GHz)\n", cf[0], stages, parts, n,
             floats_per_cta * 4.0 / mean, floats_per_cta * 4.0 / mean * n * 1.9e9 / 1e12);
    }
  }
  return 0;
}
