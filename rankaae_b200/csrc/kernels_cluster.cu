// Second compilation of the device code: thread-block cluster per trial (raae_config::ctas_per_trial 2 / 4 / 8).
// RAAE_CLUSTER == 1 gives Ctx run-time cluster rank / size and renames the namespace to raae_cn (aae_common.cuh), so both
// variants link into one library; the host side (rankaae_b200.cu) reaches this one through the three functions below.
// KParams / RunArgs are plain structs with the same layout in both translation units.
#define RAAE_CLUSTER 1
#include <cuda_runtime.h>

#include <cstring>

#include "aae_kernels.cuh"

namespace {
void fill_config(cudaLaunchConfig_t& cfg, cudaLaunchAttribute* attr, int n_clusters, int ctas, cudaStream_t stream) {
  std::memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = dim3((unsigned)(n_clusters * ctas), 1, 1);
  cfg.blockDim = dim3(raae::kThreads, 1, 1);
  cfg.dynamicSmemBytes = raae::kSmemBytes;
  cfg.stream = stream;
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = (unsigned)ctas;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
}
}  // namespace

// shared-memory opt-in + co-resident clusters of `ctas` CTAs of the train kernel on the current device
cudaError_t raae_cluster_setup(int ctas, int* max_clusters) {
  cudaError_t e = cudaFuncSetAttribute(raae::raae_train_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)raae::kSmemBytes);
  if (e != cudaSuccess) return e;
  e = cudaFuncSetAttribute(raae::raae_val_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)raae::kSmemBytes);
  if (e != cudaSuccess) return e;
  cudaLaunchConfig_t cfg;
  cudaLaunchAttribute attr[1];
  fill_config(cfg, attr, 1, ctas, nullptr);
  return cudaOccupancyMaxActiveClusters(max_clusters, raae::raae_train_kernel, &cfg);
}

// which: 0 = raae_train_kernel, 1 = raae_val_kernel; one cluster of `ctas` CTAs per trial
cudaError_t raae_cluster_launch(int which, const void* kparams, const void* run_args, int n_trials, int ctas, void* stream) {
  cudaLaunchConfig_t cfg;
  cudaLaunchAttribute attr[1];
  fill_config(cfg, attr, n_trials, ctas, (cudaStream_t)stream);
  const raae::KParams& kp = *static_cast<const raae::KParams*>(kparams);
  const raae::RunArgs& a = *static_cast<const raae::RunArgs*>(run_args);
  return cudaLaunchKernelEx(&cfg, which == 0 ? raae::raae_train_kernel : raae::raae_val_kernel, kp, a);
}
