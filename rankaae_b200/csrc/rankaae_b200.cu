// Host side of the C ABI declared in include/rankaae_b200.h.
#include <cuda_runtime.h>

#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <new>
#include <string>

#include "aae_kernels.cuh"

// kernels_cluster.cu
cudaError_t raae_cluster_setup(int ctas, int* max_clusters);
cudaError_t raae_cluster_launch(int which, const void* kparams, const void* run_args, int n_trials, int ctas, void* stream);

namespace {

thread_local std::string g_err;

int fail(int code, const std::string& msg) {
  g_err = msg;
  return code;
}

#define RAAE_CUDA(expr)                                                                           \
  do {                                                                                            \
    cudaError_t e_ = (expr);                                                                      \
    if (e_ != cudaSuccess) return fail(-2, std::string(#expr) + ": " + cudaGetErrorString(e_));   \
  } while (0)

int round4(int x) { return (x + 3) & ~3; }

int validate_config(const raae_config& c) {
  if (c.dim_in <= 0 || c.dim_in > raae::kMaxDim || (c.dim_in & 3)) return fail(-1, "dim_in must be a multiple of 4 in [4, 256]");
  if (c.dim_out <= 0 || c.dim_out > raae::kMaxDim || (c.dim_out & 3)) return fail(-1, "dim_out must be a multiple of 4 in [4, 256]");
  if (c.dim_in != c.dim_out) return fail(-1, "dim_in must equal dim_out (the mutual-information phase feeds the decoder output to the encoder)");
  if (c.nstyle <= 0 || c.nstyle > RAAE_ZPAD) return fail(-1, "nstyle must be in [1, 8]");
  if (c.n_aux < 1 || c.n_aux > c.nstyle) return fail(-1, "n_aux must be in [1, nstyle]");
  if (c.n_layers < 2 || c.n_layers > RAAE_MAX_LAYERS) return fail(-1, "n_layers must be in [2, 8]");
  if (c.dis_layers != 3) return fail(-1, "only FC_discriminator_layers == 3 is implemented");
  if (c.batch_size <= 1) return fail(-1, "batch_size must be > 1");
  if (c.n_trials <= 0) return fail(-1, "n_trials must be positive");
  if (c.max_rows < c.batch_size) return fail(-1, "max_rows must be >= batch_size");
  if (c.ctas_per_trial != 1 && c.ctas_per_trial != 2 && c.ctas_per_trial != 4 && c.ctas_per_trial != RAAE_MAX_CTAS)
    return fail(-1, "ctas_per_trial must be 1, 2, 4 or 8 (thread-block cluster size per trial)");
  if (c.tensor_cores & ~0x37)
    return fail(-1, "tensor_cores: bits 0 (hidden forward), 1 (hidden backward), 2 (input block from operand images), 4 (decoder output forward), 5 (decoder output backward) are implemented");
  return 0;
}

// parameters()-order vector of an MLP: per Linear: weight, bias, then the PReLU slopes of the block
void layout_net(raae_net_layout& n, const int* dims, int n_linear, bool last_has_prelu_none, int& cursor_params) {
  (void)last_has_prelu_none;
  n.n_linear = n_linear;
  int off = 0;
  for (int l = 0; l < RAAE_MAX_LAYERS; ++l) {
    n.in_dim[l] = n.out_dim[l] = 0;
    n.w_off[l] = n.b_off[l] = n.a_off[l] = n.rm_off[l] = n.rv_off[l] = -1;
  }
  for (int l = 0; l < n_linear; ++l) {
    n.in_dim[l] = dims[l];
    n.out_dim[l] = dims[l + 1];
    n.w_off[l] = off; off += dims[l] * dims[l + 1];
    n.b_off[l] = off; off += dims[l + 1];
    if (l < n_linear - 1) { n.a_off[l] = off; off += dims[l + 1]; }
  }
  n.n_params = off;
  n.param_off = cursor_params;
  cursor_params += round4(off);
}

void build_layout(const raae_config& c, raae_layout& L, raae::ScratchLayout& S) {
  std::memset(&L, 0, sizeof(L));
  std::memset(&S, 0, sizeof(S));
  const int H = RAAE_HIDDEN;
  int cur = 0;
  int dims[RAAE_MAX_LAYERS + 1];
  // encoder: dim_in -> H x (n_layers-1) -> nstyle            (model.py:330-378)
  dims[0] = c.dim_in;
  for (int l = 1; l < c.n_layers; ++l) dims[l] = H;
  dims[c.n_layers] = c.nstyle;
  layout_net(L.net[raae::kE], dims, c.n_layers, true, cur);
  // decoder: nstyle -> H x (n_layers-1) -> dim_out            (model.py:518-570)
  dims[0] = c.nstyle;
  for (int l = 1; l < c.n_layers; ++l) dims[l] = H;
  dims[c.n_layers] = c.dim_out;
  layout_net(L.net[raae::kD], dims, c.n_layers, true, cur);
  // discriminator: nstyle -> H x (dis_layers-1) -> 1          (model.py:631-663)
  dims[0] = c.nstyle;
  for (int l = 1; l < c.dis_layers; ++l) dims[l] = H;
  dims[c.dis_layers] = 1;
  layout_net(L.net[raae::kS], dims, c.dis_layers, true, cur);
  // BatchNorm buffers: encoder every layer, decoder hidden layers only
  for (int net = 0; net < 2; ++net) {
    raae_net_layout& n = L.net[net];
    int nbn = net == raae::kE ? n.n_linear : n.n_linear - 1;
    for (int l = 0; l < nbn; ++l) {
      n.rm_off[l] = cur; cur += round4(n.out_dim[l]);
      n.rv_off[l] = cur; cur += round4(n.out_dim[l]);
    }
    n.nbt_off = cur; cur += 4;
  }
  L.net[raae::kS].nbt_off = -1;
  // optimizers (trainer.py:333-397): adversarial (S+E), correlation (E), reconstruction (E+D),
  // mutual_info (E+D), smoothness (D); vector order = param-group order
  const int members[RAAE_NUM_PHASES][3] = {{raae::kS, raae::kE, -1}, {raae::kE, -1, -1}, {raae::kE, raae::kD, -1},
                                           {raae::kE, raae::kD, -1}, {raae::kD, -1, -1}};
  for (int o = 0; o < RAAE_NUM_PHASES; ++o) {
    raae_opt_layout& ol = L.opt[o];
    for (int k = 0; k < RAAE_NUM_NETS; ++k) ol.net_off[k] = -1;
    int n = 0;
    for (int k = 0; k < 3 && members[o][k] >= 0; ++k) { ol.net_off[members[o][k]] = n; n += L.net[members[o][k]].n_params; }
    ol.n = n;
    ol.m_off = cur; cur += round4(n);
    ol.v_off = cur; cur += round4(n);
    ol.scalar_off = cur; cur += 4;
  }
  L.misc_off = cur; cur += 16;
  L.state_floats = round4(cur);
  // scratch
  const int rows = c.max_rows;
  int s = 0;
  S.xld = round4(c.dim_in);
  S.vld = round4(c.dim_out);
  S.xn = s; s += rows * S.xld;
  S.aux = s; s += rows * RAAE_ZPAD;
  for (int l = 0; l < c.n_layers - 1; ++l) { S.uE[l] = s; s += rows * H; }
  S.zE = s; s += rows * RAAE_ZPAD;
  S.dz = s; s += rows * RAAE_ZPAD;
  for (int l = 0; l < c.n_layers - 1; ++l) { S.uD[l] = s; s += rows * H; }
  S.v = s; s += rows * S.vld;
  S.g[0] = s; s += rows * H;
  S.g[1] = s; s += rows * H;
  S.zs = s; s += rows * RAAE_ZPAD;
  S.rank = s; s += rows * RAAE_ZPAD;
  S.nch64 = (c.dim_in + 63) / 64;
  S.nch128 = (c.dim_in + 127) / 128;
  const int tiles = (rows + 127) / 128;
  s = (s + 255) & ~255;                                  // 1 KB alignment of the operand images
  S.xk = s; s += tiles * S.nch64 * 8192;
  S.xm = s; s += tiles * S.nch128 * 16384;
  S.yk = s; s += tiles * S.nch64 * 8192;
  S.ym = s; s += tiles * S.nch128 * 16384;
  S.wk = s; s += S.nch64 * 8192;
  S.wl = s; s += ((c.dim_out + 63) / 64) * 8192;
  S.xref = s; s += 256 * RAAE_MAX_CTAS;                   // one reference row per CTA of the trial's cluster
  S.yref = s; s += 256 * RAAE_MAX_CTAS;
  S.total = (s + 255) & ~255;
  L.scratch_floats = S.total;
}

}  // namespace

struct raae_handle {
  raae::KParams kp;
  int device;
  int64_t launches;
  bool bound_state, bound_data;
  int shapiro_n;
  int max_clusters;                             // co-resident clusters of the train kernel (ctas_per_trial > 1), 0 otherwise
  long long* prof;
  // peer-memory exchange of the data-parallel mode (raae_peer_*)
  struct Peer {
    int world, rank;
    bool connected;
    unsigned seq;
    int last_phase;                             // phase of the previous exchange (-1: none)
    bool rewritten[RAAE_NUM_PHASES];            // raae_train_phase wrote the phase's vector since its last exchange
    int max_blocks;                             // co-resident blocks of the exchange kernel on this device (0 = not queried yet)
    size_t bytes, grad_off[RAAE_NUM_PHASES], sum_off[RAAE_NUM_PHASES];   // sum_off == grad_off when n_trials == 1
    unsigned char* local;                       // cudaMalloc: [256 B flags + done counter][gradient vector per phase]
    unsigned char* mapped[RAAE_MAX_PEERS];      // every rank's block in this process' address space
  } peer;
};

namespace {
// One CTA per trial: the kernels of this translation unit.  ctas_per_trial > 1: the cluster build of the same device code
// (kernels_cluster.cu, namespace raae_cn), one thread-block cluster per trial.
cudaError_t launch_trials(int which, const raae_handle* h, int n_trials, const raae::RunArgs& a, void* stream) {
  const int ctas = h->kp.cfg.ctas_per_trial;
  if (ctas > 1) return raae_cluster_launch(which, &h->kp, &a, n_trials, ctas, stream);
  if (which == 0) raae::raae_train_kernel<<<n_trials, raae::kThreads, raae::kSmemBytes, (cudaStream_t)stream>>>(h->kp, a);
  else raae::raae_val_kernel<<<n_trials, raae::kThreads, raae::kSmemBytes, (cudaStream_t)stream>>>(h->kp, a);
  return cudaGetLastError();
}
constexpr int kTrainKernel = 0, kValKernel = 1;

constexpr size_t kPeerHeaderBytes = 256;        // words [0, 8): flags, word 16: finished-block counter, word 17: pre-reduction arrivals, word 18: error
int peer_release(raae_handle* h) {
  if (!h->peer.local) return 0;
  cudaSetDevice(h->device);
  cudaDeviceSynchronize();
  for (int r = 0; r < h->peer.world; ++r)
    if (r != h->peer.rank && h->peer.mapped[r]) cudaIpcCloseMemHandle(h->peer.mapped[r]);
  cudaFree(h->peer.local);
  std::memset(&h->peer, 0, sizeof(h->peer));
  return 0;
}
}  // namespace

extern "C" {

const char* raae_last_error(void) { return g_err.c_str(); }
int raae_version(void) { return 100; }

int raae_query_layout(const raae_config* cfg, raae_layout* out) {
  if (!cfg || !out) return fail(-1, "null argument");
  if (int rc = validate_config(*cfg)) return rc;
  raae::ScratchLayout sl;
  build_layout(*cfg, *out, sl);
  return 0;
}

int raae_create(const raae_config* cfg, int device, raae_handle** out) {
  if (!cfg || !out) return fail(-1, "null argument");
  if (int rc = validate_config(*cfg)) return rc;
  int ndev = 0;
  RAAE_CUDA(cudaGetDeviceCount(&ndev));
  if (device < 0 || device >= ndev) return fail(-1, "no such CUDA device");
  RAAE_CUDA(cudaSetDevice(device));
  cudaDeviceProp prop;
  RAAE_CUDA(cudaGetDeviceProperties(&prop, device));
  if (prop.major < 10) return fail(-3, "rankaae_b200 requires an sm_100a (B200) device; there is no fallback path");
  raae_handle* h = new (std::nothrow) raae_handle();
  if (!h) return fail(-4, "out of host memory");
  std::memset(&h->kp, 0, sizeof(h->kp));
  h->kp.cfg = *cfg;
  build_layout(*cfg, h->kp.lay, h->kp.sl);
  h->device = device;
  h->launches = 0;
  h->bound_state = h->bound_data = false;
  h->shapiro_n = 0;
  h->max_clusters = 0;
  h->prof = nullptr;
  std::memset(&h->peer, 0, sizeof(h->peer));
  RAAE_CUDA(cudaFuncSetAttribute(raae::raae_train_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)raae::kSmemBytes));
  RAAE_CUDA(cudaFuncSetAttribute(raae::raae_val_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)raae::kSmemBytes));
  if (cfg->ctas_per_trial > 1) {
    // a cluster needs ctas_per_trial SMs of one GPC with the kernel's whole shared-memory footprint free at the same time
    int nclusters = 0;
    cudaError_t e = raae_cluster_setup(cfg->ctas_per_trial, &nclusters);
    if (e != cudaSuccess || nclusters < 1) {
      delete h;
      return fail(-2, std::string("ctas_per_trial = ") + std::to_string(cfg->ctas_per_trial) + " cannot be scheduled on this device" +
                          (e != cudaSuccess ? std::string(": ") + cudaGetErrorString(e) : std::string()));
    }
    h->max_clusters = nclusters;
  }
  *out = h;
  return 0;
}

int raae_max_clusters(int ctas_per_trial, int device, int* out) {
  if (!out) return fail(-1, "null argument");
  if (ctas_per_trial != 2 && ctas_per_trial != 4 && ctas_per_trial != RAAE_MAX_CTAS) return fail(-1, "ctas_per_trial must be 2, 4 or 8");
  RAAE_CUDA(cudaSetDevice(device));
  int n = 0;
  RAAE_CUDA(raae_cluster_setup(ctas_per_trial, &n));
  *out = n;
  return 0;
}

int raae_destroy(raae_handle* h) {
  if (h) peer_release(h);
  delete h;
  return 0;
}

int raae_bind_state(raae_handle* h, float* state, float* scratch, const double* hp) {
  if (!h || !state || !scratch || !hp) return fail(-1, "null argument");
  if (((uintptr_t)state | (uintptr_t)scratch) & 15) return fail(-1, "state/scratch must be 16-byte aligned");
  h->kp.state = state;
  h->kp.scratch = scratch;
  h->kp.hp = hp;
  h->bound_state = true;
  return 0;
}

int raae_bind_dataset(raae_handle* h, const float* spec_train, const float* aux_train, int n_train,
                      const float* spec_val, const float* aux_val, int n_val) {
  if (!h) return fail(-1, "null handle");
  if (n_train < 0 || n_val < 0) return fail(-1, "negative row count");
  if (n_val > h->kp.cfg.max_rows) return fail(-1, "n_val exceeds max_rows");
  if (n_val > 16384) return fail(-1, "n_val > 16384 is not supported by the in-kernel sort (cap the validation split)");
  if (((uintptr_t)spec_train | (uintptr_t)spec_val) & 15) return fail(-1, "spectra must be 16-byte aligned");
  h->kp.spec_train = spec_train;
  h->kp.aux_train = aux_train;
  h->kp.n_train = n_train;
  h->kp.spec_val = spec_val;
  h->kp.aux_val = aux_val;
  h->kp.n_val = n_val;
  h->bound_data = true;
  return 0;
}

int raae_bind_shapiro_weights(raae_handle* h, const float* w, int n) {
  if (!h || !w) return fail(-1, "null argument");
  h->kp.shapiro_w = w;
  h->shapiro_n = n;
  return 0;
}

int raae_reset_optimizers(raae_handle* h, void* stream) {
  if (!h || !h->bound_state) return fail(-1, "state not bound");
  RAAE_CUDA(cudaSetDevice(h->device));
  int n = h->kp.cfg.n_trials;
  raae::raae_reset_opt_kernel<<<(n + 127) / 128, 128, 0, (cudaStream_t)stream>>>(h->kp, n);
  RAAE_CUDA(cudaGetLastError());
  h->launches++;
  return 0;
}

int raae_step_debug(raae_handle* h, int trial, const raae_debug_io* io, void* stream) {
  if (!h || !io) return fail(-1, "null argument");
  if (!h->bound_state) return fail(-1, "state not bound");
  if (trial < 0 || trial >= h->kp.cfg.n_trials) return fail(-1, "trial out of range");
  if (io->rows <= 1 || io->rows > h->kp.cfg.batch_size) return fail(-1, "rows must be in (1, batch_size]");
  if (!io->x_noisy || (h->kp.cfg.n_aux > 0 && !io->aux)) return fail(-1, "x_noisy / aux are required");
  RAAE_CUDA(cudaSetDevice(h->device));
  raae::RunArgs a;
  std::memset(&a, 0, sizeof(a));
  a.trial0 = trial;
  a.debug = 1;
  a.epoch = io->epoch;
  a.dbg = *io;
  RAAE_CUDA(launch_trials(kTrainKernel, h, 1, a, stream));
  RAAE_CUDA(cudaGetLastError());
  h->launches++;
  return 0;
}

int raae_validate(raae_handle* h, int trial, const raae_val_io* io, void* stream) {
  if (!h || !io) return fail(-1, "null argument");
  if (!h->bound_state || !h->bound_data) return fail(-1, "state / dataset not bound");
  if (h->shapiro_n != h->kp.n_val) return fail(-1, "Shapiro-Wilk weights are not bound for n_val");
  if (trial < 0 || trial >= h->kp.cfg.n_trials) return fail(-1, "trial out of range");
  if (h->kp.n_val < 3) return fail(-1, "validation split too small");
  RAAE_CUDA(cudaSetDevice(h->device));
  raae::RunArgs a;
  std::memset(&a, 0, sizeof(a));
  a.trial0 = trial;
  a.debug = 1;
  a.epoch = io->epoch;
  a.val = *io;
  RAAE_CUDA(launch_trials(kValKernel, h, 1, a, stream));
  RAAE_CUDA(cudaGetLastError());
  h->launches++;
  return 0;
}

int raae_train_epochs(raae_handle* h, int epoch_begin, int n_epochs, const int32_t* perm, float* out_losses,
                      float* out_metrics, void* stream) {
  if (!h || !perm) return fail(-1, "null argument");
  if (!h->bound_state || !h->bound_data) return fail(-1, "state / dataset not bound");
  if (h->kp.n_train <= 1) return fail(-1, "training split too small");
  if (h->kp.n_val >= 3 && h->shapiro_n != h->kp.n_val) return fail(-1, "Shapiro-Wilk weights are not bound for n_val");
  RAAE_CUDA(cudaSetDevice(h->device));
  const int nt = h->kp.cfg.n_trials, bs = h->kp.cfg.batch_size;
  const int n_steps = (h->kp.n_train + bs - 1) / bs;
  // a short last batch of one row would make BatchNorm statistics undefined (torch raises there too)
  if (h->kp.n_train - (n_steps - 1) * bs < 2) return fail(-1, "last batch has fewer than 2 rows");
  for (int e = 0; e < n_epochs; ++e) {
    raae::RunArgs a;
    std::memset(&a, 0, sizeof(a));
    a.trial0 = 0;
    a.debug = 0;
    a.epoch = epoch_begin + e;
    a.n_steps = n_steps;
    a.perm = perm + (size_t)e * nt * h->kp.n_train;
    a.out_losses = out_losses ? out_losses + (size_t)e * nt * 12 : nullptr;
    a.out_metrics = out_metrics ? out_metrics + (size_t)e * nt * 6 : nullptr;
    a.val.avg_mutual_info = -INFINITY;
    a.prof = h->prof;
    RAAE_CUDA(launch_trials(kTrainKernel, h, nt, a, stream));
    RAAE_CUDA(cudaGetLastError());
    h->launches++;
    if (h->kp.n_val >= 3) {
      RAAE_CUDA(launch_trials(kValKernel, h, nt, a, stream));
      RAAE_CUDA(cudaGetLastError());
      h->launches++;
    }
  }
  return 0;
}

int raae_train_phase(raae_handle* h, int epoch, int step, int phase_mask, const int32_t* perm, float* const* grads,
                     void* stream) {
  if (!h || !perm || !grads) return fail(-1, "null argument");
  if (!h->bound_state || !h->bound_data) return fail(-1, "state / dataset not bound");
  RAAE_CUDA(cudaSetDevice(h->device));
  const int nt = h->kp.cfg.n_trials, bs = h->kp.cfg.batch_size;
  const int n_steps = (h->kp.n_train + bs - 1) / bs;
  if (step < 0 || step >= n_steps) return fail(-1, "step out of range");
  if (h->kp.n_train - (n_steps - 1) * bs < 2) return fail(-1, "last batch has fewer than 2 rows");
  raae::RunArgs a;
  std::memset(&a, 0, sizeof(a));
  a.epoch = epoch;
  a.n_steps = n_steps;
  a.perm = perm;
  a.split = 1;
  a.step0 = step;
  a.phase_mask = phase_mask;
  for (int o = 0; o < RAAE_NUM_PHASES; ++o) {
    a.grads_out[o] = grads[o];
    if (h->peer.local && grads[o] && (phase_mask & (1 << o))) h->peer.rewritten[o] = true;
  }
  a.val.avg_mutual_info = -INFINITY;
  RAAE_CUDA(launch_trials(kTrainKernel, h, nt, a, stream));
  RAAE_CUDA(cudaGetLastError());
  h->launches++;
  return 0;
}

int raae_apply_adam(raae_handle* h, int phase, const float* grads, void* stream) {
  if (!h || !grads) return fail(-1, "null argument");
  if (phase < 0 || phase >= RAAE_NUM_PHASES) return fail(-1, "phase out of range");
  if (!h->bound_state) return fail(-1, "state not bound");
  RAAE_CUDA(cudaSetDevice(h->device));
  const int nt = h->kp.cfg.n_trials;
  dim3 grid((h->kp.lay.opt[phase].n + 255) / 256, nt);
  raae::raae_adam_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(h->kp, phase, grads);
  RAAE_CUDA(cudaGetLastError());
  raae::raae_adam_tick_kernel<<<(nt + 127) / 128, 128, 0, (cudaStream_t)stream>>>(h->kp, phase, nt);
  RAAE_CUDA(cudaGetLastError());
  h->launches += 2;
  return 0;
}

int raae_validate_epoch(raae_handle* h, int epoch, float* out_losses, float* out_metrics, void* stream) {
  if (!h) return fail(-1, "null handle");
  if (!h->bound_state || !h->bound_data) return fail(-1, "state / dataset not bound");
  if (h->kp.n_val < 3 || h->shapiro_n != h->kp.n_val) return fail(-1, "validation split / Shapiro-Wilk weights not bound");
  RAAE_CUDA(cudaSetDevice(h->device));
  raae::RunArgs a;
  std::memset(&a, 0, sizeof(a));
  a.epoch = epoch;
  a.out_losses = out_losses;
  a.out_metrics = out_metrics;
  a.val.avg_mutual_info = -INFINITY;
  RAAE_CUDA(launch_trials(kValKernel, h, h->kp.cfg.n_trials, a, stream));
  RAAE_CUDA(cudaGetLastError());
  h->launches++;
  return 0;
}

int raae_debug_plateau(raae_handle* h, int trial, const double* metrics, int n, float* out, void* stream) {
  if (!h || !metrics || !out) return fail(-1, "null argument");
  if (!h->bound_state) return fail(-1, "state not bound");
  if (trial < 0 || trial >= h->kp.cfg.n_trials || n < 0) return fail(-1, "trial / n out of range");
  RAAE_CUDA(cudaSetDevice(h->device));
  raae::raae_plateau_debug_kernel<<<1, 32, 0, (cudaStream_t)stream>>>(h->kp, trial, metrics, n, out);
  RAAE_CUDA(cudaGetLastError());
  h->launches++;
  return 0;
}

int raae_evaluate_trials(raae_handle* h, int epoch, float* z, float* row_mae, float* losses, float* metrics, void* stream) {
  if (!h) return fail(-1, "null handle");
  if (!h->bound_state || !h->bound_data) return fail(-1, "state / dataset not bound");
  if (h->kp.n_val < 3 || h->shapiro_n != h->kp.n_val) return fail(-1, "validation split / Shapiro-Wilk weights not bound");
  RAAE_CUDA(cudaSetDevice(h->device));
  raae::RunArgs a;
  std::memset(&a, 0, sizeof(a));
  a.epoch = epoch;
  a.debug = 1;                                   // no scheduler step, no losses.csv row
  a.val.epoch = epoch;
  a.val.avg_mutual_info = -INFINITY;
  a.val.z = z;
  a.val.row_mae = row_mae;
  a.val.losses = losses;
  a.val.metrics = metrics;
  a.val.per_trial = 1;
  RAAE_CUDA(launch_trials(kValKernel, h, h->kp.cfg.n_trials, a, stream));
  h->launches++;
  return 0;
}

int raae_peer_alloc(raae_handle* h, int world, int rank, unsigned char* ipc_handle_out) {
  if (!h || !ipc_handle_out) return fail(-1, "null argument");
  if (world < 1 || world > RAAE_MAX_PEERS || rank < 0 || rank >= world) return fail(-1, "world must be in [1, 8] and rank in [0, world)");
  if (h->peer.local) return fail(-1, "exchange block already allocated");
  static_assert(sizeof(cudaIpcMemHandle_t) == RAAE_IPC_HANDLE_BYTES, "IPC handle size");
  RAAE_CUDA(cudaSetDevice(h->device));
  size_t off = kPeerHeaderBytes;
  for (int o = 0; o < RAAE_NUM_PHASES; ++o) {
    h->peer.grad_off[o] = h->peer.sum_off[o] = off;
    off += (((size_t)h->kp.cfg.n_trials * h->kp.lay.opt[o].n * sizeof(float)) + 255) & ~(size_t)255;
    if (h->kp.cfg.n_trials > 1) {                // the replicas' pre-reduced sum, the vector the peers read
      h->peer.sum_off[o] = off;
      off += (((size_t)h->kp.lay.opt[o].n * sizeof(float)) + 255) & ~(size_t)255;
    }
  }
  void* ptr = nullptr;
  RAAE_CUDA(cudaMalloc(&ptr, off));
  RAAE_CUDA(cudaMemset(ptr, 0, off));
  RAAE_CUDA(cudaDeviceSynchronize());
  cudaIpcMemHandle_t ipc;
  cudaError_t e = cudaIpcGetMemHandle(&ipc, ptr);
  if (e != cudaSuccess) { cudaFree(ptr); return fail(-2, std::string("cudaIpcGetMemHandle: ") + cudaGetErrorString(e)); }
  std::memcpy(ipc_handle_out, &ipc, RAAE_IPC_HANDLE_BYTES);
  h->peer.world = world;
  h->peer.rank = rank;
  h->peer.bytes = off;
  h->peer.local = (unsigned char*)ptr;
  h->peer.mapped[rank] = h->peer.local;
  h->peer.seq = 0;
  h->peer.last_phase = -1;
  h->peer.connected = false;
  return 0;
}

int raae_peer_connect(raae_handle* h, const unsigned char* all_handles) {
  if (!h || !all_handles) return fail(-1, "null argument");
  if (!h->peer.local) return fail(-1, "raae_peer_alloc has not been called");
  if (h->peer.connected) return fail(-1, "already connected");
  RAAE_CUDA(cudaSetDevice(h->device));
  for (int r = 0; r < h->peer.world; ++r) {
    if (r == h->peer.rank) continue;
    cudaIpcMemHandle_t ipc;
    std::memcpy(&ipc, all_handles + (size_t)r * RAAE_IPC_HANDLE_BYTES, RAAE_IPC_HANDLE_BYTES);
    void* ptr = nullptr;
    cudaError_t e = cudaIpcOpenMemHandle(&ptr, ipc, cudaIpcMemLazyEnablePeerAccess);
    if (e != cudaSuccess)
      return fail(-2, "cudaIpcOpenMemHandle (rank " + std::to_string(r) + "): " + cudaGetErrorString(e) +
                          " - the peer exchange needs one process per GPU on a P2P-capable (NVLink) node");
    h->peer.mapped[r] = (unsigned char*)ptr;
  }
  h->peer.connected = true;
  return 0;
}

int raae_peer_grad_ptr(raae_handle* h, int phase, float** out) {
  if (!h || !out) return fail(-1, "null argument");
  if (phase < 0 || phase >= RAAE_NUM_PHASES) return fail(-1, "phase out of range");
  if (!h->peer.local) return fail(-1, "raae_peer_alloc has not been called");
  *out = (float*)(h->peer.local + h->peer.grad_off[phase]);
  return 0;
}

int raae_apply_adam_peer(raae_handle* h, int phase, void* stream) {
  if (!h) return fail(-1, "null handle");
  if (phase < 0 || phase >= RAAE_NUM_PHASES) return fail(-1, "phase out of range");
  if (!h->bound_state) return fail(-1, "state not bound");
  if (!h->peer.connected) return fail(-1, "peer exchange not connected (raae_peer_alloc / raae_peer_connect)");
  // Buffer reuse: a rank may still be reading this rank's vector of exchange k when this rank has already left it.  The
  // vector is safe because it is only rewritten by the NEXT raae_train_phase of the same phase, and every rank enters an
  // exchange only after its previous one (the reader) has finished - which needs at least one OTHER exchange in between.
  if (h->peer.world > 1 && h->peer.last_phase == phase && h->peer.rewritten[phase])
    return fail(-1, "raae_apply_adam_peer: the same phase twice in a row with a rewritten gradient vector - a peer may still "
                    "read the previous one; interleave at least two phases (the trainer cycles through five)");
  h->peer.last_phase = phase;
  h->peer.rewritten[phase] = false;
  RAAE_CUDA(cudaSetDevice(h->device));
  raae::PeerArgs pa;
  std::memset(&pa, 0, sizeof(pa));
  for (int r = 0; r < h->peer.world; ++r) {
    pa.sums[r] = (const float*)(h->peer.mapped[r] + h->peer.sum_off[phase]);
    pa.flags[r] = (unsigned*)h->peer.mapped[r];
  }
  pa.lgrads = (const float*)(h->peer.local + h->peer.grad_off[phase]);
  pa.lsum = (float*)(h->peer.local + h->peer.sum_off[phase]);
  pa.done = (unsigned*)h->peer.local + 16;
  pa.arrive = (unsigned*)h->peer.local + 17;
  pa.error = (unsigned*)h->peer.local + 18;
  {
    const char* ts = std::getenv("RAAE_PEER_TIMEOUT_S");
    double sec = ts ? std::atof(ts) : 120.0;
    if (!(sec > 0.0)) sec = 120.0;
    pa.timeout_ns = (unsigned long long)(sec * 1e9);
  }
  pa.world = h->peer.world;
  pa.rank = h->peer.rank;
  pa.replicas = h->kp.cfg.n_trials;
  pa.seq = ++h->peer.seq;
  // every block of the grid must be resident: block 0 waits for the other blocks' slices of the local pre-reduction and all
  // blocks wait for the peers' flags, so a block that cannot be scheduled would stall the whole exchange
  if (h->peer.max_blocks == 0) {
    int per_sm = 0, sms = 0;
    RAAE_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, raae::raae_adam_peer_kernel, 256, 0));
    RAAE_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, h->device));
    h->peer.max_blocks = (per_sm > 0 ? per_sm : 1) * (sms > 0 ? sms : 1);
  }
  int bx = (h->kp.lay.opt[phase].n + 255) / 256;
  if (bx > h->peer.max_blocks) bx = h->peer.max_blocks;
  // cooperative launch: the grid starts only when ALL its blocks can be resident at once (block 0 waits for the others'
  // slices and every block waits on the peers' flags), whatever else shares the GPU
  {
    void* kargs[3] = {(void*)&h->kp, (void*)&phase, (void*)&pa};
    RAAE_CUDA(cudaLaunchCooperativeKernel((const void*)raae::raae_adam_peer_kernel, dim3((unsigned)bx), dim3(256), kargs, 0,
                                          (cudaStream_t)stream));
  }
  RAAE_CUDA(cudaGetLastError());
  h->launches++;
  return 0;
}

int raae_peer_status(raae_handle* h, unsigned* failed_seq) {
  if (!h || !failed_seq) return fail(-1, "null argument");
  if (!h->peer.local) return fail(-1, "raae_peer_alloc has not been called");
  RAAE_CUDA(cudaSetDevice(h->device));
  RAAE_CUDA(cudaMemcpy(failed_seq, (unsigned*)h->peer.local + 18, sizeof(unsigned), cudaMemcpyDeviceToHost));
  return 0;
}

int raae_peer_free(raae_handle* h) {
  if (!h) return fail(-1, "null handle");
  return peer_release(h);
}

int64_t raae_launch_count(const raae_handle* h) { return h ? h->launches : 0; }

int raae_set_profile_buffer(raae_handle* h, long long* prof) {
  if (!h) return fail(-1, "null handle");
  h->prof = prof;
  return 0;
}

}  // extern "C"
