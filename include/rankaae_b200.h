/*
 * rankaae_b200 — C ABI of the fused sm_100a adversarial-autoencoder (AAE) train step.
 *
 * The reference (AI-multimodal/RankAAE) is pure Python and has no FFI of its own; its seams for
 * this path are the Python `Trainer` methods.  Each entry point below names the reference
 * interface it stands in for (file:line under the reference tree):
 *
 *   raae_create / raae_bind_*      Trainer.__init__ + load_optimizers + load_schedulers
 *                                  sc/clustering/trainer.py:38-62, 333-408
 *   raae_step_debug                ONE iteration of the batch loop body, teacher-forced
 *                                  sc/clustering/trainer.py:103-204
 *   raae_train_epochs              the epoch loop incl. validation, metrics and scheduler step
 *                                  sc/clustering/trainer.py:89-307
 *   raae_validate                  the eval block  sc/clustering/trainer.py:207-297
 *
 * Conventions: plain pointers and sizes only; every buffer is owned by the caller (device memory
 * unless stated otherwise) and only borrowed for the duration of a call, except the state,
 * scratch and dataset pointers which are borrowed until raae_destroy.  All functions return 0 on
 * success and a negative code on failure; raae_last_error() returns a thread-local message.
 * One handle per GPU, not thread-safe.  `stream` is a cudaStream_t passed as void*.
 */
#ifndef RANKAAE_B200_H_
#define RANKAAE_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define RAAE_MAX_LAYERS 8
#define RAAE_HIDDEN 64          /* reference hard-codes hidden_size=64: model.py:342,528,632 */
#define RAAE_NUM_PHASES 5       /* adversarial, correlation, reconstruction, mutual_info, smoothness */
#define RAAE_NUM_NETS 3         /* 0 = encoder, 1 = decoder, 2 = style discriminator */
#define RAAE_ZPAD 8             /* latent rows are padded to 8 floats */

/* per-trial hyper-parameters (float64 slots of the `hp` array, one row per trial — double so that
 * 1 - beta1, lr * weight_decay and the bias corrections are formed exactly as torch forms them);
 * fix_config.yaml keys consumed by trainer.py:89-204, 333-408 */
enum raae_hp_slot {
  RAAE_HP_LR0 = 0,             /* 5 slots: initial lr of each stepped optimizer (phase order)     */
  RAAE_HP_BETA1 = 5,           /* 5 slots                                                          */
  RAAE_HP_BETA2 = 10,          /* 5 slots                                                          */
  RAAE_HP_WD = 15,             /* 5 slots                                                          */
  RAAE_HP_DROPOUT = 20,        /* dropout_rate (encoder + decoder)                                 */
  RAAE_HP_DIS_DROPOUT = 21,    /* dis_dropout_rate                                                 */
  RAAE_HP_DIS_NOISE = 22,      /* dis_noise                                                        */
  RAAE_HP_SPEC_NOISE = 23,     /* spec_noise                                                       */
  RAAE_HP_ALPHA_FLAT_STEP = 24,
  RAAE_HP_ALPHA_LIMIT = 25,
  RAAE_HP_SCH_FACTOR = 26,
  RAAE_HP_SCH_PATIENCE = 27,
  RAAE_HP_EPOCH_STOP_SMOOTH = 28,
  RAAE_HP_MAX_EPOCH = 29,
  RAAE_HP_SEED = 30,           /* integer-valued, mixed with the trial index                       */
  RAAE_HP_COUNT = 32
};

/* structural configuration, fixed per handle (fix_config.yaml: model keys, trainer.py:442-463) */
typedef struct raae_config {
  int32_t dim_in;              /* multiple of 4, <= 256 */
  int32_t dim_out;             /* == dim_in */
  int32_t nstyle;              /* <= 8 */
  int32_t n_aux;               /* 1 .. nstyle */
  int32_t n_layers;            /* FCEncoder/FCDecoder n_layers, 2..8 */
  int32_t dis_layers;          /* FC_discriminator_layers; only 3 is implemented */
  int32_t batch_size;
  int32_t n_trials;            /* trials resident on this GPU */
  int32_t kendall_activation;  /* bool */
  int32_t use_flex_spec_target;/* bool */
  int32_t decoder_softplus;    /* 1 = Softplus(beta=2), 0 = ReLU */
  int32_t max_rows;            /* scratch rows per trial: >= max(batch_size, n_val) */
  int32_t ctas_per_trial;      /* thread-block cluster per trial: 1, 2, 4 or 8 CTAs share one trial (128-row tiles of a batch are
                                  dealt round-robin; BatchNorm statistics, weight gradients, loss and Kendall totals are reduced
                                  through distributed shared memory).  1 = one CTA per trial (ensembles of >= 148 trials) */
  int32_t tensor_cores;        /* contractions on tcgen05 (kind::tf32, 3 x TF32 round-to-nearest split, TMEM accumulators): bit 0
                                  hidden-block forward, bit 1 hidden-block backward, bit 2 input block of the encoder on the batch
                                  (forward + weight gradient from operand images in scratch, streamed with bulk copies; also the
                                  re-encoding pass of the MI phase), bit 4 decoder output forward, bit 5 decoder output backward (input
                                  gradient and weight gradient, the loss-gradient tile parked in tensor memory); bit 3 is unused;
                                  0 = everything as FP32 FMA */
  int32_t reserved[2];
} raae_config;

/* flat per-trial state block layout (all offsets in floats from the start of a trial's block) */
typedef struct raae_net_layout {
  int32_t n_linear;                       /* number of Linear layers */
  int32_t in_dim[RAAE_MAX_LAYERS], out_dim[RAAE_MAX_LAYERS];
  int32_t w_off[RAAE_MAX_LAYERS];         /* weight [out][in] row-major (nn.Linear layout)      */
  int32_t b_off[RAAE_MAX_LAYERS];         /* bias [out]                                          */
  int32_t a_off[RAAE_MAX_LAYERS];         /* PReLU slopes [out], -1 if the layer has none        */
  int32_t rm_off[RAAE_MAX_LAYERS];        /* BN running_mean [out], -1 if the layer has no BN    */
  int32_t rv_off[RAAE_MAX_LAYERS];        /* BN running_var  [out]                               */
  int32_t param_off;                      /* start of this net's parameter vector                */
  int32_t n_params;                       /* parameters()-order vector length                    */
  int32_t nbt_off;                        /* num_batches_tracked (stored as float), -1 if none   */
} raae_net_layout;

typedef struct raae_opt_layout {
  int32_t m_off, v_off;                   /* exp_avg / exp_avg_sq vectors                        */
  int32_t n;                              /* vector length                                       */
  int32_t net_off[RAAE_NUM_NETS];         /* start of each net's slice in the vector, -1 = absent*/
  int32_t scalar_off;                     /* 4 floats: lr, step count t, plateau best, num_bad   */
} raae_opt_layout;

typedef struct raae_layout {
  raae_net_layout net[RAAE_NUM_NETS];
  raae_opt_layout opt[RAAE_NUM_PHASES];
  int32_t misc_off;                       /* 16 floats: [0..4] last-batch train losses (phase order),
                                             [5] sum of train MI losses this epoch, [6] batches this epoch */
  int32_t state_floats;                   /* per-trial state block size                          */
  int32_t scratch_floats;                 /* per-trial scratch block size                        */
} raae_layout;

/* explicit random draws / outputs of one teacher-forced step (all device pointers, any may be NULL:
 * NULL draws are generated by the in-kernel counter-based RNG, NULL outputs are skipped).
 * Draw order and shapes follow SURVEY.md Appendix E.3. */
typedef struct raae_debug_io {
  const float* x_noisy;                   /* [B][dim_in]  spec_in after the noise add, trainer.py:112 */
  const float* aux;                       /* [B][n_aux]                                               */
  int32_t rows;                           /* B (<= batch_size)                                        */
  int32_t epoch;                          /* for alpha() and epoch_stop_smooth                        */
  int32_t phase_mask;                     /* bit p set = run phase p (P0 forwards always run)         */
  int32_t apply_updates;                  /* 0 = compute losses/gradients only                        */
  const uint8_t* mask_enc[6][RAAE_MAX_LAYERS];  /* keep-masks [B][64] of the 6 encoder forwards       */
  const uint8_t* mask_dec[4][RAAE_MAX_LAYERS];  /* 4 decoder forwards                                 */
  const uint8_t* mask_dis[2][RAAE_MAX_LAYERS];  /* discriminator on z_real [batch_size][64], on styles [B][64] */
  const float* z_real;                    /* [batch_size][nstyle]  functions.py:122                   */
  const float* dis_eps_real;              /* [batch_size][nstyle]  model.py:660                       */
  const float* dis_eps_fake;              /* [B][nstyle]                                              */
  const float* z_sample;                  /* [B][nstyle]           functions.py:187                   */
  float* losses;                          /* [5] phase order                                          */
  float* grads[RAAE_NUM_PHASES];          /* gradient vector in the optimizer's parameter order       */
  float* styles;                          /* [B][nstyle] encoder output of the P0 forward             */
} raae_debug_io;

/* validation inputs/outputs of one trial (device pointers unless noted) */
typedef struct raae_val_io {
  const float* z_sample;                  /* [n_val][nstyle] or NULL                                  */
  const float* z_real;                    /* [batch_size][nstyle] or NULL                             */
  int32_t epoch;
  float avg_mutual_info;                  /* used when > -1e30, else the value accumulated in-state   */
  float* losses;                          /* [5] phase order: val adversarial, Kendall, recon, MI, smooth */
  float* metrics;                         /* [6]: min Shapiro W, recon, avg MI, max |Spearman|, Kendall, combined */
  float* z;                               /* [n_val][nstyle] or NULL                                  */
  float* row_mae;                         /* [n_val] or NULL: mean |D(E(x)) - x| of every row (sc/report/analysis.py:425-428) */
  int32_t per_trial;                      /* 1: losses / metrics / z / row_mae hold one block per resident trial (raae_evaluate_trials) */
  int32_t reserved;
} raae_val_io;

typedef struct raae_handle raae_handle;

const char* raae_last_error(void);
int raae_version(void);

/* Layout of the state / scratch blocks for a configuration (host-only, no GPU needed). */
int raae_query_layout(const raae_config* cfg, raae_layout* out);

/* Handle life cycle.  `device` is the CUDA ordinal. */
int raae_create(const raae_config* cfg, int device, raae_handle** out);
int raae_destroy(raae_handle* h);
/* Clusters of `ctas_per_trial` CTAs of the train kernel that fit the device at once (a cluster lives inside one GPC and
 * every CTA needs a whole SM): more trials than this per launch run in waves. */
int raae_max_clusters(int ctas_per_trial, int device, int* out);

/* state: [n_trials][layout.state_floats]; scratch: [n_trials][layout.scratch_floats];
 * float32 device memory, 16-byte aligned; hp: [n_trials][RAAE_HP_COUNT] float64 device memory. */
int raae_bind_state(raae_handle* h, float* state, float* scratch, const double* hp);

/* Training split (row-major float32, device): spec [n_train][dim_in], aux [n_train][n_aux];
 * validation split likewise.  trainer.py:52-53 / dataloader.py:64-77. */
int raae_bind_dataset(raae_handle* h, const float* spec_train, const float* aux_train, int n_train,
                      const float* spec_val, const float* aux_val, int n_val);

/* Shapiro-Wilk weights for n_val (float32 [n_val], device), computed by the caller in float64
 * (SURVEY.md Appendix B).  Needed before raae_validate / raae_train_epochs. */
int raae_bind_shapiro_weights(raae_handle* h, const float* w, int n);

/* Initialise the optimizer scalars (lr, t=0, best=+inf, bad=0) of every trial from `hp`. */
int raae_reset_optimizers(raae_handle* h, void* stream);

/* One teacher-forced step of one trial (parity entry point). */
int raae_step_debug(raae_handle* h, int trial, const raae_debug_io* io, void* stream);

/* Validation block of one trial with explicit draws (parity entry point). */
int raae_validate(raae_handle* h, int trial, const raae_val_io* io, void* stream);

/* Production path.  Runs `n_epochs` epochs [epoch_begin, epoch_begin + n_epochs) of every resident
 * trial: per epoch all train batches (perm: int32 [n_epochs][n_trials][n_train] device, the shuffled
 * row order of each trial; batches are consecutive batch_size chunks, the last one short —
 * dataloader.py:70-71), then validation, metrics and the ReduceLROnPlateau step.
 * out_losses:  float32 [n_epochs][n_trials][12] device — the losses.csv columns after `Epoch`
 *              (trainer.py:84-87): Train_D,Val_D,Train_G,Val_G,Train_Aux,Val_Aux,Train_Recon,
 *              Val_Recon,Train_Smooth,Val_Smooth,Train_Mutual_Info,Val_Mutual_Info
 * out_metrics: float32 [n_epochs][n_trials][6] device — the 5 metrics of trainer.py:294-295 + combined.
 * Nothing is synchronised; the caller syncs `stream`. */
int raae_train_epochs(raae_handle* h, int epoch_begin, int n_epochs, const int32_t* perm,
                      float* out_losses, float* out_metrics, void* stream);

/* Split-phase entry points for the single-trial DATA-PARALLEL mode (BASELINE.json configs[3]): every rank
 * holds the same weights and its own shard of the training rows.  raae_train_phase runs the phases in `phase_mask` of
 * batch `step` of epoch `epoch` WITHOUT updating, exporting each phase's gradient vector (optimizer parameter order,
 * [n_trials][raae_layout.opt[o].n]) to grads[o] (device pointers, NULL for phases not in the mask); the caller
 * all-reduces (mean) the vector across ranks and hands it to raae_apply_adam.  Phase 0 (adversarial) must be launched
 * first for every batch: that launch also builds the noised batch and runs the P0 forwards (trainer.py:112-114).
 * BatchNorm statistics and Kendall pairs stay rank-local (DistributedDataParallel semantics).
 * raae_validate_epoch runs the validation block + scheduler step of every resident trial (what raae_train_epochs
 * does after the last batch of an epoch). */
int raae_train_phase(raae_handle* h, int epoch, int step, int phase_mask, const int32_t* perm, float* const* grads,
                     void* stream);
int raae_apply_adam(raae_handle* h, int phase, const float* grads, void* stream);
int raae_validate_epoch(raae_handle* h, int epoch, float* out_losses, float* out_metrics, void* stream);

/* Parity hook of the in-kernel ReduceLROnPlateau (trainer.py:303-304, 400-408): steps the schedulers of `trial` with the
 * scripted sequence metrics[0..n) (float64, device) and records (lr, best, num_bad_epochs) of the five optimizers after
 * every step in out [n][5][3] (float32, device).  The trial's scheduler state advances as in training. */
int raae_debug_plateau(raae_handle* h, int trial, const double* metrics, int n, float* out, void* stream);

/* Batched, side-effect-free evaluation of EVERY resident trial on the bound validation rows (bind the test split with
 * raae_bind_dataset to rank trials the way sc/report/analysis.py:394-450 `evaluate_model` does): eval-mode forward of the
 * encoder and decoder, the five validation losses and the metric vector - no scheduler step, no state change.
 * z [n_trials][n_val][nstyle], row_mae [n_trials][n_val], losses [n_trials][5], metrics [n_trials][6]; any may be NULL. */
int raae_evaluate_trials(raae_handle* h, int epoch, float* z, float* row_mae, float* losses, float* metrics, void* stream);

/* Peer-memory gradient exchange for the data-parallel mode ("one-shot ... fused into the phase epilogue"):
 * replaces `all_reduce(grads) ; grads /= world ; raae_apply_adam` by ONE launch that signals the peers, waits for their
 * gradient vectors, sums them in rank order straight out of the peers' HBM over NVLink (P2P loads, no NCCL call, no
 * staging copy), divides by `world` and applies AdamW — every rank forms bit-identical updates.
 *   raae_peer_alloc    allocates this rank's exchange block (flag words + one gradient vector per phase) with cudaMalloc
 *                      (the only device memory the library owns) and returns its 64-byte CUDA IPC handle;
 *   raae_peer_connect  maps the blocks of all ranks (handles: [world][64] bytes gathered by the caller in rank order,
 *                      e.g. with torch.distributed.all_gather_object); the caller runs a barrier afterwards;
 *   raae_peer_grad_ptr the local gradient vector of `phase` — pass it to raae_train_phase as grads[phase];
 *   raae_apply_adam_peer  the fused exchange + mean + AdamW launch.  Every rank must call it for the same phases in the
 *                      same order (the flag protocol counts calls).  A rank that waits > 10 s for a peer traps (the
 *                      stream reports a launch failure) instead of hanging.
 *   raae_peer_free     unmaps / frees (also done by raae_destroy); the caller runs a barrier before it.
 * With n_trials > 1 the resident trials are data-parallel REPLICAS of one trial (identical weights, one shard each): the
 * launch first sums the replicas' vectors locally, the peers read that one vector per rank, the mean runs over
 * world x n_trials shards and the one update is applied to every replica's state (rankaae_b200/dp.py, `replicas`).
 * world <= RAAE_MAX_PEERS ranks on one NVLink / NVSwitch domain, one process per GPU. */
#define RAAE_MAX_PEERS 8
#define RAAE_MAX_CTAS 8         /* largest thread-block cluster per trial (portable cluster size) */
#define RAAE_IPC_HANDLE_BYTES 64
int raae_peer_alloc(raae_handle* h, int world, int rank, unsigned char* ipc_handle_out);
int raae_peer_connect(raae_handle* h, const unsigned char* all_handles);
int raae_peer_grad_ptr(raae_handle* h, int phase, float** out);
int raae_apply_adam_peer(raae_handle* h, int phase, void* stream);
/* 0 in *failed_seq: every exchange so far met all its peers; otherwise the call number of the last exchange that gave up
 * waiting (RAAE_PEER_TIMEOUT_S seconds, default 120) and skipped its update - the job is then inconsistent and must stop.
 * Synchronises the device. */
int raae_peer_status(raae_handle* h, unsigned* failed_seq);
int raae_peer_free(raae_handle* h);

/* Number of kernel launches issued by this handle so far (for bench.py's gpu_launches). */
int64_t raae_launch_count(const raae_handle* h);

/* Optional instrumentation: int64 [n_trials][32] device buffer; raae_train_epochs ADDS the SM cycles each trial's
 * CTA spent per stage type (slot 15 = whole kernel; slots documented in csrc/aae_step.cuh StageId).  NULL disables. */
int raae_set_profile_buffer(raae_handle* h, long long* prof);

#ifdef __cplusplus
}
#endif
#endif /* RANKAAE_B200_H_ */
