"""BASELINE.json configs[4] / the north-star target: a hyper-parameter-sweep ensemble of 1024 trials on synthetic
100k x 256 spectra (70 000 train / 15 000 validation rows, batch 1024 -> 69 batches per epoch), partitioned across the
ranks (trial t -> rank t % world, no data-path collective, one NCCL all-gather of [5 metrics, time] per trial at the end).

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29551 \
        tools/sweep_1024.py [--trials 1024] [--epochs 4] [--out gpurun_out/sweep_1024.json]

All trials of a rank are resident at once and train inside ONE raae_train_kernel + ONE raae_val_kernel launch per epoch
(128 CTAs on 148 SMs at 8 GPUs).  Per-trial hyper-parameters are rows of the float64 hp table (no recompile): the
continuous knobs in which the reference's shipped configs differ.  Timed with CUDA events on the
launching stream after one warm-up epoch, max over ranks; every epoch includes the validation block, the metrics and the
scheduler step.  Reports samples/s, steps/s, trials/hour (for the stated epoch budget) and the spread of the final metrics.
"""
import argparse
import json
import os
import sys
import time

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from rankaae_b200.synthetic import synthetic_dataset                   # noqa: E402
from rankaae_b200.engine import Engine                                 # noqa: E402
from rankaae_b200.ensemble import gather_results, shard_trials         # noqa: E402
from rankaae_b200.trainer import init_trial_state                      # noqa: E402

BASE = dict(max_epoch=2000, batch_size=1024, gradient_reversal=True, alpha_flat_step=739, alpha_limit=0.7172,
            decoder_activation="Softplus", dis_beta=1.1, dis_dropout_rate=0.056, dis_noise=0.56, n_aux=5, nstyle=6,
            ae_form="FC", dim_in=256, dim_out=256, n_layers=5, FC_discriminator_layers=3, dropout_rate=0.04,
            lr_base=0.001, lr_ratio_Corr=10, lr_ratio_Mutual=1, lr_ratio_Reconn=10, lr_ratio_Smooth=1,
            lr_ratio_dis=1, optimizer_name="AdamW", spec_noise=0.02, use_flex_spec_target=True,
            weight_decay=0.01, kendall_activation=True, epoch_stop_smooth=1500)
N_ROWS, N_TRAIN, N_VAL = 100_000, 70_000, 15_000


def sweep_point(t):
    """Hyper-parameters of global trial t (deterministic in t, independent of the partition)."""
    r = np.random.default_rng(90_000 + t)
    lu = lambda lo, hi: float(np.exp(r.uniform(np.log(lo), np.log(hi))))
    return dict(BASE, lr_base=lu(3e-4, 3e-3), lr_ratio_Corr=lu(3, 30), lr_ratio_Reconn=lu(3, 30), lr_ratio_Mutual=lu(0.5, 2),
                lr_ratio_Smooth=lu(0.5, 2), lr_ratio_dis=lu(0.5, 2), weight_decay=lu(1e-3, 1e-1),
                dropout_rate=float(r.uniform(0.0, 0.1)), dis_dropout_rate=float(r.uniform(0.0, 0.1)),
                dis_noise=float(r.uniform(0.2, 0.8)), spec_noise=float(r.uniform(0.0, 0.05)),
                alpha_limit=float(r.uniform(0.5, 1.0)), alpha_flat_step=float(r.uniform(300, 1200)))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--trials", type=int, default=1024)
    ap.add_argument("--epochs", type=int, default=4, help="timed epochs (after one warm-up epoch)")
    ap.add_argument("--budget-epochs", type=int, default=2000, help="epoch budget trials/hour is quoted for (example max_epoch)")
    ap.add_argument("--out", default="")
    args = ap.parse_args()
    rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    dev = torch.device(f"cuda:{local}")
    torch.cuda.set_device(dev)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    t_wall0 = time.time()
    spec, aux = synthetic_dataset(N_ROWS, BASE["n_aux"], BASE["dim_in"], seed=5, dtype=np.float32)
    mine = shard_trials(args.trials, world, rank)
    per_trial = [sweep_point(t) for t in mine]
    eng = Engine(BASE, n_trials=len(mine), device=dev, max_rows=N_VAL, seeds=mine, per_trial_cfg=per_trial)
    for i, t in enumerate(mine):
        init_trial_state(eng, i, per_trial[i], seed=t)
    eng.bind_dataset(spec[:N_TRAIN], aux[:N_TRAIN], spec[N_TRAIN:N_TRAIN + N_VAL], aux[N_TRAIN:N_TRAIN + N_VAL])
    t_setup = time.time() - t_wall0

    def sync():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    n0 = eng.launch_count
    eng.train_epochs(0, 1)                                              # warm-up epoch (also epoch 0 of the trials)
    sync()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    losses, metrics = eng.train_epochs(1, args.epochs)
    b.record()
    sync()
    ms = torch.tensor([a.elapsed_time(b)], device=dev)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    sec = float(ms) / 1e3
    launches = eng.launch_count - n0
    me = metrics[-1].double().cpu().numpy()                             # [n_local][6]
    local_rows = np.concatenate([me[:, :5], np.full((len(mine), 1), sec)], axis=1)
    res = gather_results(local_rows, args.trials, world, rank, device=dev)
    if rank == 0:
        steps_per_epoch = (N_TRAIN + BASE["batch_size"] - 1) // BASE["batch_size"]
        sec_per_epoch = sec / args.epochs
        finite = bool(np.isfinite(res[:, :5]).all())
        q = lambda c: [float(v) for v in np.nanpercentile(res[:, c], [5, 50, 95])]
        out = {"workload": f"{args.trials}-trial hyper-parameter sweep, synthetic {N_ROWS} x 256 spectra ({N_TRAIN} train / {N_VAL} "
                           f"validation rows), batch 1024, {steps_per_epoch} batches per epoch, {world} GPU(s), "
                           f"{len(mine)} trials resident per GPU; every epoch incl. validation + metrics + scheduler",
               "n_gpus": world, "trials": args.trials, "timed_epochs": args.epochs, "seconds": sec,
               "sec_per_epoch": sec_per_epoch,
               "samples_per_sec": args.trials * N_TRAIN * args.epochs / sec,
               "samples_per_sec_per_gpu": args.trials * N_TRAIN * args.epochs / sec / world,
               "steps_per_sec": args.trials * steps_per_epoch * args.epochs / sec,
               "trials_per_hour": args.trials / (sec_per_epoch * args.budget_epochs / 3600.0),
               "trials_per_hour_epoch_budget": args.budget_epochs,
               "hours_for_the_sweep": sec_per_epoch * args.budget_epochs / 3600.0,
               "gpu_launches_rank0": int(launches), "setup_seconds_rank0": t_setup, "all_finite": finite,
               "after_epochs": args.epochs + 1,
               "val_recon_mse_p5_p50_p95": q(1), "min_shapiro_W_p5_p50_p95": q(0), "max_abs_spearman_p5_p50_p95": q(3),
               "val_kendall_p5_p50_p95": q(4),
               "state_plus_scratch_GB_per_gpu": (eng.state.numel() + eng.scratch.numel()) * 4 / 1e9}
        line = json.dumps(out)
        print(line)
        if args.out:
            os.makedirs(os.path.dirname(os.path.abspath(args.out)), exist_ok=True)
            with open(args.out, "w") as f:
                f.write(line + "\n")
        assert finite, "non-finite metrics in the sweep"
    eng.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
