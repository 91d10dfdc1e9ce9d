"""The BASELINE.json configurations beside the headline ensemble, measured by bench.py as sub-records of its JSON line:

  strong_64        configs[2] as written: 64 trials of the example config partitioned over the N GPUs of the run
                   (sc/cmd/train_sc.py:127-143); every trial gets a thread-block cluster sized so that the rank's trials
                   fill its SMs (ctas_per_trial 2 / 4 / 8 at 1 / 2 / 4-8 GPUs); strong scaling: the job is fixed.
  dp_single_trial  configs[3]: synthetic 1 M x 256 spectra, 6 descriptors, per-GPU batch 512, ONE trial data-parallel over the
                   N GPUs through rankaae_b200/dp.py with the fused peer-memory exchange and with NCCL.
  sweep_1024       configs[4]: 1024-trial hyper-parameter sweep on synthetic 100k x 256 spectra (trials sharded over ranks).

Every timing: CUDA events on the launching stream after warm-up, barrier + synchronize on both sides, max over ranks.
"""
import os
import sys
import time

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
from rankaae_b200 import _lib as L                                     # noqa: E402
from rankaae_b200.engine import Engine                                 # noqa: E402
from rankaae_b200.ensemble import gather_results, shard_trials         # noqa: E402
from rankaae_b200.synthetic import synthetic_dataset, synthetic_dataset_torch   # noqa: E402
from rankaae_b200.trainer import init_trial_state                      # noqa: E402


def _sync(dev, world):
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize(dev)


def _timed(fn, dev, world):
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    _sync(dev, world)
    a.record()
    out = fn()
    b.record()
    _sync(dev, world)
    t = torch.tensor([a.elapsed_time(b)], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item()), out


def pick_cluster(trials_on_rank, dev_index=0):
    """Largest cluster size whose clusters (one per trial) still fit the GPU in ONE wave: a cluster lives inside a GPC and
    every CTA takes a whole SM, so the driver's occupancy query (raae_max_clusters) decides, not 148 / C."""
    for c in (8, 4, 2):
        if trials_on_rank <= L.max_clusters(c, dev_index):
            return c
    return 1


def single_trial(cfg, data, dev, warmup, steps, n_train):
    """configs[1]: one trial of the example config resident, as a cluster of 8 CTAs and as one CTA."""
    out = {"workload": "BASELINE configs[1]: example config, trials=1; one epoch = 5 batches x 5 phases + validation block"}
    for c in (8, 1):
        eng = Engine(dict(cfg, ctas_per_trial=c), n_trials=1, device=dev, max_rows=1056, seeds=[12345])
        init_trial_state(eng, 0, cfg, seed=12345)
        eng.bind_dataset(*data)
        eng.train_epochs(0, warmup)
        ms, _ = _timed(lambda: eng.train_epochs(warmup, steps), dev, 1)
        ms /= steps
        rec = {"ms_per_epoch": ms, "samples_per_sec": n_train / (ms * 1e-3), "hours_per_2000_epochs": ms * 2000.0 / 3.6e6}
        if c == 8:
            out.update(rec, ctas_per_trial=8)
        else:
            out["one_cta"] = rec
        eng.close()
    return out


def strong_64(cfg, data, dev, rank, world, warmup, steps, n_train, trials=64):
    mine = shard_trials(trials, world, rank)
    c = pick_cluster((trials + world - 1) // world, dev.index or 0)
    eng = Engine(dict(cfg, ctas_per_trial=c), n_trials=len(mine), device=dev, max_rows=1056, seeds=mine)
    for i, t in enumerate(mine):
        init_trial_state(eng, i, cfg, seed=t)
    eng.bind_dataset(*data)
    eng.train_epochs(0, warmup)
    ms, (losses, metrics) = _timed(lambda: eng.train_epochs(warmup, steps), dev, world)
    ms /= steps
    finite = bool(torch.isfinite(metrics).all().item())
    eng.close()
    return {"workload": f"BASELINE configs[2] as written: {trials} trials of the example config partitioned over {world} GPU(s) "
                        f"(trial t -> rank t % world, no data-path collective), one {c}-CTA cluster per trial",
            "scaling": "strong", "trials": trials, "trials_per_gpu": len(mine), "ctas_per_trial": c,
            "co_resident_clusters": {str(k): L.max_clusters(k, dev.index or 0) for k in (2, 4, 8)}, "ms_per_epoch": ms,
            "samples_per_sec": trials * n_train / (ms * 1e-3), "trials_per_hour_2000_epochs": trials * 3600.0 / (ms * 1e-3 * 2000.0),
            "finite": finite,
            "limiter": "raae_train_kernel: every stage of a trial is a dependent chain (load -> MMA -> BatchNorm reduction -> "
                       "cluster barrier), so an epoch costs the same 10-28 ms whatever the number of resident clusters; no collective"}


DP_CFG = dict(max_epoch=100, batch_size=512, gradient_reversal=True, alpha_flat_step=739, alpha_limit=0.7172,
              decoder_activation="Softplus", dis_beta=1.1, dis_dropout_rate=0.056, dis_noise=0.56, n_aux=6, nstyle=6,
              ae_form="FC", dim_in=256, dim_out=256, n_layers=5, FC_discriminator_layers=3, dropout_rate=0.04,
              lr_base=0.001, lr_ratio_Corr=10, lr_ratio_Mutual=1, lr_ratio_Reconn=10, lr_ratio_Smooth=1,
              lr_ratio_dis=1, optimizer_name="AdamW", spec_noise=0.02, use_flex_spec_target=True,
              weight_decay=0.01, kendall_activation=True, epoch_stop_smooth=80)


def dp_single_trial(dev, rank, world, steps=40, n_rows=1_000_000, ctas=4):
    """configs[3].  Every rank generates ITS shard of the 700 000 training rows on the device (1 M x 256 x 4 B does not need
    to cross PCIe) and the first 15 000 validation rows (the in-kernel sort caps n_val at 16 384).
    Timed: `steps` batches per rank (5 split-phase launches + 5 exchanges each) out of the shard's 700 000 / (512 N)."""
    from rankaae_b200.dp import DataParallelTrainer
    n_train = int(n_rows * 0.7)
    per = n_train // world
    spec, aux = synthetic_dataset_torch(per, 6, 256, seed=100 + rank, device=dev)
    sv, av = synthetic_dataset_torch(15_000, 6, 256, seed=99, device=dev)
    res = {"workload": f"BASELINE configs[3]: synthetic {n_rows} x 256 spectra, 6 descriptors ({n_train} training rows sharded over "
                       f"{world} GPU(s), {per} each), per-GPU batch 512, one trial data-parallel (global batch {512 * world}), "
                       f"one {ctas}-CTA cluster per rank; timed: {steps} batches per rank x 5 phases (launch + gradient exchange "
                       "fused with AdamW)", "n_gpus": world, "ctas_per_trial": ctas, "rows_per_rank": per, "timed_batches": steps}

    def train_some(dp, epoch, perm, n):
        eng = dp.engine
        for s in range(n):
            for o in range(L.NUM_PHASES):
                for k in range(L.NUM_PHASES):
                    dp._gptr[k] = dp._grad_ptrs[k] if k == o else None
                L.check(eng.lib.raae_train_phase(eng.handle, epoch, s, 1 << o, perm.data_ptr(), dp._gptr, eng.stream))
                dp._exchange_update(o)

    for exchange in ("peer", "nccl"):
        dp = DataParallelTrainer(dict(DP_CFG, ctas_per_trial=ctas), spec, aux, sv, av, dev, rank, world, seed=1, exchange=exchange,
                                 presharded=True)
        perm = dp.make_perm()
        train_some(dp, 0, perm, 8)
        ms, _ = _timed(lambda: train_some(dp, 1, perm, steps), dev, world)
        reps = 100
        for _ in range(10):
            dp._exchange_update(2)
        us, _ = _timed(lambda: [dp._exchange_update(2) for _ in range(reps)], dev, world)
        v = dp.state_vector()
        same = True
        if world > 1:
            ref = v.clone()
            dist.broadcast(ref, 0)
            flag = torch.tensor([float(torch.equal(v, ref))], device=dev)
            dist.all_reduce(flag, op=dist.ReduceOp.MIN)
            same = bool(flag.item() == 1.0)
        res[exchange] = {"ms_per_step": ms / steps, "steps_per_sec": steps / ms * 1e3,
                         "samples_per_sec": steps * 512 * world / ms * 1e3, "exchange_update_us": us * 1e3 / reps,
                         "ranks_bit_identical": same, "finite": bool(torch.isfinite(v).all().item()),
                         "seconds_per_epoch_700k_rows": (per // 512) * (ms / steps) * 1e-3}
        dp.close()
    res["limiter"] = ("five split-phase raae_train_kernel launches per batch (one 4-CTA cluster per GPU) + five exchange launches: "
                      "launch- and latency-bound; the exchange (raae_adam_peer_kernel over NVLink peer loads vs NCCL all-reduce + "
                      "AdamW) is 5-10 % of the step")
    return res


SWEEP_BASE = dict(max_epoch=2000, batch_size=1024, gradient_reversal=True, alpha_flat_step=739, alpha_limit=0.7172,
                  decoder_activation="Softplus", dis_beta=1.1, dis_dropout_rate=0.056, dis_noise=0.56, n_aux=5, nstyle=6,
                  ae_form="FC", dim_in=256, dim_out=256, n_layers=5, FC_discriminator_layers=3, dropout_rate=0.04,
                  lr_base=0.001, lr_ratio_Corr=10, lr_ratio_Mutual=1, lr_ratio_Reconn=10, lr_ratio_Smooth=1,
                  lr_ratio_dis=1, optimizer_name="AdamW", spec_noise=0.02, use_flex_spec_target=True,
                  weight_decay=0.01, kendall_activation=True, epoch_stop_smooth=1500)


def sweep_point(t):
    """Hyper-parameters of global trial t (deterministic in t, independent of the partition): the continuous knobs in which
    the reference's shipped configs differ."""
    r = np.random.default_rng(90_000 + t)
    lu = lambda lo, hi: float(np.exp(r.uniform(np.log(lo), np.log(hi))))
    return dict(SWEEP_BASE, lr_base=lu(3e-4, 3e-3), lr_ratio_Corr=lu(3, 30), lr_ratio_Reconn=lu(3, 30), lr_ratio_Mutual=lu(0.5, 2),
                lr_ratio_Smooth=lu(0.5, 2), lr_ratio_dis=lu(0.5, 2), weight_decay=lu(1e-3, 1e-1),
                dropout_rate=float(r.uniform(0.0, 0.1)), dis_dropout_rate=float(r.uniform(0.0, 0.1)),
                dis_noise=float(r.uniform(0.2, 0.8)), spec_noise=float(r.uniform(0.0, 0.05)),
                alpha_limit=float(r.uniform(0.5, 1.0)), alpha_flat_step=float(r.uniform(300, 1200)))


def sweep_1024(dev, rank, world, trials=1024, epochs=2, budget_epochs=2000):
    n_rows, n_train, n_val = 100_000, 70_000, 15_000
    t0 = time.time()
    spec, aux = synthetic_dataset(n_rows, 5, 256, seed=5, dtype=np.float32)
    mine = shard_trials(trials, world, rank)
    per_trial = [sweep_point(t) for t in mine]
    eng = Engine(SWEEP_BASE, n_trials=len(mine), device=dev, max_rows=n_val, seeds=mine, per_trial_cfg=per_trial)
    for i, t in enumerate(mine):
        init_trial_state(eng, i, per_trial[i], seed=t)
    eng.bind_dataset(spec[:n_train], aux[:n_train], spec[n_train:n_train + n_val], aux[n_train:n_train + n_val])
    setup = time.time() - t0
    eng.train_epochs(0, 1)
    ms, (losses, metrics) = _timed(lambda: eng.train_epochs(1, epochs), dev, world)
    sec = ms / 1e3
    me = metrics[-1].double().cpu().numpy()
    rows = np.concatenate([me[:, :5], np.full((len(mine), 1), sec)], axis=1)
    res = gather_results(rows, trials, world, rank, device=dev)
    eng.close()
    spe = (n_train + 1023) // 1024
    q = lambda c: [float(v) for v in np.nanpercentile(res[:, c], [5, 50, 95])]
    return {"workload": f"BASELINE configs[4]: {trials}-trial hyper-parameter sweep, synthetic {n_rows} x 256 spectra ({n_train} train / "
                        f"{n_val} validation rows), batch 1024, {spe} batches per epoch, {world} GPU(s), {len(mine)} trials resident per GPU "
                        "(one CTA each); every epoch incl. validation + metrics + scheduler", "n_gpus": world, "trials": trials,
            "timed_epochs": epochs, "sec_per_epoch": sec / epochs, "samples_per_sec": trials * n_train * epochs / sec,
            "steps_per_sec": trials * spe * epochs / sec, "trials_per_hour": trials / (sec / epochs * budget_epochs / 3600.0),
            "trials_per_hour_epoch_budget": budget_epochs, "hours_for_the_sweep": sec / epochs * budget_epochs / 3600.0,
            "all_finite": bool(np.isfinite(res[:, :5]).all()), "setup_seconds_rank0": setup,
            "val_recon_mse_p5_p50_p95": q(1), "min_shapiro_W_p5_p50_p95": q(0),
            "limiter": "raae_train_kernel (one CTA per trial, 128 of 148 SMs per GPU busy); no data-path collective, one NCCL "
                       "all-gather of 6 floats per trial at the end"}


def reference_api(cfg, n_rows, dev_index, epochs=40, trials=148):
    """The path a user of the reference takes, timed by wall clock from the CSV on disk to final.pt on disk (N = 1):
      trainer   `Trainer.from_data(csv, ...).train()` - one trial, one 8-CTA cluster (trainer.py:411-474, 65-315);
      ensemble  `run_ensemble(...)` - what `train_sc -c fix_config.yaml` runs for `trials` trials (train_sc.py:105-156):
                binary-cache loader, all trials resident, losses.csv / messages.txt / final.pt per job."""
    import tempfile
    import yaml
    from rankaae_b200.ensemble import run_ensemble
    from rankaae_b200.logger import create_logger
    from rankaae_b200.parameter import Parameters
    from rankaae_b200.synthetic import write_csv
    from rankaae_b200.trainer import Trainer
    work = tempfile.mkdtemp(prefix="raae_api_")
    spec, aux = synthetic_dataset(n_rows, cfg["n_aux"], cfg["dim_in"], seed=0, dtype=np.float32)
    csv = os.path.join(work, "data.csv")
    write_csv(csv, spec, aux)
    n_train = int(n_rows * 0.7)
    c1 = dict(cfg, max_epoch=epochs, trials=1, timeout=1, verbose=False, data_file="data.csv")   # cluster size chosen by Trainer
    with open(os.path.join(work, "fix_config.yaml"), "w") as f:
        yaml.safe_dump(c1, f)
    out = {}
    t0 = time.time()
    job = os.path.join(work, "single")
    os.makedirs(job)
    tr = Trainer.from_data(csv, igpu=dev_index, verbose=False, work_dir=job, config_parameters=Parameters(c1),
                           logger=create_logger("b_msg", os.path.join(job, "messages.txt")),
                           loss_logger=create_logger("b_loss", os.path.join(job, "losses.csv"), simple_fmt=True))
    t1 = time.time()
    tr.train()
    torch.cuda.synchronize()
    t2 = time.time()
    out["trainer"] = {"api": "Trainer.from_data(csv).train()", "epochs": epochs, "ctas_per_trial": int(tr.engine.ccfg.ctas_per_trial), "load_s": t1 - t0, "train_s": t2 - t1,
                      "samples_per_sec_incl_load": epochs * n_train / (t2 - t0), "samples_per_sec_train": epochs * n_train / (t2 - t1),
                      "final_pt": os.path.exists(os.path.join(job, "final.pt"))}
    tr.engine.close()
    cN = dict(cfg, max_epoch=epochs, trials=trials, timeout=1, verbose=False, data_file="data.csv")
    t0 = time.time()
    res = run_ensemble(work, Parameters(cN), csv, trials, device=f"cuda:{dev_index}", epochs_per_call=epochs)
    t1 = time.time()
    out["ensemble"] = {"api": "run_ensemble (train_sc farm)", "trials": trials, "epochs": epochs, "wall_s": t1 - t0,
                       "samples_per_sec_incl_load_and_artifacts": trials * epochs * n_train / (t1 - t0),
                       "trials_per_hour_2000_epochs_extrapolated_from_wall": trials * 3600.0 / ((t1 - t0) * 2000.0 / epochs),
                       "final_pt_written": sum(os.path.exists(os.path.join(work, "training", f"job_{i + 1}", "final.pt")) for i in range(trials)),
                       "all_finite": bool(np.isfinite(np.array([m for m, _ in res])).all())}
    import shutil
    shutil.rmtree(work, ignore_errors=True)
    return out
