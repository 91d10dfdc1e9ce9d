"""Trial ensemble — the replacement of the reference's ipyparallel farm (`sc/cmd/train_sc.py:25-45,
127-143`): `trials` independent trainings of one config are partitioned across ranks (one process per
GPU, trial t -> rank t % world), and all trials of a rank train CONCURRENTLY inside one pair of kernel
launches per epoch (one CTA per trial).  There is no collective in the data path; the only exchange is
the final gather of `[5 metrics, time]` per trial (`torch.distributed`, NCCL on GPUs / gloo in tests).

Artifacts keep the reference layout: work_dir/training/job_{n+1}/{final.pt, losses.csv, messages.txt}
(train_sc.py:58-66, trainer.py:84-87, 270-283, 310).
"""
import os
import time

import numpy as np
import torch

from .logger import create_logger
from .trainer import LOSS_HEADER, build_modules, save_final


def shard_trials(trials, world=1, rank=0):
    """Global trial numbers handled by `rank`: t -> rank t % world (round robin keeps the shards within one
    trial of each other for any (trials, world))."""
    return list(range(rank, trials, world))


def format_loss_row(epoch, row12):
    """One losses.csv row exactly as trainer.py:271-279 formats it (tab after every comma, trailing comma)."""
    return f"{epoch:d},\t" + "".join(f"{v:.6f},\t" for v in row12)


def gather_results(local, trials, world=1, rank=0, device=None):
    """All-gathers per-trial result rows ([5 metrics, time_used], float64) into global trial order.
    `local` has one row per trial of shard_trials(trials, world, rank)."""
    local = np.asarray(local, dtype=np.float64).reshape(-1, 6)
    if world == 1:
        return local
    import torch.distributed as dist
    n_max = (trials + world - 1) // world
    buf = torch.full((n_max, 6), float("nan"), dtype=torch.float64, device=device)
    if len(local):
        buf[:len(local)] = torch.from_numpy(local).to(buf.device)
    out = [torch.empty_like(buf) for _ in range(world)]
    dist.all_gather(out, buf)
    res = np.full((trials, 6), np.nan)
    for r in range(world):
        mine = shard_trials(trials, world, r)
        res[mine] = out[r][:len(mine)].cpu().numpy()
    return res


def run_ensemble(work_dir, train_config, data_file, trials, verbose=False, device="cuda:0", rank=0, world=1,
                 timeout_hours=0, base_seed=0, epochs_per_call=25, write_artifacts=True, logger=None, raise_on_timeout=False):
    """Trains this rank's share of `trials` trials concurrently.  Returns [(metrics, time_used)] in the order of
    shard_trials(trials, world, rank) — the same tuples the reference's run_training returns (train_sc.py:102)."""
    from .dataloader import load_splits, to_device_pinned
    from .engine import Engine, auto_ctas_per_trial

    p = train_config
    mine = shard_trials(trials, world, rank)
    start = time.time()
    # one CSV parse per node (rank 0, binary cache next to the file), memory-mapped by every rank, pinned staging
    (st, at), (sv, av), _ = load_splits(data_file, n_aux=p.n_aux, rank=rank, world=world)
    max_epoch = int(p.max_epoch)
    if not mine:
        if world > 1 and timeout_hours:                             # keep the timeout agreement collective (see below)
            import torch.distributed as dist
            for _ in range(0, max_epoch, epochs_per_call):
                flag = torch.zeros(1, device=device if dist.get_backend() == "nccl" else "cpu")
                dist.all_reduce(flag, op=dist.ReduceOp.MAX)
                if flag.item() > 0:
                    break
        return []
    cfg = dict(p.to_dict())
    cfg.setdefault("epoch_stop_smooth", 500)
    # few trials per GPU (e.g. 64 trials over 8 GPUs): a thread-block cluster per trial, as large as still fits in one wave
    cfg["ctas_per_trial"] = auto_ctas_per_trial(cfg, (trials + world - 1) // world, device)
    eng = Engine(cfg, n_trials=len(mine), device=device, max_rows=max(int(p.batch_size), sv.shape[0]),
                 seeds=[base_seed + t for t in mine])
    st, at, sv, av = (to_device_pinned(a, eng.device) for a in (st, at, sv, av))
    modules, loggers = [], []
    for i, t in enumerate(mine):
        job_dir = f"{work_dir}/training/job_{t + 1}"
        if write_artifacts:
            os.makedirs(job_dir, exist_ok=True)
            lg = create_logger(f"subtraining_{t + 1}", os.path.join(job_dir, "messages.txt"))
            ll = create_logger(f"losses_{t + 1}", os.path.join(job_dir, "losses.csv"), simple_fmt=True)
            lg.info(f"Training started for trial {t + 1}.")
            if verbose:
                lg.info("Use GPU")
            ll.info(LOSS_HEADER)
            loggers.append((lg, ll))
        mods = build_modules(p, seed=base_seed + t)
        eng.load_modules(i, *mods)
        modules.append(mods)
    eng.bind_dataset(st, at, sv, av)
    metrics = np.zeros((len(mine), 6))
    failed = np.zeros(len(mine), dtype=bool)
    timed_out = False
    epoch = 0
    while epoch < max_epoch:
        n = min(epochs_per_call, max_epoch - epoch)
        losses, mets = eng.train_epochs(epoch, n)
        torch.cuda.synchronize(eng.device)
        losses, mets = losses.cpu().numpy(), mets.cpu().numpy()
        if write_artifacts:
            for e in range(n):
                if (epoch + e) % 10 == 0:
                    for i in range(len(mine)):
                        loggers[i][1].info(format_loss_row(epoch + e, losses[e, i]))
        metrics = mets[-1]
        # per-trial failure isolation (the reference aborts the whole map_sync when one engine raises, train_sc.py:91-97):
        # a trial whose metrics went non-finite is flagged and keeps its slot; the others are unaffected (independent CTAs)
        failed |= ~np.isfinite(mets[:, :, :6]).all(axis=(0, 2))
        epoch += n
        over = bool(timeout_hours and time.time() - start > timeout_hours * 3600)
        if world > 1 and timeout_hours:
            # the ranks agree on the decision (one flag per chunk): a rank that raised alone would leave the others
            # hanging in the final gather
            import torch.distributed as dist
            flag = torch.tensor([1.0 if over else 0.0], device=eng.device if dist.get_backend() == "nccl" else "cpu")
            dist.all_reduce(flag, op=dist.ReduceOp.MAX)
            over = bool(flag.item() > 0)
        if over:
            timed_out = True
            break
    time_used = time.time() - start
    out = []
    for i, t in enumerate(mine):
        m = [float(v) for v in metrics[i, :5]]
        if failed[i]:
            m = [float("nan")] * 5                                  # flagged: the gather and the report see NaN, not garbage
        if write_artifacts:
            if timed_out:
                loggers[i][0].warning(f"Training Overtime! Stopped after epoch {epoch} of {max_epoch}.")   # train_sc.py:21-22
            if failed[i]:
                loggers[i][0].warning("Trial failed: non-finite losses / metrics; no final.pt written.")
            else:
                eng.store_modules(i, *modules[i])
                save_final(modules[i], f"{work_dir}/training/job_{t + 1}/final.pt")
            loggers[i][0].info(m)
            loggers[i][0].info(f"Training finished. Time used: {time_used:.2f}s.\n\n")
        out.append((m, time_used))
    eng.close()
    if timed_out and raise_on_timeout:
        raise Exception("Training Overtime!")                   # after every rank has left the loop together
    return out
