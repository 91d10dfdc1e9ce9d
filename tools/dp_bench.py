"""Data-parallel single-trial mode (BASELINE.json configs[3]) on N GPUs: peer-memory exchange vs NCCL.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29541 \
        tools/dp_bench.py [--steps 40] [--out gpurun_out/dp_bench.json]

Workload: synthetic spectra x 256 points, 6 descriptors, per-rank batch 512, `--steps` batches per
rank per timed epoch (a bounded slice of the 1 M-row set: the per-step cost does not depend on the row count).  Two
measurements per exchange mode, CUDA events on the launching stream, max over ranks:
  * train epoch (all batches x 5 phases, without the validation block) -> steps/s, global samples/s;
  * the exchange + update alone (200 back-to-back calls on the phase-2 vector, 237 KB) -> microseconds per phase.
"""
import argparse
import json
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from rankaae_b200.synthetic import synthetic_dataset                   # noqa: E402
from rankaae_b200 import _lib as L                                     # noqa: E402
from rankaae_b200.dp import DataParallelTrainer                        # noqa: E402

CFG = dict(max_epoch=100, batch_size=512, gradient_reversal=True, alpha_flat_step=739, alpha_limit=0.7172,
           decoder_activation="Softplus", dis_beta=1.1, dis_dropout_rate=0.056, dis_noise=0.56, n_aux=6, nstyle=6,
           ae_form="FC", dim_in=256, dim_out=256, n_layers=5, FC_discriminator_layers=3, dropout_rate=0.04,
           lr_base=0.001, lr_ratio_Corr=10, lr_ratio_Mutual=1, lr_ratio_Reconn=10, lr_ratio_Smooth=1,
           lr_ratio_dis=1, optimizer_name="AdamW", spec_noise=0.02, use_flex_spec_target=True,
           weight_decay=0.01, kendall_activation=True, epoch_stop_smooth=80)


def timed(fn, dev, world):
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize(dev)
    a.record()
    fn()
    b.record()
    torch.cuda.synchronize(dev)
    t = torch.tensor([a.elapsed_time(b)], device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t)


def train_only(dp, epoch, perm):
    """The batch loop of DataParallelTrainer.train_epoch without the validation block."""
    eng = dp.engine
    for s in range(dp.n_steps):
        for o in range(L.NUM_PHASES):
            for k in range(L.NUM_PHASES):
                dp._gptr[k] = dp._grad_ptrs[k] if k == o else None
            L.check(eng.lib.raae_train_phase(eng.handle, epoch, s, 1 << o, perm.data_ptr(), dp._gptr, eng.stream))
            dp._exchange_update(o)


def exchange_update(dp, o):
    dp._exchange_update(o)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=40)
    ap.add_argument("--replicas", type=int, default=1, help="data-parallel replicas of the trial per GPU (one SM each)")
    ap.add_argument("--out", default="")
    args = ap.parse_args()
    rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    dev = torch.device(f"cuda:{local}")
    torch.cuda.set_device(dev)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    bs = CFG["batch_size"]
    V = args.replicas
    n_train = bs * args.steps * world * V
    spec, aux = synthetic_dataset(n_train + 1050, CFG["n_aux"], CFG["dim_in"], seed=3, dtype=np.float32)
    res = {"workload": f"config #4 shape: 256-point spectra, 6 descriptors, per-rank batch {bs}, {args.steps} batches per rank, "
                       f"{world} rank(s) x {V} replica(s) per GPU (global batch {bs * world * V}), 5 phases per batch", "n_gpus": world,
           "replicas_per_gpu": V}
    for exchange in ("peer", "nccl"):
        dp = DataParallelTrainer(CFG, spec[:n_train], aux[:n_train], spec[n_train:], aux[n_train:], dev, rank, world, seed=1,
                                 exchange=exchange, replicas=V)
        perm = dp.make_perm()
        for e in range(2):                                              # warm-up epochs
            train_only(dp, e, perm)
        ms = min(timed(lambda: train_only(dp, 2 + r, perm), dev, world) for r in range(3))
        reps = 200
        for _ in range(20):
            exchange_update(dp, 2)
        us = min(timed(lambda: [exchange_update(dp, 2) for _ in range(reps)], dev, world) for r in range(3)) * 1e3 / reps
        v = dp.state_vector()
        same = True
        if world > 1:                                                   # every rank must hold the same weights, bit for bit
            ref = v.clone()
            dist.broadcast(ref, 0)
            flag = torch.tensor([float(torch.equal(v, ref))], device=dev)
            dist.all_reduce(flag, op=dist.ReduceOp.MIN)
            same = bool(flag.item() == 1.0)
        finite = bool(torch.isfinite(v).all().item())
        res[exchange] = {"ranks_bit_identical": same, "finite": finite, "ms_per_epoch": ms, "ms_per_step": ms / dp.n_steps, "steps_per_sec": dp.n_steps / ms * 1e3,
                         "samples_per_sec": dp.n_steps * bs * world * V / ms * 1e3, "exchange_update_us": us,
                         "launches_per_phase": 2 if exchange == "peer" else 3 + (2 if world > 1 else 0) + (3 if V > 1 else 0)}
        dp.close()
    if rank == 0:
        line = json.dumps(res)
        print(line)
        if args.out:
            os.makedirs(os.path.dirname(os.path.abspath(args.out)), exist_ok=True)
            with open(args.out, "w") as f:
                f.write(line + "\n")
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
