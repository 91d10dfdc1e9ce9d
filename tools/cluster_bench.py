#!/usr/bin/env python
"""ms per epoch (train + validation launches) of T resident trials at cluster size C, example configuration.
usage: python tools/cluster_bench.py "T:C,T:C,..." [epochs]"""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__ as graft  # noqa: E402

graft.build()
from bench import EXAMPLE, N_TRAIN, synthetic_arrays  # noqa: E402
from rankaae_b200.engine import Engine  # noqa: E402
from rankaae_b200.trainer import init_trial_state  # noqa: E402

pairs = [tuple(int(v) for v in p.split(":")) for p in (sys.argv[1] if len(sys.argv) > 1 else "1:1,1:2,1:4,1:8").split(",")]
E = int(sys.argv[2]) if len(sys.argv) > 2 else 10
data = synthetic_arrays()
out = []
for T, C in pairs:
    eng = Engine(dict(EXAMPLE, ctas_per_trial=C), n_trials=T, device="cuda:0", max_rows=1056)
    for t in range(T):
        init_trial_state(eng, t, EXAMPLE, seed=t)
    eng.bind_dataset(*data)
    eng.train_epochs(0, 3)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    losses, metrics = eng.train_epochs(3, E)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / E
    rec = {"trials": T, "ctas_per_trial": C, "ms_per_epoch": ms, "samples_per_sec": T * N_TRAIN / (ms * 1e-3),
           "finite": bool(torch.isfinite(metrics).all().item()), "val_recon_last": float(metrics[-1, :, 1].mean().item())}
    print(json.dumps(rec))
    out.append(rec)
    eng.close()
