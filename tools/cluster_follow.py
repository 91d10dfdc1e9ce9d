#!/usr/bin/env python
"""How closely does one free-running epoch (20 updates) in a thread-block cluster follow the one-CTA kernel from the same warm
state?  Prints the parameter rel-L2 per (shuffle, trial, net) - the spread shows how much of the difference is the chaotic
amplification of float32 reduction-order noise (tests/test_cluster_gpu.py::test_cluster_production_epochs_follow_single_cta).
usage: python tools/cluster_follow.py [n_shuffles]"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__ as graft  # noqa: E402

graft.build()
from rankaae_b200 import _lib as L  # noqa: E402
from rankaae_b200.engine import Engine, make_config  # noqa: E402
from rankaae_b200.synthetic import synthetic_dataset  # noqa: E402
from rankaae_b200.trainer import init_trial_state  # noqa: E402
from bench import EXAMPLE  # noqa: E402

NS = int(sys.argv[1]) if len(sys.argv) > 1 else 6
cfg = dict(EXAMPLE, batch_size=512, max_epoch=40)


def engine(c, n_trials=2):
    return Engine(dict(cfg, ctas_per_trial=c), n_trials=n_trials, device="cuda:0", max_rows=512, seeds=list(range(n_trials)))


spec, aux = synthetic_dataset(2400, n_aux=cfg["n_aux"], dim=cfg["dim_in"], seed=1)
data = (spec[:1680], aux[:1680], spec[1680:2040], aux[1680:2040])
warm = engine(1)
for t in range(2):
    init_trial_state(warm, t, cfg, seed=t)
warm.bind_dataset(*data)
warm.train_epochs(0, 3)
torch.cuda.synchronize()
snapshot = warm.state.clone()
lay = L.query_layout(make_config(cfg, 2, 512))
rows = []
for s in range(NS):
    g = torch.Generator(device="cuda:0").manual_seed(1000 + s)
    perm = warm.make_perm(1, generator=g)
    outs = {}
    for c in (1, 2, 8):
        eng = engine(c)
        eng.state.copy_(snapshot)
        eng.bind_dataset(*data)
        eng.train_epochs(3, 1, perm)
        torch.cuda.synchronize()
        outs[c] = eng.state.clone().cpu().numpy().astype(np.float64)
        eng.close()
    for c in (2, 8):
        for t in range(2):
            for ni in range(3):
                n = lay.net[ni]
                a, b = outs[1][t, n.param_off:n.param_off + n.n_params], outs[c][t, n.param_off:n.param_off + n.n_params]
                rows.append((s, c, t, ni, float(np.linalg.norm(b - a) / np.linalg.norm(a))))
v = np.array([r[-1] for r in rows])
print("tensor_cores", os.environ.get("RAAE_TENSOR_CORES", "default"), "n", len(v), "median %.2e p90 %.2e max %.2e" % (np.median(v), np.quantile(v, 0.9), v.max()))
for r in sorted(rows, key=lambda r: -r[-1])[:8]:
    print("  shuffle %d ctas %d trial %d net %d: %.2e" % r)
