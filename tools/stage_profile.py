#!/usr/bin/env python
"""Per-stage-type cycle breakdown of raae_train_kernel (in-kernel clock64 counters, thread 0 of each CTA).
usage: python tools/stage_profile.py [trials] [epochs]"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__ as graft  # noqa: E402

graft.build()
from bench import EXAMPLE, synthetic_arrays  # noqa: E402
from rankaae_b200 import _lib as L  # noqa: E402
from rankaae_b200.engine import Engine  # noqa: E402
from rankaae_b200.trainer import init_trial_state  # noqa: E402

T = int(sys.argv[1]) if len(sys.argv) > 1 else 1
E = int(sys.argv[2]) if len(sys.argv) > 2 else 3
names = ["build_batch", "fwd_hidden(wide)", "fwd_hidden(64)", "fwd_hidden(latent)", "fwd_enc_last", "bwd_hidden(wide)",
         "bwd_hidden(64)", "bwd_hidden(latent)", "bwd_enc_last", "dec_last(loss+bwd)", "dec_last(storeV)", "dec_last(fromDv)",
         "dis_stage", "kendall", "mi_mse", "TOTAL"]
if os.environ.get("RAAE_NODROP"):
    EXAMPLE = dict(EXAMPLE, dropout_rate=0.0, dis_dropout_rate=0.0)
eng = Engine(EXAMPLE, n_trials=T, device="cuda:0", max_rows=1056)
for t in range(T):
    init_trial_state(eng, t, EXAMPLE, seed=t)
eng.bind_dataset(*synthetic_arrays())
eng.train_epochs(0, 2)
torch.cuda.synchronize()
VAL = bool(os.environ.get("RAAE_PROFILE_VAL"))      # build with RAAE_NVCC_EXTRA="-DRAAE_PROFILE_VAL=1": second block = validation kernel
prof = torch.zeros(2 * T if VAL else T, 32, dtype=torch.int64, device="cuda:0")
L.check(eng.lib.raae_set_profile_buffer(eng.handle, prof.data_ptr()))
eng.train_epochs(2, E)
torch.cuda.synchronize()
p_all = prof.cpu().numpy().astype(np.float64)
p = p_all[:T] / (E * 5)          # cycles per train step
if VAL:
    pv = p_all[T:] / E           # cycles per validation block
    vnames = list(names)
    vnames[0] = "latent_metrics"
    print(f"validation kernel: {pv[:, 15].mean()/1e6:.3f} Mcycles per epoch per trial")
    for i, n in enumerate(vnames[:15]):
        if pv[:, i].mean() > 0:
            print(f"  val {n:22s} {pv[:, i].mean()/1e3:10.1f} kcyc  {100*pv[:, i].mean()/pv[:, 15].mean():5.1f}%")
    print(f"  val {'(unaccounted)':22s} {(pv[:, 15].mean() - pv[:, :15].sum(1).mean())/1e3:10.1f} kcyc")
tot = p[:, 15].mean()
print(f"trials {T}: {tot/1e6:.3f} Mcycles per step per trial (mean over trials), calls per step in brackets")
calls = [1, 6, 34, 4, 6, 4, 21, 3, 4, 2, 1, 1, 1, 1, 1, 1]
for i, n in enumerate(names):
    v = p[:, i].mean()
    print(f"  {n:22s} {v/1e3:10.1f} kcyc  {100*v/tot:5.1f}%   [{calls[i]:2d} calls, {v/1e3/calls[i]:8.1f} kcyc each]")
for i, n in enumerate(["dec_last: act tile + fwd GEMM", "dec_last: recon loss pass", "dec_last: smooth loss pass", "dec_last: db column sums",
                       "dec_last: dW (mma_tn8)", "dec_last: g GEMM + epilogue"]):
    v = p[:, 16 + i].mean()
    print(f"    probe {n:30s} {v/1e3:10.1f} kcyc  per tile {v/(8 if 'loss pass' in n else 16):8.0f} cyc")
for i, n in [(23, "dis: input rows + noise"), (24, "dis: layer 0"), (25, "dis: layer 1 GEMM + H2"), (26, "dis: logits + BCE"), (16, "dis: du2 pass"), (17, "dis: dW1 + dh1 GEMMs + epilogue"), (22, "dis: dW0 + dz + loop end"),
             (28, "bwd64tc: wait prev dW + du pass"), (29, "bwd64tc: act staging"),
             (30, "bwd64tc: g_prev MMA"), (31, "bwd64tc: g epilogue + trailing sync"), (27, "bwd64tc: tail reductions + adam (per call x8)")]:
    v = p[:, i].mean()
    calls_tiles = (16 if (22 <= i < 27 or i in (16, 17)) else 21 * 8)
    print(f"    probe {n:50s} {v/1e3:10.1f} kcyc  per item/tile {v/calls_tiles:8.0f} cyc")
print(f"  {'(unaccounted)':22s} {(tot - p[:, :15].sum(1).mean())/1e3:10.1f} kcyc")
if os.environ.get("RAAE_RAW_PROBES"):
    print("raw probe slots 16..31 (kcyc per step):", " ".join(f"{i}:{p[:, i].mean()/1e3:.1f}" for i in range(16, 32)))
