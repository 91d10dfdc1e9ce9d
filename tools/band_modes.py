import json, os, sys
import numpy as np, torch
sys.path.insert(0, "/root/repo")
from oracle import aae_oracle as O
from rankaae_b200.engine import Engine
from rankaae_b200.trainer import init_trial_state
ref = json.load(open("/root/repo/tests/golden/e2e_band_ref.json"))
cfg = ref["config"]
spec, aux = O.synthetic_dataset(ref["n_rows"], O.Config.from_dict(cfg), seed=ref["data_seed"], dtype=np.float32)
n_train, n_val = int(ref["n_rows"] * 0.7), int(ref["n_rows"] * 0.15)
for tcm in (23, 7):
    for gseed in (1, 2, 3, 4, 5, 6, 8, 9):
        T = 32
        eng = Engine(dict(cfg, tensor_cores=tcm), n_trials=T, device="cuda:0", max_rows=max(cfg["batch_size"], n_val), seeds=list(range(100, 100 + T)))
        for t in range(T):
            init_trial_state(eng, t, cfg, seed=100 + t)
        eng.bind_dataset(spec[:n_train], aux[:n_train], spec[n_train:n_train + n_val], aux[n_train:n_train + n_val])
        gen = torch.Generator(device="cuda:0"); gen.manual_seed(gseed)
        _, mets = eng.train_epochs(0, cfg["max_epoch"], perm=eng.make_perm(cfg["max_epoch"], generator=gen))
        torch.cuda.synchronize()
        m = mets[-1, :, :5].cpu().numpy().astype(np.float64)
        print(tcm, gseed, "mean", np.round(m.mean(0), 4), "recon median", round(float(np.median(m[:, 1])), 4), "max", round(float(m[:, 1].max()), 4))
        eng.close()
