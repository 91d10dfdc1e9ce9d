import json,sys,os
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from oracle import aae_oracle as O
from rankaae_b200.engine import Engine
from rankaae_b200.trainer import init_trial_state
ref=json.load(open("tests/golden/e2e_band_ref.json")); cfg=ref["config"]
spec,aux=O.synthetic_dataset(ref["n_rows"],O.Config.from_dict(cfg),seed=ref["data_seed"],dtype=np.float32)
n_train,n_val=int(ref["n_rows"]*0.7),int(ref["n_rows"]*0.15)
T=16
eng=Engine(cfg,n_trials=T,device="cuda:0",max_rows=max(cfg["batch_size"],n_val),seeds=list(range(100,100+T)))
for t in range(T): init_trial_state(eng,t,cfg,seed=100+t)
eng.bind_dataset(spec[:n_train],aux[:n_train],spec[n_train:n_train+n_val],aux[n_train:n_train+n_val])
_,m=eng.train_epochs(0,cfg["max_epoch"]); torch.cuda.synchronize()
m=m.cpu().numpy()
np.set_printoptions(precision=4,suppress=True,linewidth=200)
print("fused recon per trial at epochs 10,30,59:"); print(m[10,:,1]); print(m[30,:,1]); print(m[59,:,1])
print("ref recon per seed:", [round(r["metrics"][1],4) for r in ref["runs"]])
print("fused W:", m[59,:,0]); print("ref W:", [round(r["metrics"][0],4) for r in ref["runs"]])
