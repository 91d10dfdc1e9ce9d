"""Debug: row-permutation case (fresh init, 804 rows) against the float64 oracle and the float32 yardstick, per cluster size."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__ as graft
graft.build()
from oracle import aae_oracle as O
from tests import parity_util as PU
from tests.test_parity_gpu import EXAMPLE
from rankaae_b200.engine import Engine

cfg = O.Config.from_dict(EXAMPLE)
rows = 804
rng = np.random.default_rng(7)
state = PU.f32_state(O.init_state(cfg, rng))
spec, aux = O.synthetic_dataset(rows, cfg, seed=3, dtype=np.float32)
rnd = PU.f32_rnd(O.draw_step_randoms(cfg, rows, rng))
perm = rng.permutation(rows)
def permuted(r):
    out = {}
    for k, v in r.items():
        if k in ("z_real", "S_real_eps", "S_real_masks") or v is None: out[k] = v
        elif isinstance(v, list): out[k] = [m[perm] for m in v]
        else: out[k] = v[perm]
    return out
x64, a64 = spec.astype(np.float64), aux.astype(np.float64)
for ph in ("reconstruction", "smoothness", "adversarial"):
    ref = O.train_step(O.clone_state(state), None, cfg, x64, a64, rnd, 10, apply_updates=False, phases=(ph,))
    ys = PU.f32_yardstick(cfg, state, x64, a64, rnd, 10, ph, ref)
    print(ph, "f32 yardstick", {k: f"{v:.1e}" for k, v in ys.items()})
    for C in (1, 2, 4, 8):
        eng = Engine(dict(EXAMPLE, ctas_per_trial=C), n_trials=1, device="cuda:0", max_rows=1056)
        p = O.PHASES.index(ph)
        eng.set_state(0, state)
        a = eng.step_debug(0, spec, aux, rnd, epoch=10, phase_mask=1 << p, apply_updates=False)
        eng.set_state(0, state)
        b = eng.step_debug(0, spec[perm], aux[perm], permuted(rnd), epoch=10, phase_mask=1 << p, apply_updates=False)
        for net in a["grads"][ph]:
            va, vb, vr = (PU.net_vec(g["grads"][ph][net], skip_last_bias=(net == "E")) for g in (a, b, ref))
            print(f"  C={C} {net}: a-vs-oracle {PU.rel_l2(va, vr):.1e}  b-vs-oracle {PU.rel_l2(vb, vr):.1e}  a-vs-b {PU.rel_l2(va, vb):.1e}")
        eng.close()
