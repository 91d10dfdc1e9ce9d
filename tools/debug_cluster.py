"""Debug: teacher-forced phases at several cluster sizes vs ctas_per_trial = 1 (per-tensor differences)."""
import os, sys, zlib
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__ as graft
graft.build()
from oracle import aae_oracle as O
from tests import parity_util as PU
from tests.test_parity_gpu import EXAMPLE, VARIANTS
from rankaae_b200.engine import Engine

name, rows = sys.argv[1], int(sys.argv[2])
phases = sys.argv[3].split(",") if len(sys.argv) > 3 else list(O.PHASES)
over = eval(sys.argv[4]) if len(sys.argv) > 4 else {}
cfgd = dict(EXAMPLE, batch_size=512, **VARIANTS[name])
cfgd.update(over)
print("==", name, rows, phases, over)
cfg = O.Config.from_dict(cfgd)
rng = np.random.default_rng(zlib.crc32(name.encode()) % 1000 + rows)
state = PU.f32_state(O.init_state(cfg, rng))
for net in ("E", "D", "S"):
    state[net]["a"] = [np.float32(a + rng.uniform(-0.005, 0.3, a.shape)).astype(np.float64) for a in state[net]["a"]]
spec, aux = O.synthetic_dataset(max(rows, 8), cfg, seed=rows, dtype=np.float32)
spec, aux = spec[:rows], aux[:rows]
x = np.float32(spec + cfg.spec_noise * rng.standard_normal(spec.shape)).astype(np.float64)
rnd = PU.f32_rnd(O.draw_step_randoms(cfg, rows, rng))
base = {}
for C in (1, 4):
    eng = Engine(dict(cfgd, ctas_per_trial=C), n_trials=1, device="cuda:0", max_rows=512)
    for ph in phases:
        p = O.PHASES.index(ph)
        eng.set_state(0, state, None)
        got = eng.step_debug(0, x, aux, rnd, epoch=700, phase_mask=1 << p, apply_updates=False)
        if C == 1:
            base[ph] = got
            continue
        b = base[ph]
        out = []
        for net in got["grads"][ph]:
            for k in ("W", "b", "a"):
                for i, g in enumerate(got["grads"][ph][net][k]):
                    e = PU.rel_l2(np.asarray(g, np.float64), np.asarray(b["grads"][ph][net][k][i], np.float64))
                    if e > 1e-4:
                        out.append(f"{net}.{k}{i}:{e:.1e}")
        print(f"C={C} {ph}: loss diff {abs(got['losses'][ph]-b['losses'][ph]):.2e}  bad tensors: {' '.join(out) if out else 'none'}")
    eng.close()
