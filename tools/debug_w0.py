import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__ as graft
graft.build()
from oracle import aae_oracle as O
from tests import parity_util as PU
from tests.test_parity_gpu import EXAMPLE
from rankaae_b200.engine import Engine
rows = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
phase = sys.argv[2] if len(sys.argv) > 2 else "adversarial"
cfg = O.Config.from_dict(EXAMPLE)
rng = np.random.default_rng(100 + rows)
state = PU.f32_state(O.init_state(cfg, rng))
for net in ("E", "D", "S"):
    state[net]["a"] = [np.float32(a + rng.uniform(-0.005, 0.2, a.shape)).astype(np.float64) for a in state[net]["a"]]
spec, aux = O.synthetic_dataset(rows, cfg, seed=rows, dtype=np.float32)
x = np.float32(spec + cfg.spec_noise * rng.standard_normal(spec.shape)).astype(np.float64)
rnd = PU.f32_rnd(O.draw_step_randoms(cfg, rows, rng))
p = O.PHASES.index(phase)
res = {}
for tcm in (3, 7, 7):
    eng = Engine(dict(EXAMPLE, tensor_cores=tcm), n_trials=1, device="cuda:0", max_rows=1056)
    eng.set_state(0, state, None)
    got = eng.step_debug(0, x, aux.astype(np.float64), rnd, epoch=700, phase_mask=1 << p, apply_updates=False)
    torch.cuda.synchronize()
    sc = eng.scratch[0].cpu().numpy()
    off = 1056 * 256 + 1056 * 8
    u0 = sc[off:off + rows * 64].reshape(rows, 64).copy()
    # rank panel offset: xn, aux, uE[4], zE, dz, uD[4], v, g0, g1, zs, rank
    R = 1056
    roff = R*256 + R*8 + 4*R*64 + R*8 + R*8 + 4*R*64 + R*256 + 2*R*64 + R*8
    st = sc[roff:roff+128].copy()
    res.setdefault(("st", tcm), []).append(st)
    gW0 = np.asarray(got["grads"][phase]["E"]["W"][0], np.float64)
    res.setdefault(tcm, []).append((u0, gW0))
    eng.close()
ref = O.train_step(O.clone_state(state), None, cfg, x, aux.astype(np.float64), rnd, 700, apply_updates=False, phases=(phase,))
gref = np.asarray(ref["grads"][phase]["E"]["W"][0], np.float64)
u3 = res[3][0][0]
s3 = res[("st", 3)][0]; s7 = res[("st", 7)][0]
a0 = np.float32(state["E"]["a"][0])
pu = np.where(u3 > 0, u3, a0[None, :] * u3).astype(np.float64)
print("numpy mean/var from u0 vs tc3:", np.abs(pu.mean(0) - s3[:64]).max(), np.abs(pu.var(0) - s3[64:]).max())
u7 = res[7][0][0]
pu7 = np.where(u7 > 0, u7, a0[None, :] * u7).astype(np.float64)
print("numpy mean/var from OWN u0 vs tc7:", np.abs(pu7.mean(0) - s7[:64]).max(), np.abs(pu7.var(0) - s7[64:]).max())
print("rel var err tc3", np.abs(pu.var(0) - s3[64:]).max()/1, (np.abs(pu.var(0) - s3[64:])/(pu.var(0)+1e-5)).max(), "tc7", (np.abs(pu7.var(0) - s7[64:])/(pu7.var(0)+1e-5)).max())
print("u7-u3: per-channel mean shift max", np.abs((u7-u3).mean(0)).max(), "residual after removing channel shift", np.abs((u7-u3) - (u7-u3).mean(0)).max())
print("var (numpy) min/median:", pu.var(0).min(), np.median(pu.var(0)))
i = np.argmax(np.abs(pu.var(0) - s7[64:])); print("worst channel", i, "var np", pu.var(0)[i], "tc3", s3[64+i], "tc7", s7[64+i], "mean np", pu.mean(0)[i], s3[i], s7[i])
u3, g3 = res[3][0]
print("tc3 vs oracle", np.linalg.norm(g3-gref)/np.linalg.norm(gref))
for k, (u7, g7) in enumerate(res[7]):
    du = np.abs(u7 - u3)
    bad = np.where(du.max(1) > 1e-4)[0]
    print(f"run {k}: u0 max diff {du.max():.3e}; rows with diff > 1e-4: {len(bad)}", bad[:40])
    if len(bad):
        r = bad[0]
        print(" row", r, "tc3", u3[r, :6], "tc7", u7[r, :6])
    print("  tc7 vs oracle", np.linalg.norm(g7-gref)/np.linalg.norm(gref))
    dg = np.abs(g7 - g3)
    print(f"  gW0 rel diff {np.linalg.norm(g7-g3)/np.linalg.norm(g3):.3e}; worst n rows", np.argsort(-dg.max(1))[:8], "worst k cols", np.argsort(-dg.max(0))[:8])
