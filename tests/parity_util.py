"""Teacher-forced comparison of the CUDA step (through the C ABI) with the numpy oracle.

Every comparison starts both sides from the SAME float32-representable state and feeds the SAME
explicit random draws; the oracle runs in float64.  Tolerances (see DESIGN.md "Parity"):
  * losses: |cuda - oracle| <= 1e-4 * max(1, |oracle|)   (Kendall: 5e-4, its pair classification
    flips for |s_i - s_j| ~ 1e-7)
  * gradients: rel-L2 per network per phase <= 2e-3 — the reference's OWN float32 path differs from
    its float64 run by up to 3.5e-3 on these networks (BatchNorm divides
    by sqrt(var + 1e-5) with var << eps for dead PReLU channels and amplifies rounding noise);
    measured on B200: 7e-7..4e-4 (profiles/parity_r01.md); every value is written to gpurun_out/parity_*.json
  * AdamW: parameters / moments after the update, given the kernel's own gradient, to 2e-6 rel.
"""
import json
import os

import numpy as np

from oracle import aae_oracle as O

PHASES = O.PHASES
LOSS_TOL = {"adversarial": 1e-4, "correlation": 5e-4, "reconstruction": 1e-4, "mutual_info": 1e-4, "smoothness": 1e-4}
GRAD_TOL = 2e-3
REPORT = []


def f32_state(state):
    """Rounds an oracle state to float32-representable float64."""
    return O.cast_state(O.cast_state(state, np.float32), np.float64)


def f32_opt(opt):
    out = {}
    for ph, o in opt.items():
        out[ph] = dict(t=o["t"], lr=float(np.float32(o["lr"])),
                       m={n: {k: [np.float32(x).astype(np.float64) for x in v] for k, v in d.items()} for n, d in o["m"].items()},
                       v={n: {k: [np.float32(x).astype(np.float64) for x in v] for k, v in d.items()} for n, d in o["v"].items()})
    return out


def f32_rnd(rnd):
    out = {}
    for k, v in rnd.items():
        if v is None:
            out[k] = None
        elif isinstance(v, list):
            out[k] = [np.float32(m).astype(np.float64) for m in v]
        else:
            out[k] = np.float32(v).astype(np.float64)
    return out


def net_vec(g, skip_last_bias=False):
    parts = []
    for k in ("W", "b", "a"):
        for i, x in enumerate(g[k]):
            if skip_last_bias and k == "b" and i == len(g[k]) - 1:
                continue
            parts.append(np.asarray(x, dtype=np.float64).ravel())
    return np.concatenate(parts)


def rel_l2(a, b):
    den = np.linalg.norm(b)
    return float(np.linalg.norm(a - b) / den) if den > 0 else float(np.linalg.norm(a - b))


def f32_yardstick(cfg, state, x, aux, rnd, epoch, phase, ref):
    """Error of the SAME oracle code run in float32 against its float64 run: what the reference's own precision
    (PyTorch float32) costs on this case.  Ill-conditioned cases (deep stacks, tiny batches: BatchNorm divides by
    sqrt(var + 1e-5)) are judged against this yardstick, as BASELINE.md §4 prescribes."""
    c32 = lambda v: None if v is None else ([np.float32(m) for m in v] if isinstance(v, list) else np.float32(v))
    st32 = O.cast_state(state, np.float32)
    r32 = O.train_step(st32, None, cfg, np.float32(x), np.float32(aux), {k: c32(v) for k, v in rnd.items()}, epoch,
                       apply_updates=False, phases=(phase,))
    out = {"loss": abs(float(r32["losses"][phase]) - float(ref["losses"][phase]))}
    for net in cfg.optimizer_hparams()[phase]["nets"]:
        out[net] = rel_l2(net_vec(r32["grads"][phase][net], net == "E"), net_vec(ref["grads"][phase][net], net == "E"))
    return out


def compare_phase(engine, trial, cfg, state, opt, x, aux, rnd, epoch, phase, tag, yardstick=False):
    """Runs phase `phase` teacher-forced on both sides (no optimizer update); returns a report dict."""
    p = PHASES.index(phase)
    engine.set_state(trial, state, opt)
    got = engine.step_debug(trial, x, aux, rnd, epoch=epoch, phase_mask=1 << p, apply_updates=False)
    st = O.clone_state(state)
    ref = O.train_step(st, None, cfg, x, aux, rnd, epoch, apply_updates=False, phases=(phase,))
    rep = {"tag": tag, "phase": phase, "rows": int(x.shape[0]), "loss_cuda": float(got["losses"][phase]),
           "loss_oracle": float(ref["losses"].get(phase, 0.0)), "grad_rel_l2": {}, "grad_rel_l2_tensor": {}}
    if phase not in ref["losses"]:
        return rep, got, ref
    hp = cfg.optimizer_hparams()[phase]
    for net in hp["nets"]:
        # E.b[last] feeds BatchNorm directly: its gradient is mathematically zero (rounding noise on both sides)
        a = net_vec(got["grads"][phase][net], skip_last_bias=(net == "E"))
        b = net_vec(ref["grads"][phase][net], skip_last_bias=(net == "E"))
        rep["grad_rel_l2"][net] = rel_l2(a, b)
        for k in ("W", "b", "a"):
            for i, xr in enumerate(ref["grads"][phase][net][k]):
                if net == "E" and k == "b" and i == len(ref["grads"][phase][net][k]) - 1:
                    continue
                rep["grad_rel_l2_tensor"][f"{net}.{k}{i}"] = rel_l2(np.asarray(got["grads"][phase][net][k][i], np.float64), xr)
    # BN buffers after the forwards of this phase
    got_state, _ = engine.get_state(trial)
    bn = 0.0
    for net in ("E", "D"):
        for k in ("rm", "rv"):
            for u, w in zip(got_state[net][k], st[net][k]):
                bn = max(bn, float(np.abs(u - w).max() / max(1.0, np.abs(w).max())))
        rep[f"nbt_{net}"] = [int(got_state[net]["nbt"]), int(st[net]["nbt"])]
    rep["bn_buffer_err"] = bn
    if yardstick:
        rep["f32_yardstick"] = f32_yardstick(cfg, state, x, aux, rnd, epoch, phase, ref)
    REPORT.append(rep)
    return rep, got, ref


def check_phase_report(rep):
    """Bands: losses LOSS_TOL, gradients GRAD_TOL; when the report carries a float32 yardstick (ill-conditioned cases) the
    band widens to 3x what the reference's own float32 arithmetic loses on that case."""
    ph = rep["phase"]
    lo = rep["loss_oracle"]
    ys = rep.get("f32_yardstick", {})
    assert abs(rep["loss_cuda"] - lo) <= max(LOSS_TOL[ph] * max(1.0, abs(lo)), 3.0 * ys.get("loss", 0.0)), rep
    for net, e in rep["grad_rel_l2"].items():
        assert e <= max(GRAD_TOL, 3.0 * ys.get(net, 0.0)), (rep["tag"], ph, net, e, ys, rep["grad_rel_l2_tensor"])
    assert rep["bn_buffer_err"] <= 1e-5, rep
    for net in ("E", "D"):
        assert rep[f"nbt_{net}"][0] == rep[f"nbt_{net}"][1], rep


def check_adam(engine, trial, cfg, state, opt, x, aux, rnd, epoch, phase):
    """Applies phase `phase` on the device, then replays AdamW in float64 on the kernel's own gradient."""
    p = PHASES.index(phase)
    engine.set_state(trial, state, opt)
    got = engine.step_debug(trial, x, aux, rnd, epoch=epoch, phase_mask=1 << p, apply_updates=True)
    new_state, new_opt = engine.get_state(trial)
    hp = cfg.optimizer_hparams()[phase]
    t = opt[phase]["t"] + 1
    assert new_opt[phase]["t"] == t
    worst = 0.0
    for net in hp["nets"]:
        for k in ("W", "b", "a"):
            for i in range(len(state[net][k])):
                g = np.asarray(got["grads"][phase][net][k][i], dtype=np.float64)
                pe, me, ve = O.adamw_update(state[net][k][i], g, opt[phase]["m"][net][k][i], opt[phase]["v"][net][k][i],
                                            t, opt[phase]["lr"], hp["betas"][0], hp["betas"][1], hp["wd"])
                for name, a, b in (("p", new_state[net][k][i], pe), ("m", new_opt[phase]["m"][net][k][i], me),
                                   ("v", new_opt[phase]["v"][net][k][i], ve)):
                    err = float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-30))
                    worst = max(worst, err)
                    assert err <= 2e-6, (phase, net, k, i, name, err)
    # nets outside the optimizer must be untouched
    for net in ("E", "D", "S"):
        if net in hp["nets"]:
            continue
        for k in ("W", "b", "a"):
            for a, b in zip(new_state[net][k], state[net][k]):
                assert np.array_equal(a, np.float32(b)), (phase, net, k)
    REPORT.append({"tag": "adam", "phase": phase, "worst_rel": worst})
    return worst


def dump_report(name="parity_report.json"):
    out_dir = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "gpurun_out")
    os.makedirs(out_dir, exist_ok=True)
    with open(os.path.join(out_dir, name), "w") as f:
        json.dump(REPORT, f, indent=1, default=float)
    del REPORT[:]
