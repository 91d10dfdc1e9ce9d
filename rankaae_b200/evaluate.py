"""Batched, device-side `evaluate_model` for ranking the trials of an ensemble (the reference evaluates one pickled model at a
time on the CPU: `sc/report/analysis.py:394-450`, called per job by `evaluate_all_models` :104-121 and ranked by
`sort_all_models` :130-231).

`evaluate_trials` runs ONE launch of the validation kernel over every resident trial (`raae_evaluate_trials`: eval-mode
encoder + decoder on the test split, no state change) and turns the latents / per-row reconstruction errors into the
reference's result dictionaries with batched torch operations on the device (rank statistics are library sorts - this is
report tooling, not the train path):

  "Reconstruct Err"        [mean, std] of the per-row mean absolute error, rounded to 4 digits       (analysis.py:424-432)
  "Style-descriptor Corr"  i != 1: {"Spearman", "Linear": {"R2", "slope", "intercept"}}              (analysis.py:328-391)
                           i == 1 (coordination number): {"F1 score", "CN45 Threshold", "CN56 Threshold"} (analysis.py:234-311)
  "Inter-style Corr"       max_i |spearman(style_i, style_last)|                                     (analysis.py:313-325)

(The quadratic fit of analysis.py:377-383 is not used by the ranking and is left to the reference's own tooling.)
"""
import ctypes as C

import numpy as np
import torch

from . import _lib as L


def _avg_ranks(x):
    """scipy.stats.rankdata(method='average') along dim 1 of [T, n] (ties share the mean of their positions)."""
    T, n = x.shape
    order = torch.argsort(x, dim=1, stable=True)
    xs = torch.gather(x, 1, order)
    pos = torch.arange(1, n + 1, device=x.device, dtype=torch.float64).expand(T, n)
    new = torch.ones_like(xs, dtype=torch.bool)
    new[:, 1:] = xs[:, 1:] != xs[:, :-1]
    gid = torch.cumsum(new.to(torch.int64), dim=1) - 1                    # tie-group index per sorted position
    gsum = torch.zeros(T, n, device=x.device, dtype=torch.float64).scatter_add_(1, gid, pos)
    gcnt = torch.zeros(T, n, device=x.device, dtype=torch.float64).scatter_add_(1, gid, torch.ones_like(pos))
    mean_rank = torch.gather(gsum / gcnt.clamp(min=1), 1, gid)
    ranks = torch.empty_like(mean_rank)
    ranks.scatter_(1, order, mean_rank)
    return ranks


def _pearson(a, b):
    a = a - a.mean(dim=1, keepdim=True)
    b = b - b.mean(dim=1, keepdim=True)
    return (a * b).sum(1) / torch.sqrt((a * a).sum(1) * (b * b).sum(1))


def _cn_f1(style, cn):
    """analysis.get_confusion_matrix without the plots, batched over trials: style [T, n], cn [n] (4 / 5 / 6)."""
    T, n = style.shape
    dev = style.device
    grid = torch.linspace(-3.5, 3.5, 700, device=dev, dtype=torch.float64)
    cls = (cn - 4).to(torch.int64)
    if len(torch.unique(cls)) > 3:
        return None
    inf = torch.tensor(float("inf"), device=dev, dtype=torch.float64)
    s_all = torch.sort(style, dim=1).values
    g = grid.expand(T, -1).contiguous()
    # CN4: predicted = style < th, target = class < 1
    t4 = cls < 1
    s4 = torch.sort(torch.where(t4, style, inf), dim=1).values
    tp4 = torch.searchsorted(s4, g).double()
    pp4 = torch.searchsorted(s_all, g).double()
    d4 = pp4 + float(t4.sum())
    f4 = torch.where(d4 > 0, 2 * tp4 / d4.clamp(min=1), torch.zeros_like(d4))
    # CN6: predicted = style > th, target = class > 1
    t6 = cls > 1
    s6 = torch.sort(torch.where(t6, style, inf), dim=1).values
    n6 = float(t6.sum())
    tp6 = n6 - torch.searchsorted(s6, g, right=True).clamp(max=int(n6)).double()
    pp6 = n - torch.searchsorted(s_all, g, right=True).double()
    d6 = pp6 + n6
    f6 = torch.where(d6 > 0, 2 * tp6 / d6.clamp(min=1), torch.zeros_like(d6))
    th45 = grid[torch.argmax(f4, dim=1)]
    th56 = grid[torch.argmax(f6, dim=1)]
    pred = (style > th45[:, None]).to(torch.int64) + (style > th56[:, None]).to(torch.int64)
    f1 = torch.zeros(T, device=dev, dtype=torch.float64)
    for c in range(3):
        tc = (cls == c)[None, :]
        pc = pred == c
        tp = (pc & tc).sum(1).double()
        den = pc.sum(1).double() + float(tc.sum())
        f1c = torch.where(den > 0, 2 * tp / den.clamp(min=1), torch.zeros_like(den))
        f1 += f1c * float(tc.sum()) / n                                  # average='weighted'
    return f1, th45, th56


def evaluate_trials(engine, spec_test, aux_test, epoch=0):
    """Returns one result dict per resident trial of `engine`, in the reference's `evaluate_model` format.  The engine's
    bound dataset is restored afterwards; the trials' state is not touched."""
    dev = engine.device
    old = getattr(engine, "_data", None)
    old_rows = engine.n_train
    st = old[0] if old is not None else torch.zeros(2, engine.ccfg.dim_in, device=dev)
    at = old[1] if old is not None else torch.zeros(2, engine.ccfg.n_aux, device=dev)
    xs = torch.as_tensor(spec_test, dtype=torch.float32).to(dev).contiguous()
    au = torch.as_tensor(aux_test, dtype=torch.float32).to(dev).contiguous()
    n, T, ns = xs.shape[0], engine.n_trials, engine.ccfg.nstyle
    if n > engine.ccfg.max_rows:
        raise ValueError(f"the test split ({n} rows) exceeds the engine's max_rows ({engine.ccfg.max_rows})")
    engine.bind_dataset(st, at, xs, au, rows_per_trial=old_rows if old is not None else None)
    z = torch.zeros(T, n, ns, dtype=torch.float32, device=dev)
    mae = torch.zeros(T, n, dtype=torch.float32, device=dev)
    losses = torch.zeros(T, L.NUM_PHASES, dtype=torch.float32, device=dev)
    metrics = torch.zeros(T, 6, dtype=torch.float32, device=dev)
    L.check(engine.lib.raae_evaluate_trials(engine.handle, int(epoch), z.data_ptr(), mae.data_ptr(), losses.data_ptr(),
                                            metrics.data_ptr(), engine.stream))
    torch.cuda.synchronize(dev)
    if old is not None:
        engine.bind_dataset(*old, rows_per_trial=old_rows)
    zd, desc = z.double(), au.double()
    n_aux = desc.shape[1]
    mae_d = mae.double()
    rec_mean, rec_std = mae_d.mean(1), mae_d.std(1, unbiased=False)
    zr = torch.stack([_avg_ranks(zd[:, :, k]) for k in range(ns)], dim=2)             # [T, n, ns]
    dr = _avg_ranks(desc.t().contiguous())                                            # [n_aux, n]
    out = [{"Style-descriptor Corr": {}, "Reconstruct Err": None, "Inter-style Corr": None} for _ in range(T)]
    for i in range(min(n_aux, ns)):
        if i == 1:
            f = _cn_f1(zd[:, :, i], desc[:, i])
            for t in range(T):
                out[t]["Style-descriptor Corr"][i] = None if f is None else {
                    "F1 score": round(float(f[0][t]), 4), "CN45 Threshold": round(float(f[1][t]), 4),
                    "CN56 Threshold": round(float(f[2][t]), 4)}
            continue
        sp = _pearson(zr[:, :, i], dr[i].expand(T, -1))
        x, y = zd[:, :, i], desc[:, i].expand(T, -1)
        r = _pearson(x, y)
        xm, ym = x.mean(1), y.mean(1)
        slope = ((x - xm[:, None]) * (y - ym[:, None])).sum(1) / ((x - xm[:, None]) ** 2).sum(1)
        icpt = ym - slope * xm
        for t in range(T):
            out[t]["Style-descriptor Corr"][i] = {
                "Spearman": round(float(sp[t]), 4),
                "Linear": {"slope": round(float(slope[t]), 4), "intercept": round(float(icpt[t]), 4), "R2": round(float(r[t] ** 2), 4)}}
    inter = torch.stack([_pearson(zr[:, :, i], zr[:, :, ns - 1]).abs() for i in range(ns - 1)], dim=1).max(dim=1).values
    lo, me = losses.cpu().numpy(), metrics.cpu().numpy()
    for t in range(T):
        out[t]["Reconstruct Err"] = [round(float(rec_mean[t]), 4), round(float(rec_std[t]), 4)]
        out[t]["Inter-style Corr"] = round(float(inter[t]), 4)
        out[t]["Validation losses"] = dict(zip(L.PHASES, (float(v) for v in lo[t])))
        out[t]["Trainer metrics"] = [float(v) for v in me[t, :5]]
    return out, z
