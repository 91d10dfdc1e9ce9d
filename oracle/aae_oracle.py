"""CPU oracle for the RankAAE adversarial-autoencoder train step (TEST INFRASTRUCTURE ONLY).

This file is a numpy restatement of the reference's training hot path
(`/root/reference/sc/clustering/trainer.py:103-304`) with hand-derived backward
passes.  It exists to CHECK the CUDA product path; it is never shipped and never
measured as the product.  Only `tests/`, `__graft_entry__.smoke()` and the
`cpu_baseline` / `--impl reference` legs of `bench.py` may import it.

Parity status: PINNED.  `oracle/make_golden.py` runs the unmodified reference
modules (imported from /root/reference, fp64, with recorded RNG draws) and writes
`tests/golden/*.npz`; `tests/test_oracle_golden.py` requires this restatement to
reproduce every loss, gradient, post-step parameter, Adam moment, BatchNorm
buffer and validation metric in those files to ~1e-9.

All arithmetic of the reference lives in PyTorch / scipy; the op semantics
restated here are those listed in SURVEY.md Appendix A, each function cites the
reference call site it follows.

State representation (plain dicts of numpy arrays, one per network):
    E: {"W": [W1..WL], "b": [b1..bL], "a": [a1..a(L-1)],
        "rm": [..L], "rv": [..L], "nbt": int}          FCEncoder  model.py:330-378
    D: {"W": [W1..WL], "b": [...], "a": [a1..a(L-1)],
        "rm": [..L-1], "rv": [..L-1], "nbt": int}      FCDecoder  model.py:518-570
    S: {"W": [W1..Wn], "b": [...], "a": [a1..a(n-1)]}   DiscriminatorFC model.py:631-663
W is [out, in] row-major exactly like nn.Linear.
"""
from __future__ import annotations

import copy
import math
from dataclasses import dataclass, field
from typing import Dict, List, Optional

import numpy as np

BN_EPS = 1e-5       # nn.BatchNorm1d default eps            model.py:349
BN_MOMENTUM = 0.1   # nn.BatchNorm1d default momentum
ADAM_EPS = 1e-8     # torch.optim.AdamW default
ADAMW_DEFAULT_WD = 1e-2
GAU_KERNEL_SIZE = 17  # Trainer.gau_kernel_size              trainer.py:36
GAU_SIGMA = 3.0       # smoothness_loss                      functions.py:204
METRIC_WEIGHTS = (1.0, -1.0, -0.01, -1.0, -1.0)  # Trainer.metric_weights trainer.py:35

PHASES = ("adversarial", "correlation", "reconstruction", "mutual_info", "smoothness")


# --------------------------------------------------------------------------------------
# configuration (the fix_config.yaml keys the step consumes)
# --------------------------------------------------------------------------------------
@dataclass
class Config:
    dim_in: int = 256
    dim_out: int = 256
    nstyle: int = 6
    n_aux: int = 5
    n_layers: int = 5
    hidden: int = 64
    dis_layers: int = 3
    batch_size: int = 1024
    max_epoch: int = 2000
    dropout_rate: float = 0.04
    dis_dropout_rate: float = 0.056
    dis_noise: float = 0.56
    spec_noise: float = 0.02
    alpha_flat_step: float = 739
    alpha_limit: float = 0.7172
    kendall_activation: bool = True
    use_flex_spec_target: bool = True
    decoder_activation: str = "Softplus"
    epoch_stop_smooth: int = 1500
    lr_base: float = 1e-3
    lr_ratio_Reconn: float = 10
    lr_ratio_Mutual: float = 1
    lr_ratio_Smooth: float = 1
    lr_ratio_Corr: float = 10
    lr_ratio_dis: float = 1
    weight_decay: float = 1e-2
    dis_beta: float = 1.1
    sch_factor: float = 0.1
    sch_patience: int = 100

    @classmethod
    def from_dict(cls, d):
        names = {f for f in cls.__dataclass_fields__}
        kw = {k: v for k, v in d.items() if k in names}
        if "FC_discriminator_layers" in d:
            kw["dis_layers"] = d["FC_discriminator_layers"]
        return cls(**kw)

    # Trainer.load_optimizers trainer.py:333-397 (GRL branch: 5 of the 7 are stepped)
    def optimizer_hparams(self):
        dflt = (0.9, 0.999)
        return {
            "adversarial": dict(lr=self.lr_ratio_dis * self.lr_base,
                                betas=(self.dis_beta * 0.9, self.dis_beta * 0.009 + 0.99),
                                wd=ADAMW_DEFAULT_WD, nets=("S", "E")),
            "correlation": dict(lr=self.lr_ratio_Corr * self.lr_base, betas=dflt,
                                wd=self.weight_decay, nets=("E",)),
            "reconstruction": dict(lr=self.lr_ratio_Reconn * self.lr_base, betas=dflt,
                                   wd=self.weight_decay, nets=("E", "D")),
            "mutual_info": dict(lr=self.lr_ratio_Mutual * self.lr_base, betas=dflt,
                                wd=ADAMW_DEFAULT_WD, nets=("E", "D")),
            "smoothness": dict(lr=self.lr_ratio_Smooth * self.lr_base, betas=dflt,
                               wd=self.weight_decay, nets=("D",)),
        }


# --------------------------------------------------------------------------------------
# elementary ops
# --------------------------------------------------------------------------------------
def alpha(epoch_percentage, step=800, limit=0.7):
    """functions.py:214-219 (float64 on host)."""
    return (2.0 / (1.0 + np.exp(-1.0e4 / step * epoch_percentage)) - 1) * limit


def prelu(u, a):
    """nn.PReLU(num_parameters=C): y = u>0 ? u : a_c*u   (model.py:348)."""
    return np.where(u > 0, u, a[None, :] * u)


def bn_train(h, rm, rv):
    """nn.BatchNorm1d(affine=False) in train mode (model.py:349).

    Returns xhat, invstd and the UPDATED running buffers (biased var for the
    normalisation, unbiased for the running estimate)."""
    B = h.shape[0]
    mu = h.mean(axis=0)
    var = ((h - mu[None, :]) ** 2).mean(axis=0)
    invstd = 1.0 / np.sqrt(var + BN_EPS)
    xhat = (h - mu[None, :]) * invstd[None, :]
    unbiased = var * (B / (B - 1.0)) if B > 1 else var
    rm_new = (1 - BN_MOMENTUM) * rm + BN_MOMENTUM * mu
    rv_new = (1 - BN_MOMENTUM) * rv + BN_MOMENTUM * unbiased
    return xhat, invstd, rm_new, rv_new


def bn_eval(h, rm, rv):
    return (h - rm[None, :]) / np.sqrt(rv[None, :] + BN_EPS)


def bn_train_backward(g, xhat, invstd):
    """d/dh of train-mode BN: (g - mean(g) - xhat*mean(g*xhat)) * invstd."""
    return (g - g.mean(axis=0)[None, :] - xhat * (g * xhat).mean(axis=0)[None, :]) * invstd[None, :]


def softplus2(v):
    """nn.Softplus(beta=2, threshold=20)  (model.py:535)."""
    bv = 2.0 * v
    return np.where(bv > 20.0, v, np.log1p(np.exp(np.minimum(bv, 20.0))) / 2.0)


def softplus2_grad(v):
    bv = 2.0 * v
    return np.where(bv > 20.0, 1.0, 1.0 / (1.0 + np.exp(-bv)))


def bce_with_logits(x, y):
    """nn.BCEWithLogitsLoss() mean reduction (trainer.py:73). Returns loss, dloss/dx."""
    n = x.shape[0]
    loss = np.mean(np.maximum(x, 0) - x * y + np.log1p(np.exp(-np.abs(x))))
    grad = (1.0 / (1.0 + np.exp(-x)) - y) / n
    return loss, grad


# GaussianSmoothing(channels=1, kernel_size=17, sigma=3.0, dim=1).weight as torch builds it in
# float32 (model.py:186-206; kernel_size and sigma are fixed at trainer.py:36, functions.py:204).
# The exact float32 values are tabulated (first 9 of the symmetric 17) because numpy's and
# torch's float32 exp differ in the last bit; tests/test_oracle_golden.py pins the table.
_GAU17_HALF_F32 = (
    0.0038155282381922007, 0.008779441937804222, 0.018076900392770767, 0.03330628201365471,
    0.05491277202963829, 0.08101504296064377, 0.10695548355579376, 0.1263529658317566,
    0.13357123732566833)


def gaussian_taps(kernel_size=GAU_KERNEL_SIZE, sigma=GAU_SIGMA, dtype=np.float64):
    if kernel_size == 17 and sigma == 3.0:
        half = np.array(_GAU17_HALF_F32, dtype=np.float64)
        return np.concatenate([half, half[-2::-1]]).astype(dtype)
    x = np.arange(kernel_size, dtype=np.float32)
    mean = (kernel_size - 1) / 2
    k = (np.float32(1 / (sigma * math.sqrt(2 * math.pi)))
         * np.exp(-(((x - np.float32(mean)) / np.float32(sigma)) ** 2) / 2).astype(np.float32))
    k = (k / k.sum(dtype=np.float32)).astype(np.float32)
    return k.astype(dtype)


# --------------------------------------------------------------------------------------
# networks: forward caches and backward
# --------------------------------------------------------------------------------------
def _hidden_block_fwd(a_in, W, b, slope, rm, rv, mask, p, train):
    u = a_in @ W.T + b[None, :]
    h = prelu(u, slope)
    if train:
        xhat, invstd, rm_n, rv_n = bn_train(h, rm, rv)
        out = xhat * mask / (1.0 - p) if mask is not None else xhat
    else:
        xhat, invstd, rm_n, rv_n = bn_eval(h, rm, rv), None, rm, rv
        out = xhat
    return out, dict(a_in=a_in, u=u, xhat=xhat, invstd=invstd, mask=mask), rm_n, rv_n


def _hidden_block_bwd(d_out, cache, W, slope, p, need_dx=True):
    g = d_out * cache["mask"] / (1.0 - p) if cache["mask"] is not None else d_out
    dh = bn_train_backward(g, cache["xhat"], cache["invstd"])
    u = cache["u"]
    pos = u > 0
    du = np.where(pos, dh, slope[None, :] * dh)
    dslope = np.where(pos, 0.0, u * dh).sum(axis=0)
    dW = du.T @ cache["a_in"]
    db = du.sum(axis=0)
    dx = du @ W if need_dx else None
    return dW, db, dslope, dx


def encoder_forward(E, x, masks, p, train=True, update=True):
    """FCEncoder.forward (model.py:330-378).  masks: list of (n_layers-1) arrays
    [B, hidden] of {0,1} keep-masks, or None for p == 0 / eval."""
    L = len(E["W"])
    caches = []
    a = x
    new_rm, new_rv = [], []
    for l in range(L - 1):
        m = None if (masks is None or not train) else masks[l]
        a, c, rm_n, rv_n = _hidden_block_fwd(a, E["W"][l], E["b"][l], E["a"][l],
                                             E["rm"][l], E["rv"][l], m, p, train)
        caches.append(c)
        new_rm.append(rm_n)
        new_rv.append(rv_n)
    u = a @ E["W"][L - 1].T + E["b"][L - 1][None, :]
    if train:
        z, invstd, rm_n, rv_n = bn_train(u, E["rm"][L - 1], E["rv"][L - 1])
    else:
        z, invstd, rm_n, rv_n = bn_eval(u, E["rm"][L - 1], E["rv"][L - 1]), None, E["rm"][L - 1], E["rv"][L - 1]
    new_rm.append(rm_n)
    new_rv.append(rv_n)
    caches.append(dict(a_in=a, xhat=z, invstd=invstd))
    if train and update:
        E["rm"], E["rv"] = new_rm, new_rv
        E["nbt"] = E.get("nbt", 0) + 1
    return z, caches


def encoder_backward(E, caches, dz, p, need_dx=False):
    L = len(E["W"])
    g = {"W": [None] * L, "b": [None] * L, "a": [None] * (L - 1)}
    c = caches[L - 1]
    du = bn_train_backward(dz, c["xhat"], c["invstd"])
    g["W"][L - 1] = du.T @ c["a_in"]
    g["b"][L - 1] = du.sum(axis=0)
    d = du @ E["W"][L - 1]
    for l in range(L - 2, -1, -1):
        dW, db, da, d = _hidden_block_bwd(d, caches[l], E["W"][l], E["a"][l], p,
                                          need_dx=(l > 0 or need_dx))
        g["W"][l], g["b"][l], g["a"][l] = dW, db, da
    return g, d


def decoder_forward(D, z, masks, p, train=True, update=True, activation="Softplus"):
    """FCDecoder.forward (model.py:518-570)."""
    L = len(D["W"])
    caches = []
    a = z
    new_rm, new_rv = [], []
    for l in range(L - 1):
        m = None if (masks is None or not train) else masks[l]
        a, c, rm_n, rv_n = _hidden_block_fwd(a, D["W"][l], D["b"][l], D["a"][l],
                                             D["rm"][l], D["rv"][l], m, p, train)
        caches.append(c)
        new_rm.append(rm_n)
        new_rv.append(rv_n)
    v = a @ D["W"][L - 1].T + D["b"][L - 1][None, :]
    if activation == "Softplus":
        y = softplus2(v)
    elif activation == "ReLu":
        y = np.maximum(v, 0)
    else:
        raise ValueError(f'Unknow activation function "{activation}"')
    caches.append(dict(a_in=a, v=v))
    if train and update:
        D["rm"], D["rv"] = new_rm, new_rv
        D["nbt"] = D.get("nbt", 0) + 1
    return y, caches


def decoder_backward(D, caches, dy, p, need_dz=False, activation="Softplus"):
    L = len(D["W"])
    g = {"W": [None] * L, "b": [None] * L, "a": [None] * (L - 1)}
    c = caches[L - 1]
    dv = dy * (softplus2_grad(c["v"]) if activation == "Softplus" else (c["v"] > 0))
    g["W"][L - 1] = dv.T @ c["a_in"]
    g["b"][L - 1] = dv.sum(axis=0)
    d = dv @ D["W"][L - 1]
    for l in range(L - 2, -1, -1):
        dW, db, da, d = _hidden_block_bwd(d, caches[l], D["W"][l], D["a"][l], p,
                                          need_dx=(l > 0 or need_dz))
        g["W"][l], g["b"][l], g["a"][l] = dW, db, da
    return g, d


def discriminator_forward(S, x, noise_eps, masks, p, noise, train=True):
    """DiscriminatorFC.forward (model.py:658-663): input noise (train only), GRL
    (identity forward), [Linear, PReLU, Dropout] x (layers-1), Linear(hidden, 1)."""
    n = len(S["W"])
    a = x + noise * noise_eps if (train and noise_eps is not None) else x
    caches = []
    for l in range(n - 1):
        u = a @ S["W"][l].T + S["b"][l][None, :]
        h = prelu(u, S["a"][l])
        m = None if (masks is None or not train) else masks[l]
        out = h * m / (1.0 - p) if m is not None else h
        caches.append(dict(a_in=a, u=u, mask=m))
        a = out
    logit = a @ S["W"][n - 1].T + S["b"][n - 1][None, :]
    caches.append(dict(a_in=a))
    return logit[:, 0], caches


def discriminator_backward(S, caches, dlogit, p, beta):
    """Backward incl. GradientReversalLayer.backward (model.py:16-22): dx = -beta * g."""
    n = len(S["W"])
    g = {"W": [None] * n, "b": [None] * n, "a": [None] * (n - 1)}
    du = dlogit[:, None]
    g["W"][n - 1] = du.T @ caches[n - 1]["a_in"]
    g["b"][n - 1] = du.sum(axis=0)
    d = du @ S["W"][n - 1]
    for l in range(n - 2, -1, -1):
        c = caches[l]
        gg = d * c["mask"] / (1.0 - p) if c["mask"] is not None else d
        pos = c["u"] > 0
        du = np.where(pos, gg, S["a"][l][None, :] * gg)
        g["a"][l] = np.where(pos, 0.0, c["u"] * gg).sum(axis=0)
        g["W"][l] = du.T @ c["a_in"]
        g["b"][l] = du.sum(axis=0)
        d = du @ S["W"][l]
    dx = d if beta is None else -beta * d
    return g, dx


# --------------------------------------------------------------------------------------
# losses (sc/utils/functions.py)
# --------------------------------------------------------------------------------------
def kendall_constraint(descriptors, styles, activate=False):
    """functions.py:37-79, restated the way the reference executes it: the full
    [B, B, n_aux] pair tensors are materialised and the activation rescales the
    concordant entries per descriptor.  Returns (loss, dloss/dstyles)."""
    n_aux = styles.shape[1]
    aux_target = np.sign(descriptors[:, None, :] - descriptors[None, :, :])
    aux_pred = styles[:, None, :] - styles[None, :, :]
    aux_len = aux_pred.shape[0]
    product = aux_pred * aux_target
    weight = np.ones_like(product)
    if activate:
        full_same_sel = product > 0
        full_opp_sel = product < 0
        for i in range(n_aux):
            n_same = max(int(full_same_sel[:, :, i].sum()), 1)
            n_opp = max(int(full_opp_sel[:, :, i].sum()), 1)
            w = n_opp / max(n_same, n_opp)
            weight[:, :, i] = np.where(full_same_sel[:, :, i], w, 1.0)
        product = product * weight
    norm = (aux_len ** 2 - aux_len) * n_aux
    loss = -product.sum() / norm
    # d/ds_i of sum_ij w_ij t_ij (s_i - s_j): row sum minus column sum; t antisymmetric,
    # w symmetric  ->  2 * row sum.  (verified against autograd.)
    wt = weight * aux_target
    grad = -(wt.sum(axis=1) - wt.sum(axis=0)) / norm
    return loss, grad


def recon_loss(spec_in, spec_out, scale=False):
    """functions.py:81-107.  Returns (loss, dloss/dspec_out)."""
    B, Ls = spec_out.shape
    if not scale:
        diff = spec_out - spec_in
        return np.mean(diff ** 2), 2.0 * diff / (B * Ls)
    m_out = spec_out.mean(axis=1)
    m_in = spec_in.mean(axis=1)
    r = np.abs(m_out) / np.abs(m_in)
    loss = np.mean((r - 1.0) ** 2) * 0.1
    c = np.clip(r, 0.7, 1.3)
    target = spec_in * c[:, None]
    diff = spec_out - target
    loss = loss + np.mean(diff ** 2)
    dr = (0.2 / B) * (r - 1.0) * np.sign(m_out) / (np.abs(m_in) * Ls)
    grad = dr[:, None] + 2.0 * diff / (B * Ls)
    return loss, grad


def gaussian_smooth(y, taps):
    """ReplicationPad1d(8) + depthwise conv1d (functions.py:203-209)."""
    half = (len(taps) - 1) // 2
    yp = np.concatenate([np.repeat(y[:, :1], half, axis=1), y, np.repeat(y[:, -1:], half, axis=1)], axis=1)
    out = np.zeros_like(y)
    n = y.shape[1]
    for t in range(len(taps)):
        out += taps[t] * yp[:, t:t + n]
    return out


def gaussian_smooth_transpose(e, taps):
    """Adjoint of gaussian_smooth: zero-padded full correlation, overhang folded into
    the first / last element."""
    half = (len(taps) - 1) // 2
    n = e.shape[1]
    full = np.zeros((e.shape[0], n + 2 * half), dtype=e.dtype)
    for t in range(len(taps)):
        full[:, t:t + n] += taps[t] * e
    out = full[:, half:half + n].copy()
    out[:, 0] += full[:, :half].sum(axis=1)
    out[:, -1] += full[:, half + n:].sum(axis=1)
    return out


def smoothness_loss(spec_out, taps=None):
    """functions.py:194-212.  Returns (loss, dloss/dspec_out) with the gradient through
    both operands of the MSE."""
    if taps is None:
        taps = gaussian_taps(dtype=spec_out.dtype)
    B, Ls = spec_out.shape
    e = spec_out - gaussian_smooth(spec_out, taps)
    loss = np.mean(e ** 2)
    grad = (2.0 / (B * Ls)) * (e - gaussian_smooth_transpose(e, taps))
    return loss, grad


def mse(a, b):
    d = a - b
    return np.mean(d ** 2), 2.0 * d / d.size


# --------------------------------------------------------------------------------------
# AdamW (torch/optim/adam.py single-tensor path) and ReduceLROnPlateau
# --------------------------------------------------------------------------------------
def adamw_update(p, g, m, v, t, lr, beta1, beta2, wd, eps=ADAM_EPS):
    """One AdamW update of one tensor at step count t (already incremented)."""
    p = p * (1.0 - lr * wd)
    m = m + (g - m) * (1.0 - beta1)
    v = beta2 * v + (1.0 - beta2) * g * g
    bc1 = 1.0 - beta1 ** t
    bc2 = 1.0 - beta2 ** t
    denom = np.sqrt(v) / math.sqrt(bc2) + eps
    p = p - (lr / bc1) * (m / denom)
    return p, m, v


class ReduceLROnPlateau:
    """torch.optim.lr_scheduler.ReduceLROnPlateau(mode="min", threshold_mode="rel",
    threshold=0.01, cooldown=0, min_lr=0, eps=1e-8) as constructed at trainer.py:400-408."""

    def __init__(self, lr, factor=0.1, patience=100, threshold=0.01):
        self.lr = lr
        self.factor = factor
        self.patience = patience
        self.threshold = threshold
        self.best = math.inf
        self.num_bad_epochs = 0

    def step(self, metric):
        current = float(metric)
        if current < self.best * (1.0 - self.threshold):
            self.best = current
            self.num_bad_epochs = 0
        else:
            self.num_bad_epochs += 1
        if self.num_bad_epochs > self.patience:
            new_lr = max(self.lr * self.factor, 0.0)
            if self.lr - new_lr > 1e-8:
                self.lr = new_lr
            self.num_bad_epochs = 0
        return self.lr


# --------------------------------------------------------------------------------------
# per-epoch host metrics (trainer.py:285-297)
# --------------------------------------------------------------------------------------
def _norm_ppf(p):
    """Inverse normal CDF (Wichura AS241 PPND16, ~1e-16 relative), vectorised."""
    p = np.asarray(p, dtype=np.float64)
    q = p - 0.5
    out = np.empty_like(p)
    central = np.abs(q) <= 0.425
    r = 0.180625 - q[central] ** 2
    num = (((((((2509.0809287301226727 * r + 33430.575583588128105) * r + 67265.770927008700853) * r
               + 45921.953931549871457) * r + 13731.693765509461125) * r + 1971.5909503065514427) * r
            + 133.14166789178437745) * r + 3.387132872796366608)
    den = (((((((5226.495278852545925 * r + 28729.085735721942674) * r + 39307.89580009271061) * r
               + 21213.794301586595867) * r + 5394.1960214247511077) * r + 687.1870074920579083) * r
            + 42.313330701600911252) * r + 1.0)
    out[central] = q[central] * num / den
    tail = ~central
    pt = np.where(q[tail] < 0, p[tail], 1.0 - p[tail])
    r = np.sqrt(-np.log(pt))
    res = np.empty_like(r)
    mid = r <= 5.0
    rr = r[mid] - 1.6
    num = (((((((7.7454501427834140764e-4 * rr + 0.0227238449892691845833) * rr + 0.24178072517745061177) * rr
               + 1.27045825245236838258) * rr + 3.64784832476320460504) * rr + 5.7694972214606914055) * rr
            + 4.6303378461565452959) * rr + 1.42343711074968357734)
    den = (((((((1.05075007164441684324e-9 * rr + 5.475938084995344946e-4) * rr + 0.0151986665636164571966) * rr
               + 0.14810397642748007459) * rr + 0.68976733498510000455) * rr + 1.6763848301838038494) * rr
            + 2.05319162663775882187) * rr + 1.0)
    res[mid] = num / den
    far = ~mid
    rr = r[far] - 5.0
    num = (((((((2.01033439929228813265e-7 * rr + 2.71155556874348757815e-5) * rr + 0.0012426609473880784386) * rr
               + 0.026532189526576123093) * rr + 0.29656057182850489123) * rr + 1.7848265399172913358) * rr
            + 5.4637849111641143699) * rr + 6.6579046435011037772)
    den = (((((((2.04426310338993978564e-15 * rr + 1.4215117583164458887e-7) * rr + 1.8463183175100546818e-5) * rr
               + 7.868691311456132591e-4) * rr + 0.0148753612908506148525) * rr + 0.13692988092273580531) * rr
            + 0.59983224205350205389) * rr + 1.0)
    res[far] = num / den
    out[tail] = np.where(q[tail] < 0, -res, res)
    return out


def shapiro_weights(n):
    """Royston AS R94 coefficients for sample size n.  Returns
    the full length-n weight vector w (sum w^2 = 1), ascending-order convention."""
    n2 = n // 2
    i = np.arange(1, n2 + 1, dtype=np.float64)
    m = _norm_ppf((i - 0.375) / (n + 0.25))
    S = 2.0 * np.sum(m ** 2)
    r = 1.0 / math.sqrt(n)
    c1 = [0.0, 0.221157, -0.147981, -2.07119, 4.434685, -2.706056]
    c2 = [0.0, 0.042981, -0.293762, -1.752461, 5.682633, -3.582633]
    poly = lambda c: sum(ck * r ** k for k, ck in enumerate(c))
    a = np.zeros(n2)
    a1 = poly(c1) - m[0] / math.sqrt(S)
    if n > 5:
        a2 = poly(c2) - m[1] / math.sqrt(S)
        fac = math.sqrt((S - 2 * m[0] ** 2 - 2 * m[1] ** 2) / (1 - 2 * a1 ** 2 - 2 * a2 ** 2))
        a[0], a[1] = a1, a2
        a[2:] = -m[2:] / fac
    else:
        fac = math.sqrt((S - 2 * m[0] ** 2) / (1 - 2 * a1 ** 2))
        a[0] = a1
        a[1:] = -m[1:] / fac
    w = np.zeros(n)
    w[:n2] = -a
    w[n - n2:] = a[::-1]
    return w


def shapiro_w(x, w=None):
    """scipy.stats.shapiro(x).statistic (trainer.py:287)."""
    x = np.sort(np.asarray(x, dtype=np.float64))
    if w is None:
        w = shapiro_weights(len(x))
    xc = x - x.mean()
    ssq = np.sum(xc ** 2)
    return float(np.dot(w, x) ** 2 / ssq)


def average_ranks(x):
    """scipy.stats.rankdata(method='average') for a 1-D array."""
    order = np.argsort(x, kind="stable")
    xs = x[order]
    n = len(x)
    ranks = np.empty(n, dtype=np.float64)
    i = 0
    while i < n:
        j = i
        while j + 1 < n and xs[j + 1] == xs[i]:
            j += 1
        ranks[order[i:j + 1]] = 0.5 * (i + j) + 1.0
        i = j + 1
    return ranks


def spearman_max_coupling(z):
    """max_{j1<j2} |spearmanr(z[:, j1], z[:, j2])|  (trainer.py:288-293)."""
    R = np.stack([average_ranks(z[:, k]) for k in range(z.shape[1])], axis=1)
    C = np.corrcoef(R.T)
    iu = np.triu_indices(z.shape[1], 1)
    return float(np.max(np.abs(C[iu])))


# --------------------------------------------------------------------------------------
# the train step (trainer.py:103-204) and the validation block (trainer.py:207-304)
# --------------------------------------------------------------------------------------
def _net_param_list(net):
    """Flat ordered parameter list of one network, nn.Module.parameters() order:
    per block Linear.weight, Linear.bias, PReLU.weight."""
    out = []
    L = len(net["W"])
    for l in range(L):
        out.append(("W", l))
        out.append(("b", l))
        if l < len(net["a"]):
            out.append(("a", l))
    return out


def new_opt_state(state, cfg: Config):
    """Zeroed AdamW state for the 5 stepped optimizers."""
    opt = {}
    for name, hp in cfg.optimizer_hparams().items():
        o = {"t": 0, "lr": hp["lr"], "m": {}, "v": {}}
        for net in hp["nets"]:
            o["m"][net] = {k: [np.zeros_like(x) for x in state[net][k]] for k in ("W", "b", "a")}
            o["v"][net] = {k: [np.zeros_like(x) for x in state[net][k]] for k in ("W", "b", "a")}
        opt[name] = o
    return opt


def _apply_optimizer(state, opt, name, grads, cfg: Config):
    hp = cfg.optimizer_hparams()[name]
    o = opt[name]
    o["t"] += 1
    b1, b2 = hp["betas"]
    for net in hp["nets"]:
        for k in ("W", "b", "a"):
            for i in range(len(state[net][k])):
                p, m, v = adamw_update(state[net][k][i], grads[net][k][i], o["m"][net][k][i], o["v"][net][k][i],
                                       o["t"], o["lr"], b1, b2, hp["wd"])
                state[net][k][i], o["m"][net][k][i], o["v"][net][k][i] = p, m, v


def _scale_grads(g, s):
    return {k: [x * s for x in v] for k, v in g.items()}


def train_step(state, opt, cfg: Config, x_noisy, aux, rnd, epoch, apply_updates=True, phases=PHASES):
    """One batch of the GRL-branch loop body, trainer.py:112-204.

    x_noisy : spec_in AFTER `spec_in += randn_like * spec_noise` (trainer.py:112), [B, dim_in]
    aux     : [B, n_aux]
    rnd     : dict of explicit random draws:
        "E0".."E5": lists of (n_layers-1) keep-masks [B,hidden] for the six encoder forwards
                    (P0, P2, P3, P4-discarded, P4-on-D(z_s), P5)
        "D0".."D3": same for the four decoder forwards (P0, P3, P4, P5)
        "z_real" [batch_size, nstyle], "S_real_eps", "S_real_masks", "S_fake_eps", "S_fake_masks",
        "z_sample" [B, nstyle]
      any mask entry may be None (dropout p == 0).
    Returns dict(losses=..., grads={phase: {net: ...}}).  `state` / `opt` are updated in place
    when apply_updates is True (teacher-forced callers pass copies).
    """
    E, D, S = state["E"], state["D"], state["S"]
    pE, pS = cfg.dropout_rate, cfg.dis_dropout_rate
    act = cfg.decoder_activation
    losses, grads = {}, {}
    alpha_ = alpha(epoch / cfg.max_epoch, cfg.alpha_flat_step, cfg.alpha_limit)
    B = x_noisy.shape[0]

    # P0: trainer.py:113-114 (spec_out unused, but both nets' BN buffers advance)
    styles, cE0 = encoder_forward(E, x_noisy, rnd.get("E0"), pE)
    decoder_forward(D, styles, rnd.get("D0"), pE, activation=act)

    # P1: adversarial_loss functions.py:109-132 through GRL; trainer.py:118-127
    if "adversarial" in phases:
        lr_, cr = discriminator_forward(S, rnd["z_real"], rnd.get("S_real_eps"), rnd.get("S_real_masks"), pS, cfg.dis_noise)
        lf_, cf = discriminator_forward(S, styles, rnd.get("S_fake_eps"), rnd.get("S_fake_masks"), pS, cfg.dis_noise)
        l_real, g_real = bce_with_logits(lr_, np.ones_like(lr_))
        l_fake, g_fake = bce_with_logits(lf_, np.zeros_like(lf_))
        gS_r, _ = discriminator_backward(S, cr, g_real, pS, alpha_)
        gS_f, dstyles = discriminator_backward(S, cf, g_fake, pS, alpha_)
        gS = {k: [a + b for a, b in zip(gS_r[k], gS_f[k])] for k in gS_r}
        gE, _ = encoder_backward(E, cE0, dstyles, pE)
        losses["adversarial"] = l_real + l_fake
        grads["adversarial"] = {"S": gS, "E": gE}
        if apply_updates:
            _apply_optimizer(state, opt, "adversarial", grads["adversarial"], cfg)

    # P2: kendall_constraint trainer.py:153-161
    if "correlation" in phases:
        styles, cE = encoder_forward(E, x_noisy, rnd.get("E1"), pE)
        l, gk = kendall_constraint(aux, styles[:, :cfg.n_aux], activate=cfg.kendall_activation)
        dz = np.zeros_like(styles)
        dz[:, :cfg.n_aux] = gk
        gE, _ = encoder_backward(E, cE, dz, pE)
        losses["correlation"] = l
        grads["correlation"] = {"E": gE}
        if apply_updates:
            _apply_optimizer(state, opt, "correlation", grads["correlation"], cfg)

    # P3: recon_loss trainer.py:164-172 (the noised input is the target)
    if "reconstruction" in phases:
        styles, cE = encoder_forward(E, x_noisy, rnd.get("E2"), pE)
        y, cD = decoder_forward(D, styles, rnd.get("D1"), pE, activation=act)
        l, dy = recon_loss(x_noisy, y, scale=cfg.use_flex_spec_target)
        gD, dz = decoder_backward(D, cD, dy, pE, need_dz=True, activation=act)
        gE, _ = encoder_backward(E, cE, dz, pE)
        losses["reconstruction"] = l
        grads["reconstruction"] = {"E": gE, "D": gD}
        if apply_updates:
            _apply_optimizer(state, opt, "reconstruction", grads["reconstruction"], cfg)

    # P4: mutual_info_loss trainer.py:175-186, functions.py:174-192
    if "mutual_info" in phases:
        encoder_forward(E, x_noisy, rnd.get("E3"), pE)       # trainer.py:176 (result unused)
        zs = rnd["z_sample"]
        y, cD = decoder_forward(D, zs, rnd.get("D2"), pE, activation=act)
        zr, cE = encoder_forward(E, y, rnd.get("E4"), pE)
        l, dzr = mse(zr, zs)
        gE, dy = encoder_backward(E, cE, dzr, pE, need_dx=True)
        gD, _ = decoder_backward(D, cD, dy, pE, activation=act)
        losses["mutual_info"] = l
        grads["mutual_info"] = {"E": gE, "D": gD}
        if apply_updates:
            _apply_optimizer(state, opt, "mutual_info", grads["mutual_info"], cfg)

    # P5: smoothness_loss trainer.py:189-200 (only the decoder is stepped)
    if "smoothness" in phases and epoch < cfg.epoch_stop_smooth:
        styles, cE = encoder_forward(E, x_noisy, rnd.get("E5"), pE)
        y, cD = decoder_forward(D, styles, rnd.get("D3"), pE, activation=act)
        l, dy = smoothness_loss(y, gaussian_taps(dtype=y.dtype))
        gD, dz = decoder_backward(D, cD, dy, pE, need_dz=True, activation=act)
        gE, _ = encoder_backward(E, cE, dz, pE)
        losses["smoothness"] = l
        grads["smoothness"] = {"E": gE, "D": gD}
        if apply_updates:
            _apply_optimizer(state, opt, "smoothness", {"D": gD}, cfg)
    return dict(losses=losses, grads=grads, alpha=alpha_)


def validate(state, cfg: Config, spec_val, aux_val, z_real, z_sample, epoch, avg_mutual_info=0.0):
    """The eval-mode block trainer.py:207-297: five validation losses, the 5-vector of
    metrics and the combined metric that feeds ReduceLROnPlateau."""
    E, D, S = state["E"], state["D"], state["S"]
    act = cfg.decoder_activation
    alpha_ = alpha(epoch / cfg.max_epoch, cfg.alpha_flat_step, cfg.alpha_limit)
    z, _ = encoder_forward(E, spec_val, None, 0.0, train=False)
    y, _ = decoder_forward(D, z, None, 0.0, train=False, activation=act)
    recon, _ = recon_loss(spec_val, y, scale=False)
    aux, _ = kendall_constraint(aux_val, z[:, :cfg.n_aux], activate=cfg.kendall_activation)
    smooth, _ = smoothness_loss(y, gaussian_taps(dtype=y.dtype))
    ys, _ = decoder_forward(D, z_sample, None, 0.0, train=False, activation=act)
    zr, _ = encoder_forward(E, ys, None, 0.0, train=False)
    mi, _ = mse(zr, z_sample)
    lr_, _ = discriminator_forward(S, z_real, None, None, 0.0, 0.0, train=False)
    lf_, _ = discriminator_forward(S, z, None, None, 0.0, 0.0, train=False)
    dis = bce_with_logits(lr_, np.ones_like(lr_))[0] + bce_with_logits(lf_, np.zeros_like(lf_))[0]
    n = z.shape[0]
    w = shapiro_weights(n)
    shap = [shapiro_w(z[:, k], w) for k in range(z.shape[1])]
    coupling = spearman_max_coupling(z)
    metrics = [min(shap), float(recon), float(avg_mutual_info), coupling, float(aux)]
    combined = -float(np.sum(np.array(METRIC_WEIGHTS) * np.array(metrics)))
    return dict(losses=dict(adversarial=dis, correlation=aux, reconstruction=recon,
                            smoothness=smooth, mutual_info=mi),
                metrics=metrics, combined=combined, z=z, shapiro=shap)


# --------------------------------------------------------------------------------------
# helpers for tests / baselines
# --------------------------------------------------------------------------------------
def cast_state(state, dtype):
    out = {}
    for net, d in state.items():
        out[net] = {}
        for k, v in d.items():
            out[net][k] = [np.asarray(x, dtype=dtype).copy() for x in v] if isinstance(v, list) else v
    return out


def clone_state(state):
    return copy.deepcopy(state)


def init_state(cfg: Config, rng: np.random.Generator, dtype=np.float64):
    """nn.Linear default init (kaiming_uniform(a=sqrt(5)) -> U(-1/sqrt(fan_in), 1/sqrt(fan_in))
    for both weight and bias), PReLU slopes 0.01, BN buffers (0, 1).  Used for synthetic
    baselines only; parity tests take their weights from torch modules."""
    H = cfg.hidden

    def lin(o, i):
        k = 1.0 / math.sqrt(i)
        return rng.uniform(-k, k, size=(o, i)).astype(dtype), rng.uniform(-k, k, size=(o,)).astype(dtype)

    def mlp(dims):
        W, b = [], []
        for i, o in zip(dims[:-1], dims[1:]):
            w_, b_ = lin(o, i)
            W.append(w_)
            b.append(b_)
        return W, b

    L = cfg.n_layers
    We, be = mlp([cfg.dim_in] + [H] * (L - 1) + [cfg.nstyle])
    Wd, bd = mlp([cfg.nstyle] + [H] * (L - 1) + [cfg.dim_out])
    Ws, bs = mlp([cfg.nstyle] + [H] * (cfg.dis_layers - 1) + [1])
    slopes = lambda n: [np.full(H, 0.01, dtype=dtype) for _ in range(n)]
    E = dict(W=We, b=be, a=slopes(L - 1),
             rm=[np.zeros(H, dtype) for _ in range(L - 1)] + [np.zeros(cfg.nstyle, dtype)],
             rv=[np.ones(H, dtype) for _ in range(L - 1)] + [np.ones(cfg.nstyle, dtype)], nbt=0)
    D = dict(W=Wd, b=bd, a=slopes(L - 1),
             rm=[np.zeros(H, dtype) for _ in range(L - 1)],
             rv=[np.ones(H, dtype) for _ in range(L - 1)], nbt=0)
    S = dict(W=Ws, b=bs, a=slopes(cfg.dis_layers - 1))
    return dict(E=E, D=D, S=S)


def draw_step_randoms(cfg: Config, B, rng: np.random.Generator, dtype=np.float64):
    """Explicit random draws of one step in the order of SURVEY.md Appendix E.3."""
    H, L = cfg.hidden, cfg.n_layers

    def masks(n, rows, p):
        if p <= 0:
            return None
        return [(rng.random((rows, H)) >= p).astype(dtype) for _ in range(n)]

    r = {}
    for i in range(6):
        r[f"E{i}"] = masks(L - 1, B, cfg.dropout_rate)
    for i in range(4):
        r[f"D{i}"] = masks(L - 1, B, cfg.dropout_rate)
    r["z_real"] = rng.standard_normal((cfg.batch_size, cfg.nstyle)).astype(dtype)
    r["S_real_eps"] = rng.standard_normal((cfg.batch_size, cfg.nstyle)).astype(dtype)
    r["S_real_masks"] = masks(cfg.dis_layers - 1, cfg.batch_size, cfg.dis_dropout_rate)
    r["S_fake_eps"] = rng.standard_normal((B, cfg.nstyle)).astype(dtype)
    r["S_fake_masks"] = masks(cfg.dis_layers - 1, B, cfg.dis_dropout_rate)
    r["z_sample"] = rng.standard_normal((B, cfg.nstyle)).astype(dtype)
    return r


def synthetic_dataset(n, cfg: Config, seed=0, dtype=np.float64):
    """Seeded synthetic spectra in the reference CSV's value range:
    descriptors ~ N(0,1) with column 1 an integer in {4,5,6}; every descriptor modulates a
    distinct spectral feature so the Kendall term has something to learn."""
    rng = np.random.default_rng(seed)
    K = cfg.n_aux
    d = rng.standard_normal((n, max(K, 1)))
    if K > 1:
        d[:, 1] = rng.integers(4, 7, size=n)
    grid = np.linspace(0.0, 1.0, cfg.dim_in)[None, :]
    dd = d.copy()
    if K > 1:
        dd[:, 1] = dd[:, 1] - 5.0
    g = lambda k: dd[:, k % max(K, 1)][:, None]
    edge = 1.0 / (1.0 + np.exp(-(grid - 0.18 - 0.02 * g(0)) * 40.0))
    peak1 = (0.9 + 0.25 * g(1)) * np.exp(-0.5 * ((grid - 0.27) / 0.035) ** 2)
    peak2 = 0.35 * np.exp(-0.5 * ((grid - 0.5 - 0.04 * g(2)) / 0.06) ** 2)
    osc = (0.12 + 0.04 * g(3)) * np.sin(2 * np.pi * (grid - 0.3) * (3.0 + 0.3 * g(4))) * (grid > 0.3)
    spec = edge * (1.0 + osc) + peak1 * (grid > 0.1) + peak2
    spec = spec + 0.02 * rng.standard_normal(spec.shape)
    spec = np.clip(spec, 0.0, None)
    return spec.astype(dtype), d[:, :K].astype(dtype)
