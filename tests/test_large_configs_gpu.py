"""BASELINE.json configs[4] at its real data shape (synthetic 100k x 256 spectra: 70 000 train rows = 68 full batches + one
of 368, 15 000 validation rows), with a handful of trials instead of 1024: what changes with the size is the machinery -
69 batches per launch, a validation block far larger than a batch (multi-chunk Kendall pairs, a 16 384-key bitonic sort
for Shapiro-Wilk / Spearman, 118 row tiles of scratch) and per-trial hyper-parameter rows - so that is what is exercised
here; the arithmetic is covered by the teacher-forced parity tests.  Size-independent properties checked: everything
finite and in range, the validation metrics of one trial reproduce bit-for-bit, and the validation block agrees with the
float64 oracle on a 15 000-row set (no pair tensor on either side)."""
import numpy as np
import pytest

from oracle import aae_oracle as O

pytestmark = pytest.mark.gpu


def _cfg():
    from tests.test_parity_gpu import EXAMPLE
    return dict(EXAMPLE, max_epoch=2)


def test_config5_shape_two_epochs_sweep():
    import torch
    from rankaae_b200.engine import Engine
    from rankaae_b200.trainer import init_trial_state
    cfg = _cfg()
    n, n_train, n_val = 100_000, 70_000, 15_000
    spec, aux = O.synthetic_dataset(n, O.Config.from_dict(cfg), seed=5, dtype=np.float32)
    T = 4
    per_trial = [dict(cfg, lr_base=cfg["lr_base"] * (0.5 + 0.5 * t), dropout_rate=0.02 * (t + 1), dis_noise=0.3 + 0.1 * t,
                      weight_decay=0.005 * (t + 1)) for t in range(T)]
    eng = Engine(cfg, n_trials=T, device="cuda:0", max_rows=n_val, seeds=list(range(T)), per_trial_cfg=per_trial)
    for t in range(T):
        init_trial_state(eng, t, per_trial[t], seed=t)
    eng.bind_dataset(spec[:n_train], aux[:n_train], spec[n_train:n_train + n_val], aux[n_train:n_train + n_val])
    losses, metrics = eng.train_epochs(0, 2)
    torch.cuda.synchronize()
    lo, me = losses.cpu().numpy(), metrics.cpu().numpy()
    assert np.isfinite(lo).all() and np.isfinite(me).all(), (lo, me)
    assert ((me[..., 0] > 0.0) & (me[..., 0] <= 1.0)).all(), me[..., 0]            # min Shapiro-Wilk W
    assert ((me[..., 3] >= 0.0) & (me[..., 3] <= 1.0)).all(), me[..., 3]            # max |Spearman|
    # the trials learn: after two epochs (138 steps) the validation reconstruction MSE is far below that of the initial
    # network (~0.3 on these spectra); epoch-to-epoch monotonicity is NOT required (adversarial training is noisy and the
    # sweep includes a 2x learning rate)
    assert (me[1, :, 1] < 0.15).all(), me[..., 1]
    assert len({float(x) for x in me[1, :, 1]}) == T                                # the trials really differ (hyper-parameters)
    # the validation block is a pure function of the state: two calls agree bit-for-bit
    a = eng.validate(0, epoch=1)
    b = eng.validate(0, epoch=1)
    assert np.array_equal(a["metrics"], b["metrics"]) and np.array_equal(a["z"], b["z"])
    eng.close()


def _kendall_chunked(descriptors, styles, activate=False, rows=500):
    """oracle.kendall_constraint (functions.py:37-79) evaluated in row blocks: same pair arithmetic, no [n, n, K] tensor."""
    n, K = styles.shape
    same = np.zeros(K); opp = np.zeros(K); sp = np.zeros(K); sn = np.zeros(K)
    for i0 in range(0, n, rows):
        t = np.sign(descriptors[i0:i0 + rows, None, :] - descriptors[None, :, :])
        p = (styles[i0:i0 + rows, None, :] - styles[None, :, :]) * t
        same += (p > 0).sum((0, 1)); opp += (p < 0).sum((0, 1))
        sp += np.where(p > 0, p, 0.0).sum((0, 1)); sn += np.where(p < 0, p, 0.0).sum((0, 1))
    w = np.ones(K)
    if activate:
        ns, no = np.maximum(same, 1), np.maximum(opp, 1)
        w = no / np.maximum(ns, no)
    return -float((w * sp + sn).sum()) / ((n * n - n) * K), None


def test_validation_block_15000_rows_matches_oracle(monkeypatch):
    """Validation losses / metrics on a 15 000-row split against the float64 oracle (explicit draws)."""
    from rankaae_b200.engine import Engine
    from tests import parity_util as PU
    monkeypatch.setattr(O, "kendall_constraint", _kendall_chunked)
    cfg_d = _cfg()
    cfg = O.Config.from_dict(cfg_d)
    n_val = 15_000
    rng = np.random.default_rng(77)
    state = PU.f32_state(O.init_state(cfg, rng))
    spec, aux = O.synthetic_dataset(n_val + 1024, cfg, seed=6, dtype=np.float32)
    eng = Engine(cfg_d, n_trials=1, device="cuda:0", max_rows=n_val)
    eng.set_state(0, state, None)
    eng.bind_dataset(spec[n_val:], aux[n_val:], spec[:n_val], aux[:n_val])
    z_sample = np.float32(rng.standard_normal((n_val, cfg.nstyle)))
    z_real = np.float32(rng.standard_normal((cfg.batch_size, cfg.nstyle)))
    got = eng.validate(0, z_sample=z_sample, z_real=z_real, epoch=0, avg_mutual_info=0.5)
    ref = O.validate(O.clone_state(state), cfg, spec[:n_val].astype(np.float64), aux[:n_val].astype(np.float64),
                     z_real.astype(np.float64), z_sample.astype(np.float64), epoch=0, avg_mutual_info=0.5)
    m, r = np.asarray(got["metrics"], np.float64), np.asarray(ref["metrics"], np.float64)
    assert abs(m[0] - r[0]) <= 2e-4, (m, r)          # min Shapiro-Wilk W (float32 sort keys vs float64)
    assert abs(m[1] - r[1]) <= 1e-4 * max(1.0, abs(r[1])), (m, r)
    assert abs(m[3] - r[3]) <= 2e-3, (m, r)          # max |Spearman|
    assert abs(m[4] - r[4]) <= 5e-4, (m, r)          # Kendall on 15 000^2 x 5 pairs
    eng.close()
