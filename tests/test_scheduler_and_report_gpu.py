"""GPU tests of (a) the in-kernel ReduceLROnPlateau against torch's (trainer.py:303-304, 400-408) on scripted metric
sequences and end to end with a short patience, and (b) report compatibility: final.pt written by a fused trial is evaluated
by the reference's own `evaluate_model` (sc/report/analysis.py:394-450) and the batched device-side
`rankaae_b200.evaluate.evaluate_trials` must give the same numbers."""
import math
import os
import sys
import types

import numpy as np
import pytest

from oracle import aae_oracle as O
from tests.test_parity_gpu import EXAMPLE

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_DIRS = [os.path.join(ROOT, "baseline", "_ref"), "/root/reference"]


def scripted_sequences():
    rng = np.random.default_rng(0)
    seqs = {}
    # improving, then a plateau longer than the patience, improving again, second plateau
    seqs["plateaus"] = np.concatenate([np.linspace(1.0, 0.5, 12), np.full(9, 0.499), np.linspace(0.45, 0.2, 6), np.full(14, 0.21)])
    # negative metric: `metric < best * (1 - 0.01)` demands a value 1 % CLOSER to zero's far side, i.e. a more negative best
    # makes the bar HARDER (-0.50 -> must beat -0.495 is wrong way round: the reference's quirk, trainer.py:297 + rel mode)
    seqs["negative"] = np.concatenate([np.linspace(-0.1, -0.6, 15), np.full(8, -0.6), -0.6 - 0.001 * np.arange(10)])
    seqs["noisy"] = 0.3 + 0.05 * rng.standard_normal(60)
    seqs["sign_change"] = np.concatenate([np.linspace(0.2, -0.2, 20), np.full(10, -0.2)])
    return seqs


@pytest.mark.parametrize("name", sorted(scripted_sequences()))
@pytest.mark.parametrize("lr0,factor,patience", [(1e-2, 0.1, 3), (1e-3, 0.5, 5), (3e-8, 0.1, 2)])
def test_plateau_scheduler_matches_torch(name, lr0, factor, patience):
    """lr / best / num_bad_epochs after EVERY step against torch.optim.lr_scheduler.ReduceLROnPlateau(mode='min',
    threshold 0.01 rel, cooldown 0, eps 1e-8) as the reference constructs it; (3e-8, 0.1): the update lr - new_lr = 2.7e-8 is
    above eps once and below it afterwards (the minimum-change rule)."""
    import ctypes as C
    import torch
    from rankaae_b200 import _lib as L
    from rankaae_b200.engine import Engine
    seq = scripted_sequences()[name]
    cfg = dict(EXAMPLE, lr_base=lr0, lr_ratio_Corr=1, lr_ratio_Reconn=1, lr_ratio_Mutual=1, lr_ratio_Smooth=1, lr_ratio_dis=1,
               sch_factor=factor, sch_patience=patience)
    eng = Engine(cfg, n_trials=1, device="cuda:0")
    m = torch.tensor(seq, dtype=torch.float64, device="cuda:0")
    out = torch.zeros(len(seq), L.NUM_PHASES, 3, dtype=torch.float32, device="cuda:0")
    L.check(eng.lib.raae_debug_plateau(eng.handle, 0, m.data_ptr(), len(seq), out.data_ptr(), eng.stream))
    torch.cuda.synchronize()
    got = out.cpu().numpy()
    eng.close()
    par = torch.nn.Parameter(torch.zeros(1))
    opt = torch.optim.AdamW([par], lr=lr0)
    sch = torch.optim.lr_scheduler.ReduceLROnPlateau(opt, mode="min", factor=factor, patience=patience, cooldown=0, threshold=0.01)
    orc = O.ReduceLROnPlateau(lr0, factor=factor, patience=patience)
    for e, v in enumerate(seq):
        sch.step(float(v))
        orc.step(float(v))
        lr = opt.param_groups[0]["lr"]
        assert orc.lr == pytest.approx(lr, rel=1e-12) and orc.num_bad_epochs == sch.num_bad_epochs
        for o in range(L.NUM_PHASES):
            assert got[e, o, 0] == pytest.approx(lr, rel=2e-7), (name, e, o, got[e, o], lr)
            assert got[e, o, 2] == sch.num_bad_epochs, (name, e, o, got[e, o], sch.num_bad_epochs)
            assert (math.isinf(sch.best) and math.isinf(got[e, o, 1])) or got[e, o, 1] == pytest.approx(sch.best, rel=2e-7)
    if name == "plateaus":
        assert opt.param_groups[0]["lr"] < lr0 or lr0 < 1e-7    # the plateaus are longer than every patience used here


def test_short_patience_end_to_end():
    """Production epochs with sch_patience 2: the learning rates in the state block follow torch's scheduler stepped on
    the combined metric the kernel reported each epoch."""
    import torch
    from rankaae_b200.engine import Engine
    from rankaae_b200.trainer import init_trial_state
    cfg = dict(EXAMPLE, batch_size=256, max_epoch=40, sch_patience=2, sch_factor=0.5)
    spec, aux = O.synthetic_dataset(1400, O.Config.from_dict(cfg), seed=1, dtype=np.float32)
    eng = Engine(cfg, n_trials=2, device="cuda:0", max_rows=512)
    for t in range(2):
        init_trial_state(eng, t, cfg, seed=t)
    eng.bind_dataset(spec[:980], aux[:980], spec[980:1190], aux[980:1190])
    n_ep = 30
    _, metrics = eng.train_epochs(0, n_ep)
    torch.cuda.synchronize()
    combined = metrics.cpu().numpy()[:, :, 5]
    for t in range(2):
        _, opt_state = eng.get_state(t)
        for ph, lr0 in zip(O.PHASES, (1e-3, 1e-2, 1e-2, 1e-3, 1e-3)):
            par = torch.nn.Parameter(torch.zeros(1))
            opt = torch.optim.AdamW([par], lr=lr0)
            sch = torch.optim.lr_scheduler.ReduceLROnPlateau(opt, mode="min", factor=0.5, patience=2, cooldown=0, threshold=0.01)
            for e in range(n_ep):
                sch.step(float(combined[e, t]))
            assert opt_state[ph]["lr"] == pytest.approx(opt.param_groups[0]["lr"], rel=1e-6), (t, ph)
            assert opt_state[ph]["bad"] == sch.num_bad_epochs
        assert opt_state["reconstruction"]["lr"] < 1e-2         # the short patience did trigger
    eng.close()


def _reference_analysis():
    """The reference's sc.report.analysis with its plotting imports stubbed (matplotlib / seaborn / plotly are not installed)."""
    ref = next((d for d in REF_DIRS if os.path.exists(os.path.join(d, "sc", "report", "analysis.py"))), None)
    if ref is None:
        pytest.skip("reference sources not available (baseline/_ref)")

    class Stub(types.ModuleType):
        def __getattr__(self, k):
            if k.startswith("__"):
                raise AttributeError(k)
            return Stub(self.__name__ + "." + k)

        def __call__(self, *a, **k):
            return None

    for name in ("seaborn", "matplotlib", "matplotlib.pyplot", "plotly", "plotly.express", "torch_optimizer", "ipyparallel"):
        sys.modules.setdefault(name, Stub(name))
    if ref not in sys.path:
        sys.path.insert(0, ref)
    import sc.report.analysis as A
    import sklearn.metrics
    # scikit-learn >= 1.x returns a Python float from f1_score; the reference calls .tolist() on it (analysis.py:267)
    A.f1_score = lambda *a, **k: np.float64(sklearn.metrics.f1_score(*a, **k))
    return A


def test_reference_evaluate_model_on_fused_trials(tmp_path):
    """Three fused trials -> final.pt (reference classes) -> the reference's evaluate_model on the test split; the batched
    device-side evaluate_trials reproduces its numbers from ONE launch over all trials."""
    import torch
    A = _reference_analysis()
    from rankaae_b200.engine import Engine
    from rankaae_b200.evaluate import evaluate_trials
    from rankaae_b200.trainer import build_modules, save_final
    cfg = dict(EXAMPLE, batch_size=256, max_epoch=30)
    spec, aux = O.synthetic_dataset(1600, O.Config.from_dict(cfg), seed=2, dtype=np.float32)
    n_tr, n_va = 1120, 240
    eng = Engine(cfg, n_trials=3, device="cuda:0", max_rows=512)
    mods = []
    for t in range(3):
        m = build_modules(cfg, seed=t)
        eng.load_modules(t, *m)
        mods.append(m)
    eng.bind_dataset(spec[:n_tr], aux[:n_tr], spec[n_tr:n_tr + n_va], aux[n_tr:n_tr + n_va])
    eng.train_epochs(0, 30)
    torch.cuda.synchronize()
    test_spec, test_aux = spec[n_tr + n_va:], aux[n_tr + n_va:]
    ours, z = evaluate_trials(eng, test_spec, test_aux)
    ds = types.SimpleNamespace(spec=test_spec, aux=test_aux)
    for t in range(3):
        eng.store_modules(t, *mods[t])
        path = str(tmp_path / f"final_{t}.pt")
        assert save_final(mods[t], path, reference_classes=True)
        model = torch.load(path, map_location="cpu", weights_only=False)
        assert type(model["Encoder"]).__module__ == "sc.clustering.model"
        with torch.no_grad():
            ref = A.evaluate_model(ds, model)
        got = ours[t]
        assert got["Reconstruct Err"][0] == pytest.approx(ref["Reconstruct Err"][0], abs=2e-4)
        assert got["Reconstruct Err"][1] == pytest.approx(ref["Reconstruct Err"][1], abs=2e-4)
        assert got["Inter-style Corr"] == pytest.approx(ref["Inter-style Corr"], abs=2e-3)
        for i in range(5):
            r, g = ref["Style-descriptor Corr"][i], got["Style-descriptor Corr"][i]
            if i == 1:
                assert g["F1 score"] == pytest.approx(r["F1 score"], abs=5e-3), (t, r, g)
                continue
            assert g["Spearman"] == pytest.approx(r["Spearman"], abs=2e-3), (t, i, r, g)
            assert g["Linear"]["R2"] == pytest.approx(r["Linear"]["R2"], abs=2e-3)
        # the latents of the device-side evaluation are the pickled encoder's
        with torch.no_grad():
            z_ref = model["Encoder"](torch.from_numpy(test_spec)).numpy()
        np.testing.assert_allclose(z[t].cpu().numpy(), z_ref, atol=3e-4, rtol=0)
    eng.close()
