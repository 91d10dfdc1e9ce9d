"""GPU tests of the thread-block cluster per trial (`ctas_per_trial` 2 / 4 / 8): 128-row tiles of a batch are dealt
round-robin to the CTAs of the trial's cluster; BatchNorm statistics, weight gradients, BN-backward sums, losses and Kendall
totals are reduced through distributed shared memory in rank order.

The cluster path must satisfy the SAME oracle parity as one CTA per trial (teacher-forced, float64 oracle), on the golden
fixtures of the unmodified reference, on full-size batches (1024 rows, the ragged 804) and on batches that leave CTAs of
the cluster without rows; the validation block must give the reference's numbers; repeated calls must agree bit for bit;
and the production path (in-kernel generator keyed by the GLOBAL row, so the draws do not depend on the cluster size)
must follow the one-CTA trajectory."""
import numpy as np
import pytest

from oracle import aae_oracle as O
from tests import parity_util as PU
from tests.golden_util import Golden
from tests.test_parity_gpu import EXAMPLE, _golden_states

pytestmark = pytest.mark.gpu
CLUSTERS = [2, 4, 8]


def _engine(cfg_dict, ctas, n_trials=1, max_rows=None):
    from rankaae_b200.engine import Engine
    return Engine(dict(cfg_dict, ctas_per_trial=ctas), n_trials=n_trials, device="cuda:0", max_rows=max_rows)


def test_cluster_sizes_validated():
    from rankaae_b200 import _lib as L
    from rankaae_b200.engine import Engine
    with pytest.raises(L.RaaeError):
        Engine(dict(EXAMPLE, ctas_per_trial=3), n_trials=1, device="cuda:0")


@pytest.mark.parametrize("ctas", CLUSTERS)
def test_cluster_golden_parity_and_validation(ctas):
    """Golden fixtures of the unmodified reference (batches of 200 and 80 rows: CTAs beyond the second own no rows), every
    phase + AdamW + the validation block against the reference's own float64 numbers."""
    g = Golden("step_warm.npz")
    eng = _engine(g.cfg_dict, ctas, max_rows=max(g.cfg.batch_size, g.n_val))
    try:
        states, end_state = _golden_states(g)
        for batch in (0, 1):
            state, opt = states[batch]
            x, aux, rnd = g.batch(batch)
            x = np.float32(x).astype(np.float64)
            for phase in O.PHASES:
                rep, _, _ = PU.compare_phase(eng, 0, g.cfg, state, opt, x, aux, PU.f32_rnd(rnd), g.record_epoch, phase,
                                             tag=f"cluster{ctas}-golden-b{batch}")
                PU.check_phase_report(rep)
        state, opt = states[1]
        x, aux, rnd = g.batch(1)
        x = np.float32(x).astype(np.float64)
        for phase in O.PHASES:
            PU.check_adam(eng, 0, g.cfg, state, opt, x, aux, PU.f32_rnd(rnd), g.record_epoch, phase)
        v = g.val()
        eng.set_state(0, PU.f32_state(end_state))
        eng.bind_dataset(g.spec[:g.n_train], g.aux[:g.n_train], v["spec"], v["aux"])
        out = eng.validate(0, z_sample=v["z_sample"], z_real=v["z_real"], epoch=g.record_epoch,
                           avg_mutual_info=float(v["metrics"][2]))
        for ph in O.PHASES:
            assert abs(out["losses"][ph] - v["losses"][ph]) <= 5e-4 * max(1.0, abs(v["losses"][ph])), (ph, out["losses"])
        m = out["metrics"]
        assert abs(m[0] - v["metrics"][0]) <= 2e-4                                   # min Shapiro W
        assert abs(m[1] - v["metrics"][1]) <= 5e-4 * max(1.0, abs(v["metrics"][1]))
        assert abs(m[3] - v["metrics"][3]) <= 1e-3                                   # Spearman coupling
        assert abs(m[4] - v["metrics"][4]) <= 5e-4
    finally:
        PU.dump_report(f"parity_cluster{ctas}_golden.json")
        eng.close()


@pytest.mark.parametrize("tensor_cores", [23, 0], ids=["tcgen05", "fp32fma"])
@pytest.mark.parametrize("ctas", CLUSTERS)
def test_cluster_fullsize_phase_parity(ctas, tensor_cores):
    """BASELINE configs[1] batches: 1024 rows (one tile per CTA at 8 CTAs) and the ragged 804 (the last CTA of 8 owns no
    rows, the seventh a 36-row tile), trained-like state, all five phases against the float64 oracle."""
    eng = _engine(dict(EXAMPLE, tensor_cores=tensor_cores), ctas, max_rows=1056)
    cfg = O.Config.from_dict(EXAMPLE)
    try:
        for rows in (1024, 804):
            rng = np.random.default_rng(100 + rows)
            state = PU.f32_state(O.init_state(cfg, rng))
            for net in ("E", "D", "S"):
                state[net]["a"] = [np.float32(a + rng.uniform(-0.005, 0.2, a.shape)).astype(np.float64) for a in state[net]["a"]]
            spec, aux = O.synthetic_dataset(rows, cfg, seed=rows, dtype=np.float32)
            x = np.float32(spec + cfg.spec_noise * rng.standard_normal(spec.shape)).astype(np.float64)
            rnd = PU.f32_rnd(O.draw_step_randoms(cfg, rows, rng))
            for phase in O.PHASES:
                rep, _, _ = PU.compare_phase(eng, 0, cfg, state, None, x, aux.astype(np.float64), rnd, 700, phase,
                                             tag=f"cluster{ctas}-full-{rows}")
                PU.check_phase_report(rep)
    finally:
        PU.dump_report(f"parity_cluster{ctas}_fullsize_tc{tensor_cores}.json")
        eng.close()


@pytest.mark.parametrize("ctas", CLUSTERS)
def test_cluster_matches_single_cta(ctas):
    """Same step at ctas_per_trial = 1 and in a cluster: losses to 5e-6, gradients to 1e-4 rel-L2 per network (only the
    order of the reductions differs), applied AdamW state to 1e-5; twice the same cluster call: identical bits."""
    cfg = O.Config.from_dict(EXAMPLE)
    rows = 804
    rng = np.random.default_rng(55)
    state = PU.f32_state(O.init_state(cfg, rng))
    for net in ("E", "D", "S"):
        state[net]["a"] = [np.float32(a + rng.uniform(0.02, 0.2, a.shape)).astype(np.float64) for a in state[net]["a"]]
    spec, aux = O.synthetic_dataset(rows, cfg, seed=12, dtype=np.float32)
    rnd = PU.f32_rnd(O.draw_step_randoms(cfg, rows, rng))
    res = {}
    for c in (1, ctas, ctas):
        eng = _engine(EXAMPLE, c, max_rows=1056)
        eng.set_state(0, state)
        out = eng.step_debug(0, spec, aux, rnd, epoch=30, apply_updates=True)
        res.setdefault(c, []).append((out, eng.state[0].clone().cpu().numpy()))
        eng.close()
    (a, sa), (b, sb), (b2, sb2) = res[1][0], res[ctas][0], res[ctas][1]
    assert np.array_equal(sb, sb2) and all(b["losses"][ph] == b2["losses"][ph] for ph in O.PHASES)
    for ph in O.PHASES:
        assert abs(a["losses"][ph] - b["losses"][ph]) <= 5e-6 * max(1.0, abs(a["losses"][ph])), ph
    # the phases run back to back with their updates: later phases see the earlier updates of their own path
    for ph, nets in a["grads"].items():
        for net in nets:
            va = PU.net_vec(a["grads"][ph][net], skip_last_bias=(net == "E"))
            vb = PU.net_vec(b["grads"][ph][net], skip_last_bias=(net == "E"))
            assert PU.rel_l2(va, vb) <= 1e-3, (ph, net, PU.rel_l2(va, vb))
    assert PU.rel_l2(sb.astype(np.float64), sa.astype(np.float64)) <= 1e-4


@pytest.mark.parametrize("ctas", [2, 8])
def test_cluster_production_epochs_follow_single_cta(ctas):
    """Production path (device-resident dataset, in-kernel generator, validation + metrics + scheduler in-kernel) for
    three epochs of two trials: the draws are keyed by the global row, so the cluster run follows the one-CTA run (float32
    reduction order apart); BN counters and optimizer step counts are identical."""
    import torch
    from rankaae_b200.trainer import init_trial_state
    cfg = dict(EXAMPLE, batch_size=512, max_epoch=40)
    ocfg = O.Config.from_dict(cfg)
    spec, aux = O.synthetic_dataset(2400, ocfg, seed=1, dtype=np.float32)
    outs = {}
    perm = None
    for c in (1, ctas):
        eng = _engine(cfg, c, n_trials=2, max_rows=512)
        for t in range(2):
            init_trial_state(eng, t, cfg, seed=t)
        eng.bind_dataset(spec[:1680], aux[:1680], spec[1680:2040], aux[1680:2040])
        if perm is None:
            perm = eng.make_perm(3)
        losses, metrics = eng.train_epochs(0, 3, perm)
        torch.cuda.synchronize()
        outs[c] = (losses.cpu().numpy(), metrics.cpu().numpy(), eng.get_state(0), eng.get_state(1))
        eng.close()
    l1, m1, s1, _ = outs[1]
    lc, mc, sc, _ = outs[ctas]
    assert np.isfinite(lc).all() and np.isfinite(mc).all()
    assert sc[0]["E"]["nbt"] == s1[0]["E"]["nbt"] == 3 * 4 * 6 and sc[0]["D"]["nbt"] == s1[0]["D"]["nbt"]
    assert all(sc[1][ph]["t"] == s1[1][ph]["t"] for ph in O.PHASES)
    # epoch 0 starts from identical weights: its train losses (last batch) and validation numbers agree closely; AdamW's
    # first sign-like steps then amplify float32 noise, so later epochs are compared loosely
    np.testing.assert_allclose(lc[0], l1[0], rtol=2e-2, atol=2e-3)
    np.testing.assert_allclose(mc[-1][:, 1], m1[-1][:, 1], rtol=0.25)        # validation reconstruction after 3 epochs
    assert (mc[-1][:, 1] < mc[0][:, 1]).all()
