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


@pytest.mark.parametrize("tensor_cores", [55, 0], ids=["tcgen05", "fp32fma"])
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
    """Same teacher-forced step at ctas_per_trial = 1 and in a cluster: losses to 5e-6, gradients to 5e-3 rel-L2 per network
    (only the order of the reductions differs); one applied phase: state blocks agree entry-wise (see below); the same cluster call twice (all
    five phases with their updates): identical bits."""
    cfg = O.Config.from_dict(EXAMPLE)
    rows = 804
    rng = np.random.default_rng(55)
    state = PU.f32_state(O.init_state(cfg, rng))
    for net in ("E", "D", "S"):
        state[net]["a"] = [np.float32(a + rng.uniform(0.02, 0.2, a.shape)).astype(np.float64) for a in state[net]["a"]]
    spec, aux = O.synthetic_dataset(rows, cfg, seed=12, dtype=np.float32)
    rnd = PU.f32_rnd(O.draw_step_randoms(cfg, rows, rng))
    res = {}
    for c in (1, ctas):
        eng = _engine(EXAMPLE, c, max_rows=1056)
        eng.set_state(0, state)
        forced = eng.step_debug(0, spec, aux, rnd, epoch=30, apply_updates=False)
        eng.set_state(0, state)
        eng.step_debug(0, spec, aux, rnd, epoch=30, phase_mask=1 << 2, apply_updates=True)
        one_phase = eng.state[0].clone().cpu().numpy()
        full = []
        for _ in range(2):
            eng.state.zero_()
            eng.reset_optimizers()
            eng.set_state(0, state)
            out = eng.step_debug(0, spec, aux, rnd, epoch=30, apply_updates=True)
            full.append((out["losses"], eng.state[0].clone().cpu().numpy()))
        res[c] = (forced, one_phase, full)
        eng.close()
    (a, sa, _), (b, sb, full) = res[1], res[ctas]
    assert np.array_equal(full[0][1], full[1][1]) and all(full[0][0][ph] == full[1][0][ph] for ph in O.PHASES)
    for ph in O.PHASES:
        assert abs(a["losses"][ph] - b["losses"][ph]) <= 5e-6 * max(1.0, abs(a["losses"][ph])), (ph, a["losses"], b["losses"])
    for ph, nets in a["grads"].items():
        for net in nets:
            va = PU.net_vec(a["grads"][ph][net], skip_last_bias=(net == "E"))
            vb = PU.net_vec(b["grads"][ph][net], skip_last_bias=(net == "E"))
            assert PU.rel_l2(va, vb) <= 5e-3, (ph, net, PU.rel_l2(va, vb))     # a PReLU kink flip moves it by ~1/rows
    # applied update (fresh AdamW state: the first step is sign-like, -lr sign(g), so entries with |g| ~ 0 may land on
    # either side; everything else agrees): at most 0.5 % of the state block differs by more than 1e-5
    ok = np.isfinite(sa)                                  # the plateau scheduler's `best` starts at +inf
    assert np.array_equal(ok, np.isfinite(sb))
    assert float(np.mean(np.abs(sb[ok] - sa[ok]) > 1e-5)) <= 5e-3


@pytest.mark.parametrize("ctas", [2, 8])
def test_cluster_production_epochs_follow_single_cta(ctas):
    """Production path (device-resident dataset, in-kernel generator keyed by the GLOBAL row, validation + metrics +
    scheduler in-kernel): from the same WARM state (weights, BN buffers, AdamW moments after three epochs - with fresh
    moments AdamW's first steps are sign-like and amplify float32 reduction-order noise into different trajectories) one
    more epoch in a cluster ends in the one-CTA state.  The 20 free-running updates of that epoch amplify reduction-order
    noise chaotically (a PReLU kink or a dropout-adjacent sign flips), so the comparison is statistical: over four seeded
    shuffles x 2 trials x 3 networks the MEDIAN parameter rel-L2 stays below 1e-3 (measured 1.5e-5 ... 8e-5, 90th percentile
    3e-4 ... 8e-4 over the builds of the round, tools/cluster_follow.py -> profiles/cluster_follow_r02.txt) and no single case
    exceeds 5e-2 (measured maxima 3.6e-3 and 6.5e-3; one unseeded run hit 1.7e-2); BN
    counters / optimizer step counts are identical, losses and metrics of the epoch close."""
    import torch
    from rankaae_b200.trainer import init_trial_state
    from rankaae_b200 import _lib as L
    from rankaae_b200.engine import make_config
    cfg = dict(EXAMPLE, batch_size=512, max_epoch=40)
    ocfg = O.Config.from_dict(cfg)
    spec, aux = O.synthetic_dataset(2400, ocfg, seed=1, dtype=np.float32)
    data = (spec[:1680], aux[:1680], spec[1680:2040], aux[1680:2040])
    warm = _engine(cfg, 1, n_trials=2, max_rows=512)
    for t in range(2):
        init_trial_state(warm, t, cfg, seed=t)
    warm.bind_dataset(*data)
    warm.train_epochs(0, 3)
    torch.cuda.synchronize()
    snapshot = warm.state.clone()
    perms = [warm.make_perm(1, generator=torch.Generator(device=warm.device).manual_seed(1000 + s)) for s in range(4)]
    warm.close()
    lay = L.query_layout(make_config(cfg, 2, 512))
    rels, close = [], []
    for si, perm in enumerate(perms):
        outs = {}
        for c in (1, ctas):
            eng = _engine(cfg, c, n_trials=2, max_rows=512)
            eng.state.copy_(snapshot)
            eng.bind_dataset(*data)
            losses, metrics = eng.train_epochs(3, 1, perm)
            torch.cuda.synchronize()
            outs[c] = (losses.cpu().numpy()[0], metrics.cpu().numpy()[0], eng.state.clone().cpu().numpy(), eng.get_state(0))
            eng.close()
        l1, m1, b1, s1 = outs[1]
        lc, mc, bc, sc = outs[ctas]
        assert np.isfinite(lc).all() and np.isfinite(mc).all()
        assert sc[0]["E"]["nbt"] == s1[0]["E"]["nbt"] == 4 * 4 * 6 and sc[0]["D"]["nbt"] == s1[0]["D"]["nbt"] == 4 * 4 * 4
        assert all(sc[1][ph]["t"] == s1[1][ph]["t"] == 16 for ph in O.PHASES)
        for t in range(2):
            for ni in range(3):
                n = lay.net[ni]
                a, b = b1[t, n.param_off:n.param_off + n.n_params], bc[t, n.param_off:n.param_off + n.n_params]
                rels.append(PU.rel_l2(b.astype(np.float64), a.astype(np.float64)))
        close.append(np.isclose(lc, l1, rtol=5e-2, atol=5e-3).ravel())
        close.append(np.isclose(mc[:, :5], m1[:, :5], rtol=5e-2, atol=5e-3).ravel())
    assert float(np.median(rels)) <= 1e-3, rels
    assert max(rels) <= 5e-2, rels
    # losses / metrics of the epoch: the same chaotic tail - at most 5 % of the 4 x (24 + 10) numbers may leave the band
    assert float(np.mean(np.concatenate(close))) >= 0.95, [float(np.mean(c)) for c in close]
