"""CPU tests of the host side: config namespace (the reference's own sc/tests/test_parameters.py restated),
CSV loader schema/splits, state-block layout, the C-ABI exports, trial sharding and the gloo metric gather."""
import ctypes
import os
import re
import subprocess
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def built_lib():
    import __graft_entry__ as graft
    return graft.build()


# ---- Parameters: sc/tests/test_parameters.py:6-50 -------------------------------------------------------
def test_parameters_from_yaml(tmp_path):
    from rankaae_b200.parameter import Parameters
    f = tmp_path / "fix_config.yaml"
    f.write_text("ae_form: FC\nalpha_limit: 0.7172\ntrials: 32\n")
    p = Parameters.from_yaml(str(f))
    assert p.ae_form == "FC"
    assert p.alpha_limit == 0.7172


def test_parameters_update_and_immutability():
    from rankaae_b200.parameter import Parameters
    p = Parameters(dict(nstyle=2, weight_decay=1e-2, lr_ratio_Reconn=2.0, optimizer_name="AdamW",
                        aux_weights=None, kendall_activation=False))
    assert p.get('nstyle', 0) == 2
    assert p.get('nstyll', 0) == 0
    with pytest.raises(TypeError):
        p.nstyle = 3
    p.update({"nstyle": 5, "new_key": "x"})
    assert p.nstyle == 5 and p.new_key == "x" and p.to_dict()["nstyle"] == 5


def test_registries_reject_unimplemented_families():
    from rankaae_b200.parameter import AE_CLS_DICT, OPTIM_DICT
    from rankaae_b200.trainer import build_modules
    assert set(AE_CLS_DICT) == {"FC"} and set(OPTIM_DICT) == {"AdamW"}
    with pytest.raises(NotImplementedError):
        build_modules(dict(ae_form="normal"))
    with pytest.raises(ValueError):
        build_modules(dict(ae_form="FC", decoder_activation="Tanh"))


def test_module_structure_matches_reference_counts():
    """FCEncoder/FCDecoder(n_layers=5)/DiscriminatorFC: 29574 / 29824 / 4801 parameters, reference state_dict keys
    (SURVEY.md §4 fixture caveats)."""
    from rankaae_b200.trainer import build_modules
    e, d, s = build_modules(dict(nstyle=6, n_layers=5, dim_in=256, dim_out=256, decoder_activation="Softplus"), seed=0)
    assert [sum(p.numel() for p in m.parameters()) for m in (e, d, s)] == [29574, 29824, 4801]
    assert "main.16.weight" in e.state_dict() and "main.17.running_mean" in e.state_dict()
    assert "main.16.bias" in d.state_dict() and d.nstyle == 6
    e2, _, _ = build_modules(dict(nstyle=6, n_layers=5, dim_in=256, dim_out=256, decoder_activation="Softplus"), seed=0)
    assert all(torch.equal(a, b) for a, b in zip(e.parameters(), e2.parameters()))


# ---- loader: dataloader.py:8-77 ----------------------------------------------------------------------
def test_csv_loader_schema_and_splits(tmp_path):
    from oracle import ref_shim
    from oracle.aae_oracle import Config, synthetic_dataset
    from rankaae_b200.dataloader import get_datasets
    spec, aux = synthetic_dataset(101, Config(dim_in=32, dim_out=32), seed=2, dtype=np.float32)
    csv = str(tmp_path / "d.csv")
    ref_shim.write_csv(csv, spec, aux)
    tr, va, te = get_datasets(csv, n_aux=5)
    assert (len(tr), len(va), len(te)) == (70, 15, 16)          # int(101*.7), int(101*.15), remainder
    s, a = va.tensors()
    assert s.dtype == np.float32 and s.shape == (15, 32) and a.shape == (15, 5)
    np.testing.assert_array_equal(s, spec[70:85])
    np.testing.assert_array_equal(a, aux[70:85])
    assert len(tr.grid) == 32
    with pytest.raises(AssertionError):
        get_datasets(csv, n_aux=6)                               # column n_aux-1 must be AUX_*, column n_aux ENE_*


# ---- C ABI -------------------------------------------------------------------------------------------
def test_c_abi_exports_every_declared_symbol(built_lib):
    hdr = open(os.path.join(ROOT, "include", "rankaae_b200.h")).read()
    declared = set(re.findall(r"\b(raae_[a-z_]+)\s*\(", hdr))
    from rankaae_b200 import _lib
    assert declared == set(_lib.EXPORTS), declared ^ set(_lib.EXPORTS)
    lib = ctypes.CDLL(built_lib)
    for name in declared:
        assert hasattr(lib, name), name
    assert _lib.load().raae_version() == 100


def test_layout_query_example_config(built_lib):
    from rankaae_b200 import _lib as L
    from rankaae_b200.engine import make_config
    cfg = make_config(dict(dim_in=256, dim_out=256, nstyle=6, n_aux=5, n_layers=5, FC_discriminator_layers=3,
                           batch_size=1024, decoder_activation="Softplus", ae_form="FC"), n_trials=3, max_rows=1056)
    lay = L.query_layout(cfg)
    assert [lay.net[i].n_params for i in range(3)] == [29574, 29824, 4801]
    assert [lay.opt[o].n for o in range(5)] == [34375, 29574, 59398, 59398, 29824]      # 425 138 moment floats / 2
    assert sum(2 * lay.opt[o].n for o in range(5)) == 425138                           # SURVEY.md §7 "State residency"
    assert lay.state_floats % 4 == 0 and lay.scratch_floats % 4 == 0
    for i in range(3):
        assert lay.net[i].param_off % 4 == 0
    # every region is disjoint
    spans = []
    for i in range(3):
        spans.append((lay.net[i].param_off, lay.net[i].param_off + lay.net[i].n_params))
    for o in range(5):
        spans += [(lay.opt[o].m_off, lay.opt[o].m_off + lay.opt[o].n), (lay.opt[o].v_off, lay.opt[o].v_off + lay.opt[o].n),
                  (lay.opt[o].scalar_off, lay.opt[o].scalar_off + 4)]
    spans.append((lay.misc_off, lay.misc_off + 16))
    spans.sort()
    assert all(a[1] <= b[0] for a, b in zip(spans, spans[1:]))
    assert spans[-1][1] <= lay.state_floats


def test_config_errors_are_reported(built_lib):
    from rankaae_b200 import _lib as L
    from rankaae_b200.engine import make_config
    base = dict(dim_in=256, dim_out=256, nstyle=6, n_aux=5, n_layers=5, batch_size=64, ae_form="FC")
    for bad, msg in ((dict(dim_in=250), "multiple of 4"), (dict(nstyle=9), "nstyle"), (dict(n_aux=7), "n_aux"),
                     (dict(FC_discriminator_layers=4), "FC_discriminator_layers"), (dict(dim_out=128), "dim_in must equal")):
        with pytest.raises(L.RaaeError, match=msg):
            L.query_layout(make_config(dict(base, **bad), n_trials=1))
    with pytest.raises(NotImplementedError):
        make_config(dict(base, gradient_reversal=False), n_trials=1)
    # contraction groups: bits 0, 1, 2, 4, 5 exist (55 = all of them, the default); bit 3 and anything above bit 5 do not
    from rankaae_b200.engine import DEFAULT_TENSOR_CORES
    assert DEFAULT_TENSOR_CORES == 55
    for ok in (0, 7, 23, 55):
        L.query_layout(make_config(dict(base, tensor_cores=ok), n_trials=1))
    for bad in (8, 64, 63):
        with pytest.raises(L.RaaeError, match="tensor_cores"):
            L.query_layout(make_config(dict(base, tensor_cores=bad), n_trials=1))


def test_no_cpu_fallback(built_lib):
    """The product path must fail loudly without a GPU instead of computing elsewhere."""
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from rankaae_b200 import _lib as L
    from rankaae_b200.engine import Engine
    with pytest.raises(L.RaaeError, match="no CPU path"):
        Engine(dict(dim_in=256, dim_out=256, nstyle=6, n_aux=5, n_layers=5, batch_size=64, ae_form="FC"))


def test_hp_row_matches_reference_optimizer_table():
    """trainer.py:333-397: lrs, betas and the weight decays (mutual_info / adversarial use AdamW's default 1e-2)."""
    from rankaae_b200 import _lib as L
    from rankaae_b200.engine import hp_row
    cfg = dict(lr_base=0.002, lr_ratio_dis=3, lr_ratio_Corr=10, lr_ratio_Reconn=5, lr_ratio_Mutual=0.5, lr_ratio_Smooth=2,
               weight_decay=0.02, dis_beta=1.1, dropout_rate=0.04, dis_dropout_rate=0.056, max_epoch=77)
    hp = hp_row(cfg, seed=5)
    np.testing.assert_allclose(hp[L.HP_LR0:L.HP_LR0 + 5], [0.006, 0.02, 0.01, 0.001, 0.004])
    np.testing.assert_allclose(hp[L.HP_BETA1:L.HP_BETA1 + 5], [0.99, 0.9, 0.9, 0.9, 0.9])
    np.testing.assert_allclose(hp[L.HP_BETA2:L.HP_BETA2 + 5], [0.9999, 0.999, 0.999, 0.999, 0.999])
    np.testing.assert_allclose(hp[L.HP_WD:L.HP_WD + 5], [0.01, 0.02, 0.02, 0.01, 0.02])
    assert hp[L.HP_MAX_EPOCH] == 77 and hp[L.HP_SEED] == 5 and hp[L.HP_EPOCH_STOP_SMOOTH] == 500


def test_shapiro_weights_match_oracle_and_scipy():
    from scipy.stats import shapiro
    from oracle.aae_oracle import shapiro_w, shapiro_weights as ow
    from rankaae_b200.engine import shapiro_weights
    rng = np.random.default_rng(0)
    for n in (60, 1050):
        w = shapiro_weights(n)
        np.testing.assert_allclose(w, ow(n), rtol=0, atol=1e-12)
        assert abs(np.sum(w ** 2) - 1.0) < 1e-9
        x = rng.standard_normal(n) ** 3
        assert abs(shapiro_w(x, w) - shapiro(x).statistic) < 1e-10


# ---- ensemble sharding + gather (world_size 2, gloo) ------------------------------------------------
def test_shard_trials_partition():
    from rankaae_b200.ensemble import shard_trials
    for trials, world in ((64, 8), (7, 2), (3, 4), (1024, 8)):
        parts = [shard_trials(trials, world, r) for r in range(world)]
        assert sorted(sum(parts, [])) == list(range(trials))
        assert max(map(len, parts)) - min(map(len, parts)) <= 1


def test_loss_row_format_is_the_reference_one():
    from rankaae_b200.ensemble import format_loss_row
    row = format_loss_row(10, [-0.999078, -1.139364, 0, 0, -0.141979, -0.133747, 0.029688, 0.02685, 0.002777, 0.001395,
                               1.197696, 1.469422])
    # sc/tests/data/training/job_1/losses.csv:3
    assert row == ("10,\t-0.999078,\t-1.139364,\t0.000000,\t0.000000,\t-0.141979,\t-0.133747,\t0.029688,\t0.026850,\t"
                   "0.002777,\t0.001395,\t1.197696,\t1.469422,\t")


GLOO_WORKER = r"""
import os, sys
import numpy as np
import torch.distributed as dist
sys.path.insert(0, {root!r})
from rankaae_b200.ensemble import gather_results, shard_trials
rank, world, trials = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), 5
dist.init_process_group("gloo")
mine = shard_trials(trials, world, rank)
rows = [[t + 0.1, t + 0.2, t + 0.3, t + 0.4, t + 0.5, 100.0 + t] for t in mine]
res = gather_results(rows, trials, world, rank)
assert res.shape == (trials, 6)
for t in range(trials):
    assert abs(res[t, 0] - (t + 0.1)) < 1e-12 and abs(res[t, 5] - (100.0 + t)) < 1e-12, (rank, t, res[t])
dist.destroy_process_group()
open(os.path.join({out!r}, f"ok{{rank}}"), "w").write("ok")
"""


def test_metric_gather_world2_gloo(tmp_path):
    script = tmp_path / "w.py"
    script.write_text(GLOO_WORKER.format(root=ROOT, out=str(tmp_path)))
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
           "--master-port", "29531", str(script)]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=240)
    assert r.returncode == 0, r.stdout + r.stderr
    assert (tmp_path / "ok0").exists() and (tmp_path / "ok1").exists()


def test_data_parallel_shard_layout_partitions_rows():
    """world x replicas equally sized, disjoint, contiguous shards; 2 ranks x 1 replica and 1 rank x 2 replicas cover the
    same rows in the same shard order (what makes the two configurations bit-identical on the GPU)."""
    from rankaae_b200.dp import shard_layout
    for n_rows, world, V in ((840, 2, 1), (840, 1, 2), (1_000_000, 8, 16), (1001, 3, 2)):
        per = n_rows // (world * V)
        seen = []
        for r in range(world):
            p, lo, hi = shard_layout(n_rows, world, r, V)
            assert p == per and hi - lo == V * per
            seen += [(lo + v * per, lo + (v + 1) * per) for v in range(V)]
        assert seen == [(q * per, (q + 1) * per) for q in range(world * V)]
    assert shard_layout(840, 2, 1, 1) == (420, 420, 840) and shard_layout(840, 1, 0, 2) == (420, 0, 840)


def test_sweep_points_are_deterministic_and_in_range():
    """tools/sweep_1024.py: the hyper-parameters of global trial t depend on t only (not on the partition over ranks)."""
    import importlib.util
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    spec = importlib.util.spec_from_file_location("sweep_1024", os.path.join(root, "tools", "sweep_1024.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    a, b = mod.sweep_point(17), mod.sweep_point(17)
    assert a == b and a != mod.sweep_point(18)
    for t in range(0, 1024, 37):
        p = mod.sweep_point(t)
        assert 3e-4 <= p["lr_base"] <= 3e-3 and 1e-3 <= p["weight_decay"] <= 1e-1 and 0.0 <= p["dropout_rate"] <= 0.1
        assert p["n_layers"] == mod.BASE["n_layers"] and p["batch_size"] == 1024          # structural keys are fixed per launch
    from rankaae_b200.ensemble import shard_trials
    assert sorted(sum((shard_trials(1024, 8, r) for r in range(8)), [])) == list(range(1024))
    assert all(len(shard_trials(1024, 8, r)) == 128 for r in range(8))


# ------------------------------------------------------------------------------------------
# round 2: loader front end, report compatibility, generator, scheduler restatement, reference install
# ------------------------------------------------------------------------------------------
def test_package_generator_is_the_oracle_generator():
    """bench.py and the tools draw their data from rankaae_b200.synthetic (no oracle import on the product side); it must be
    the generator the tests use."""
    from oracle import aae_oracle as O
    from rankaae_b200.synthetic import synthetic_dataset
    cfg = O.Config.from_dict(dict(n_aux=5, dim_in=256, dim_out=256, nstyle=6))
    a = O.synthetic_dataset(300, cfg, seed=7, dtype=np.float32)
    b = synthetic_dataset(300, 5, 256, seed=7, dtype=np.float32)
    assert np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1])


def test_binary_cache_loader_one_parse(tmp_path, monkeypatch):
    """load_splits == the reference-rule splits of the direct parse; the second call maps the cache without touching the CSV
    parser; a stale cache (file changed) is rebuilt; the schema asserts of dataloader.py:21-25 still fire."""
    import pandas as pd
    from rankaae_b200 import dataloader as D
    from rankaae_b200.synthetic import synthetic_dataset, write_csv
    spec, aux = synthetic_dataset(203, 5, 256, seed=1)
    csv = str(tmp_path / "data.csv")
    write_csv(csv, spec, aux)
    splits = D.load_splits(csv, n_aux=5)
    for (s1, a1), ds in zip(splits, D.get_datasets(csv, n_aux=5)):
        s2, a2 = ds.tensors()
        assert np.array_equal(np.asarray(s1), s2) and np.array_equal(np.asarray(a1), a2)
    assert [len(s) for s, _ in splits] == [142, 30, 31]
    real = pd.read_csv
    monkeypatch.setattr(pd, "read_csv", lambda *a, **k: (_ for _ in ()).throw(AssertionError("parsed twice")))
    again = D.load_splits(csv, n_aux=5)
    assert np.array_equal(np.asarray(again[0][0]), np.asarray(splits[0][0]))
    monkeypatch.setattr(pd, "read_csv", real)
    write_csv(csv, spec[:100], aux[:100])                       # the file changed: size / mtime differ -> rebuilt
    assert [len(s) for s, _ in D.load_splits(csv, n_aux=5)] == [70, 15, 15]
    with pytest.raises(AssertionError):
        D.load_splits(csv, n_aux=6, cache_dir=str(tmp_path))   # column 5 is an ENE_ column, not AUX_
    t = D.to_device_pinned(splits[0][1], "cpu", chunk_bytes=256)
    assert np.array_equal(t.numpy(), np.asarray(splits[0][1]))


LOADER_WORKER = r"""
import os, sys
import numpy as np
import torch.distributed as dist
sys.path.insert(0, {root!r})
import pandas as pd
from rankaae_b200 import dataloader as D
rank = int(os.environ["RANK"])
dist.init_process_group("gloo")
real = pd.read_csv
def counted(*a, **k):
    open({marker!r} + f".{{rank}}", "a").write("x")
    return real(*a, **k)
pd.read_csv = counted
splits = D.load_splits({csv!r}, n_aux=5, rank=rank, world=2)
assert [len(s) for s, _ in splits] == [140, 30, 30]
dist.destroy_process_group()
"""


def test_loader_parses_once_for_two_ranks_gloo(tmp_path):
    from rankaae_b200.synthetic import synthetic_dataset, write_csv
    spec, aux = synthetic_dataset(200, 5, 256, seed=2)
    csv = str(tmp_path / "data.csv")
    write_csv(csv, spec, aux)
    script = tmp_path / "w.py"
    script.write_text(LOADER_WORKER.format(root=ROOT, marker=str(tmp_path / "parsed"), csv=csv))
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr",
                        "127.0.0.1", "--master-port", "29577", str(script)], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr[-2000:]
    assert os.path.exists(str(tmp_path / "parsed") + ".0") and not os.path.exists(str(tmp_path / "parsed") + ".1")


def _reference_dir():
    for d in (os.path.join(ROOT, "baseline", "_ref"), "/root/reference"):
        if os.path.exists(os.path.join(d, "sc", "clustering", "model.py")):
            return d
    return None


def test_final_pt_loads_with_only_the_reference_importable(tmp_path):
    """sc/report/analysis.py:115-121 unpickles final.pt with the reference on the path and nothing else: the pickle must
    name sc.clustering.model classes, be in eval mode and reproduce the encoder / decoder outputs."""
    ref = _reference_dir()
    if ref is None:
        pytest.skip("reference sources not available")
    from rankaae_b200.trainer import build_modules, save_final
    cfg = dict(ae_form="FC", nstyle=6, dropout_rate=0.04, dim_in=256, dim_out=256, n_layers=5, decoder_activation="Softplus",
               dis_dropout_rate=0.05, dis_noise=0.5, FC_discriminator_layers=3)
    mods = build_modules(cfg, seed=1)
    for m in mods[:2]:                                          # non-trivial BatchNorm buffers
        for b in m.modules():
            if isinstance(b, torch.nn.BatchNorm1d):
                b.running_mean.uniform_(-0.3, 0.3)
                b.running_var.uniform_(0.5, 1.5)
    sys.path.insert(0, ref)
    try:
        assert save_final(mods, str(tmp_path / "final.pt")) is True
    finally:
        sys.path.remove(ref)
        for k in [k for k in sys.modules if k == "sc" or k.startswith("sc.")]:
            del sys.modules[k]
    x = torch.randn(7, 256, generator=torch.Generator().manual_seed(0))
    with torch.no_grad():
        z = mods[0].eval()(x)
        y = mods[1].eval()(z)
    torch.save({"x": x, "z": z, "y": y}, str(tmp_path / "io.pt"))
    code = (f"import sys, torch\nsys.path.insert(0, {ref!r})\n"
            f"m = torch.load({str(tmp_path / 'final.pt')!r}, weights_only=False)\n"
            f"io = torch.load({str(tmp_path / 'io.pt')!r})\n"
            "assert 'rankaae_b200' not in sys.modules\n"
            "assert type(m['Encoder']).__module__ == 'sc.clustering.model' and not m['Encoder'].training\n"
            "assert m['Decoder'].nstyle == 6\n"
            "with torch.no_grad():\n    z = m['Encoder'](io['x']); y = m['Decoder'](z)\n"
            "assert torch.equal(z, io['z']) and torch.equal(y, io['y'])\nprint('ok')\n")
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, cwd=str(tmp_path), timeout=300)
    assert r.returncode == 0 and "ok" in r.stdout, r.stderr[-2000:]


def test_oracle_plateau_scheduler_matches_torch():
    """The oracle's ReduceLROnPlateau restatement (and with it the in-kernel one, checked on the GPU) against torch's on
    scripted sequences: plateaus longer than the patience, negative metrics (relative threshold quirk), the 1e-8 rule."""
    from oracle import aae_oracle as O
    rng = np.random.default_rng(0)
    seqs = [np.concatenate([np.linspace(1.0, 0.5, 12), np.full(9, 0.499), np.linspace(0.45, 0.2, 6), np.full(14, 0.21)]),
            np.concatenate([np.linspace(-0.1, -0.6, 15), np.full(8, -0.6), -0.6 - 0.001 * np.arange(10)]),
            0.3 + 0.05 * rng.standard_normal(80)]
    for seq in seqs:
        for lr0, factor, patience in ((1e-2, 0.1, 3), (1e-3, 0.5, 5), (3e-8, 0.1, 2), (1e-3, 0.1, 100)):
            par = torch.nn.Parameter(torch.zeros(1))
            opt = torch.optim.AdamW([par], lr=lr0)
            sch = torch.optim.lr_scheduler.ReduceLROnPlateau(opt, mode="min", factor=factor, patience=patience, cooldown=0,
                                                             threshold=0.01)
            orc = O.ReduceLROnPlateau(lr0, factor=factor, patience=patience)
            for v in seq:
                sch.step(float(v))
                orc.step(float(v))
                assert orc.lr == pytest.approx(opt.param_groups[0]["lr"], rel=1e-12)
                assert orc.num_bad_epochs == sch.num_bad_epochs and (orc.best == sch.best or (np.isinf(orc.best) and np.isinf(sch.best)))


def test_reference_runner_times_the_unmodified_trainer(tmp_path):
    """bench.py's reference arm: baseline/ref_runner.py drives the reference's own Trainer.from_data(...).train() (from
    baseline/_ref) on a small CSV, on the CPU, with the reference's callback hook delimiting the epochs."""
    sys.path.insert(0, os.path.join(ROOT, "baseline"))
    import ref_runner
    if not ref_runner.available():
        pytest.skip("baseline/_ref not installed (build() installs it where /root/reference exists)")
    import bench
    from rankaae_b200.synthetic import synthetic_dataset, write_csv
    spec, aux = synthetic_dataset(400, 5, 256, seed=3)
    csv = str(tmp_path / "data.csv")
    write_csv(csv, spec, aux)
    cfg = dict(bench.EXAMPLE, **bench.REFERENCE_ONLY_KEYS, batch_size=128)
    res = ref_runner.run(csv, cfg, n_procs=2, n_epochs=2, threads=1, anomaly=True, timeout=600)
    assert len(res) == 2 and all(len(r["epoch_s"]) == 2 and all(t > 0 for t in r["epoch_s"]) for r in res)
    assert all(len(r["metrics"]) == 5 and np.isfinite(r["metrics"]).all() for r in res)
