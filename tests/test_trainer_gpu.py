"""GPU tests of the reference-facing surface: Trainer.from_data / Trainer.train artifacts and the trial ensemble
that replaces the ipyparallel farm."""
import os

import numpy as np
import pytest
import yaml

from oracle import aae_oracle as O
from oracle import ref_shim
from tests.test_parity_gpu import EXAMPLE

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def workdir(tmp_path_factory):
    d = tmp_path_factory.mktemp("work")
    cfg = dict(EXAMPLE, batch_size=128, max_epoch=21, trials=3, timeout=1, verbose=True, data_file="data.csv")
    spec, aux = O.synthetic_dataset(600, O.Config.from_dict(cfg), seed=4, dtype=np.float32)
    ref_shim.write_csv(str(d / "data.csv"), spec, aux)
    (d / "fix_config.yaml").write_text(yaml.safe_dump(cfg))
    return d, cfg


def test_trainer_from_data_train_artifacts(workdir):
    import torch
    from rankaae_b200.logger import create_logger
    from rankaae_b200.parameter import Parameters
    from rankaae_b200.trainer import Trainer
    d, cfg = workdir
    job = d / "single"
    job.mkdir()
    p = Parameters.from_yaml(str(d / "fix_config.yaml"))
    tr = Trainer.from_data(str(d / "data.csv"), igpu=0, verbose=False, work_dir=str(job), config_parameters=p,
                           logger=create_logger("t_msg", str(job / "messages.txt")),
                           loss_logger=create_logger("t_loss", str(job / "losses.csv"), simple_fmt=True), seed=3)
    seen = []
    metrics = tr.train(callback=lambda e, m: seen.append((e, list(m))))
    assert len(metrics) == 5 and all(np.isfinite(metrics))
    assert [e for e, _ in seen] == list(range(21)) and seen[-1][1] == metrics
    # losses.csv: header + rows at epochs 0, 10, 20 in the reference's format (trainer.py:84-87, 270-279)
    lines = (job / "losses.csv").read_text().rstrip("\n").split("\n")
    assert lines[0].startswith("Epoch,Train_D,Val_D,Train_G,Val_G,Train_Aux,Val_Aux,Train_Recon,")
    assert [int(l.split(",")[0]) for l in lines[1:]] == [0, 10, 20]
    assert all(len(l.split(",\t")) == 14 and l.endswith(",\t") for l in lines[1:])   # 13 fields + trailing ",\t"
    # final.pt: dict of three nn.Modules with BN buffers advanced: 6 encoder / 4 decoder forwards per batch
    fin = torch.load(str(job / "final.pt"), weights_only=False)
    assert set(fin) == {"Encoder", "Decoder", "Style Discriminator"}
    n_batches = -(-420 // 128)
    assert int(fin["Encoder"].main[2].num_batches_tracked) == 21 * n_batches * 6
    assert int(fin["Decoder"].main[2].num_batches_tracked) == 21 * n_batches * 4
    # the pickled modules reproduce the kernel's validation latents in eval mode (report tooling path)
    spec = np.loadtxt(str(d / "data.csv"), delimiter=",", skiprows=2, usecols=range(7, 263), dtype=np.float32)[420:510]
    enc = fin["Encoder"].eval()
    with torch.no_grad():
        z_t = enc(torch.from_numpy(spec)).numpy()
    out = tr.engine.validate(0, epoch=20)
    np.testing.assert_allclose(out["z"], z_t, rtol=0, atol=2e-4)
    assert fin["Decoder"].nstyle == 6


def test_ensemble_writes_reference_layout(workdir):
    import torch
    from rankaae_b200.ensemble import run_ensemble
    from rankaae_b200.parameter import Parameters
    d, cfg = workdir
    p = Parameters.from_yaml(str(d / "fix_config.yaml"))
    res = run_ensemble(str(d), p, str(d / "data.csv"), trials=3, verbose=True, device="cuda:0", epochs_per_call=8)
    assert len(res) == 3
    for n in (1, 2, 3):
        job = d / "training" / f"job_{n}"
        assert (job / "final.pt").exists() and (job / "losses.csv").exists() and (job / "messages.txt").exists()
        assert "Training finished. Time used:" in (job / "messages.txt").read_text()
        assert len((job / "losses.csv").read_text().strip().split("\n")) == 4
    m = np.array([r[0] for r in res])
    assert np.isfinite(m).all() and len(np.unique(np.round(m[:, 1], 7))) == 3     # independent trials
    a = torch.load(str(d / "training" / "job_1" / "final.pt"), weights_only=False)
    b = torch.load(str(d / "training" / "job_2" / "final.pt"), weights_only=False)
    assert not torch.equal(a["Encoder"].main[0].weight, b["Encoder"].main[0].weight)


def test_cli_two_gpus_when_available(workdir):
    """`torchrun --nproc-per-node 2 -m rankaae_b200.cmd.train_sc`: trials sharded over 2 GPUs, NCCL gather."""
    import subprocess
    import sys
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    d, cfg = workdir
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
           "--master-port", "29533", "-m", "rankaae_b200.cmd.train_sc", "-c", "fix_config.yaml", "-w", str(d)]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600, cwd=root, env=dict(os.environ, PYTHONPATH=root))
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    log = (d / "main_process_message.txt").read_text()
    assert "Running with 2 process(es)." in log and "for 3 trails" in log
