#!/usr/bin/env python
"""Times the UNMODIFIED reference trainer on host cores (bench.py's reference arm and cpu_baseline leg).

The reference itself runs here: `sc.clustering.trainer.Trainer.from_data(...).train()` imported from baseline/_ref (installed
by baseline/install_ref.py from /root/reference, byte for byte), on the synthetic CSV in the reference's schema, with the
reference's own deployment settings (sc/cmd/run_training.sh:3-9: one thread per engine process; sc/cmd/train_sc.py:68-70) -
N independent single-thread trials in N processes, which is how the reference fills a host - or one trial with all threads.

Nothing of the product (rankaae_b200/, the CUDA library) and nothing of oracle/ is imported.  Not touched in the reference:
the arithmetic, the loop, the loader, the optimizers.  Shims, all outside the arithmetic:
  * empty stub modules for seaborn / matplotlib(.pyplot) / torch_optimizer / ipyparallel (plot helpers and optimizers no config
    uses; not installed in this image);
  * torch >= 2.4 removed ReduceLROnPlateau(verbose=...) (trainer.py:403-406): the name is rebound to a wrapper dropping it;
  * CUDA_VISIBLE_DEVICES is emptied so that Trainer.from_data (trainer.py:430-439) takes its CPU branch on the GPU box
    (BASELINE.json configs[0]: "reference PyTorch on CPU");
  * `torch.autograd.set_detect_anomaly(True)` is switched on by the reference at import (trainer.py:11): timed as shipped
    (anomaly "on") and, separately, switched off again after the import (anomaly "off").
Epoch boundaries are taken from the reference's own `callback(epoch, metrics)` hook (trainer.py:306-307).
"""
import json
import os
import subprocess
import sys
import tempfile
import time

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.path.join(HERE, "_ref")


def available():
    return os.path.exists(os.path.join(REF, "sc", "clustering", "trainer.py"))


def _worker(csv, cfg_json, n_epochs, threads, anomaly, seed):
    import types
    for name in ("seaborn", "matplotlib", "matplotlib.pyplot", "ipyparallel"):
        sys.modules.setdefault(name, types.ModuleType(name))
    sys.modules["matplotlib"].pyplot = sys.modules["matplotlib.pyplot"]
    m = types.ModuleType("torch_optimizer")
    m.AdaBound = m.RAdam = None
    sys.modules.setdefault("torch_optimizer", m)
    sys.path.insert(0, REF)
    import logging
    import torch
    import torch.optim.lr_scheduler as lrs
    torch.set_num_threads(threads)
    import sc.clustering.trainer as T
    from sc.utils.parameter import Parameters
    base = lrs.ReduceLROnPlateau
    T.ReduceLROnPlateau = lambda optimizer, *a, verbose=None, **kw: base(optimizer, *a, **kw)
    if not anomaly:
        torch.autograd.set_detect_anomaly(False)
    cfg = json.loads(cfg_json)
    cfg["max_epoch"] = n_epochs
    torch.manual_seed(seed)
    tmp = tempfile.mkdtemp(prefix="raae_ref_")
    quiet = logging.getLogger("raae_ref_quiet")
    quiet.addHandler(logging.NullHandler())
    quiet.propagate = False
    t_load = time.perf_counter()
    trainer = T.Trainer.from_data(csv, igpu=0, verbose=False, work_dir=tmp, config_parameters=Parameters(cfg), logger=quiet,
                                  loss_logger=quiet)
    assert str(trainer.device) == "cpu"
    stamps = [time.perf_counter()]
    metrics = trainer.train(callback=lambda e, mt: stamps.append(time.perf_counter()))
    print("RAAE_REF " + json.dumps({"epoch_s": [b - a for a, b in zip(stamps[:-1], stamps[1:])],
                                    "load_s": stamps[0] - t_load, "metrics": [float(v) for v in metrics]}))


def run(csv, cfg, n_procs, n_epochs, threads=1, anomaly=True, timeout=3000):
    """n_procs concurrent processes, one trial each, `threads` torch threads each.  Returns the per-process lists of epoch
    wall times (seconds)."""
    env = dict(os.environ, CUDA_VISIBLE_DEVICES="", OMP_NUM_THREADS=str(threads), MKL_NUM_THREADS=str(threads),
               OMP_DYNAMIC="FALSE", MKL_DYNAMIC="FALSE", NUMEXPR_NUM_THREADS=str(threads))
    procs = [subprocess.Popen([sys.executable, os.path.abspath(__file__), "--worker", csv, json.dumps(cfg), str(n_epochs),
                               str(threads), str(int(anomaly)), str(1000 + i)], env=env, stdout=subprocess.PIPE,
                              stderr=subprocess.PIPE, text=True) for i in range(n_procs)]
    out = []
    for p in procs:
        so, se = p.communicate(timeout=timeout)
        line = [l for l in so.splitlines() if l.startswith("RAAE_REF ")]
        if p.returncode != 0 or not line:
            raise RuntimeError("reference worker failed:\n" + se[-2000:])
        out.append(json.loads(line[-1][9:]))
    return out


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "--worker":
        _worker(sys.argv[2], sys.argv[3], int(sys.argv[4]), int(sys.argv[5]), bool(int(sys.argv[6])), int(sys.argv[7]))
