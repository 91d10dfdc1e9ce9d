"""Import shims for running the UNMODIFIED reference (`/root/reference`, Python) in this
container (TEST INFRASTRUCTURE ONLY — see oracle/aae_oracle.py header).

Three obstacles, none of them in the arithmetic:
  1. `seaborn`, `matplotlib(.pyplot)`, `torch_optimizer`, `ipyparallel` are not installed
     (trainer.py:5,7; parameter.py:11; train_sc.py:12) -> empty stub modules;
  2. torch >= 2.4 removed `ReduceLROnPlateau(verbose=...)` (trainer.py:403-406) -> the name
     `sc.clustering.trainer.ReduceLROnPlateau` is rebound to a wrapper dropping that kwarg;
  3. the dataset CSV is a missing blob -> `write_csv` synthesises one in the schema asserted
     at dataloader.py:12-25.
Nothing under /root/reference is copied or modified.  This module is only usable where
/root/reference exists (the build container); nothing that runs on the GPU box imports it.
"""
import os
import sys
import types

REFERENCE_ROOT = os.environ.get("RANKAAE_REFERENCE", "/root/reference")


def reference_available():
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "sc", "clustering"))


def import_reference():
    """Returns the reference's `sc.clustering.trainer` module with the shims applied."""
    if not reference_available():
        raise RuntimeError(f"reference tree not found at {REFERENCE_ROOT}")
    for name in ("seaborn", "matplotlib", "matplotlib.pyplot", "ipyparallel"):
        if name not in sys.modules:
            sys.modules[name] = types.ModuleType(name)
    sys.modules["matplotlib"].pyplot = sys.modules["matplotlib.pyplot"]
    if "torch_optimizer" not in sys.modules:
        m = types.ModuleType("torch_optimizer")
        m.AdaBound = None
        m.RAdam = None
        sys.modules["torch_optimizer"] = m
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    import torch.optim.lr_scheduler as lrs
    import sc.clustering.trainer as ref_trainer

    if not getattr(ref_trainer, "_graft_shimmed", False):
        base = lrs.ReduceLROnPlateau

        def _plateau(optimizer, *args, verbose=None, **kw):
            return base(optimizer, *args, **kw)

        ref_trainer.ReduceLROnPlateau = _plateau
        ref_trainer._graft_shimmed = True
    return ref_trainer


def write_csv(path, spec, aux, grid=None):
    """Writes spectra + descriptors in the reference CSV schema: two index columns, AUX_* x
    n_aux, ENE_<energy> x dim (dataloader.py:12-25).  Values are printed with repr(float(v)),
    which round-trips the double exactly (so float32 inputs are float32-representable doubles
    after pandas parses them)."""
    import numpy as np

    n, dim = spec.shape
    n_aux = aux.shape[1]
    if grid is None:
        grid = np.linspace(5460.0, 5520.0, dim)
    names = ["AUX_CT", "AUX_CN", "AUX_OCN", "AUX_RSTD", "AUX_MOOD", "AUX_X5", "AUX_X6", "AUX_X7"]
    cols = ["mp_id", "site"] + names[:n_aux] + [f"ENE_{e:.3f}" for e in grid]
    with open(path, "w") as f:
        f.write("# synthetic spectra, reference CSV schema\n")
        f.write(",".join(cols) + "\n")
        for i in range(n):
            vals = [f"mp-{i}", "0"] + [repr(float(v)) for v in aux[i]] + [repr(float(v)) for v in spec[i]]
            f.write(",".join(vals) + "\n")
