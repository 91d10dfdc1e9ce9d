"""Installs the UNMODIFIED reference into baseline/_ref (git-ignored; travels to the GPU box with gpurun) for bench.py's
reference arm.  Runs only where /root/reference exists (the build container); on the GPU box the prebuilt tree is used.

Recipe (recorded in DESIGN.md §8): `pip install --no-index --no-build-isolation --no-deps --find-links /opt/wheelhouse
--target baseline/_ref <copy of /root/reference under /tmp>` (the source tree is read-only and dependency resolution
fails offline: torch_optimizer, ipyparallel, seaborn ... are not in the wheelhouse).  The reference's setup.py lists
`packages=['sc']` only, so pip installs `sc/__init__.py` without the sub-packages; the install is completed by copying the
reference's own `sc/{clustering,utils,cmd,report}` modules and `example/fix_config.yaml` next to it, byte for byte.  Nothing of it
is tracked by git and nothing in the product imports it.
"""
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.environ.get("RANKAAE_REFERENCE", "/root/reference")
DST = os.path.join(HERE, "_ref")
NEEDED = ["sc/__init__.py", "sc/clustering/__init__.py", "sc/clustering/trainer.py", "sc/clustering/model.py",
          "sc/clustering/dataloader.py", "sc/utils/__init__.py", "sc/utils/functions.py", "sc/utils/parameter.py",
          "sc/utils/logger.py", "sc/report/analysis.py", "example/fix_config.yaml"]


def installed():
    return all(os.path.exists(os.path.join(DST, f)) for f in NEEDED)


def install(force=False):
    if installed() and not force:
        return DST
    if not os.path.isdir(os.path.join(REF, "sc", "clustering")):
        raise RuntimeError(f"reference tree not found at {REF} and baseline/_ref is incomplete")
    tmp = "/tmp/rankaae_ref_copy"
    shutil.rmtree(tmp, ignore_errors=True)
    shutil.copytree(REF, tmp)
    subprocess.run([sys.executable, "-m", "pip", "install", "-q", "--no-index", "--no-build-isolation", "--no-deps",
                    "--find-links", "/opt/wheelhouse", "--upgrade", "--target", DST, tmp], check=False)
    for sub in ("clustering", "utils", "cmd", "report"):
        src = os.path.join(REF, "sc", sub)
        dst = os.path.join(DST, "sc", sub)
        os.makedirs(dst, exist_ok=True)
        for f in os.listdir(src):
            if f.endswith(".py") or f.endswith(".sh") or f.endswith(".yaml"):
                shutil.copy2(os.path.join(src, f), os.path.join(dst, f))
    os.makedirs(os.path.join(DST, "sc"), exist_ok=True)
    if not os.path.exists(os.path.join(DST, "sc", "__init__.py")):
        shutil.copy2(os.path.join(REF, "sc", "__init__.py"), os.path.join(DST, "sc", "__init__.py"))
    os.makedirs(os.path.join(DST, "example"), exist_ok=True)
    shutil.copy2(os.path.join(REF, "example", "fix_config.yaml"), os.path.join(DST, "example", "fix_config.yaml"))
    if not installed():
        raise RuntimeError("baseline/_ref is incomplete after the install")
    return DST


if __name__ == "__main__":
    print(install(force="--force" in sys.argv))
