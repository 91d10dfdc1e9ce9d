#!/usr/bin/env python
"""Re-pickles `final.pt` files written without the reference on the path (rankaae_b200.model classes) as the reference's
own `sc.clustering.model` classes, so that `sc_generate_report` (sc/report/analysis.py:115-121) loads them with only the
reference installed.  Needs both packages importable.

    PYTHONPATH=/path/to/RankAAE python tools/convert_final_pt.py work_dir/training/job_*/final.pt
"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from rankaae_b200.trainer import save_final  # noqa: E402


def convert(path):
    d = torch.load(path, map_location="cpu", weights_only=False)
    mods = (d["Encoder"], d["Decoder"], d["Style Discriminator"])
    if all(type(m).__module__.startswith("sc.") for m in mods):
        return False
    save_final(mods, path, reference_classes=True)
    return True


if __name__ == "__main__":
    for p in sys.argv[1:]:
        print(p, "converted" if convert(p) else "already reference classes")
