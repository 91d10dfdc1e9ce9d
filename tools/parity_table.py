#!/usr/bin/env python
"""Builds profiles/parity_rNN.md from the gpurun_out/parity_*.json reports written by the last `pytest -m gpu` run.
usage: python tools/parity_table.py [gpurun_out] > profiles/parity_r02.md"""
import glob
import json
import os
import sys

src = sys.argv[1] if len(sys.argv) > 1 else "gpurun_out"
print("# Parity results on B200 (from gpurun_out/parity_*.json of the final `pytest -m gpu` run of the round)\n")
print("Teacher-forced, same float32-representable state and explicit draws on both sides; oracle in float64.")
print("`tc55` = default (hidden blocks, encoder input block, decoder output forward AND backward on tcgen05, 3xTF32 split with a")
print("rounded high half), `tc23` = the same with the decoder output backward as FP32 FMA, `tc7` = hidden + input blocks only,")
print("`tc0` = all-FP32-FMA path.  `f32 yardstick` = error of the SAME oracle code run")
print("in float32 against its float64 run (what the reference's own precision costs on that case); the test band is")
print("max(2e-3, 3 x yardstick) for gradients.\n")
print("| suite | case | phase | rows | abs loss err | worst per-network grad rel-L2 | f32 yardstick (worst net) | BN buffer err |")
print("|---|---|---|---|---|---|---|---|")
worst = {}
for f in sorted(glob.glob(os.path.join(src, "parity_*.json"))):
    suite = os.path.basename(f)[len("parity_"):-len(".json")]
    for r in json.load(open(f)):
        if "grad_rel_l2" not in r:
            if r.get("tag") == "adam":
                worst["adam"] = max(worst.get("adam", 0.0), r["worst_rel"])
            continue
        g = max(r["grad_rel_l2"].values()) if r["grad_rel_l2"] else 0.0
        ys = r.get("f32_yardstick", {})
        ysg = max([v for k, v in ys.items() if k != "loss"], default=None)
        worst[suite.split("_")[0]] = max(worst.get(suite.split("_")[0], 0.0), g)
        print(f"| {suite} | {r['tag']} | {r['phase']} | {r['rows']} | {abs(r['loss_cuda'] - r['loss_oracle']):.1e} | {g:.1e} | "
              f"{'-' if ysg is None else f'{ysg:.1e}'} | {r['bn_buffer_err']:.1e} |")
print("\nWorst AdamW parameter / moment error after one update (relative): "
      f"{worst.get('adam', float('nan')):.1e} (tolerance 2e-6).")
