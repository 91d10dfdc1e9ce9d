#!/usr/bin/env python
"""Copies / derives the round's evidence from gpurun_out/ into profiles/ after a final GPU run:
  prof_r01_final.ncu-rep -> ncu_train_r01_traffic.json, ncu_train_r01_raw_t148.txt, ncu_train_r01_lines.txt
  bench_r01.json (+ roofline.traffic from the capture), bench_n2.json, launches_r01.csv, stage_profile_*.txt,
  e2e_band.json, parity_*.json -> parity_r01.md
usage: python tools/refresh_profiles.py [round_tag]"""
import csv
import io
import json
import os
import shutil
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
G, P = os.path.join(ROOT, "gpurun_out"), os.path.join(ROOT, "profiles")
tag = sys.argv[1] if len(sys.argv) > 1 else "r01"
rep = os.path.join(G, f"prof_{tag}_final.ncu-rep")
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, vals = rows[0], rows[1], rows[2]
d = {h: vals[i] for i, h in enumerate(hdr)}
g = lambda k: float(d[k].replace(",", ""))
traffic = {"kernel": "raae_train_kernel",
           "launch": "5 train batches x 148 trials (bench.py --steps 1 --warmup 3 --trials 148 --no-cpu --no-single --no-configs, 4th launch)",
           "duration_ms": g("gpu__time_duration.sum"), "dram_bytes_read": g("dram__bytes_read.sum") * 1e9,
           "dram_bytes_write": g("dram__bytes_write.sum") * 1e9,
           "dram_pct_of_peak": g("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"),
           "dram_gbs": (g("dram__bytes_read.sum") + g("dram__bytes_write.sum")) * 1e9 / (g("gpu__time_duration.sum") * 1e-3) / 1e9,
           "tensor_pipe_pct_active": g("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"),
           "issue_active_pct": g("smsp__issue_active.avg.pct_of_peak_sustained_active"),
           "warps_active_pct": g("sm__warps_active.avg.pct_of_peak_sustained_active"),
           "registers_per_thread": g("launch__registers_per_thread"), "l2_hit_pct": g("lts__t_sector_hit_rate.pct"),
           "inst_per_cycle_per_sm": g("sm__inst_executed.sum.per_cycle_elapsed") / 148}
json.dump(traffic, open(os.path.join(P, f"ncu_train_{tag}_traffic.json"), "w"), indent=1)
keys = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__mem_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_tensor_subpipe_hmma.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size",
        "launch__block_size", "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__inst_executed.sum.per_cycle_elapsed",
        "lts__t_sector_hit_rate.pct", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
        "sm__cycles_elapsed.avg", "sm__cycles_elapsed.avg.per_second"]
with open(os.path.join(P, f"ncu_train_{tag}_raw_t148.txt"), "w") as f:
    f.write(f"# ncu -i prof_{tag}_final.ncu-rep --page raw --csv (selected metrics); raae_train_kernel, 148 trials, 5 train batches\n")
    for i, h in enumerate(hdr):
        if h in keys:
            f.write(f"{h:95s} {units[i]:16s} {vals[i]}\n")
with open(os.path.join(P, f"ncu_train_{tag}_lines.txt"), "w") as f:
    f.write(subprocess.run([sys.executable, os.path.join(ROOT, "tools", "ncu_lines.py"), rep, "40"], capture_output=True, text=True).stdout)
b = json.load(open(os.path.join(G, f"bench_{tag}.json")))
b["roofline"]["traffic"] = traffic["dram_bytes_read"] + traffic["dram_bytes_write"]
b["roofline"].update(tensor_pipe_pct=traffic["tensor_pipe_pct_active"], dram_gbs=traffic["dram_gbs"],
                     dram_pct_of_peak=traffic["dram_pct_of_peak"], occupancy_warps_active_pct=traffic["warps_active_pct"],
                     issue_active_pct=traffic["issue_active_pct"], ncu_capture=f"profiles/ncu_train_{tag}_traffic.json")
json.dump(b, open(os.path.join(P, f"bench_{tag}.json"), "w"))
for src, dst in (("stage_profile_t148.txt", f"stage_profile_{tag}_t148.txt"), ("stage_profile_t1.txt", f"stage_profile_{tag}_t1.txt"),
                 ("stage_profile_t1_c8.txt", f"stage_profile_{tag}_t1_c8.txt"),
                 (f"launches_{tag}.csv", f"launches_{tag}.csv"), ("e2e_band.json", f"e2e_band_{tag}.json"),
                 ("e2e_band_long.json", f"e2e_band_long_{tag}.json"), (f"bench_ref_{tag}.json", f"bench_ref_{tag}.json"),
                 (f"bench_{tag}_n2.json", f"bench_{tag}_n2.json"), (f"bench_{tag}_n4.json", f"bench_{tag}_n4.json"),
                 (f"bench_{tag}_n8.json", f"bench_{tag}_n8.json")):
    if os.path.exists(os.path.join(G, src)):
        shutil.copy(os.path.join(G, src), os.path.join(P, dst))
crep = os.path.join(G, f"prof_{tag}_cluster8.ncu-rep")
if os.path.exists(crep):                      # the 8-CTA-cluster build, one trial
    craw = list(csv.reader(io.StringIO(subprocess.run(["ncu", "-i", crep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout)))
    with open(os.path.join(P, f"ncu_train_{tag}_raw_cluster8.txt"), "w") as f:
        f.write(f"# ncu -i prof_{tag}_cluster8.ncu-rep --page raw --csv (selected metrics); raae_cn::raae_train_kernel, 1 trial as an 8-CTA cluster, 5 train batches\n")
        for i, h in enumerate(craw[0]):
            if h in keys or h in ("launch__cluster_dim_x", "launch__cluster_scheduling_policy", "launch__cluster_max_active"):
                f.write(f"{h:95s} {craw[1][i]:16s} {craw[2][i]}\n")
    with open(os.path.join(P, f"ncu_train_{tag}_lines_cluster8.txt"), "w") as f:
        f.write(subprocess.run([sys.executable, os.path.join(ROOT, "tools", "ncu_lines.py"), crep, "30"], capture_output=True, text=True).stdout)
with open(os.path.join(P, f"sass_{tag}.txt"), "w") as f:
    f.write(subprocess.run([sys.executable, os.path.join(ROOT, "tools", "sass_summary.py")], capture_output=True, text=True).stdout)
with open(os.path.join(P, f"parity_{tag}.md"), "w") as f:
    f.write(subprocess.run([sys.executable, os.path.join(ROOT, "tools", "parity_table.py"), G], capture_output=True, text=True).stdout)
print(json.dumps({"ms_per_step": b["ms_per_step"], "value": b["value"], "e2e": b["e2e"]["value"], "frac": b["roofline"]["frac"],
                  "achieved": b["roofline"]["achieved"], "launch_ms": b["roofline"]["launch_ms"], "single": b.get("single_trial", {}).get("ms_per_epoch"),
                  "cpu": b.get("cpu_baseline", {}).get("value"), "tph": b["trials_per_hour_2000_epochs"], **{k: traffic[k] for k in
                  ("duration_ms", "dram_pct_of_peak", "tensor_pipe_pct_active", "issue_active_pct", "l2_hit_pct", "inst_per_cycle_per_sm")}}, indent=1))
