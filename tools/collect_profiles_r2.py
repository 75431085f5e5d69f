"""Tracked artefacts of a round-2 ncu session: raw metric rows of the captures under gpurun_out/ (profiles/r2_ncu_raw.csv),
the DRAM traffic per timestep-KKT of the first k_backward launch (profiles/ncu_traffic.json, read by bench.py), and the
per-region instruction attribution of the backward captures (profiles/r2_bw_regions.txt).
    python tools/collect_profiles_r2.py"""
import csv
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
G = os.path.join(ROOT, "gpurun_out")
P = os.path.join(ROOT, "profiles")
caps = [("r2_bw_bulk", "first k_backward launch of a solve: 16384 x 101 timestep-KKTs, one sweep each"),
        ("r2_bw_mid", "37th k_backward launch (mid-solve: 2x2 pivots, restarted sweeps)"),
        ("r2_fw_bulk", "second k_forward launch (16384 instances, TMA-staged rollout)"),
        ("r2_derivs", "second k_derivs launch"),
        ("r2_init", "k_init: 16384 instances, one warp each (rollout, record writes, merit terms)")]
outrows, vals = [], {}
for name, what in caps:
    rep = os.path.join(G, name + ".ncu-rep")
    if not os.path.exists(rep):
        continue
    txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rr = list(csv.reader(txt.splitlines()))
    if not outrows:
        outrows += [["capture", "what"] + rr[0], ["unit", ""] + rr[1]]
    outrows.append([name, what] + rr[2])
    vals[name] = dict(zip(rr[0], rr[2]))
csv.writer(open(os.path.join(P, "r2_ncu_raw.csv"), "w")).writerows(outrows)
b = vals["r2_bw_bulk"]
kkt = 16384 * 101
traffic = {"cartpole": {
    "dram_bytes_per_kkt_step": (float(b["dram__bytes_read.sum"]) + float(b["dram__bytes_write.sum"])) * 1e9 / kkt,
    "kkt_steps_in_launch": kkt,
    "source": "profiles/r2_ncu_raw.csv row r2_bw_bulk (ncu --set full --clock-control none, B=16384, first k_backward launch of a "
              "solve = 16384 x 101 timestep-KKTs, one sweep each; dram__bytes_read.sum + dram__bytes_write.sum)"}}
json.dump(traffic, open(os.path.join(P, "ncu_traffic.json"), "w"), indent=1)
with open(os.path.join(P, "r2_bw_regions.txt"), "w") as fh:
    for name in ("r2_bw_bulk", "r2_bw_mid"):
        fh.write(f"== {name}: executed warp-instructions of k_backward<Model_cartpole> by source region and opcode\n")
        fh.write(subprocess.run([sys.executable, os.path.join(ROOT, "tools", "ncu_regions.py"), os.path.join(G, name + ".ncu-rep"),
                                 "cartpole", "k_backward"], capture_output=True, text=True).stdout + "\n")
show = ["gpu__time_duration.sum", "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active"]
for r, v in vals.items():
    print(r, {k: v[k] for k in show if k in v})
print(traffic)
