"""Turns the raw outputs of tools/profile_round.sh (gpurun_out/*_<tag>.*) into the tracked artefacts under profiles/:
condensed launch list, raw ncu metric rows, DRAM traffic per timestep-KKT, bench lines, per-round series.
    python tools/collect_profiles.py <tag>"""
import collections
import csv
import json
import os
import shutil
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag = sys.argv[1]
G = os.path.join(ROOT, "gpurun_out")
P = os.path.join(ROOT, "profiles")

rows = list(csv.reader(open(os.path.join(G, f"launches_{tag}.csv"))))
hi = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
hdr = rows[hi]
idx = {h: i for i, h in enumerate(hdr)}
out = [["id", "kernel", "grid_x", "block_x", "duration_us"]]
tot, cnt = collections.Counter(), collections.Counter()
for r in rows[hi + 1:]:
    if len(r) < len(hdr):
        continue
    name = r[idx["Kernel Name"]].replace("void ", "").split("(")[0]
    v = float(r[idx["Metric Value"]].replace(",", ""))
    u = r[idx["Metric Unit"]]
    us = v / 1e3 if u.startswith("ns") else (v if u.startswith("us") else v * 1e3)
    out.append([r[idx["ID"]], name, r[idx["Grid Size"]].strip("()").split(",")[0], r[idx["Block Size"]].strip("()").split(",")[0], f"{us:.2f}"])
    k = name.split("<")[0]
    tot[k] += us / 1e3
    cnt[k] += 1
csv.writer(open(os.path.join(P, "r1_final_launches.csv"), "w")).writerows(out)
T = sum(tot.values())
print(f"launch list: {len(out) - 1} launches, {T:.1f} ms")
for k, v in tot.most_common():
    print(f"| `{k}` | {cnt[k]} | {v:.0f} | {100 * v / T:.1f} % |")

keys, outrows, vals = None, [], {}
for r in ("bw_bulk", "bw_mid", "fw_bulk", "bw_lone", "derivs"):
    rep = os.path.join(G, f"{r}_{tag}.ncu-rep")
    txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rr = list(csv.reader(txt.splitlines()))
    if keys is None:
        keys = rr[0]
        outrows += [["capture"] + rr[0], ["unit"] + rr[1]]
    outrows.append([f"{r}_{tag}"] + rr[2])
    vals[r] = dict(zip(rr[0], rr[2]))
csv.writer(open(os.path.join(P, "r1_final_ncu_raw.csv"), "w")).writerows(outrows)
show = ["gpu__time_duration.sum", "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.per_cycle_active", "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic",
        "launch__occupancy_limit_shared_mem", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active"]
for r, v in vals.items():
    print(r, {k: v[k] for k in show if k in v})
b = vals["bw_bulk"]
kkt = 16384 * 101
traffic = {"cartpole": {
    "dram_bytes_per_kkt_step": (float(b["dram__bytes_read.sum"]) + float(b["dram__bytes_write.sum"])) * 1e9 / kkt,
    "kkt_steps_in_launch": kkt,
    "source": f"profiles/r1_final_ncu_raw.csv row bw_bulk_{tag} (ncu --set full --clock-control none, B=16384, first k_backward "
              "launch of a solve = 16384 x 101 timestep-KKTs, one sweep each; dram__bytes_read.sum + dram__bytes_write.sum)"}}
json.dump(traffic, open(os.path.join(P, "ncu_traffic.json"), "w"), indent=1)
print(traffic)
for src, dst in ((f"bench_{tag}.json", "BENCH_r1_final.json"), (f"bench_ref_{tag}.json", "BENCH_r1_final_reference.json"),
                 (f"series_{tag}.log", "r1_final_round_series.jsonl"), (f"configs_{tag}.log", "r1_final_other_configs.jsonl")):
    shutil.copy(os.path.join(G, src), os.path.join(P, dst))
