"""bench.py's algorithmic work model against SURVEY.md section 8(d): FLOPs and bytes per timestep-KKT, and the bench
line's contract keys (parsed from a committed line: no GPU needed)."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402

DIMS = {"cartpole": (4, 21, 14), "acrobot": (4, 9, 6), "concar": (4, 10, 4), "pushing": (4, 11, 6), "double_integrator": (2, 3, 1)}


def test_flops_per_timestep_kkt_match_the_survey():
    want = {"cartpole": 35961, "acrobot": 6290, "concar": 6015, "pushing": 8187, "double_integrator": 418}
    for wl, (n, m, p) in DIMS.items():
        assert round(bench.kkt_flops(n, m, p)) == want[wl], wl


def test_bytes_per_timestep_kkt_match_the_survey():
    want = {"cartpole": 16832, "acrobot": 5184, "concar": 5472, "pushing": 6464, "double_integrator": 896}
    for wl, (n, m, p) in DIMS.items():
        assert bench.kkt_bytes_dense(n, m, p) == want[wl], wl
    assert bench.kkt_bytes_compact(4, 21, 14, 48) == 4760          # the compact-tile layout, cartpole


def test_committed_bench_line_carries_the_contract_keys():
    path = os.path.join(ROOT, "profiles", "BENCH_r2_final.json")
    assert os.path.exists(path), "no round-2 bench line committed under profiles/"
    d = json.loads(open(path).read().strip().splitlines()[-1])
    for k in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
              "vs_baseline", "dtype", "data", "config", "e2e", "gpu_launches", "clocks", "roofline", "cpu_baseline"):
        assert k in d, k
    assert d["dtype"] == "f64" and d["unit"] == "solves/s" and "workload" in d["config"]
    for k in ("bound", "achieved", "peak", "unit", "frac", "traffic"):
        assert k in d["roofline"], k
    assert abs(d["roofline"]["frac"] - d["roofline"]["achieved"] / d["roofline"]["peak"]) < 1e-12
    for k in ("value", "unit", "h2d_bytes_per_step", "d2h_bytes_per_step"):
        assert k in d["e2e"], k
    assert d["e2e"]["h2d_bytes_per_step"] > 0 and d["e2e"]["d2h_bytes_per_step"] > 0 and d["gpu_launches"] > 0
    assert d["parity_checked"] >= 256 and d["parity_mismatches"] == 0
    assert {c["name"] for c in d["configs"]} == {"config3", "config4a", "config4b", "config5"}
    ref = json.loads(open(os.path.join(ROOT, "profiles", "BENCH_r2_final_reference.json")).read().strip().splitlines()[-1])
    assert ref["impl"] == "reference" and ref["metric"] == d["metric"] and ref["unit"] == d["unit"]
    assert ref["e2e"]["h2d_bytes_per_step"] == 0 and ref["cpu_baseline"]["kind"] == "port"
