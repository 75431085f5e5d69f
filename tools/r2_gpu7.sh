set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2_pytest_gpu_3.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_pytest_gpu_3.log
tail -5 gpurun_out/r2_pytest_gpu_3.log
IPDDP_SERIES=1 python tools/phase_bench.py cartpole 16384 70 > gpurun_out/r2_series_chainrefactor.log 2>&1
python - <<PY
import json
tot=0; fw=0; dg=None
for l in open("gpurun_out/r2_series_chainrefactor.log"):
    try: d=json.loads(l)
    except Exception: continue
    if "round" in d: tot+=d["bw_ms"]; fw+=d["fw_ms"]
    if "digest" in d: dg=d["digest"]
print("after the stage-chain refactor: sum70",round(tot,1),"fw70",round(fw,1),"digest",dg)
PY
python tools/queue_bench.py cartpole 131072 65536 > gpurun_out/r2_queue_bench_4.log 2>&1; cat gpurun_out/r2_queue_bench_4.log
