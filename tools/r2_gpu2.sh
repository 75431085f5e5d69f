set -x
mkdir -p gpurun_out
summ() { python - "$1" <<PY
import json,sys
tot=0; fw=0; r0=r16=r40=None; dg=None
for l in open(sys.argv[1]):
    try: d=json.loads(l)
    except Exception: continue
    if "round" in d:
        tot+=d["bw_ms"]; fw+=d["fw_ms"]
        if d["round"]==0: r0=d["Mkkt_per_s"]
        if d["round"]==16: r16=d["Mkkt_per_s"]
        if d["round"]==40: r40=d["Mkkt_per_s"]
    if "digest" in d: dg=d["digest"]
print("series",sys.argv[1],"sum70",round(tot,1),"fw70",round(fw,1),"r0",r0,"r16",r16,"r40",r40,"digest",dg)
PY
}
IPDDP_SERIES=1 python tools/phase_bench.py cartpole 16384 70 > gpurun_out/r2_series_sorted.log 2>&1; summ gpurun_out/r2_series_sorted.log
IPDDP_TUNE=list_sort=0 IPDDP_SERIES=1 python tools/phase_bench.py cartpole 16384 70 > gpurun_out/r2_series_unsorted.log 2>&1; summ gpurun_out/r2_series_unsorted.log
python tools/queue_bench.py cartpole 131072 16384,32768 > gpurun_out/r2_queue_bench_2.log 2>&1
IPDDP_TUNE=list_sort=0 python tools/queue_bench.py cartpole 131072 16384 >> gpurun_out/r2_queue_bench_2.log 2>&1
cat gpurun_out/r2_queue_bench_2.log
timeout 900 python bench.py --steps 8 --warmup 1 > gpurun_out/r2_bench_2.json 2> gpurun_out/r2_bench_2.err; echo "bench rc=$?"
tail -c 6000 gpurun_out/r2_bench_2.json; tail -5 gpurun_out/r2_bench_2.err
