for v in _nofast2 ""; do
IPDDP_SERIES=1 python tools/phase_bench.py cartpole 16384 70 interiorpointddp.jl_b200/libipddp_b200$v.so > gpurun_out/ab$v.log 2>&1
python - "$v" <<PY
import json,sys
tot=0; r0=r16=r40=None
for l in open("gpurun_out/ab"+sys.argv[1]+".log"):
    try: d=json.loads(l)
    except Exception: continue
    if "round" in d:
        tot+=d["bw_ms"]
        if d["round"]==0: r0=d["Mkkt_per_s"]
        if d["round"]==16: r16=d["Mkkt_per_s"]
        if d["round"]==40: r40=d["Mkkt_per_s"]
print("variant",sys.argv[1] or "default","sum70",round(tot,1),"r0",r0,"r16",r16,"r40",r40)
PY
done
