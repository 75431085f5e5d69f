#!/bin/bash
# A/B of kernel variants built with build.build_variant(name, flags): k_backward time over the first 70 rounds of one
# cartpole batch (B = 16 384), plus a digest of the state after those rounds (variants must agree bit for bit).
#   gpurun --timeout 150 -- 'bash tools/ab_variants.sh base oneloop lean22'
mkdir -p gpurun_out
for v in "$@"; do
IPDDP_SERIES=1 python tools/phase_bench.py cartpole 16384 70 interiorpointddp.jl_b200/libipddp_b200_$v.so > gpurun_out/ab_$v.log 2>&1
python - "$v" <<PY
import json,sys
tot=0; fw=0; r0=r16=r40=None; dg=None
for l in open("gpurun_out/ab_"+sys.argv[1]+".log"):
    try: d=json.loads(l)
    except Exception: continue
    if "round" in d:
        tot+=d["bw_ms"]; fw+=d["fw_ms"]
        if d["round"]==0: r0=d["Mkkt_per_s"]
        if d["round"]==16: r16=d["Mkkt_per_s"]
        if d["round"]==40: r40=d["Mkkt_per_s"]
    if "digest" in d: dg=d["digest"]
print("variant",sys.argv[1],"sum70",round(tot,1),"fw70",round(fw,1),"r0",r0,"r16",r16,"r40",r40,"digest",dg)
PY
done
