#!/bin/bash
# usage: tools/ncu_report.sh <report.ncu-rep> <model> <kernel mangled substring> [top]
# prints key raw metrics and the per-source-line instruction / stall attribution of one kernel
REP=$1; MODEL=$2; KERN=$3; TOP=${4:-40}
TMP=$(mktemp -d)
ncu -i $REP --page source --csv 2>/dev/null > $TMP/src.csv
ncu -i $REP --page raw --csv 2>/dev/null > $TMP/raw.csv
(cd $TMP && cuobjdump -xelf $MODEL /root/repo/interiorpointddp.jl_b200/libipddp_b200.so >/dev/null 2>&1 && nvdisasm -g $MODEL.sm_100a.cubin > all.sass 2>/dev/null)
python - $TMP/raw.csv <<'PY'
import csv,sys
rows=list(csv.reader(open(sys.argv[1]))); hdr=rows[0]; vals=rows[2]
for k in ['gpu__time_duration.sum','smsp__inst_executed.sum','smsp__issue_active.avg.pct_of_peak_sustained_active','launch__grid_size','launch__registers_per_thread','launch__occupancy_limit_shared_mem','launch__occupancy_limit_registers','sm__warps_active.avg.pct_of_peak_sustained_active','launch__shared_mem_per_block_dynamic','sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active','sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active','dram__bytes_read.sum','dram__bytes_write.sum','l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum','dram__throughput.avg.pct_of_peak_sustained_elapsed']:
    if k in hdr: print(f"{k:70s} {vals[hdr.index(k)]} {rows[1][hdr.index(k)]}")
PY
python /root/repo/tools/sass_by_line.py $TMP/src.csv $TMP/all.sass $KERN $TOP
rm -rf $TMP
