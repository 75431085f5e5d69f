#!/bin/bash
# The round's `ncu --set full` captures (one GPU; each capture after the same command exited 0 without ncu):
#   gpurun --timeout 1500 -- 'bash tools/ncu_captures_r2.sh'   then   python tools/collect_profiles_r2.py
# phase_bench drives single lock-step rounds of a 16 384-instance cartpole batch through the phase-level C ABI, so launch
# number s of a kernel is round s of the solve.
O=gpurun_out; mkdir -p $O
CMD="python tools/phase_bench.py cartpole 16384"
$CMD 40 > $O/r2_phase_plain.log 2>&1; echo "plain rc=$?"; tail -c 400 $O/r2_phase_plain.log
NCU="ncu --set full --clock-control none --import-source on -f"
cap() {  # name, kernel name (exact match of the base name), launches to skip, rounds to run
  SECONDS=0
  $NCU -k "$2" -s $3 -c 1 -o $O/$1 $CMD $4 > $O/$1.log 2>&1
  echo "$1 rc=$? ${SECONDS}s $(ls -la $O/$1.ncu-rep 2>/dev/null | awk '{print $5}') bytes"
}
cap r2_bw_bulk k_backward 0 2
cap r2_bw_mid k_backward 36 38
cap r2_fw_bulk k_forward 1 3
cap r2_derivs k_derivs 1 3
cap r2_init k_init 0 1
