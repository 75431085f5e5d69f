#!/bin/bash
# One GPU session: tests, bench, ncu launch list (same command as the bench, reduced steps), full captures.
# usage (on the GPU box, from the repo root): bash tools/profile_round.sh <tag>
TAG=${1:-r1}
O=gpurun_out
python -m pytest tests -m gpu -x -q > $O/pytest_gpu_$TAG.log 2>&1; tail -2 $O/pytest_gpu_$TAG.log
python bench.py > $O/bench_$TAG.json 2> $O/bench_$TAG.err; tail -c 600 $O/bench_$TAG.json; echo
python bench.py --impl reference --steps 3 --warmup 1 > $O/bench_ref_$TAG.json 2>> $O/bench_$TAG.err
# launch list: one sequential step of the bench workload (after the bench exited 0 without ncu)
ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file $O/launches_$TAG.csv \
    python tools/phase_times.py cartpole 16384 > $O/ncu_launches_$TAG.log 2>&1
# full captures: first bulk launches (known KKT count: one sweep unless restarted), and one lone-warp round
ncu --set full --import-source on --clock-control none -k regex:k_backward --launch-skip 0 --launch-count 1 -f \
    -o $O/bw_bulk_$TAG python tools/phase_bench.py cartpole 16384 2 > $O/ncu_bw_bulk_$TAG.log 2>&1
ncu --set full --import-source on --clock-control none -k regex:k_forward --launch-skip 1 --launch-count 1 -f \
    -o $O/fw_bulk_$TAG python tools/phase_bench.py cartpole 16384 3 > $O/ncu_fw_bulk_$TAG.log 2>&1
ncu --set full --import-source on --clock-control none -k regex:k_backward --launch-skip 6 --launch-count 1 -f \
    -o $O/bw_lone_$TAG python tools/phase_bench.py cartpole 8 8 > $O/ncu_bw_lone_$TAG.log 2>&1
ncu --set full --clock-control none -k regex:k_derivs --launch-skip 1 --launch-count 1 -f \
    -o $O/derivs_$TAG python tools/phase_bench.py cartpole 16384 3 > $O/ncu_derivs_$TAG.log 2>&1
ncu --set full --import-source on --clock-control none -k regex:k_backward --launch-skip 36 --launch-count 1 -f \
    -o $O/bw_mid_$TAG python tools/phase_bench.py cartpole 16384 38 > $O/ncu_bw_mid_$TAG.log 2>&1
# per-round series of the headline batch and the other BASELINE configs (one sequential solve each)
IPDDP_SERIES=1 python tools/phase_bench.py cartpole 16384 100 > $O/series_$TAG.log 2>&1
python tools/phase_bench.py cartpole 8 12 > $O/lone_$TAG.log 2>&1
{ IPDDP_KNOTS=201 python tools/phase_times.py acrobot 8192; python tools/phase_times.py acrobot 8192; python tools/phase_times.py concar 4096;
  python tools/phase_times.py concar_quad 4096; IPDDP_KNOTS=141 IPDDP_VARY_HORIZON=1 python tools/phase_times.py pushing 2048; } > $O/configs_$TAG.log 2>&1
tail -2 $O/ncu_bw_bulk_$TAG.log; cat $O/configs_$TAG.log; tail -1 $O/lone_$TAG.log; ls -la $O/*_$TAG.ncu-rep
