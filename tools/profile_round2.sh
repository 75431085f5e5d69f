#!/bin/bash
# One GPU session of round 2: GPU tests, smoke, the bench line (timed as the driver times it), the reference arm, the ncu
# launch list of the same bench command at reduced steps, the regenerated results tables.
#   gpurun --timeout 3000 -- 'bash tools/profile_round2.sh final'
TAG=${1:-final}
O=gpurun_out
mkdir -p $O
python -m pytest tests -m gpu -x -q > $O/r2_pytest_gpu_$TAG.log 2>&1; tail -3 $O/r2_pytest_gpu_$TAG.log
python -c "import __graft_entry__ as g; g.smoke()" > $O/r2_smoke_$TAG.log 2>&1; tail -1 $O/r2_smoke_$TAG.log
SECONDS=0; python bench.py --steps 20 --warmup 3 > $O/r2_bench_$TAG.json 2> $O/r2_bench_$TAG.err; echo "bench rc=$? wall_s=$SECONDS" | tee $O/r2_bench_wall_$TAG.txt

tail -c 800 $O/r2_bench_$TAG.json; echo
python bench.py --impl reference --steps 3 --warmup 1 > $O/r2_bench_ref_$TAG.json 2>> $O/r2_bench_$TAG.err; tail -c 600 $O/r2_bench_ref_$TAG.json; echo
# launch list of the bench command (after it exited 0 without ncu), reduced steps
ncu --metrics gpu__time_duration.sum --clock-control none -c 12000 --csv --log-file $O/r2_launches_$TAG.csv \
    python bench.py --steps 2 --warmup 1 --no-configs --no-cpu-baseline --no-single > $O/r2_ncu_launches_$TAG.log 2>&1
tail -2 $O/r2_ncu_launches_$TAG.log | cut -c1-300
python tools/run_experiments.py --out $O/r2_experiments --compare > $O/r2_experiments_$TAG.log 2>&1; tail -7 $O/r2_experiments_$TAG.log
