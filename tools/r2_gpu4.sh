set -x
mkdir -p gpurun_out
bash tools/ab_variants.sh c_base c_tma > gpurun_out/r2_ab_tma.log 2>&1
cat gpurun_out/r2_ab_tma.log
python tools/queue_bench.py cartpole 131072 65536 101 interiorpointddp.jl_b200/libipddp_b200_c_base.so > gpurun_out/r2_queue_tma.log 2>&1
python tools/queue_bench.py cartpole 131072 65536 101 interiorpointddp.jl_b200/libipddp_b200_c_tma.so >> gpurun_out/r2_queue_tma.log 2>&1
cat gpurun_out/r2_queue_tma.log
# lone-instance regime (8 instances): latency of the tail rounds with and without TMA staging
python tools/phase_bench.py cartpole 8 40 interiorpointddp.jl_b200/libipddp_b200_c_base.so > gpurun_out/r2_lone_base.log 2>&1
python tools/phase_bench.py cartpole 8 40 interiorpointddp.jl_b200/libipddp_b200_c_tma.so > gpurun_out/r2_lone_tma.log 2>&1
tail -1 gpurun_out/r2_lone_base.log; tail -1 gpurun_out/r2_lone_tma.log
