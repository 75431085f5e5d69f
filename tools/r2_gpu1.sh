set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2_pytest_gpu_1.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_pytest_gpu_1.log
tail -5 gpurun_out/r2_pytest_gpu_1.log
timeout 600 python tools/queue_bench.py cartpole 65536 16384,32768,8192 > gpurun_out/r2_queue_bench_1.log 2>&1
cat gpurun_out/r2_queue_bench_1.log
timeout 600 bash tools/ab_variants.sh c_base c_tight c_nnz1 c_nanmax c_udiv c_nnz1_nanmax c_all c_occ16 c_occ12 c_occ8 > gpurun_out/r2_ab_1.log 2>&1
cat gpurun_out/r2_ab_1.log
