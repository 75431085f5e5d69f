set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2_pytest_gpu_2.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_pytest_gpu_2.log
tail -5 gpurun_out/r2_pytest_gpu_2.log
timeout 1200 python bench.py --steps 20 --warmup 3 > gpurun_out/r2_bench_k20.json 2> gpurun_out/r2_bench_k20.err; echo "bench rc=$?"
tail -c 1500 gpurun_out/r2_bench_k20.json; tail -5 gpurun_out/r2_bench_k20.err
