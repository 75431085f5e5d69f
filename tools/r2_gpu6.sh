set -x
mkdir -p gpurun_out
nvidia-smi -L
timeout 900 python -m pytest tests/test_gpu_multi.py -x -q -m gpu > gpurun_out/r2_pytest_gpu_multi.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_pytest_gpu_multi.log
tail -5 gpurun_out/r2_pytest_gpu_multi.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611 bench.py --gpus 2 --steps 4 --warmup 1 > gpurun_out/r2_bench_2gpu.json 2> gpurun_out/r2_bench_2gpu.err; echo "bench rc=$?"
tail -c 3000 gpurun_out/r2_bench_2gpu.json; tail -5 gpurun_out/r2_bench_2gpu.err
