set -x
mkdir -p gpurun_out
rm -f gpurun_out/r2_qlog_*.jsonl
IPDDP_QUEUE_LOG=gpurun_out/r2_qlog_16384.jsonl python tools/queue_bench.py cartpole 131072 16384 > gpurun_out/r2_queue_bench_3.log 2>&1
IPDDP_QUEUE_LOG=gpurun_out/r2_qlog_32768.jsonl python tools/queue_bench.py cartpole 131072 32768 >> gpurun_out/r2_queue_bench_3.log 2>&1
python tools/queue_bench.py cartpole 131072 65536 >> gpurun_out/r2_queue_bench_3.log 2>&1
IPDDP_TUNE=fw_spec_max=0 python tools/queue_bench.py cartpole 131072 32768 >> gpurun_out/r2_queue_bench_3.log 2>&1
cat gpurun_out/r2_queue_bench_3.log
