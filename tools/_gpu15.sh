python tools/parity_sweep.py 384 > gpurun_out/r2_parity_sweep.jsonl 2> gpurun_out/r2_parity_sweep.err; echo "rc=$?"; cat gpurun_out/r2_parity_sweep.jsonl; tail -3 gpurun_out/r2_parity_sweep.err
