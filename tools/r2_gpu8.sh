set -x
O=gpurun_out
mkdir -p $O
python tools/phase_bench.py cartpole 16384 2 > /dev/null 2>&1
ncu --set full --import-source on --clock-control none -k regex:k_backward --launch-skip 0 --launch-count 1 -f \
    -o $O/r2_bw_bulk python tools/phase_bench.py cartpole 16384 2 > $O/r2_ncu_bw_bulk.log 2>&1
ncu --set full --import-source on --clock-control none -k regex:k_backward --launch-skip 36 --launch-count 1 -f \
    -o $O/r2_bw_mid python tools/phase_bench.py cartpole 16384 38 > $O/r2_ncu_bw_mid.log 2>&1
ncu --set full --import-source on --clock-control none -k regex:k_forward --launch-skip 1 --launch-count 1 -f \
    -o $O/r2_fw_bulk python tools/phase_bench.py cartpole 16384 3 > $O/r2_ncu_fw_bulk.log 2>&1
ncu --set full --clock-control none -k regex:k_derivs --launch-skip 1 --launch-count 1 -f \
    -o $O/r2_derivs python tools/phase_bench.py cartpole 16384 3 > $O/r2_ncu_derivs.log 2>&1
tail -2 $O/r2_ncu_bw_mid.log; ls -la $O/*.ncu-rep
timeout 900 python -m pytest tests/test_gpu_api.py -x -q -m gpu > $O/r2_pytest_gpu_api.log 2>&1; tail -3 $O/r2_pytest_gpu_api.log
