set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_windows.py tests/test_gpu_hgate.py -m gpu -q 2>&1 | tail -20 > gpurun_out/r02i_tests.log
for a in "0 16" "2 16" "0 32" "2 32" "0 64" "2 64"; do python tools/prof_tc2.py $a --time; done > gpurun_out/r02i_tc2_times.log 2>&1
python tools/prof_bwd_one.py 2 > gpurun_out/plain_l2.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:"attn_bwd|gemm_tc_tn|gemm_nt_epi" -s 3 -c 3 -o gpurun_out/prof_r02i_k3_l2 python tools/prof_bwd_one.py 2 > gpurun_out/ncu_l2.log 2>&1
tail -n 4 gpurun_out/r02i_tests.log; grep lvl gpurun_out/r02i_tc2_times.log
