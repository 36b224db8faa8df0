set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_windows.py tests/test_gpu_hgate.py -m gpu -q 2>&1 | tail -20 > gpurun_out/r02g_tests.log
for a in "0 16" "2 16" "0 32" "2 32" "0 64" "2 64"; do python tools/prof_tc2.py $a --time; done > gpurun_out/r02g_tc2_times.log 2>&1
tail -n 4 gpurun_out/r02g_tests.log; grep lvl gpurun_out/r02g_tc2_times.log
