set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_windows.py tests/test_gpu_round2.py -m gpu -q -s 2>&1 | tail -80 > gpurun_out/r02d_tests.log
for a in "0 16" "1 16" "2 16" "0 32" "2 32" "0 64" "2 64"; do python tools/prof_tc2.py $a --time; done > gpurun_out/r02d_tc2_times.log 2>&1
python tools/prof_tc2.py 0 16 > gpurun_out/plain_tc2.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:attn_core -s 4 -c 2 -o gpurun_out/prof_r02d_tc2_l0_w16 python tools/prof_tc2.py 0 16 > gpurun_out/ncu_tc2_a.log 2>&1
python tools/prof_tc2.py 2 64 > gpurun_out/plain_tc2b.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:attn_core -s 4 -c 2 -o gpurun_out/prof_r02d_tc2_l2_w64 python tools/prof_tc2.py 2 64 > gpurun_out/ncu_tc2_b.log 2>&1
tail -n 4 gpurun_out/r02d_tests.log; cat gpurun_out/r02d_tc2_times.log
