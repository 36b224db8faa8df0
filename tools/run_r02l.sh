set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_windows.py -m gpu -q -k "hybrid or tc2_attention_vs_oracle or empty" 2>&1 | tail -30 > gpurun_out/r02l_tests.log
python tools/prof_hybrid.py > gpurun_out/r02l_hybrid_times.log 2>&1
tail -n 5 gpurun_out/r02l_tests.log; cat gpurun_out/r02l_hybrid_times.log
