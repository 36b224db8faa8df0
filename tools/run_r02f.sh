set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_hgate.py -m gpu -q -s 2>&1 | tail -80 > gpurun_out/r02f_tests_hgate.log
timeout 1200 python -m pytest tests -m gpu -q --deselect tests/test_gpu_hgate.py 2>&1 | tail -30 > gpurun_out/r02f_tests_all.log
tail -n 6 gpurun_out/r02f_tests_hgate.log gpurun_out/r02f_tests_all.log
