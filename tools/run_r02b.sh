set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_round2.py -m gpu -q -s 2>&1 | tail -60 > gpurun_out/r02b_tests_round2.log
timeout 900 python -m pytest tests/test_gpu_windows.py -m gpu -q -s -x 2>&1 | tail -60 > gpurun_out/r02b_tests_windows.log
timeout 900 python -m pytest tests -m gpu -q -x --deselect tests/test_gpu_round2.py --deselect tests/test_gpu_windows.py 2>&1 | tail -15 > gpurun_out/r02b_tests_rest.log
python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-eager-baseline > gpurun_out/bench_r02b.json 2> gpurun_out/bench_r02b.err
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02b_smoke.log 2>&1
tail -n 5 gpurun_out/r02b_tests_round2.log gpurun_out/r02b_tests_windows.log gpurun_out/r02b_tests_rest.log gpurun_out/r02b_smoke.log
