set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_round2.py -m gpu -q -s 2>&1 | tail -150 > gpurun_out/r02c_tests_round2.log
timeout 900 python -m pytest tests/test_gpu_windows.py -m gpu -q -s 2>&1 | tail -150 > gpurun_out/r02c_tests_windows.log
timeout 900 python -m pytest tests -m gpu -q --deselect tests/test_gpu_round2.py --deselect tests/test_gpu_windows.py 2>&1 | tail -40 > gpurun_out/r02c_tests_rest.log
python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-eager-baseline > gpurun_out/bench_r02c.json 2> gpurun_out/bench_r02c.err
HWGAT_ATTN_IMPL=tc2 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-eager-baseline > gpurun_out/bench_r02c_tc2.json 2> gpurun_out/bench_r02c_tc2.err
python bench.py --config train_t256_w32 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_r02c_t256_w32.json 2> gpurun_out/bench_r02c_t256_w32.err
python bench.py --config train_t256_w64 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_r02c_t256_w64.json 2> gpurun_out/bench_r02c_t256_w64.err
tail -n 3 gpurun_out/r02c_tests_round2.log gpurun_out/r02c_tests_windows.log gpurun_out/r02c_tests_rest.log
