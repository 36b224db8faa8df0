set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_windows.py tests/test_gpu_hgate.py tests/test_gpu_round2.py -m gpu -q 2>&1 | tail -40 > gpurun_out/r02j_tests.log
python bench.py --config train_t256_w32 --steps 5 --warmup 3 --no-cpu-baseline --no-eager-baseline > gpurun_out/bench_r02j_t256_w32.json 2> gpurun_out/bench_r02j_t256_w32.err
python bench.py --config train_t256_w64 --steps 5 --warmup 3 --no-cpu-baseline --no-eager-baseline > gpurun_out/bench_r02j_t256_w64.json 2> gpurun_out/bench_r02j_t256_w64.err
python bench.py --config hgate_train512 --steps 10 --warmup 3 > gpurun_out/bench_r02j_hgate.json 2> gpurun_out/bench_r02j_hgate.err
tail -n 6 gpurun_out/r02j_tests.log; tail -n 2 gpurun_out/bench_r02j_hgate.err; cut -c1-300 gpurun_out/bench_r02j_hgate.json
