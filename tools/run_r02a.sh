set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv
python -m pytest tests/test_gpu_round2.py -m gpu -x -q -s 2>&1 | tail -40 > gpurun_out/r02a_tests_round2.log
python -m pytest tests -m gpu -q --deselect tests/test_gpu_round2.py 2>&1 | tail -8 > gpurun_out/r02a_tests_rest.log
python bench.py --steps 10 --warmup 3 > gpurun_out/bench_r02a.json 2> gpurun_out/bench_r02a.err
python bench.py --config infer256_t192 --steps 10 --warmup 3 > gpurun_out/bench_r02a_infer.json 2> gpurun_out/bench_r02a_infer.err
python bench.py --config train_t256 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_r02a_t256.json 2> gpurun_out/bench_r02a_t256.err
python tools/prof_bwd_one.py 0 > gpurun_out/plain_l0.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:attn_ -s 2 -c 2 -o gpurun_out/prof_r02a_k2k3_l0 python tools/prof_bwd_one.py 0 > gpurun_out/ncu_l0.log 2>&1
python tools/prof_bwd_one.py 1 > gpurun_out/plain_l1.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:attn_ -s 2 -c 2 -o gpurun_out/prof_r02a_k2k3_l1 python tools/prof_bwd_one.py 1 > gpurun_out/ncu_l1.log 2>&1
tail -3 gpurun_out/r02a_tests_round2.log gpurun_out/r02a_tests_rest.log
