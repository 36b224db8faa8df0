set -x
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -x -q -m gpu 2>&1 | tail -8 > gpurun_out/r02ak_tests.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02ak_smoke.log 2>&1
python bench.py > gpurun_out/r02ak_bench.json 2> gpurun_out/r02ak_bench.err
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r02ak_bench_reference.json 2> gpurun_out/r02ak_bench_reference.err
for c in infer256_t192 train128_fp32 infer256_fp32; do
  timeout 900 python bench.py --config $c > gpurun_out/r02ak_bench_$c.json 2> gpurun_out/r02ak_bench_$c.err
  cut -c1-200 gpurun_out/r02ak_bench_$c.json
done
python tools/time_fp32.py > gpurun_out/r02ak_time_fp32.log 2>&1
python bench.py --config train128_fp32 --steps 2 --warmup 1 --no-cpu-baseline --no-eager-baseline > gpurun_out/plain_launch_f.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/r02ak_launches_train128_fp32.csv python bench.py --config train128_fp32 --steps 2 --warmup 1 --no-cpu-baseline --no-eager-baseline > gpurun_out/ncu_launch_f.log 2>&1
tail -n 4 gpurun_out/r02ak_tests.log gpurun_out/r02ak_smoke.log; cut -c1-250 gpurun_out/r02ak_bench.json
