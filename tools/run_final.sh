set -x
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -x -q -m gpu 2>&1 | tail -8 > gpurun_out/final_tests.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/final_smoke.log 2>&1
python bench.py > gpurun_out/final_bench.json 2> gpurun_out/final_bench.err
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/final_bench_reference.json 2> gpurun_out/final_bench_reference.err
python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-eager-baseline > gpurun_out/plain_launch.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/final_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-eager-baseline > gpurun_out/ncu_launch.log 2>&1
tail -n 4 gpurun_out/final_tests.log gpurun_out/final_smoke.log; cut -c1-250 gpurun_out/final_bench.json
for c in wgate_train512 gate_train512; do
  timeout 900 python bench.py --config $c --steps 10 --warmup 3 > gpurun_out/final_bench_$c.json 2> gpurun_out/final_bench_$c.err
  cut -c1-200 gpurun_out/final_bench_$c.json
done
