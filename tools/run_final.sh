set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -8 > gpurun_out/final_tests.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/final_smoke.log 2>&1
python bench.py > gpurun_out/final_bench.json 2> gpurun_out/final_bench.err
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/final_bench_reference.json 2> gpurun_out/final_bench_reference.err
tail -n 4 gpurun_out/final_tests.log gpurun_out/final_smoke.log; cut -c1-250 gpurun_out/final_bench.json
