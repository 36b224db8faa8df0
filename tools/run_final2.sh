#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -x -q -m gpu > gpurun_out/final2_tests.log 2>&1
tail -4 gpurun_out/final2_tests.log
python -c 'import __graft_entry__ as g; g.smoke()' > gpurun_out/final2_smoke.log 2>&1; tail -2 gpurun_out/final2_smoke.log
python bench.py > gpurun_out/final2_bench.json 2> gpurun_out/final2_bench.err; cut -c1-300 gpurun_out/final2_bench.json
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/final2_bench_reference.json 2> gpurun_out/final2_bench_reference.err; cut -c1-200 gpurun_out/final2_bench_reference.json
timeout 600 python bench.py --config infer256_t192 > gpurun_out/final2_bench_infer256_t192.json 2>/dev/null; cut -c1-200 gpurun_out/final2_bench_infer256_t192.json
for c in train128_fp32 infer256_fp32; do timeout 600 python bench.py --config $c --steps 5 --warmup 3 > gpurun_out/final2_bench_$c.json 2>/dev/null; cut -c1-200 gpurun_out/final2_bench_$c.json; done
