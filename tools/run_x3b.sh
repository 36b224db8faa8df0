#!/bin/bash
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -x -q -m gpu > gpurun_out/r02y_tests.log 2>&1
tail -5 gpurun_out/r02y_tests.log
python -c 'import __graft_entry__ as g; g.smoke()' > gpurun_out/r02y_smoke.log 2>&1; tail -2 gpurun_out/r02y_smoke.log
for c in train128_fp32 infer256_fp32; do
  timeout 900 python bench.py --config $c --steps 5 --warmup 3 > gpurun_out/r02y_bench_$c.json 2> gpurun_out/r02y_bench_$c.err
  python - <<P
import json
d=json.load(open("gpurun_out/r02y_bench_$c.json"))
print("$c", d["value"], d["ms_per_step"], d["e2e"]["value"], d["dtype"], d.get("gpu_eager_baseline"), d.get("speedup_vs_gpu_eager_fp32"), d["config"]["precision"])
P
done
HWGAT_FP32=ffma timeout 900 python bench.py --config train128_fp32 --steps 3 --warmup 2 --no-cpu-baseline --no-eager-baseline > gpurun_out/r02y_bench_train128_fp32_ffma.json 2>/dev/null
cut -c1-330 gpurun_out/r02y_bench_train128_fp32_ffma.json
