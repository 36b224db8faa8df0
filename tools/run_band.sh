#!/bin/bash
# band attention (WGATE / GATE): parity tests + bench lines
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_band.py -x -q -m gpu > gpurun_out/r02p_band_tests.log 2>&1
tail -5 gpurun_out/r02p_band_tests.log
for c in wgate_train512 gate_train512; do
  timeout 900 python bench.py --config $c --steps 10 --warmup 3 > gpurun_out/r02p_bench_$c.log 2>&1
  tail -c 300 gpurun_out/r02p_bench_$c.log; echo
done
