#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_band.py -x -q -m gpu > gpurun_out/r02t_band_tests.log 2>&1; tail -3 gpurun_out/r02t_band_tests.log
for i in 1 2; do
  for m in wgate gate; do
    echo "$(python tools/prof_band.py $m --time 2>&1 | grep hwgat_ | tr '\n' ' ')"
  done
done > gpurun_out/r02t_band_time.log 2>&1
cat gpurun_out/r02t_band_time.log
