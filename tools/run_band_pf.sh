#!/bin/bash
mkdir -p gpurun_out
for i in 1 2; do
for pf in 0 1; do
  for m in wgate gate; do
    echo "prefetch=$pf $(HWGAT_BAND_PREFETCH=$pf python tools/prof_band.py $m --time 2>&1 | grep hwgat_ | tr '\n' ' ')"
  done
done
done > gpurun_out/r02r_band_prefetch.log 2>&1
cat gpurun_out/r02r_band_prefetch.log
timeout 900 python -m pytest tests/test_gpu_band.py -x -q -m gpu > gpurun_out/r02r_band_tests.log 2>&1; tail -3 gpurun_out/r02r_band_tests.log
