#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_band.py -x -q -m gpu > gpurun_out/r02o_band_tests.log 2>&1; tail -3 gpurun_out/r02o_band_tests.log
python tools/prof_band.py wgate --time > gpurun_out/r02o_band_time_wgate.log 2>&1
python tools/prof_band.py gate --time > gpurun_out/r02o_band_time_gate.log 2>&1
cat gpurun_out/r02o_band_time_*.log
python tools/prof_band.py wgate > gpurun_out/plain_band.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:"band_attn" -s 4 -c 2 -o gpurun_out/prof_r02o_band_wgate -f python tools/prof_band.py wgate > gpurun_out/ncu_band.log 2>&1
tail -3 gpurun_out/ncu_band.log
