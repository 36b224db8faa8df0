#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_windows.py tests/test_gpu_hgate.py -x -q -m gpu -k "f32 or fp32 or refused" -s > gpurun_out/r02u_f32_tests.log 2>&1
grep -E "max-rel|passed|failed|Error|error|assert" gpurun_out/r02u_f32_tests.log | tail -40
