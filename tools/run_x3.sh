#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_x3.py -x -q -m gpu -s > gpurun_out/r02x_x3_tests.log 2>&1
grep -E "^x3|passed|failed|Error|error|assert" gpurun_out/r02x_x3_tests.log | tail -30
python tools/time_fp32.py > gpurun_out/r02x_time_fp32.log 2>&1
grep -v "^void\|^hwgat::\|^---\|Memset\|^ *Name" gpurun_out/r02x_time_fp32.log | tail -20
