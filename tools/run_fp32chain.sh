#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_x3.py -x -q -m gpu > gpurun_out/r02ad_x3_tests.log 2>&1; tail -25 gpurun_out/r02ad_x3_tests.log
timeout 1800 python -m pytest tests -x -q -m gpu > gpurun_out/r02ad_fp32_suite.log 2>&1; tail -25 gpurun_out/r02ad_fp32_suite.log
python tools/time_fp32.py > gpurun_out/r02ad_time_fp32.log 2>&1
grep -av "^void\|^hwgat::\|^---\|Memset\|^ *Name" gpurun_out/r02ad_time_fp32.log | tail -12
grep -a "^void\|^hwgat::" gpurun_out/r02ad_time_fp32.log | cut -c1-62,150-230 | head -16
