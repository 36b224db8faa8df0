#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_x3.py -x -q -m gpu -s > gpurun_out/r02z_x3_tests.log 2>&1
grep -a "x3 n=" gpurun_out/r02z_x3_tests.log | sed 's/^\.*//'; tail -3 gpurun_out/r02z_x3_tests.log
timeout 900 python -m pytest tests -x -q -m gpu -k "f32 or fp32" > gpurun_out/r02z_fp32_suite.log 2>&1; tail -3 gpurun_out/r02z_fp32_suite.log
python tools/time_fp32.py > gpurun_out/r02z_time_fp32.log 2>&1
grep -av "^void\|^hwgat::\|^---\|Memset\|^ *Name" gpurun_out/r02z_time_fp32.log | tail -12
grep -a "^void\|^hwgat::" gpurun_out/r02z_time_fp32.log | cut -c1-62,150-230 | head -12
