#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu -k "f32 or fp32 or x3" > gpurun_out/r02ae_fp32_tests.log 2>&1; tail -4 gpurun_out/r02ae_fp32_tests.log
python tools/time_fp32.py > gpurun_out/r02ae_time_fp32.log 2>&1
grep -a "^fp32 mode\|^eval logits\|^eager" gpurun_out/r02ae_time_fp32.log
grep -a "^void\|^hwgat::" gpurun_out/r02ae_time_fp32.log | cut -c1-62,150-230 | head -8
