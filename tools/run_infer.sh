#!/bin/bash
mkdir -p gpurun_out
timeout 900 python bench.py --config infer256_t192 --steps 10 --warmup 3 --no-eager-baseline > gpurun_out/r02v_infer.log 2>&1
tail -c 300 gpurun_out/r02v_infer.log
