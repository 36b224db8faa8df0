#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "feed_forward" > gpurun_out/r02ab_ffn_tests.log 2>&1; tail -15 gpurun_out/r02ab_ffn_tests.log
timeout 120 python tools/time_ffn_eval.py > gpurun_out/r02ab_ffn_eval.log 2>&1; cat gpurun_out/r02ab_ffn_eval.log
HWGAT_FFN_FUSED=0 timeout 120 python tools/time_ffn_eval.py > gpurun_out/r02ab_ffn_eval_unfused.log 2>&1; cat gpurun_out/r02ab_ffn_eval_unfused.log
timeout 600 python bench.py --config infer256_t192 --steps 10 --warmup 3 --no-cpu-baseline --no-eager-baseline > gpurun_out/r02ab_bench_infer256_t192.json 2>gpurun_out/r02ab_bench_infer.err
cut -c1-200 gpurun_out/r02ab_bench_infer256_t192.json
