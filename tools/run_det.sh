#!/bin/bash
# deterministic mode: the parity test + what it costs on the headline step
set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_round2.py -x -q -m gpu -k "determin" > gpurun_out/r02m_det_tests.log 2>&1
tail -5 gpurun_out/r02m_det_tests.log
HWGAT_DETERMINISTIC=1 timeout 600 python bench.py --steps 5 --warmup 3 --no-eager-baseline > gpurun_out/r02m_bench_det.log 2>&1
tail -2 gpurun_out/r02m_bench_det.log
timeout 600 python bench.py --steps 5 --warmup 3 --no-eager-baseline > gpurun_out/r02m_bench_default.log 2>&1
tail -2 gpurun_out/r02m_bench_default.log
