#!/bin/bash
# band attention (WGATE / GATE): parity tests
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_band.py -x -q -m gpu -s > gpurun_out/r02s_band_tests.log 2>&1
grep -E "fp32 max-rel|passed|failed|Error|error" gpurun_out/r02s_band_tests.log | tail -20
