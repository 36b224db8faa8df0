#!/bin/bash
# WGATE / GATE: bench lines (with the eager-PyTorch incumbent) 
mkdir -p gpurun_out
for c in wgate_train512 gate_train512; do
  timeout 900 python bench.py --config $c --steps 5 --warmup 3 > gpurun_out/r02n_bench_$c.log 2>&1
  tail -c 600 gpurun_out/r02n_bench_$c.log; echo
done
