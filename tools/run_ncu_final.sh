#!/bin/bash
mkdir -p gpurun_out
python tools/prof_x3_one.py x3 2 > gpurun_out/plain_x3.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:"gemm_.._x3" -s 2 -c 3 -o gpurun_out/prof_r02aj_x3_l2 python tools/prof_x3_one.py x3 2 > gpurun_out/ncu_x3.log 2>&1
python tools/prof_f32_core.py 2 > gpurun_out/plain_f32core.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:"attn_core" -s 2 -c 2 -o gpurun_out/prof_r02aj_f32core_l2 python tools/prof_f32_core.py 2 > gpurun_out/ncu_f32core.log 2>&1
ls -la gpurun_out/prof_r02aj*.ncu-rep
