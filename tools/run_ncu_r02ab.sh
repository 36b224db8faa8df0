#!/bin/bash
mkdir -p gpurun_out
python tools/prof_x3_one.py x3 2 > gpurun_out/plain_x3.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:"x3|split3" -s 4 -c 6 -o gpurun_out/prof_r02ab_x3_l2 python tools/prof_x3_one.py x3 2 > gpurun_out/ncu_x3.log 2>&1
python tools/prof_x3_one.py ffn 0 > gpurun_out/plain_ffn0.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:"ffn_eval_fused" -s 1 -c 1 -o gpurun_out/prof_r02ab_k10f_l0 python tools/prof_x3_one.py ffn 0 > gpurun_out/ncu_ffn0.log 2>&1
python tools/prof_x3_one.py ffn 1 > gpurun_out/plain_ffn1.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:"ffn_eval_fused" -s 1 -c 1 -o gpurun_out/prof_r02ab_k10f_l1 python tools/prof_x3_one.py ffn 1 > gpurun_out/ncu_ffn1.log 2>&1
ls -la gpurun_out/*.ncu-rep | tail -5
