set -x
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q 2>&1 | tail -15 > gpurun_out/r02e_tests_all.log
python bench.py --steps 20 --warmup 5 > gpurun_out/bench_r02e.json 2> gpurun_out/bench_r02e.err
python tools/prof_tc2.py 0 16 > gpurun_out/plain_tc2.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:attn_core -s 2 -c 2 -o gpurun_out/prof_r02e_tc2_l0_w16 python tools/prof_tc2.py 0 16 > gpurun_out/ncu_tc2_a.log 2>&1
python tools/prof_tc2.py 2 64 > gpurun_out/plain_tc2b.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:attn_core -s 2 -c 2 -o gpurun_out/prof_r02e_tc2_l2_w64 python tools/prof_tc2.py 2 64 > gpurun_out/ncu_tc2_b.log 2>&1
python tools/prof_misc.py > gpurun_out/plain_misc.log 2>&1 && ncu --set full --clock-control none -k regex:"adjacency|mask_bits|bda_ln_fwd|ln_bwd|merge_kernel|ln_fwd" -s 10 -c 8 -o gpurun_out/prof_r02e_misc python tools/prof_misc.py > gpurun_out/ncu_misc.log 2>&1
python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-eager-baseline > gpurun_out/plain_launch.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/r02e_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-eager-baseline > gpurun_out/ncu_launch.log 2>&1
tail -n 4 gpurun_out/r02e_tests_all.log
