set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "gemm_tn or feed_forward or attention_full_size or output_projection" 2>&1 | tail -6 > gpurun_out/r02h_tests.log
python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-eager-baseline > gpurun_out/bench_r02h.json 2> gpurun_out/bench_r02h.err
python tools/prof_bwd_one.py 2 > gpurun_out/plain_l2.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:"attn_bwd|gemm_tc_tn|gemm_nt_epi" -s 6 -c 3 -o gpurun_out/prof_r02h_k3_l2 python tools/prof_bwd_one.py 2 > gpurun_out/ncu_l2.log 2>&1
tail -n 3 gpurun_out/r02h_tests.log
