set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_round2.py tests/test_gpu_x3.py tests/test_gpu_parity.py -m gpu -q -k "second_device or guard or outside" 2>&1 | tail -4 > gpurun_out/r02ac_n2_tests.log
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/r02ac_bench_n2.json 2> gpurun_out/r02ac_bench_n2.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 10 --warmup 3 --config infer256_t192 > gpurun_out/r02ac_bench_n2_infer.json 2> gpurun_out/r02ac_bench_n2_infer.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 2 --steps 5 --warmup 3 --config train128_fp32 > gpurun_out/r02ac_bench_n2_fp32.json 2> gpurun_out/r02ac_bench_n2_fp32.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29514 bench.py --impl reference --gpus 2 --steps 1 --warmup 0 > gpurun_out/r02ac_bench_n2_ref.json 2> gpurun_out/r02ac_bench_n2_ref.err
tail -n 3 gpurun_out/r02ac_n2_tests.log; for f in n2 n2_infer n2_fp32 n2_ref; do tail -n 2 gpurun_out/r02ac_bench_$f.err; cut -c1-260 gpurun_out/r02ac_bench_$f.json; done
