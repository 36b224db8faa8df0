set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_round2.py -m gpu -q -k "second_device" 2>&1 | tail -4 > gpurun_out/r02_n2_tests.log
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 --trace-allreduce > gpurun_out/bench_r02_n2.json 2> gpurun_out/bench_r02_n2.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 10 --warmup 3 --config infer256_t192 > gpurun_out/bench_r02_n2_infer.json 2> gpurun_out/bench_r02_n2_infer.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 2 --steps 10 --warmup 3 --strong --batch 512 > gpurun_out/bench_r02_n2_strong.json 2> gpurun_out/bench_r02_n2_strong.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29514 bench.py --impl reference --gpus 2 --steps 1 --warmup 0 > gpurun_out/bench_r02_n2_ref.json 2> gpurun_out/bench_r02_n2_ref.err
tail -n 3 gpurun_out/r02_n2_tests.log; tail -n 3 gpurun_out/bench_r02_n2.err; cut -c1-400 gpurun_out/bench_r02_n2.json
