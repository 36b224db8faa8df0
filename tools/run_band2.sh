#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_band.py -x -q -m gpu -k "graph" > gpurun_out/r02q_band_graph.log 2>&1; tail -5 gpurun_out/r02q_band_graph.log
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --config wgate_train512 --steps 5 --warmup 3 --no-eager-baseline > gpurun_out/r02q_wgate_n2.log 2>&1; tail -c 400 gpurun_out/r02q_wgate_n2.log; echo
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --config gate_train512 --steps 5 --warmup 3 --no-eager-baseline > gpurun_out/r02q_gate_n2.log 2>&1; tail -c 400 gpurun_out/r02q_gate_n2.log; echo
