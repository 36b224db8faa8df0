set -x
mkdir -p gpurun_out
run() { name=$1; shift; python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port $((29600 + RANDOM % 300)) bench.py --gpus 8 "$@" > gpurun_out/bench_r02_n8_$name.json 2> gpurun_out/bench_r02_n8_$name.err; }
run weak --steps 20 --warmup 5 --trace-allreduce
run strong --steps 20 --warmup 5 --strong --batch 512
run infer --steps 10 --warmup 3 --config infer256_t192
run t256 --steps 5 --warmup 3 --config train_t256
run t256_w32 --steps 5 --warmup 3 --config train_t256_w32
run t256_w64 --steps 5 --warmup 3 --config train_t256_w64
for f in gpurun_out/bench_r02_n8_*.json; do echo $f; cut -c1-300 $f; done
