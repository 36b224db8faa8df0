set -x
mkdir -p gpurun_out
run() { name=$1; shift; python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port $((29600 + RANDOM % 300)) bench.py --gpus 8 "$@" > gpurun_out/r02ag_bench_n8_$name.json 2> gpurun_out/r02ag_bench_n8_$name.err; }
run weak --steps 10 --warmup 3 --trace-allreduce
run infer --steps 10 --warmup 3 --config infer256_t192
run fp32 --steps 5 --warmup 3 --config train128_fp32
for f in gpurun_out/r02ag_bench_n8_*.json; do echo $f; cut -c1-260 $f; done
