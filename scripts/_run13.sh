mkdir -p gpurun_out
N=${1:-8}
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/r2_bench_n$N.json 2> gpurun_out/r2_bench_n$N.err; tail -2 gpurun_out/r2_bench_n$N.err; cut -c1-600 gpurun_out/r2_bench_n$N.json
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 20 --warmup 5 --no-numa > gpurun_out/r2_bench_n${N}_nonuma.json 2> gpurun_out/r2_bench_n${N}_nonuma.err; cut -c1-300 gpurun_out/r2_bench_n${N}_nonuma.json
python -m pytest tests/test_gpu_multi.py -m gpu -q -p no:cacheprovider 2>&1 | tail -3
