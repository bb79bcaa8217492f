mkdir -p gpurun_out
N=8
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/r2_bench_n$N.json 2> gpurun_out/r2_bench_n$N.err; tail -2 gpurun_out/r2_bench_n$N.err; python -c "
import json; d=json.loads(open('gpurun_out/r2_bench_n8.json').read().strip().splitlines()[-1]); print(d['value'], d['clocks'], d['e2e']['ms_per_step'])"
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29513 scripts/pcie_concurrent.py > gpurun_out/r2_pcie_concurrent_n8.json 2> gpurun_out/pcie.err; tail -2 gpurun_out/pcie.err; cut -c1-1500 gpurun_out/r2_pcie_concurrent_n8.json
lscpu | grep -i "numa\|socket\|model name" > gpurun_out/r2_lscpu_n8.txt; cat gpurun_out/r2_lscpu_n8.txt
