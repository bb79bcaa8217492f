mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/r2_topo.txt 2>&1
python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 4 --steps 20 --warmup 5 > gpurun_out/r2_bench_n4.json 2> gpurun_out/r2_bench_n4.err; tail -3 gpurun_out/r2_bench_n4.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29518 bench.py --gpus 4 --steps 20 --warmup 5 --no-numa --no-optic-trace > gpurun_out/r2_bench_n4_nonuma.json 2> gpurun_out/r2_bench_n4_nonuma.err; tail -3 gpurun_out/r2_bench_n4_nonuma.err
python - <<'PY'
import json
for f in ("r2_bench_n4","r2_bench_n4_nonuma"):
    b=json.loads(open(f"gpurun_out/{f}.json").read().strip().splitlines()[-1])
    print(f, b["value"], b["ms_per_step"], {k:(round(b[k]["ms_per_step"],3)) for k in ("e2e","e2e_launch_arrays","e2e_spot_rms")}, b["e2e_optic_trace"].get("ms_per_step"), b["config"]["numa_binding"])
    print({k:round(v["ms_per_step"],4) for k,v in b["sharded_fixed_total"].items()})
PY
