set -x
mkdir -p gpurun_out
python -m pytest tests -m gpu -q -p no:cacheprovider -k "not cuda_engine" > gpurun_out/r2_gputests_3.log 2>&1; echo "rc=$?" >> gpurun_out/r2_gputests_3.log
tail -15 gpurun_out/r2_gputests_3.log
python scripts/bench_configs.py > gpurun_out/r2_configs.jsonl 2> gpurun_out/r2_configs.err; tail -3 gpurun_out/r2_configs.err; cat gpurun_out/r2_configs.jsonl
python scripts/host_overhead_live.py > gpurun_out/r2_host_overhead_live.json 2> gpurun_out/r2_host_overhead_live.err; tail -3 gpurun_out/r2_host_overhead_live.err; cat gpurun_out/r2_host_overhead_live.json
python bench.py --steps 20 --warmup 5 --no-cpu-reference > gpurun_out/r2_bench_n1_b.json 2> gpurun_out/r2_bench_n1_b.err; tail -3 gpurun_out/r2_bench_n1_b.err; python -c "
import json; b=json.loads(open('gpurun_out/r2_bench_n1_b.json').read().strip().splitlines()[-1]); print({k:b[k] for k in ('value','ms_per_step')}, b['e2e_optic_trace'], b['e2e']['ms_per_step'])"
for c in c5pol zern c3grad; do
ncu --set full --clock-control none --import-source on -k regex:trace -s 8 -c 2 -o gpurun_out/r2_$c -f python scripts/bench_configs.py $c > gpurun_out/ncu_$c.log 2>&1
ncu -i gpurun_out/r2_$c.ncu-rep --page raw --csv > gpurun_out/r2_$c.raw.csv 2>/dev/null
done
ls -la gpurun_out/*.ncu-rep
