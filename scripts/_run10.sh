mkdir -p gpurun_out
python -m pytest tests -m gpu -q -p no:cacheprovider -k "not cuda_engine" > gpurun_out/r2_gputests_6.log 2>&1; echo "rc=$?" >> gpurun_out/r2_gputests_6.log; tail -4 gpurun_out/r2_gputests_6.log
python bench.py --steps 20 --warmup 5 > gpurun_out/r2_bench_n1.json 2> gpurun_out/r2_bench_n1.err; tail -2 gpurun_out/r2_bench_n1.err; cut -c1-300 gpurun_out/r2_bench_n1.json
python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/r2_bench_ref.json 2>/dev/null
python scripts/bench_configs.py > gpurun_out/r2_configs_final.jsonl 2> gpurun_out/r2_configs_final.err; cut -c1-330 gpurun_out/r2_configs_final.jsonl
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_launches_bench.csv python bench.py --steps 5 --warmup 3 --no-optic-trace --no-cpu-reference > gpurun_out/ncu_bench.log 2>&1
for spec in "c2 8 c2_f32" "c2 31 c2_f64" "zern 8 zern_f32" "zern 31 zern_f64" "c5pol 8 c5pol_f32" "c5pol 31 c5pol_f64"; do set -- $spec
ncu --set full --clock-control none --import-source on -k regex:trace_kernel -s $2 -c 1 -o gpurun_out/r2f_$3 -f python scripts/bench_configs.py $1 > gpurun_out/ncu_$3.log 2>&1
ncu -i gpurun_out/r2f_$3.ncu-rep --page raw --csv > gpurun_out/r2f_$3.raw.csv 2>/dev/null
done
ncu --set full --clock-control none --import-source on -k regex:trace_bwd -s 4 -c 1 -o gpurun_out/r2f_bwd_f32 -f python scripts/bench_configs.py c3grad > gpurun_out/ncu_bwd.log 2>&1
ncu -i gpurun_out/r2f_bwd_f32.ncu-rep --page raw --csv > gpurun_out/r2f_bwd_f32.raw.csv 2>/dev/null
ls -la gpurun_out/r2f_*.raw.csv
