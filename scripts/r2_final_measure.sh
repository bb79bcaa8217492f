#!/bin/bash
# Round-2 final measurement on one B200 (run under gpurun): the whole -m gpu suite, both bench arms, the per-config
# kernel timings, the ncu launch list of the bench command and one `ncu --set full` capture per headline kernel.
# Only the exported CSV pages are kept (gpurun_out/ is capped at 64 MiB); scripts/collect_r2_profiles.py turns them
# into the tracked files under profiles/.
mkdir -p gpurun_out
python -m pytest tests -m gpu -q -p no:cacheprovider > gpurun_out/r2_gputests_final.log 2>&1; echo "rc=$?" >> gpurun_out/r2_gputests_final.log; tail -6 gpurun_out/r2_gputests_final.log
python bench.py --steps 20 --warmup 5 > gpurun_out/r2_bench_n1.json 2> gpurun_out/r2_bench_n1.err; tail -2 gpurun_out/r2_bench_n1.err; cut -c1-300 gpurun_out/r2_bench_n1.json
python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/r2_bench_ref.json 2>/dev/null
python scripts/bench_configs.py > gpurun_out/r2_configs_final.jsonl 2> gpurun_out/r2_configs_final.err; cut -c1-330 gpurun_out/r2_configs_final.jsonl
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_launches_bench.csv python bench.py --steps 5 --warmup 3 --no-optic-trace --no-cpu-reference > gpurun_out/ncu_bench.log 2>&1
for spec in "c2 8 c2_f32" "c2 31 c2_f64" "zern 8 zern_f32" "zern 31 zern_f64" "c5pol 8 c5pol_f32" "c5pol 31 c5pol_f64"; do set -- $spec
ncu --set full --clock-control none --import-source on -k regex:trace_kernel -s $2 -c 1 -o /tmp/r2f_$3 -f python scripts/bench_configs.py $1 > gpurun_out/ncu_$3.log 2>&1
ncu -i /tmp/r2f_$3.ncu-rep --page raw --csv > gpurun_out/r2f_$3.raw.csv 2>/dev/null
done
for spec in "4 bwd_f32" "16 bwd_f64"; do set -- $spec
ncu --set full --clock-control none --import-source on -k regex:trace_bwd -s $1 -c 1 -o /tmp/r2f_$2 -f python scripts/bench_configs.py c3grad > gpurun_out/ncu_$2.log 2>&1
ncu -i /tmp/r2f_$2.ncu-rep --page raw --csv > gpurun_out/r2f_$2.raw.csv 2>/dev/null
done
ls -la gpurun_out/r2f_*.raw.csv; du -sh gpurun_out
