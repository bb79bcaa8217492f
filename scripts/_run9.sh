mkdir -p gpurun_out
python scripts/host_overhead_live.py > gpurun_out/r2_host_overhead_live.json 2> gpurun_out/r2_host_overhead_live.err; tail -2 gpurun_out/r2_host_overhead_live.err; cat gpurun_out/r2_host_overhead_live.json
for w in optic changed; do python scripts/profile_small_trace.py $w > gpurun_out/r2_profile_small_$w.txt 2>&1; head -22 gpurun_out/r2_profile_small_$w.txt | cut -c1-150; done
