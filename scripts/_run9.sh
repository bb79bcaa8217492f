mkdir -p gpurun_out
for w in optic group changed; do python scripts/profile_small_trace.py $w > gpurun_out/r2_profile_small_$w.txt 2>&1; head -60 gpurun_out/r2_profile_small_$w.txt | cut -c1-160; done
python scripts/bench_configs.py c3 c3grad > gpurun_out/r2_configs_d.jsonl 2>&1; cat gpurun_out/r2_configs_d.jsonl | cut -c1-500
python -m pytest tests/test_gpu_parity.py tests/test_gpu_fuzz.py tests/test_config_shapes.py -m gpu -q -p no:cacheprovider -x 2>&1 | tail -3
