mkdir -p gpurun_out
python -m pytest tests/test_gpu_parity.py -m gpu -q -p no:cacheprovider -k "unsupported_tables or optiland_shaped_rays or polynomial_family" 2>&1 | tail -4
python scripts/profile_c3_step.py > gpurun_out/r2_c3_step_breakdown.json 2> gpurun_out/c3step.err; tail -3 gpurun_out/c3step.err; cat gpurun_out/r2_c3_step_breakdown.json
ncu --set full --clock-control none --import-source on -k regex:trace_bwd -s 4 -c 1 -o /tmp/bwd32 -f python scripts/bench_configs.py c3grad > gpurun_out/ncu_bwd32.log 2>&1
ncu -i /tmp/bwd32.ncu-rep --page source --csv --print-source cuda,sass > gpurun_out/r2g_bwd_f32.src.csv 2>/dev/null
ls -la gpurun_out/r2g_bwd_f32.src.csv
