mkdir -p gpurun_out
python -m pytest tests/test_gpu_parity.py tests/test_gpu_fuzz.py tests/test_config_shapes.py tests/test_wavefront.py -m gpu -q -p no:cacheprovider -x > gpurun_out/r2_gputests_5.log 2>&1; echo "rc=$?" >> gpurun_out/r2_gputests_5.log
tail -4 gpurun_out/r2_gputests_5.log
python scripts/bench_configs.py zern c5pol c5shape > gpurun_out/r2_configs_c.jsonl 2> gpurun_out/r2_configs_c.err; tail -3 gpurun_out/r2_configs_c.err; cat gpurun_out/r2_configs_c.jsonl
