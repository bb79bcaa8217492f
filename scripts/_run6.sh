mkdir -p gpurun_out
python -m pytest tests/test_gpu_multi.py -m gpu -q -p no:cacheprovider > gpurun_out/r2_multi_gpu_test.log 2>&1; echo "rc=$?" >> gpurun_out/r2_multi_gpu_test.log
tail -5 gpurun_out/r2_multi_gpu_test.log; cat gpurun_out/multi_gpu_correctness.json
