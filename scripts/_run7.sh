mkdir -p gpurun_out
rm -f gpurun_out/ref_sweep_cuda.txt
python -m pytest tests -m gpu -q -p no:cacheprovider --durations=8 > gpurun_out/r2_gputests_full.log 2>&1; echo "rc=$?" >> gpurun_out/r2_gputests_full.log
tail -25 gpurun_out/r2_gputests_full.log
cat gpurun_out/ref_sweep_cuda.txt
