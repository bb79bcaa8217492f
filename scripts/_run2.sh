set -x
mkdir -p gpurun_out
python -m pytest tests -m gpu -q -p no:cacheprovider -k "not cuda_engine" -x > gpurun_out/r2_gputests_2.log 2>&1; echo "rc=$?" >> gpurun_out/r2_gputests_2.log
tail -30 gpurun_out/r2_gputests_2.log
python scripts/f32_achieved.py --gpu > gpurun_out/f32_achieved_gpu.log 2>&1; tail -3 gpurun_out/f32_achieved_gpu.log
python bench.py --steps 20 --warmup 5 > gpurun_out/r2_bench_n1.json 2> gpurun_out/r2_bench_n1.err; tail -5 gpurun_out/r2_bench_n1.err; cat gpurun_out/r2_bench_n1.json
python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/r2_bench_ref.json 2> gpurun_out/r2_bench_ref.err; tail -3 gpurun_out/r2_bench_ref.err; cat gpurun_out/r2_bench_ref.json
