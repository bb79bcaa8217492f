mkdir -p gpurun_out
nvidia-smi -L | head -3
python -m pytest tests/test_gpu_multi.py -m gpu -q -p no:cacheprovider > gpurun_out/r2_multi_gpu_test.log 2>&1; echo "rc=$?" >> gpurun_out/r2_multi_gpu_test.log
tail -5 gpurun_out/r2_multi_gpu_test.log; cat gpurun_out/multi_gpu_correctness.json
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/r2_bench_n2.json 2> gpurun_out/r2_bench_n2.err; tail -5 gpurun_out/r2_bench_n2.err; cat gpurun_out/r2_bench_n2.json | cut -c1-400
