mkdir -p gpurun_out
python -m pytest tests/test_gpu_parity.py tests/test_config_shapes.py -m gpu -q -p no:cacheprovider -k "autograd or adjoint or gradient or grad or shaped_rays" 2>&1 | tail -4
python scripts/profile_c3_step.py > gpurun_out/r2_c3_step_breakdown.json 2> gpurun_out/c3step.err; tail -3 gpurun_out/c3step.err; cat gpurun_out/r2_c3_step_breakdown.json
python scripts/profile_c3_step.py 4000000 zernike_fringe > gpurun_out/r2_zern_step_breakdown.json 2> gpurun_out/zstep.err; tail -3 gpurun_out/zstep.err; cat gpurun_out/r2_zern_step_breakdown.json
python scripts/bench_configs.py c3grad zerngrad
OLB_BWD_ACC=0 python scripts/bench_configs.py c3grad
