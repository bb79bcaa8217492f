#!/bin/bash
# Round 2, second session: verification of what this session added, on one B200 (run under gpurun; short on purpose --
# 14 GPU-minutes were left): the new GPU tests (FFT-PSF gridding kernels, Chebyshev adjoint, mixed-wavelength
# differentiable trace, the plugin's FFTPSF path), smoke(), the FFT-gridding timing, then the reference's own
# test_fft_psf.py on the device stock vs. plugin.
mkdir -p gpurun_out
timeout 330 python -m pytest -m gpu -q -p no:cacheprovider -x \
  tests/test_psf.py \
  "tests/test_gpu_parity.py::test_polynomial_family_adjoint_kernel" \
  "tests/test_gpu_parity.py::test_autograd_ray_input_gradients_and_unsupported_tables" \
  tests/test_plugin_reference.py -k "test_psf or adjoint_kernel or unsupported_tables or fft_psf or several_wavelengths or zernike_and_polynomial or autograd_through or huygens" \
  > gpurun_out/r2b_gputests_new.log 2>&1; echo "rc=$?" >> gpurun_out/r2b_gputests_new.log; tail -8 gpurun_out/r2b_gputests_new.log
timeout 120 python scripts/bench_fft_psf.py > gpurun_out/r2b_fft_psf.jsonl 2> gpurun_out/r2b_fft_psf.err; tail -3 gpurun_out/r2b_fft_psf.err; cat gpurun_out/r2b_fft_psf.jsonl
timeout 90 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2b_smoke.log 2>&1; tail -3 gpurun_out/r2b_smoke.log
rm -f gpurun_out/ref_sweep_cuda.txt
timeout 300 python -m pytest -m gpu -q -p no:cacheprovider tests/test_reference_sweep.py -k "cuda_engine and fft_psf" > gpurun_out/r2b_sweep_fft.log 2>&1; echo "rc=$?" >> gpurun_out/r2b_sweep_fft.log; tail -4 gpurun_out/r2b_sweep_fft.log; cat gpurun_out/ref_sweep_cuda.txt
