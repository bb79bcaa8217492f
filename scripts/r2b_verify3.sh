#!/bin/bash
# Round 2, second session, last (very short) GPU call: the two fixes the second call's findings led to --
# (1) a batch with ONE wavelength value (RealRays(..., wavelength=0.55): the iterative aimer) is accepted by the CUDA
# engine, so the aimer's subset traces run on the kernel; (2) a NaN chief-ray reference sphere declines to the reference.
mkdir -p gpurun_out /tmp/olb_sweep_root
(cd /tmp/olb_sweep_root; OLB_SWEEP_INSTALL=1 OLB_SWEEP_DEVICE=cuda OLB_SWEEP_NOGRAD=1 PYTHONPATH=$GRAFT_REPO_ROOT timeout 60 python -m pytest -p oracle.sweep_plugin -p no:cacheprovider -q --no-header -rfE --rootdir=/tmp/olb_sweep_root -c /dev/null $GRAFT_REPO_ROOT/oracle/_ref/tests/test_fft_psf.py -k "test_invalid_working_FNO and torch" > $GRAFT_REPO_ROOT/gpurun_out/r2b_invalid_fno2.log 2>&1; tail -4 $GRAFT_REPO_ROOT/gpurun_out/r2b_invalid_fno2.log) &
timeout 75 python -m pytest -m gpu -q -p no:cacheprovider tests/test_plugin_reference.py -k "ray_aimers" > gpurun_out/r2b_gputests_new3.log 2>&1; echo "rc=$?" >> gpurun_out/r2b_gputests_new3.log; tail -25 gpurun_out/r2b_gputests_new3.log
wait
