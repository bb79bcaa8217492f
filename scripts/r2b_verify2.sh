#!/bin/bash
# Round 2, second session, second (short) GPU call: the two plugin tests the first call's time limit cut off, and the
# traceback of the one reference test (test_fft_psf.py::test_invalid_working_FNO) that differed with the CUDA engine.
mkdir -p gpurun_out
timeout 140 python -m pytest -m gpu -q -p no:cacheprovider --durations=5 tests/test_plugin_reference.py -k "fft_psf or huygens or ray_aimers" > gpurun_out/r2b_gputests_new2.log 2>&1; echo "rc=$?" >> gpurun_out/r2b_gputests_new2.log; tail -12 gpurun_out/r2b_gputests_new2.log
mkdir -p /tmp/olb_sweep_root; cd /tmp/olb_sweep_root
OLB_SWEEP_INSTALL=1 OLB_SWEEP_DEVICE=cuda OLB_SWEEP_NOGRAD=1 PYTHONPATH=$GRAFT_REPO_ROOT timeout 100 python -m pytest -p oracle.sweep_plugin -p no:cacheprovider -q --no-header -rfE --rootdir=/tmp/olb_sweep_root -c /dev/null $GRAFT_REPO_ROOT/oracle/_ref/tests/test_fft_psf.py -k "test_invalid_working_FNO and torch" > $GRAFT_REPO_ROOT/gpurun_out/r2b_invalid_fno.log 2>&1
tail -60 $GRAFT_REPO_ROOT/gpurun_out/r2b_invalid_fno.log
