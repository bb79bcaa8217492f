#!/bin/bash
# Round-1 sweep 7: rays/thread and register budget of the Newton-capable fp32/fp64 kernels.
for case in telephoto_c3_tol1e-6 zernike_fringe cheb_biconic_toroidal aspheric_singlet; do
  run() { lib=$1; rpt=$2
    if [ "$lib" = default ]; then unset OLB_LIB; else export OLB_LIB=$PWD/$lib; fi
    if [ $rpt = 0 ]; then unset OLB_FORCE_RPT; else export OLB_FORCE_RPT=$rpt; fi
    python scripts/tune_kernel.py $case 4e6; }
  run default 0; run default 1
  run build/variants/libolb_nrpt4.so 4
  run build/variants/libolb_m2.so 2; run build/variants/libolb_m2.so 1
done
