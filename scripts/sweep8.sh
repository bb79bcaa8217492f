#!/bin/bash
# Round-1 sweep 8: Newton solve inlined vs out-of-line call, 1 vs 2 rays/thread.
for case in telephoto_c3_tol1e-6 zernike_fringe cheb_biconic_toroidal aspheric_singlet; do
  for lib in inl call call_m2; do for rpt in 1 2; do
    OLB_LIB=$PWD/build/variants/libolb_$lib.so OLB_FORCE_RPT=$rpt python scripts/tune_kernel.py $case 4e6
  done; done
done
