#!/bin/bash
# Run a set of the reference's own test files twice (stock, and with the optiland_b200 plugin installed
# over the test-only oracle engine) and compare the outcomes.  Needs /root/reference.
# Runs from a neutral directory so that this repo's `tests` package does not shadow the reference's.
REPO="$(cd "$(dirname "$0")/.." && pwd)"
REFTESTS="$(cd "$REPO" && python -c "from oracle.ref_import import REFERENCE_TESTS as t; print(t)")"
FILES="${@:-test_surface_group.py test_standard_surface.py test_rays.py test_analysis.py test_wavefront.py test_optic.py test_torch_optimization.py test_operand.py}"
mkdir -p /tmp/olb_sweep_root && cd /tmp/olb_sweep_root
for f in $FILES; do
  for mode in 0 1; do
    OLB_SWEEP_INSTALL=$mode PYTHONPATH="$REPO" python -m pytest -p oracle.sweep_plugin -p no:cacheprovider -q --no-header \
      --rootdir=/tmp/olb_sweep_root -c /dev/null "$REFTESTS/$f" -k "torch and not view and not draw and not plot" \
      2>&1 | grep -E "passed|failed|error|olb sweep" | tr '\n' ' '
    echo " <- $f install=$mode"
  done
done
