"""Drop-in check at the scale of the reference's OWN test-suite: selected reference test files
(/root/reference/tests/*.py, torch backend, fp64, grad mode on -- the reference's conftest) are run
twice in a subprocess, stock and with the optiland_b200 plugin installed over the test-only oracle
engine, and must give identical outcomes while the capability is really exercised.  These files hold
the reference's golden numbers for the consumers of the path (wavefront OPD RMS, Hubble operands,
spot references, ray/record semantics).  scripts/ref_sweep.sh runs the long ones (test_analysis.py:
1280 capability calls, identical results)."""
import os
import re
import subprocess
import sys

import pytest

from oracle.ref_import import reference_available

pytestmark = pytest.mark.skipif(not reference_available(), reason="reference not present on this box")
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
FILES = ["test_surface_group.py", "test_wavefront.py", "analysis/test_spot_reference.py"]  # scripts/ref_sweep.sh: all


def _run(fname, install, nograd=False):
    env = dict(os.environ, OLB_SWEEP_INSTALL="1" if install else "0", PYTHONPATH=ROOT,
               OLB_SWEEP_NOGRAD="1" if nograd else "0")
    os.makedirs("/tmp/olb_sweep_root", exist_ok=True)
    out = subprocess.run(
        [sys.executable, "-m", "pytest", "-p", "oracle.sweep_plugin", "-p", "no:cacheprovider", "-q", "--no-header",
         "--rootdir=/tmp/olb_sweep_root", "-c", "/dev/null", f"/root/reference/tests/{fname}",
         "-k", "torch and not view and not draw and not plot"],
        cwd="/tmp/olb_sweep_root", env=env, capture_output=True, text=True, timeout=600).stdout
    counts = {k: int(v) for v, k in re.findall(r"(\d+) (passed|failed|error)", out)}
    calls = re.search(r"capability calls: (\d+) \(differentiable: (\d+), fused launch: (\d+)\)", out)
    return counts, tuple(int(v) for v in calls.groups()) if calls else (0, 0, 0)


@pytest.mark.timeout(900)
@pytest.mark.parametrize("fname", FILES)
def test_reference_tests_unchanged_with_plugin(fname):
    stock, _ = _run(fname, install=False)
    ours, calls = _run(fname, install=True)
    assert stock.get("passed", 0) > 0
    assert ours == stock, (fname, stock, ours)
    assert calls[0] > 0 and calls[1] > 0, "the (differentiable) capability was never exercised"


@pytest.mark.timeout(900)
@pytest.mark.parametrize("fname", ["analysis/test_spot_reference.py", "test_wavefront.py"])
def test_reference_tests_unchanged_with_plugin_grad_mode_off(fname):
    """Same with be.grad_mode left off (the reference's conftest normally turns it on): now the plain trace
    and the fused in-kernel launch generation (RealRayTracer.trace wrapper) carry the calls; for
    test_wavefront.py the reference's OPD goldens are then computed by the fused wavefront epilogue."""
    stock, _ = _run(fname, install=False, nograd=True)
    ours, calls = _run(fname, install=True, nograd=True)
    assert stock.get("passed", 0) > 0 and ours == stock, (fname, stock, ours)
    assert calls[1] == 0 and calls[2] > 0, calls
