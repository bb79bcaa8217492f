"""pytest plugin (TEST INFRASTRUCTURE) used to run the REFERENCE's own test files with the
optiland_b200 plugin installed: `-p oracle.sweep_plugin`.

It (1) makes the reference importable (stubs for matplotlib / vtk / seaborn), and (2) when
OLB_SWEEP_INSTALL=1 installs `optiland_b200.plugin` -- with the TEST-ONLY oracle engine of
oracle/oracle_engine.py in the GPU-less build container, or, with OLB_SWEEP_DEVICE=cuda (the B200 box), with the
product `CudaEngine` and the torch backend moved to the device -- so that every `SurfaceGroup.trace` / `Surface.trace` issued by the
reference's tests on the torch backend goes through the capability (packing of live objects, record
hand-back, autograd Function, declines).  The sweep compares pass/fail sets with and without it.
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from oracle.ref_import import import_reference  # noqa: E402

import_reference()

if os.environ.get("OLB_SWEEP_NOGRAD") == "1":
    # the reference's conftest switches grad mode ON for the torch backend; with this switch it stays off
    # (in both the stock and the plugin run) so the non-differentiable capability paths -- plain trace
    # and in-kernel launch generation -- are the ones exercised
    from optiland.backend.torch_backend import GradMode as _GradMode

    _GradMode.enable = lambda self: None

CUDA = os.environ.get("OLB_SWEEP_DEVICE") == "cuda"
if CUDA:
    # the reference's conftest pins the torch backend to the CPU (tests/conftest.py:15); on the GPU box the sweep
    # runs BOTH arms (stock and plugin) on the device instead, so that the plugin arm exercises the real kernels
    from optiland.backend.torch_backend import TorchBackend as _TB

    _orig_set_device = _TB.set_device
    _TB.set_device = lambda self, device: _orig_set_device(self, "cuda")

ENGINE = None
if os.environ.get("OLB_SWEEP_INSTALL") == "1":
    from optiland_b200 import plugin as _P

    if CUDA:
        ENGINE = _P.CudaEngine()          # the product engine: libolb.so
    else:
        from oracle.oracle_engine import OracleEngine

        ENGINE = OracleEngine()
    _P.install(engine=ENGINE)


def pytest_terminal_summary(terminalreporter):
    if ENGINE is not None:
        n_grad = sum(1 for c in ENGINE.calls if c and c[0] == "grad")
        n_pupil = sum(1 for c in ENGINE.calls if c and c[0] in ("pupil", "wavefront"))
        extra = ""
        if CUDA:
            from optiland_b200 import _lib

            extra = f", libolb launches: {_lib.load().olb_launch_count()}"
        terminalreporter.write_line(f"[olb sweep] capability calls: {len(ENGINE.calls)} (differentiable: {n_grad}, "
                                    f"fused launch: {n_pupil}){extra}")
        from optiland_b200 import plugin as _P2

        why = _P2.stats()
        if why:
            terminalreporter.write_line("[olb sweep] declines: " + "; ".join(f"{k} x{v}" for k, v in sorted(why.items())))
