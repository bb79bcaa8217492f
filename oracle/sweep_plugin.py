"""pytest plugin (TEST INFRASTRUCTURE) used to run the REFERENCE's own test files with the
optiland_b200 plugin installed: `-p oracle.sweep_plugin`.

It (1) makes the reference importable (stubs for matplotlib / vtk / seaborn), and (2) when
OLB_SWEEP_INSTALL=1 installs `optiland_b200.plugin` with the TEST-ONLY oracle engine of
oracle/oracle_engine.py, so that every `SurfaceGroup.trace` / `Surface.trace` issued by the
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

ENGINE = None
if os.environ.get("OLB_SWEEP_INSTALL") == "1":
    from optiland_b200 import plugin as _P
    from oracle.oracle_engine import OracleEngine

    ENGINE = OracleEngine()
    _P.install(engine=ENGINE)


def pytest_terminal_summary(terminalreporter):
    if ENGINE is not None:
        n_grad = sum(1 for c in ENGINE.calls if c and c[0] == "grad")
        n_pupil = sum(1 for c in ENGINE.calls if c and c[0] in ("pupil", "wavefront"))
        terminalreporter.write_line(f"[olb sweep] capability calls: {len(ENGINE.calls)} (differentiable: {n_grad}, "
                                    f"fused launch: {n_pupil})")
