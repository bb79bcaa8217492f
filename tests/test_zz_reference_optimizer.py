"""The reference's OWN torch optimiser (optimization/optimizer/torch/base.py:96-156: ``TorchAdamOptimizer.optimize`` ->
``var.update(param)`` -> ``problem.update_optics()`` -> ``problem.sum_squared()`` -> ``loss.backward()`` -> ``step()``),
unchanged, with real-ray operands, over the plugin: every merit-function evaluation traces through the capability (forward
kernel + adjoint kernel behind one autograd Function) and the optimiser reaches the same iterates as over the reference's
eager graph.

``[oracle]``: CPU, test-only oracle engine + the CPU instantiation of the adjoint.  ``[cuda]``: the product engine on the
B200.  The file sorts last because its ``[cuda]`` variant was written after the round's GPU budget had been spent (the paths
it drives -- ``trace_grad`` with live objects -- are exercised on hardware by tests/test_plugin_reference.py); if the STOCK
reference's optimiser cannot run on a CUDA device by itself, that variant skips and says why."""
import numpy as np
import pytest

from oracle.ref_import import reference_available
from tests.test_plugin_reference import plugin  # noqa: F401  (fixture)

pytestmark = pytest.mark.skipif(not reference_available(), reason="reference not present on this box")


def _problem(be):
    from optiland.optimization import OptimizationProblem
    from optiland.samples.objectives import CookeTriplet

    lens = CookeTriplet()
    r1 = float(np.asarray(be.to_numpy(lens.surfaces.surfaces[1].geometry.radius)).reshape(-1)[0])
    lens.updater.set_radius(r1 * 1.04, 1)            # a detuned start: sum of squares 0.148, six Adam steps bring it to 1.2e-3
    problem = OptimizationProblem()
    problem.add_variable(lens, "radius", surface_number=1)
    problem.add_variable(lens, "radius", surface_number=4)
    problem.add_variable(lens, "conic", surface_number=2)
    problem.add_variable(lens, "thickness", surface_number=3)
    for hy in (0.0, 0.7, 1.0):
        problem.add_operand(operand_type="rms_spot_size", target=0.0, weight=1.0,
                            input_data={"optic": lens, "surface_number": -1, "Hx": 0.0, "Hy": hy, "num_rays": 5,
                                        "wavelength": 0.55, "distribution": "hexapolar"})
    problem.update_optics()
    return problem, lens


def _optimise(be, steps=6):
    from optiland.optimization import TorchAdamOptimizer

    problem, lens = _problem(be)
    start = float(be.to_numpy(problem.sum_squared()))
    losses = []
    res = TorchAdamOptimizer(problem).optimize(n_steps=steps, lr=1e-3, disp=False, callback=lambda i, v: losses.append(v))
    values = [float(np.asarray(be.to_numpy(v.value)).reshape(-1)[0]) for v in problem.variables]
    return {"start": start, "losses": losses, "final": float(res.fun), "x": [float(v) for v in res.x], "values": values}


def test_reference_torch_optimizer_runs_unchanged_over_the_capability(plugin):
    P, eng, be = plugin
    on_device = type(eng).__name__ == "CudaEngine"
    P.uninstall()
    try:
        ref = _optimise(be)                          # the reference's own eager graph (same backend / device)
    except Exception as e:                           # noqa: BLE001
        if on_device:
            pytest.skip(f"the stock reference's optimiser does not run on a CUDA device by itself: {type(e).__name__}: {e}")
        raise
    finally:
        P.install(engine=eng)
    P.stats(reset=True)
    n0 = len(eng.calls)
    got = _optimise(be)
    grads = [c for c in eng.calls[n0:] if c[0] == "grad"]
    # 6 steps x 3 real-ray operands (+ the evaluations before and after the loop): every one a differentiable trace
    assert len(grads) >= 6 * 3, (len(grads), P.stats())
    assert not any(k.startswith("gradients wanted") or k.startswith("unsupported") for k in P.stats()), P.stats()
    assert got["start"] == pytest.approx(ref["start"], rel=1e-9)
    assert ref["final"] < 0.05 * ref["start"]        # the optimiser really optimised (0.148 -> 1.2e-3)
    assert got["final"] == pytest.approx(ref["final"], rel=1e-6)
    np.testing.assert_allclose(got["losses"], ref["losses"], rtol=1e-6)
    np.testing.assert_allclose(got["x"], ref["x"], rtol=1e-6, atol=1e-9)
    np.testing.assert_allclose(got["values"], ref["values"], rtol=1e-6, atol=1e-9)
