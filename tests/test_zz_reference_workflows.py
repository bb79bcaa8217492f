"""Whole reference WORKFLOWS, unchanged, over the plugin -- the consumers SURVEY.md 8(f-4) names -- against the same workflow
over the reference's own eager path (same backend, same seeds):

* tolerancing Monte-Carlo (tolerancing/monte_carlo.py:59-123) with real-ray operands, a perturbed radius and tilt and a
  thickness compensator (each sample: perturb -> an optimiser run for the compensator -> operands);
* sensitivity analysis (tolerancing/sensitivity_analysis.py) over the same problem;
* the SciPy-driven ``OptimizerGeneric`` (optimization/optimizer/scipy/base.py) on a real-ray merit function.

``[oracle]`` on the CPU; ``[cuda]`` on the B200 (product engine).  The file sorts last: its ``[cuda]`` variants were written
after the round's GPU budget was spent -- they drive only code paths the earlier files verify on hardware -- and skip, saying
why, if the STOCK reference cannot run the workflow on a CUDA device by itself."""
import numpy as np
import pytest

from oracle.ref_import import reference_available
from tests.test_plugin_reference import plugin  # noqa: F401  (fixture)

pytestmark = pytest.mark.skipif(not reference_available(), reason="reference not present on this box")


def _both(plugin, workflow):
    """(stock result, plugin result, engine calls of the plugin arm)."""
    P, eng, be = plugin
    on_device = type(eng).__name__ == "CudaEngine"
    P.uninstall()
    try:
        ref = workflow(be)
    except Exception as e:                           # noqa: BLE001
        if on_device:
            pytest.skip(f"the stock reference does not run this workflow on a CUDA device by itself: {type(e).__name__}: {e}")
        raise
    finally:
        P.install(engine=eng)
    P.stats(reset=True)
    n0 = len(eng.calls)
    got = workflow(be)
    return ref, got, eng.calls[n0:]


def _tolerancing(be, ranges=False):
    from optiland.samples.objectives import CookeTriplet
    from optiland.tolerancing.core import Tolerancing
    from optiland.tolerancing.perturbation import DistributionSampler, RangeSampler

    optic = CookeTriplet()
    tol = Tolerancing(optic)
    for hy in (0.0, 1.0):
        tol.add_operand(operand_type="rms_spot_size",
                        input_data={"optic": optic, "surface_number": -1, "Hx": 0.0, "Hy": hy, "num_rays": 4,
                                    "wavelength": 0.55, "distribution": "hexapolar"})
    tol.add_operand(operand_type="real_y_intercept",
                    input_data={"optic": optic, "surface_number": -1, "Hx": 0.0, "Hy": 1.0, "Px": 0.0, "Py": 0.0, "wavelength": 0.55})
    if ranges:                                       # (the sensitivity analysis steps through ranges)
        tol.add_perturbation("radius", RangeSampler(21.9, 22.1, 3), surface_number=1)
        tol.add_perturbation("tilt", RangeSampler(-2e-3, 2e-3, 3), surface_number=3, axis="x")
    else:
        tol.add_perturbation("radius", DistributionSampler("normal", seed=11, loc=22.01359, scale=0.05), surface_number=1)
        tol.add_perturbation("tilt", DistributionSampler("uniform", seed=12, low=-2e-3, high=2e-3), surface_number=3, axis="x")
    tol.add_compensator("thickness", surface_number=6)
    return tol


def test_monte_carlo_tolerancing_over_the_capability(plugin):
    def workflow(be):
        from optiland.tolerancing.monte_carlo import MonteCarlo

        mc = MonteCarlo(_tolerancing(be))
        mc.run(5)
        return mc.get_results().to_numpy(dtype=float)

    ref, got, calls = _both(plugin, workflow)
    assert ref.shape == got.shape == (5, 6)          # 2 perturbations, 3 operands, 1 compensator
    assert len(calls) >= 5 * 3                       # every operand of every sample was served by the capability
    np.testing.assert_allclose(got, ref, rtol=2e-6, atol=1e-9)   # (the compensator is an optimiser run: its stopping
    # point moves with the last digits of the merit function)


def test_sensitivity_analysis_over_the_capability(plugin):
    def workflow(be):
        from optiland.tolerancing.sensitivity_analysis import SensitivityAnalysis

        sa = SensitivityAnalysis(_tolerancing(be, ranges=True))
        sa.run()
        return sa.get_results().select_dtypes("number").to_numpy(dtype=float)

    ref, got, calls = _both(plugin, workflow)
    assert ref.shape == got.shape and ref.size > 0 and len(calls) >= 6
    np.testing.assert_allclose(got, ref, rtol=2e-6, atol=1e-9)


def test_scipy_optimizer_on_a_real_ray_merit_function_over_the_capability(plugin):
    def workflow(be):
        from optiland.optimization import OptimizationProblem, OptimizerGeneric
        from optiland.samples.objectives import CookeTriplet

        lens = CookeTriplet()
        r1 = float(np.asarray(be.to_numpy(lens.surfaces.surfaces[1].geometry.radius)).reshape(-1)[0])
        lens.updater.set_radius(r1 * 1.03, 1)
        problem = OptimizationProblem()
        problem.add_variable(lens, "radius", surface_number=1)
        problem.add_variable(lens, "thickness", surface_number=6)
        for hy in (0.0, 0.7, 1.0):
            problem.add_operand(operand_type="rms_spot_size", target=0.0, weight=1.0,
                                input_data={"optic": lens, "surface_number": -1, "Hx": 0.0, "Hy": hy, "num_rays": 4,
                                            "wavelength": 0.55, "distribution": "hexapolar"})
        problem.update_optics()
        start = float(np.asarray(be.to_numpy(problem.sum_squared())).reshape(-1)[0])
        res = OptimizerGeneric(problem).optimize(maxiter=15, disp=False, tol=1e-9)
        return np.array([start, float(res.fun)] + [float(np.asarray(be.to_numpy(v.value)).reshape(-1)[0]) for v in problem.variables])

    ref, got, calls = _both(plugin, workflow)
    assert ref[1] < 0.2 * ref[0]                     # it optimised
    assert len(calls) >= 15
    assert got[0] == pytest.approx(ref[0], rel=1e-9)
    np.testing.assert_allclose(got, ref, rtol=5e-5)  # (a quasi-Newton path: last digits of f move the line searches)
