"""Host logic of the plugin's packing fast path (optiland_b200/pack.py): one stacked device->host copy for
all scalar parameters and the direct pose read must give byte-identical tables to the scalar-by-scalar path."""
import types

import numpy as np
import pytest
import torch

from optiland_b200 import pack as PK
from optiland_b200 import table as T
from tests._fake_optiland import fake_surfaces
from tests._util import Case

CASES = ["cooke_c1", "dgauss_c2", "dgauss_multiwl", "telephoto_c3_tol1e-6", "hubble_c4"]


def _group(table, device, angles=None):
    return types.SimpleNamespace(surfaces=fake_surfaces(table, device, angles))


def _same(a: T.SurfaceTable, b: T.SurfaceTable):
    sa, pa = a.pack()
    sb, pb = b.pack()
    assert sa.tobytes() == sb.tobytes() and pa.tobytes() == pb.tobytes()
    np.testing.assert_array_equal(a.wavelengths, b.wavelengths)


@pytest.mark.parametrize("name", CASES)
@pytest.mark.parametrize("force", [False, True])
def test_fake_optiland_objects_pack_to_the_golden_table(name, force, monkeypatch):
    c = Case(name)
    monkeypatch.setattr(PK._Prefetch, "force", force)
    PK._Prefetch.last_count = 0
    got = PK.pack_surface_group(_group(c.table, "cpu"), c.table.wavelengths)
    _same(got, c.table)
    if force:
        assert PK._Prefetch.last_count >= 6 * (c.table.num_surfaces - 1)   # poses, radii, conics, indices
    assert getattr(PK._tls, "resolved", None) is None                      # nothing leaks out of the context


def test_tilted_pose_is_read_directly(monkeypatch):
    c = Case("cooke_c1")
    ang = {2: (0.01, -0.02, 0.3), 4: (0.0, 0.0, 1e-3)}
    for force in (False, True):
        monkeypatch.setattr(PK._Prefetch, "force", force)
        got = PK.pack_surface_group(_group(c.table, "cpu", ang), c.table.wavelengths)
        for s, spec in enumerate(got.surfaces):
            if s in ang:
                np.testing.assert_array_equal(spec.R, T.rotation_matrix(*ang[s]) + 0.0)
            elif spec.kind != T.GEOM_NOOP:
                np.testing.assert_array_equal(spec.R, np.eye(3))
            np.testing.assert_array_equal(spec.t, c.table.surfaces[s].t)


@pytest.mark.gpu
@pytest.mark.parametrize("name", CASES)
def test_cuda_resident_parameters_are_fetched_in_one_copy(name):
    c = Case(name)
    PK._Prefetch.last_count = 0
    got = PK.pack_surface_group(_group(c.table, "cuda"), c.table.wavelengths)
    _same(got, c.table)
    assert PK._Prefetch.last_count >= 6 * (c.table.num_surfaces - 1)
    # and the packed table traces like the golden one through the plugin's engine
    from optiland_b200.plugin import CudaEngine

    eng = CudaEngine()
    rays = types.SimpleNamespace(**{k: torch.from_numpy(c.rays[k]).cuda() for k in ("x", "y", "z", "L", "M", "N", "i", "w")})
    rays.opd = torch.zeros_like(rays.x)
    rec = eng.trace(got, rays, 0, got.num_surfaces)
    np.testing.assert_allclose(rec["x"].cpu().numpy(), c.rec["x"], rtol=0, atol=1e-11 * c.scale, equal_nan=True)


@pytest.mark.parametrize("name", ["telephoto_c3_tol1e-6", "dgauss_c2", "hubble_c4", "tilted_fold", "misc_apertures_coatings",
                                  "zernike_fringe", "aspheric_singlet", "cooke_polarized"])
def test_vectorised_parameter_packer_equals_params_to_table(name):
    """autograd._ParamPacker (the per-step host path of the differentiable trace) writes the parameter values into the
    template's packed arrays: byte-identical to rebuilding the table surface by surface."""
    from optiland_b200 import autograd as AG

    c = Case(name)
    if c.table.n_wl != 1:
        pytest.skip("one wavelength per differentiable trace")
    rng = np.random.default_rng(3)
    p0 = AG.table_to_params(c.table)
    for trial in range(3):
        p = p0.clone()
        if trial:
            p = p * torch.from_numpy(1 + 1e-3 * rng.standard_normal(p.shape)) + torch.from_numpy(1e-4 * rng.standard_normal(p.shape))
            p[:, AG.GP_CURV][rng.random(len(p)) < 0.2] = 0.0        # flat surfaces: radius = inf
        slow = AG.params_to_table(c.table, p)
        sa, pa = slow.pack()
        sb, pb = AG._packed_from_params(c.table, p)
        assert sa.tobytes() == sb.tobytes()
        assert pa.tobytes() == pb.tobytes()


def test_cuda_engine_accepts_rule_on_host_tensors(monkeypatch):
    """``CudaEngine.accepts`` (which rays may be handed to the kernels) exercised without a GPU: the residency predicate
    is swapped for one that takes host tensors.  The rule that round 2's first GPU run of the ray-aimer test uncovered --
    ``RealRays(..., wavelength=0.55)`` keeps a 1-element ``w`` (real_rays.py:79), which used to be rejected, so every
    trace of the iterative aimer silently went back to the reference's eager ops on a GPU -- is pinned here."""
    from optiland_b200 import plugin as P

    monkeypatch.setattr(P.CudaEngine, "_on_device", staticmethod(lambda t: True))
    eng = P.CudaEngine()
    n = 61

    def rays(**over):
        r = {k: torch.zeros(n, dtype=torch.float64) for k in ("x", "y", "z", "L", "M", "N", "i", "w", "opd")}
        r.update(over)
        return types.SimpleNamespace(**r)

    assert eng.accepts(rays())
    assert eng.accepts(rays(w=torch.tensor([0.55], dtype=torch.float64)))              # one wavelength for the batch
    assert not eng.accepts(rays(w=torch.tensor(0.55, dtype=torch.float64)))            # 0-d: the reference makes it 1-D
    assert not eng.accepts(rays(w=torch.zeros(7, dtype=torch.float64)))                # neither 1 nor n values
    assert not eng.accepts(rays(w=torch.tensor([0.55], dtype=torch.float32)))          # mixed precision
    assert not eng.accepts(rays(i=torch.zeros(1, dtype=torch.float64)))                # only w may be broadcast
    assert not eng.accepts(rays(x=np.zeros(n)))                                        # NumPy-resident rays
    assert not eng.accepts(rays(x=torch.zeros(n, dtype=torch.float16)))
    assert not eng.accepts(types.SimpleNamespace(x=torch.zeros(n)))                    # ParaxialRays-like objects
    monkeypatch.setattr(P.CudaEngine, "_on_device", staticmethod(lambda t: t.is_cuda))
    assert not P.CudaEngine().accepts(rays())                                          # host tensors, real predicate
