"""TEST INFRASTRUCTURE ONLY -- stand-in for optiland_b200.plugin.CudaEngine on boxes without a GPU:
evaluates the packed table with the NumPy oracle (forward) and the CPU instantiation of the device
adjoint (backward).  Used by tests/test_plugin_reference.py and by the reference-test sweep."""
import numpy as np


def raise_status(status: int) -> None:
    """The reference's ValueErrors for out-of-range freeform coordinates -- what optiland_b200.trace._raise_status does with
    the kernels' status word, from EVERY entry point (plain, pupil launch, moments, wavefront)."""
    from optiland_b200 import table as T

    if status & T.ST_ZERNIKE_RANGE:
        raise ValueError("Zernike coordinates must be normalized to [-1, 1]. Consider updating the normalization "
                         "radius to 1.1x the surface aperture.")
    if status & T.ST_CHEBYSHEV_RANGE:
        raise ValueError("Chebyshev input coordinates must be normalized to [-1, 1]. Consider updating the "
                         "normalization factors.")


class OracleEngine:
    """TEST-ONLY stand-in for optiland_b200.plugin.CudaEngine."""

    def __init__(self):
        self.calls = []

    def accepts(self, rays):
        import torch

        return all(torch.is_tensor(getattr(rays, k)) for k in ("x", "y", "z", "L", "M", "N", "i", "w", "opd"))

    def trace(self, table, rays, first, last):
        import torch

        from oracle import trace_oracle as O

        self.calls.append((table.num_surfaces, int(rays.x.numel())))
        inp = {k: getattr(rays, k).detach().double().numpy() for k in ("x", "y", "z", "L", "M", "N", "i", "w", "opd")}
        polarized = type(rays).__name__ == "PolarizedRays"
        if polarized:
            inp["p"] = rays.p.detach().numpy().astype(np.complex128)
        out, rec, status = O.trace(table, inp, first, last, polarized=polarized)
        if polarized:
            rays.p = torch.from_numpy(out["p"])
        raise_status(status)
        dt = rays.x.dtype
        for k in ("x", "y", "z", "L", "M", "N", "i", "opd"):
            setattr(rays, k, torch.from_numpy(out[k]).to(dt))
        return {k: torch.from_numpy(v).to(dt) for k, v in rec.items()}


    def accepts_tensor(self, t):
        import torch

        return torch.is_tensor(t) and t.ndim == 1

    def trace_pupil(self, table, Px, Py, affine, wavelength=None, polarization=False):
        import torch

        from oracle import trace_oracle as O
        from optiland_b200.launch import launch_from_affine

        self.calls.append(("pupil", table.num_surfaces, int(Px.numel())))
        px, py = Px.detach().double().numpy(), Py.detach().double().numpy()
        aff = dict(affine)
        if aff.get("fields") is not None:
            aff["fields"] = tuple(t.detach().double().numpy() for t in aff["fields"])
        x, y, z, L, M, N = launch_from_affine(px, py, aff)
        w = wavelength.detach().double().numpy() if wavelength is not None else np.full_like(px, table.wavelengths[0])
        i0 = np.full_like(px, affine.get("intensity", 1.0))
        inp = dict(x=x, y=y, z=z, L=L, M=M, N=N, i=i0, w=w)
        if polarization is False:
            _, rec, status = O.trace(table, inp)
            raise_status(status)
            return {k: torch.from_numpy(v).to(Px.dtype) for k, v in rec.items()}
        inp["p"] = np.tile(np.eye(3, dtype=np.complex128), (px.size, 1, 1))
        out, rec, status = O.trace(table, inp, polarized=True)
        raise_status(status)
        res = {k: torch.from_numpy(v).to(Px.dtype) for k, v in rec.items()}
        cdt = torch.complex128 if Px.dtype == torch.float64 else torch.complex64
        res["p"] = torch.from_numpy(out["p"]).to(cdt)
        if polarization == "matrix":
            res["i_pol"] = res["intensity"][-1]
        else:
            res["i_pol"] = torch.from_numpy(O.polarized_intensity(out["p"], L, M, N, i0, polarization)).to(Px.dtype)
        return res

    def spot_moments(self, table, Px, Py, affine, center=(0.0, 0.0), last=None, global_xy=False, every_ray=False):
        from oracle import trace_oracle as O
        from optiland_b200.launch import launch_from_affine

        self.calls.append(("moments", table.num_surfaces, int(Px.numel())))
        px, py = Px.detach().double().numpy(), Py.detach().double().numpy()
        x, y, z, L, M, N = launch_from_affine(px, py, affine)
        inp = dict(x=x, y=y, z=z, L=L, M=M, N=N, i=np.full_like(px, affine.get("intensity", 1.0)),
                   w=np.full_like(px, table.wavelengths[0]))
        last = table.num_surfaces if last is None else last
        fin, rec, status = O.trace(table, inp, 0, last)
        raise_status(status)
        gx, gy, ii, oo = rec["x"][-1], rec["y"][-1], rec["intensity"][-1], rec["opd"][-1]
        if not global_xy:
            s = table.surfaces[last - 1]
            d = np.stack([gx - s.t[0], gy - s.t[1], rec["z"][-1] - s.t[2]])
            loc = np.asarray(s.R).T @ d
            gx, gy = loc[0], loc[1]
        dx, dy = gx - center[0], gy - center[1]
        fin_ok = np.isfinite(dx) & np.isfinite(dy)
        keep = np.ones_like(fin_ok) if every_ray else ((ii > 0) & fin_ok)
        m = [float(keep.sum()), float(dx[keep].sum()), float(dy[keep].sum()), float((dx[keep] ** 2 + dy[keep] ** 2).sum()),
             float(ii[keep].sum()), float(oo[keep].sum()), float((oo[keep] ** 2).sum()),
             0.0 if every_ray else float(((ii > 0) & ~fin_ok).sum())]
        return m

    def trace_wavefront(self, table, Px, Py, affine, ref, polarized=False):
        import torch

        from oracle import trace_oracle as O
        from optiland_b200.launch import launch_from_affine

        self.calls.append(("wavefront", table.num_surfaces, int(Px.numel())))
        px, py = Px.detach().double().numpy(), Py.detach().double().numpy()
        x, y, z, L, M, N = launch_from_affine(px, py, affine)
        inp = dict(x=x, y=y, z=z, L=L, M=M, N=N, i=np.full_like(px, affine.get("intensity", 1.0)),
                   w=np.full_like(px, table.wavelengths[0]))
        if polarized:
            inp["p"] = np.tile(np.eye(3, dtype=np.complex128), (px.size, 1, 1))
        fin, _, status = O.trace(table, inp, polarized=polarized)
        raise_status(status)
        out = O.wavefront_reference_sphere(fin, px, py, ref)
        res = {k: torch.from_numpy(np.asarray(v)).to(Px.dtype) for k, v in out.items()}
        if polarized:
            res["p"] = torch.from_numpy(fin["p"]).to(torch.complex128 if Px.dtype == torch.float64 else torch.complex64)
        return res

    def huygens_psf(self, image_x, image_y, image_z, pupil_x, pupil_y, pupil_z, pupil_amp, pupil_opd, wavelength, Rp):
        import torch

        from oracle import trace_oracle as O

        self.calls.append(("psf", int(image_x.numel()), int(pupil_x.numel())))
        amp = pupil_amp.detach().numpy() if torch.is_tensor(pupil_amp) else np.asarray(pupil_amp)
        psf, _ = O.huygens_fresnel_psf(*[t.detach().double().numpy() for t in (image_x, image_y, image_z, pupil_x, pupil_y, pupil_z)],
                                       amp, pupil_opd.detach().double().numpy(), wavelength, Rp)
        return torch.from_numpy(psf).to(image_x.dtype)

    def fft_pupil(self, opd_waves, intensity, cell_ray, num_rays, grid_size):
        """TEST-ONLY: the gridding kernel's per-cell function (csrc/olb_fftpsf.cuh) looped on the CPU (tests/hostcheck)."""
        import ctypes as C

        import torch

        from oracle.hostcheck_api import load

        hc = load()
        self.calls.append(("fft_pupil", int(num_rays), int(grid_size)))
        f64 = opd_waves.dtype == torch.float64
        npdt = np.float64 if f64 else np.float32
        opd = np.ascontiguousarray(opd_waves.detach().cpu().numpy().astype(npdt))
        inten = np.ascontiguousarray(intensity.detach().cpu().numpy().astype(npdt))
        cell = np.ascontiguousarray(cell_ray.detach().cpu().numpy().astype(np.int32))
        out = np.empty((grid_size, grid_size, 2), dtype=npdt)
        fn = hc.olbhc_fft_pupil_f64 if f64 else hc.olbhc_fft_pupil_f32
        fn(C.c_void_p(opd.ctypes.data), C.c_void_p(inten.ctypes.data), C.c_void_p(cell.ctypes.data), C.c_int32(num_rays),
           C.c_int32(grid_size), C.c_void_p(out.ctypes.data))
        return torch.view_as_complex(torch.from_numpy(out))

    def fft_psf_accumulate(self, amp, psf, first, last, div, mul):
        import ctypes as C

        import torch

        from oracle.hostcheck_api import load

        hc = load()
        self.calls.append(("fft_psf", int(amp.shape[-1])))
        a = np.ascontiguousarray(torch.view_as_real(amp.detach().to(torch.complex128).contiguous()).numpy())
        acc = np.ascontiguousarray(psf.detach().double().numpy())
        hc.olbhc_fft_psf_accumulate_f64(C.c_void_p(a.ctypes.data), C.c_int32(amp.shape[-1]), C.c_int32(int(first)),
                                        C.c_int32(int(last)), C.c_double(div), C.c_double(mul), C.c_void_p(acc.ctypes.data))
        psf.copy_(torch.from_numpy(acc).to(psf.dtype))
        return psf

    def trace_grad(self, table, params, rays, coefs=None):
        """TEST-ONLY differentiable engine: oracle forward + the CPU instantiation of the device adjoint
        (tests/hostcheck) -- the arithmetic of olb_trace_bwd_* without a GPU."""
        import torch

        from oracle import trace_oracle as O
        from optiland_b200 import autograd as AG
        from oracle.hostcheck_api import load, run_backward

        hc = load()
        import ctypes as C

        from optiland_b200 import _lib

        ht = _lib.HostTable(table)  # keep the packed arrays alive across the call
        if not hc.olbhc_bwd_supported(C.byref(ht.c)):
            return None  # same scope rule as CudaEngine (OlbDeviceTable.bwd_supported): the plugin declines
        self.calls.append(("grad", table.num_surfaces, int(rays.x.numel())))
        keys = ("x", "y", "z", "L", "M", "N", "i", "opd")

        has_tables = any(sp.kind in AG.POLY_KINDS or sp.kind == AG.T.GEOM_FORBES_QBFS for sp in table.surfaces)
        cf_in = coefs if coefs is not None else torch.zeros((table.num_surfaces, 1), dtype=torch.float64)

        class Fn(torch.autograd.Function):
            @staticmethod
            def forward(ctx, params, cf, *ins):
                ctx.set_materialize_grads(False)
                tab = AG.params_to_table(table, params, cf if coefs is not None else None)
                inp = {k: t.detach().double().numpy() for k, t in zip(keys, ins)}
                inp["w"] = rays.w.detach().double().numpy()
                _, rec, _ = O.trace(tab, inp)
                ctx.tab, ctx.inp, ctx.rec = tab, inp, rec
                return tuple(torch.from_numpy(rec[k]) for k in ("x", "y", "z", "L", "M", "N", "intensity", "opd"))

            @staticmethod
            def backward(ctx, *grads):
                grec = {k: (None if g is None else g.double().numpy()) for k, g in
                        zip(("x", "y", "z", "L", "M", "N", "intensity", "opd"), grads)}
                if has_tables:
                    gin, gpar, gtab = run_backward(hc, ctx.tab, ctx.inp, ctx.rec, grec, tables=True)
                    gcf = torch.from_numpy(AG.tables_to_coef_grads(ctx.tab, gtab, cf_in.shape[1])) if coefs is not None else None
                else:
                    gin, gpar = run_backward(hc, ctx.tab, ctx.inp, ctx.rec, grec)
                    gcf = None
                for s_, sp in enumerate(ctx.tab.surfaces):      # Forbes: dLoss/db (Clenshaw basis) -> dLoss/da, as _TraceFn.backward
                    if sp.kind == AG.T.GEOM_FORBES_QBFS and len(sp.coefficients):
                        nc = len(sp.coefficients)
                        gpar[s_, AG.GP_COEF:AG.GP_COEF + nc] = AG.forbes_coef_grads(gpar[s_, AG.GP_COEF:AG.GP_COEF + nc])
                return (torch.from_numpy(gpar), gcf, *[torch.from_numpy(gin[k]) for k in keys])

        outs = Fn.apply(params, cf_in, *[getattr(rays, k) for k in keys])
        rec = dict(zip(("x", "y", "z", "L", "M", "N", "intensity", "opd"), outs))
        for k, key in zip(keys, ("x", "y", "z", "L", "M", "N", "intensity", "opd")):
            setattr(rays, k, rec[key][-1])
        return rec


