"""TEST INFRASTRUCTURE ONLY -- a second stand-in for optiland_b200.plugin.CudaEngine on boxes without a GPU: the
forward trace runs the DEVICE ARITHMETIC itself (csrc/olb_math.cuh + olb_prep.h compiled for the host,
tests/hostcheck) on the packed table the product would upload (``_lib.HostTable``), where ``OracleEngine`` evaluates
the NumPy restatement of the reference.  With it a CPU test chains: live reference objects -> ``plugin`` / ``pack`` ->
prepared table -> the kernel's per-ray code, and compares with the unmodified reference -- everything of the product
path but the CUDA launch wrapper.  The wavefront epilogue (``wavefront_point``) and the polarized intensity epilogue
(``polarized_intensity``) run through their host instantiations as well; moments, PSF and the adjoint are inherited from
``OracleEngine`` (PSF gridding and the adjoint are host instantiations of device code there too)."""
import numpy as np

from oracle.oracle_engine import OracleEngine, raise_status

_KEYS = ("x", "y", "z", "L", "M", "N", "i", "w", "opd")


class DeviceMathEngine(OracleEngine):
    """TEST-ONLY: ``trace`` / ``trace_pupil`` through the host instantiation of the device math."""

    def trace(self, table, rays, first, last):
        import torch

        from oracle.hostcheck_api import load, run_hostcheck

        self.calls.append((table.num_surfaces, int(rays.x.numel())))
        n = int(rays.x.numel())
        inp = {k: np.broadcast_to(getattr(rays, k).detach().double().numpy(), (n,)).copy() for k in _KEYS}
        polarized = type(rays).__name__ == "PolarizedRays"
        pmat = rays.p.detach().numpy().astype(np.complex128) if polarized else None
        out, rec, status = run_hostcheck(load(), table, inp, np.float64, first, last, pmat=pmat)
        raise_status(status)
        if polarized:
            rays.p = torch.from_numpy(out["p"])
        dt = rays.x.dtype
        for k in ("x", "y", "z", "L", "M", "N", "i", "opd"):
            setattr(rays, k, torch.from_numpy(out[k]).to(dt))
        return {k: torch.from_numpy(v).to(dt) for k, v in rec.items()}

    def trace_pupil(self, table, Px, Py, affine, wavelength=None, polarization=False):
        import torch

        from oracle import trace_oracle as O
        from oracle.hostcheck_api import load, run_hostcheck
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
        pmat = None if polarization is False else np.tile(np.eye(3, dtype=np.complex128), (px.size, 1, 1))
        out, rec, status = run_hostcheck(load(), table, inp, np.float64, pmat=pmat)
        raise_status(status)
        res = {k: torch.from_numpy(v).to(Px.dtype) for k, v in rec.items()}
        if polarization is False:
            return res
        res["p"] = torch.from_numpy(out["p"]).to(torch.complex128 if Px.dtype == torch.float64 else torch.complex64)
        if polarization == "matrix":
            res["i_pol"] = res["intensity"][-1]
        else:
            from oracle.hostcheck_api import run_pol_intensity
            from optiland_b200 import table as T

            ipol, st = run_pol_intensity(load(), out["p"], (L, M, N), i0, polarization)
            if st & T.ST_K_PARALLEL_X:
                raise ValueError("k-vector parallel to x-axis is not currently supported.")
            res["i_pol"] = torch.from_numpy(ipol).to(Px.dtype)
        return res

    def trace_wavefront(self, table, Px, Py, affine, ref, polarized=False):
        import torch

        from oracle.hostcheck_api import load, run_hostcheck, run_wavefront
        from optiland_b200.launch import launch_from_affine

        self.calls.append(("wavefront", table.num_surfaces, int(Px.numel())))
        px, py = Px.detach().double().numpy(), Py.detach().double().numpy()
        x, y, z, L, M, N = launch_from_affine(px, py, affine)
        inp = dict(x=x, y=y, z=z, L=L, M=M, N=N, i=np.full_like(px, affine.get("intensity", 1.0)),
                   w=np.full_like(px, table.wavelengths[0]))
        pmat = np.tile(np.eye(3, dtype=np.complex128), (px.size, 1, 1)) if polarized else None
        fin, _, status = run_hostcheck(load(), table, inp, np.float64, pmat=pmat)
        raise_status(status)
        out = run_wavefront(load(), fin, px, py, ref)
        out["intensity"] = fin["i"]
        res = {k: torch.from_numpy(np.asarray(v)).to(Px.dtype) for k, v in out.items()}
        if polarized:
            res["p"] = torch.from_numpy(fin["p"]).to(torch.complex128 if Px.dtype == torch.float64 else torch.complex64)
        return res
