"""Huygens-Fresnel summation timing: CUDA kernel vs an eager-torch restatement of the reference's
TorchSummation.compute (batched, /root/reference/optiland/psf/huygens_fresnel_strategies.py:183-273) on the
same device, and the NumPy oracle on the host for a small sample."""
import json
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from optiland_b200.psf import huygens_fresnel_psf  # noqa: E402


def eager_torch(ix, iy, iz, px, py, pz, amp, opd, wl, Rp, batch=1024):
    k = 2.0 * torch.pi / wl
    px, py, pz, amp, opd = (t.reshape(1, -1) for t in (px, py, pz, amp, opd))
    out = torch.zeros(ix.numel(), dtype=torch.complex128, device=ix.device)
    fx, fy, fz = ix.flatten(), iy.flatten(), iz.flatten()
    for i in range(0, fx.numel(), batch):
        x, y, z = (t[i:i + batch].reshape(-1, 1) for t in (fx, fy, fz))
        dx, dy, dz = x - px, y - py, z - pz
        R = torch.sqrt(dx**2 + dy**2 + dz**2)
        wave = torch.exp(1j * k * R) / R
        dot = dx * (px / Rp) + dy * (py / Rp) + dz * (pz / Rp)
        out[i:i + batch] = torch.sum(amp * torch.exp(-1j * k * opd) * wave * (0.5 * (1.0 + dot / R)), dim=1)
    return (out.abs() ** 2).reshape(ix.shape)


def main():
    dev = torch.device("cuda:0")
    res = []
    for n_img, n_pup_side in ((128, 128), (256, 256)):
        g = torch.Generator(device=dev).manual_seed(0)
        lin = torch.linspace(-1, 1, n_pup_side, device=dev, dtype=torch.float64)
        U, V = torch.meshgrid(lin, lin, indexing="ij")
        m = (U**2 + V**2) <= 1
        Rp = -100.0
        pu, pv = 10 * U[m], 10 * V[m]
        pw = -torch.sqrt(Rp**2 - pu**2 - pv**2)
        amp = torch.ones_like(pu)
        opd = 1e-4 * torch.randn(pu.numel(), generator=g, device=dev, dtype=torch.float64)
        gl = torch.linspace(-0.03, 0.03, n_img, device=dev, dtype=torch.float64)
        X, Y = torch.meshgrid(gl, gl, indexing="ij")
        Z = torch.zeros_like(X)
        args = (X, Y, Z, pu, pv, pw, amp, opd, 0.55e-3, Rp)
        for _ in range(2):
            ours = huygens_fresnel_psf(*args)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        K = 5
        for _ in range(K):
            ours = huygens_fresnel_psf(*args)
        torch.cuda.synchronize()
        t_ours = (time.perf_counter() - t0) / K
        ref = eager_torch(*args)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        ref = eager_torch(*args)
        torch.cuda.synchronize()
        t_ref = time.perf_counter() - t0
        pairs = X.numel() * pu.numel()
        res.append({"image": f"{n_img}x{n_img}", "pupil_points": int(pu.numel()), "pairs": pairs, "cuda_kernel_ms": round(t_ours * 1e3, 3),
                    "pairs_per_s": round(pairs / t_ours, 0), "eager_torch_same_gpu_ms": round(t_ref * 1e3, 2),
                    "speedup_vs_eager_torch": round(t_ref / t_ours, 1),
                    "max_rel_diff_vs_eager": float(((ours - ref).abs().max() / ref.max()).item())})
        print(json.dumps(res[-1]), flush=True)


if __name__ == "__main__":
    main()
