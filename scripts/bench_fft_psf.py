"""FFT-PSF gridding timing (second half): the two gridding kernels (olb_fft_pupil_*,
olb_fft_psf_accumulate_*) against the reference's element-wise op sequence around the same library FFT
(/root/reference/optiland/psf/fft.py:123-227, restated with eager torch ops on the same GPU), CUDA events.
One JSON line per (num_rays, grid_size, dtype): ms per PSF for each arm split into {gridding before, fft, after},
and the achieved GB/s of the two kernels against their algorithmic bytes
(pupil: grid^2 complex written + num_rays^2 int32 + 2 n reals read; psf: grid^2 complex read + grid^2 real written)."""
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from optiland_b200.psf import fft_psf_accumulate, fft_pupil  # noqa: E402


def timed(fn, reps=20, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def main():
    dev = torch.device("cuda:0")
    for num_rays, grid in ((128, 1024), (256, 2048), (512, 4096)):
        for dtype in (torch.float64, torch.float32):
            cdt = torch.complex128 if dtype == torch.float64 else torch.complex64
            x = torch.linspace(-1, 1, num_rays, device=dev, dtype=dtype)
            X, Y = torch.meshgrid(x, x, indexing="xy")
            R2 = X.ravel() ** 2 + Y.ravel() ** 2
            mask = R2 <= 1
            count = int(mask.sum())
            g = torch.Generator(device=dev).manual_seed(0)
            opd = 2.0 * torch.randn(count, generator=g, device=dev, dtype=dtype)
            inten = torch.rand(count, generator=g, device=dev, dtype=dtype)
            idx = torch.cumsum(mask.to(torch.int32), 0, dtype=torch.int32) - 1
            cell = torch.where(mask, idx, torch.full_like(idx, -1)).contiguous()
            pb = (grid - num_rays) // 2
            pa = pb + (grid - num_rays) % 2
            norm = float(count) ** 2

            def ref_before():
                P = torch.zeros_like(R2).to(cdt)
                P[mask] = (torch.sqrt(inten) * torch.exp(-1j * 2 * torch.pi * opd)).to(cdt)
                return torch.nn.functional.pad(P.reshape(num_rays, num_rays), (pb, pa, pb, pa))

            padded_ref = ref_before()

            def ref_after(amp):
                a = torch.fft.fftshift(amp)
                return torch.real(torch.sum(torch.stack([torch.real(a * torch.conj(a))]), dim=0)) / norm * 100

            amp_ref = torch.fft.fft2(padded_ref)
            psf = torch.empty((grid, grid), dtype=dtype, device=dev)
            padded = fft_pupil(opd, inten, cell, num_rays, grid)
            assert float((padded - padded_ref).abs().max()) < (1e-12 if dtype == torch.float64 else 1e-4)
            fft_psf_accumulate(amp_ref, psf, True, True, norm, 100.0)
            want = ref_after(amp_ref)
            assert float((psf - want).abs().max()) <= (1e-11 if dtype == torch.float64 else 1e-4) * float(want.max())
            b = 8 if dtype == torch.float64 else 4
            t_k1 = timed(lambda: fft_pupil(opd, inten, cell, num_rays, grid))
            t_k2 = timed(lambda: fft_psf_accumulate(amp_ref, psf, True, True, norm, 100.0))
            t_r1 = timed(ref_before)
            t_r2 = timed(lambda: ref_after(amp_ref))
            t_fft = timed(lambda: torch.fft.fft2(padded_ref))
            bytes1 = grid * grid * 2 * b + num_rays * num_rays * 4 + 2 * count * b
            bytes2 = grid * grid * 3 * b
            print(json.dumps({
                "num_rays": num_rays, "grid_size": grid, "dtype": str(dtype).split(".")[-1],
                "kernel_pupil_ms": round(t_k1, 4), "kernel_psf_ms": round(t_k2, 4), "library_fft_ms": round(t_fft, 4),
                "eager_before_ms": round(t_r1, 4), "eager_after_ms": round(t_r2, 4),
                "pupil_GBps": round(bytes1 / t_k1 / 1e6, 1), "psf_GBps": round(bytes2 / t_k2 / 1e6, 1),
                "gridding_speedup": round((t_r1 + t_r2) / (t_k1 + t_k2), 1),
                "whole_psf_speedup": round((t_r1 + t_r2 + t_fft) / (t_k1 + t_k2 + t_fft), 2)}), flush=True)


if __name__ == "__main__":
    main()
