"""Huygens-Fresnel PSF summation: the NumPy oracle against the reference's own Numba
kernel and HuygensPSF (fixture tests/golden/huygens_psf_ref.npz, oracle/make_golden.py), and the CUDA
kernel against both."""
import os

import numpy as np
import pytest
import torch

from oracle import trace_oracle as O
from tests._util import GOLDEN


@pytest.fixture(scope="module")
def g():
    return np.load(os.path.join(GOLDEN, "huygens_psf_ref.npz"))


def test_oracle_matches_reference_numba_summation(g):
    psf, _ = O.huygens_fresnel_psf(g["image_x"], g["image_y"], g["image_z"], g["pupil_x"], g["pupil_y"], g["pupil_z"],
                                   g["pupil_amp"], g["pupil_opd"], float(g["wavelength"]), float(g["Rp"]))
    assert psf.max() > 1e-3
    # the reference kernel is compiled with fastmath=True: agreement to ~1e-9 of the peak
    assert np.max(np.abs(psf - g["psf"])) <= 1e-8 * psf.max()
    psf2, _ = O.huygens_fresnel_psf(g["sys_image_x"], g["sys_image_y"], g["sys_image_z"], g["sys_pupil_x"], g["sys_pupil_y"],
                                    g["sys_pupil_z"], g["sys_pupil_amp"], g["sys_pupil_opd"], 0.55e-3, float(g["sys_Rp"]))
    ours = psf2 / float(g["sys_norm"]) * 100.0
    assert np.max(np.abs(ours - g["sys_psf"])) <= 1e-9 * g["sys_psf"].max()


@pytest.mark.gpu
def test_cuda_summation_matches_reference(g):
    from optiland_b200.psf import huygens_fresnel_psf

    psf, field = huygens_fresnel_psf(*[torch.from_numpy(g[k]).cuda() for k in
                                       ("image_x", "image_y", "image_z", "pupil_x", "pupil_y", "pupil_z", "pupil_amp", "pupil_opd")],
                                     float(g["wavelength"]), float(g["Rp"]), return_field=True)
    assert psf.shape == g["psf"].shape
    ref, ref_field = O.huygens_fresnel_psf(g["image_x"], g["image_y"], g["image_z"], g["pupil_x"], g["pupil_y"], g["pupil_z"],
                                           g["pupil_amp"], g["pupil_opd"], float(g["wavelength"]), float(g["Rp"]))
    assert float((psf.cpu() - torch.from_numpy(g["psf"])).abs().max()) <= 1e-8 * g["psf"].max()
    assert float((field.cpu() - torch.from_numpy(ref_field)).abs().max()) <= 1e-9 * np.abs(ref_field).max()
    # system level: Cooke triplet HuygensPSF (32 x 32 pupil, 32 x 32 image), Strehl-normalised
    psf2 = huygens_fresnel_psf(*[torch.from_numpy(g["sys_" + k]).cuda() for k in
                                 ("image_x", "image_y", "image_z", "pupil_x", "pupil_y", "pupil_z", "pupil_amp", "pupil_opd")],
                               0.55e-3, float(g["sys_Rp"]))
    ours = psf2.cpu().numpy() / float(g["sys_norm"]) * 100.0
    assert np.max(np.abs(ours - g["sys_psf"])) <= 1e-9 * g["sys_psf"].max()
    # complex amplitudes (vectorial case) and ragged sizes
    amp = torch.from_numpy(g["pupil_amp"]).cuda() * torch.exp(1j * torch.linspace(0, 3, g["pupil_amp"].size, device="cuda"))
    n = 37
    p3 = huygens_fresnel_psf(torch.from_numpy(g["image_x"].ravel()[:n]).cuda(), torch.from_numpy(g["image_y"].ravel()[:n]).cuda(),
                             torch.from_numpy(g["image_z"].ravel()[:n]).cuda(), torch.from_numpy(g["pupil_x"]).cuda(),
                             torch.from_numpy(g["pupil_y"]).cuda(), torch.from_numpy(g["pupil_z"]).cuda(), amp,
                             torch.from_numpy(g["pupil_opd"]).cuda(), float(g["wavelength"]), float(g["Rp"]))
    r3, _ = O.huygens_fresnel_psf(g["image_x"].ravel()[:n], g["image_y"].ravel()[:n], g["image_z"].ravel()[:n], g["pupil_x"],
                                  g["pupil_y"], g["pupil_z"], amp.cpu().numpy(), g["pupil_opd"], float(g["wavelength"]), float(g["Rp"]))
    assert np.max(np.abs(p3.cpu().numpy() - r3)) <= 1e-9 * r3.max()


# ---- FFT-PSF gridding (second half): olb_fft_pupil_* / olb_fft_psf_accumulate_* -----------------------
def _fft_reference(opd_list, inten_list, num_rays, grid, dtype=np.float64):
    """NumPy restatement of ScalarFFTPSF._generate_pupils / _pad_pupils / _compute_psf
    (/root/reference/optiland/psf/fft.py:123-227) for a list of wavelengths."""
    x = np.linspace(-1, 1, num_rays).astype(dtype)
    x, y = np.meshgrid(x, x)
    R2 = x.ravel() ** 2 + y.ravel() ** 2
    pupils = []
    for opd, inten in zip(opd_list, inten_list):
        P = np.zeros(num_rays * num_rays, dtype=np.complex128)
        with np.errstate(invalid="ignore"):
            P[R2 <= 1] = np.sqrt(inten) * np.exp(-1j * 2 * np.pi * opd)
        pupils.append(P.reshape(num_rays, num_rays))
    pb = (grid - num_rays) // 2
    pa = pb + (grid - num_rays) % 2
    padded = [np.pad(p, ((pb, pa), (pb, pa))) for p in pupils]
    if not pupils:
        return [], None, 0.0, (R2 <= 1)
    norm = float(np.sum(np.abs(pupils[0]) > 0) ** 2)
    psf = []
    for p in padded:
        amp = np.fft.fftshift(np.fft.fft2(p))
        psf.append(np.real(amp * np.conj(amp)))
    return padded, np.real(np.sum(np.stack(psf), axis=0)) / norm * 100, norm, (R2 <= 1)


def _cell_map(mask):
    idx = np.cumsum(mask.astype(np.int32), dtype=np.int32) - 1
    return np.where(mask, idx, -1).astype(np.int32)


@pytest.mark.parametrize("num_rays,grid", [(32, 64), (33, 77), (17, 17), (40, 41)])
def test_fft_gridding_cell_functions_match_the_reference_formulas(num_rays, grid):
    """The per-cell functions the two kernels run (csrc/olb_fftpsf.cuh), looped on the CPU by tests/hostcheck: padded
    pupil function, |.|^2 + fftshift + accumulation over three wavelengths + normalisation, even and odd sizes."""
    import ctypes as C

    from oracle.hostcheck_api import load

    hc = load()
    rng = np.random.default_rng(num_rays)
    _, _, _, mask = _fft_reference([], [], num_rays, grid)
    count = int(mask.sum())
    opds = [rng.normal(scale=3.0, size=count) for _ in range(3)]
    intens = [rng.uniform(0.0, 1.0, size=count) for _ in range(3)]
    intens[0][:3] = 0.0                      # vignetted samples: zero amplitude, not counted by the normalisation
    padded_ref, psf_ref, norm, _ = _fft_reference(opds, intens, num_rays, grid)
    cell = _cell_map(mask)
    psf = np.full((grid, grid), np.nan)
    for j, (opd, inten) in enumerate(zip(opds, intens)):
        out = np.empty((grid, grid, 2))
        hc.olbhc_fft_pupil_f64(C.c_void_p(opd.ctypes.data), C.c_void_p(inten.ctypes.data), C.c_void_p(cell.ctypes.data),
                               C.c_int32(num_rays), C.c_int32(grid), C.c_void_p(out.ctypes.data))
        got = out[..., 0] + 1j * out[..., 1]
        assert np.max(np.abs(got - padded_ref[j])) <= 1e-13
        amp = np.ascontiguousarray(np.fft.fft2(got))
        a2 = np.ascontiguousarray(np.stack([amp.real, amp.imag], axis=-1))
        hc.olbhc_fft_psf_accumulate_f64(C.c_void_p(a2.ctypes.data), C.c_int32(grid), C.c_int32(int(j == 0)),
                                        C.c_int32(int(j == 2)), C.c_double(norm), C.c_double(100.0), C.c_void_p(psf.ctypes.data))
    assert np.max(np.abs(psf - psf_ref)) <= 1e-12 * psf_ref.max()
    # fp32 instantiation of the pupil cell
    o32, i32 = opds[1].astype(np.float32), intens[1].astype(np.float32)
    out32 = np.empty((grid, grid, 2), dtype=np.float32)
    hc.olbhc_fft_pupil_f32(C.c_void_p(o32.ctypes.data), C.c_void_p(i32.ctypes.data), C.c_void_p(cell.ctypes.data),
                           C.c_int32(num_rays), C.c_int32(grid), C.c_void_p(out32.ctypes.data))
    assert np.max(np.abs((out32[..., 0] + 1j * out32[..., 1]) - padded_ref[1])) <= 2e-5   # fp32 rounding of 2 pi opd ~ 1e-6 * 20 rad


@pytest.mark.gpu
@pytest.mark.parametrize("num_rays,grid", [(32, 64), (33, 77), (128, 1024), (257, 600)])
@pytest.mark.parametrize("dtype", [torch.float64, torch.float32])
def test_cuda_fft_gridding_kernels_match_the_reference_formulas(num_rays, grid, dtype):
    """olb_fft_pupil_* / olb_fft_psf_accumulate_* through the C ABI against the NumPy restatement of fft.py:123-227:
    padded pupil function bit-close, PSF of three accumulated wavelengths, even / odd / large grids, fp64 and fp32, NaN
    intensity propagating like the reference's sqrt."""
    from optiland_b200 import _lib
    from optiland_b200.psf import fft_psf_accumulate, fft_pupil

    rng = np.random.default_rng(num_rays + grid)
    _, _, _, mask = _fft_reference([], [], num_rays, grid)
    count = int(mask.sum())
    npdt = np.float64 if dtype == torch.float64 else np.float32
    opds = [rng.normal(scale=3.0, size=count).astype(npdt) for _ in range(3)]
    intens = [rng.uniform(0.0, 1.0, size=count).astype(npdt) for _ in range(3)]
    intens[0][:3] = 0.0
    padded_ref, psf_ref, norm, _ = _fft_reference([o.astype(np.float64) for o in opds], [i.astype(np.float64) for i in intens],
                                                  num_rays, grid)
    cell = torch.from_numpy(_cell_map(mask)).cuda()
    l0 = _lib.load().olb_launch_count()
    psf = torch.full((grid, grid), float("nan"), dtype=dtype, device="cuda")
    tol_p = 1e-13 if dtype == torch.float64 else 3e-5
    for j in range(3):
        got = fft_pupil(torch.from_numpy(opds[j]).cuda(), torch.from_numpy(intens[j]).cuda(), cell, num_rays, grid)
        assert got.shape == (grid, grid) and got.dtype == (torch.complex128 if dtype == torch.float64 else torch.complex64)
        assert np.max(np.abs(got.cpu().numpy() - padded_ref[j])) <= tol_p
        fft_psf_accumulate(torch.fft.fft2(got), psf, j == 0, j == 2, norm, 100.0)
    assert _lib.load().olb_launch_count() - l0 == 6
    tol = 1e-11 if dtype == torch.float64 else 2e-4
    assert np.max(np.abs(psf.cpu().numpy() - psf_ref)) <= tol * psf_ref.max()
    # NaN intensity -> NaN cell (the reference's sqrt / exp), everything else untouched
    bad = intens[1].copy()
    bad[5] = np.nan
    got = fft_pupil(torch.from_numpy(opds[1]).cuda(), torch.from_numpy(bad).cuda(), cell, num_rays, grid).cpu().numpy()
    assert np.isnan(got).sum() == 1 and np.nanmax(np.abs(got - padded_ref[1])) <= tol_p
    # argument checks come back as error codes, not crashes
    with pytest.raises(_lib.OlbError):
        fft_pupil(torch.from_numpy(opds[1]).cuda(), torch.from_numpy(intens[1]).cuda(), cell, num_rays, num_rays - 1)
