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
