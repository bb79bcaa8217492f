"""Huygens-Fresnel PSF summation on the GPU -- host-side mirror of the reference's
summation strategies (/root/reference/optiland/psf/huygens_fresnel_strategies.py:63-273): same argument
list as ``HuygensFresnelSummation.compute``.  CUDA only (libolb ``olb_huygens_psf_f64``)."""
from __future__ import annotations

import ctypes as C

import torch

from . import _lib


def _dev64(t, device):
    return torch.as_tensor(t, device=device).to(torch.float64).contiguous()


def huygens_fresnel_psf(image_x, image_y, image_z, pupil_x, pupil_y, pupil_z, pupil_amp, pupil_opd, wavelength, Rp,
                        return_field: bool = False):
    """``compute(image_x, image_y, image_z, pupil_x, pupil_y, pupil_z, pupil_amp, pupil_opd, wavelength, Rp)``:
    image arrays of any (common) shape, pupil arrays 1-D, ``pupil_opd`` and ``wavelength`` in mm.  Returns the
    PSF with the image arrays' shape (and the complex field when asked)."""
    if not torch.cuda.is_available():
        raise _lib.OlbError("optiland_b200.psf needs a CUDA device; there is no CPU fallback")
    lib = _lib.load()
    device = image_x.device if torch.is_tensor(image_x) and image_x.is_cuda else torch.device("cuda", torch.cuda.current_device())
    ix, iy, iz = (_dev64(t, device).reshape(-1) for t in (image_x, image_y, image_z))
    px, py, pz, popd = (_dev64(t, device).reshape(-1) for t in (pupil_x, pupil_y, pupil_z, pupil_opd))
    amp = torch.as_tensor(pupil_amp, device=device).reshape(-1)
    if amp.is_complex():
        ar, ai = amp.real.to(torch.float64).contiguous(), amp.imag.to(torch.float64).contiguous()
    else:
        ar, ai = amp.to(torch.float64).contiguous(), None
    n_img, n_pup = ix.numel(), px.numel()
    psf = torch.empty(n_img, dtype=torch.float64, device=device)
    field = torch.empty(2 * n_img, dtype=torch.float64, device=device)  # also the scratch of the split-pupil mode
    with torch.cuda.device(device):
        stream = torch.cuda.current_stream(device).cuda_stream
        rc = lib.olb_huygens_psf_f64(ix.data_ptr(), iy.data_ptr(), iz.data_ptr(), n_img, px.data_ptr(), py.data_ptr(),
                                     pz.data_ptr(), ar.data_ptr(), ai.data_ptr() if ai is not None else None,
                                     popd.data_ptr(), n_pup, float(wavelength), float(Rp), psf.data_ptr(),
                                     field.data_ptr(), C.c_void_p(stream))
    _lib.check(rc, "olb_huygens_psf_f64")
    shape = tuple(image_x.shape) if hasattr(image_x, "shape") else (n_img,)
    psf = psf.reshape(shape)
    if return_field:
        return psf, torch.view_as_complex(field.reshape(n_img, 2)).reshape(shape)
    return psf


# ---- FFT-PSF gridding (second half): the element-wise passes on either side of the library FFT ----------
_SFX = {torch.float64: "f64", torch.float32: "f32"}
_CPLX = {torch.float64: torch.complex128, torch.float32: torch.complex64}


def fft_pupil(opd_waves, intensity, cell_ray, num_rays: int, grid_size: int):
    """The zero-PADDED complex pupil function of one wavelength, (grid_size, grid_size) complex, in ONE pass
    (``olb_fft_pupil_*``): what ``ScalarFFTPSF._generate_pupils`` + ``_pad_pupils`` build with a masked assignment,
    a reshape and ``be.pad`` (/root/reference/optiland/psf/fft.py:123-227).  ``opd_waves`` / ``intensity``: the
    WavefrontData arrays (CUDA, fp32 or fp64); ``cell_ray``: int32 (num_rays^2) sample index per pupil-grid cell, -1
    outside the unit disk."""
    dtype = opd_waves.dtype
    if dtype not in _SFX or not opd_waves.is_cuda:
        raise _lib.OlbError("fft_pupil: CUDA fp32 / fp64 arrays expected; there is no CPU fallback")
    lib = _lib.load()
    opd = opd_waves.detach().contiguous()
    inten = intensity.detach().to(dtype).contiguous()
    cell = cell_ray.contiguous()
    out = torch.empty((grid_size, grid_size), dtype=_CPLX[dtype], device=opd.device)
    with torch.cuda.device(opd.device):
        stream = torch.cuda.current_stream(opd.device).cuda_stream
        rc = getattr(lib, f"olb_fft_pupil_{_SFX[dtype]}")(opd.data_ptr(), inten.data_ptr(), opd.numel(), cell.data_ptr(),
                                                          int(num_rays), int(grid_size), out.data_ptr(), C.c_void_p(stream))
    _lib.check(rc, "olb_fft_pupil")
    return out


def fft_psf_accumulate(amp, psf, first: bool, last: bool, div: float = 1.0, mul: float = 1.0):
    """``psf[fftshift] (+)= |amp|^2`` (and ``/ div * mul`` on the last wavelength) in one pass over the spectrum
    (``olb_fft_psf_accumulate_*``; fft.py:184-191).  ``amp``: (g, g) complex CUDA tensor, ``psf``: (g, g) real, updated
    in place and returned."""
    rdt = psf.dtype
    if rdt not in _SFX or not amp.is_cuda or amp.dtype != _CPLX[rdt]:
        raise _lib.OlbError("fft_psf_accumulate: CUDA complex64 / complex128 spectrum with a matching real psf expected")
    lib = _lib.load()
    amp = amp.contiguous()
    g = amp.shape[-1]
    with torch.cuda.device(amp.device):
        stream = torch.cuda.current_stream(amp.device).cuda_stream
        rc = getattr(lib, f"olb_fft_psf_accumulate_{_SFX[rdt]}")(amp.data_ptr(), int(g), int(bool(first)), int(bool(last)),
                                                                 float(div), float(mul), psf.data_ptr(), C.c_void_p(stream))
    _lib.check(rc, "olb_fft_psf_accumulate")
    return psf
