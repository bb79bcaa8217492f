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
