"""SURVEY.md 8f-3 (second half) wired into the reference's class: the FFT-PSF's gridding passes as two kernels behind

* ``ScalarFFTPSF._generate_pupils``  /root/reference/optiland/psf/fft.py:123-157
* ``ScalarFFTPSF._pad_pupils``       fft.py:193-227
* ``ScalarFFTPSF._compute_psf``      fft.py:159-191

The reference builds, per wavelength, ``zeros -> sqrt -> exp -> masked assignment -> reshape -> pad`` before and
``fftshift -> conj -> multiply -> real -> stack -> sum -> divide -> multiply`` after ONE ``fft2``.  Here the padded pupil
function is written in one pass (``olb_fft_pupil_*``: the masked scatter as a gather through a cell -> sample map) and
the spectrum is read once (``olb_fft_psf_accumulate_*``: |.|^2, the shift, the sum over wavelengths and the
normalisation).  The FFT itself stays the library's (``torch.fft.fft2`` -> cuFFT).

``self.pupils`` keeps its meaning -- a list of (num_rays, num_rays) complex arrays, here VIEWS of the centre of the
padded buffers -- so ``_get_normalization`` and anything else that reads it is untouched; if a caller replaces
``self.pupils`` the wrappers notice (identity check) and run the reference's own code on it.  The wavefront data itself
comes from the fused wavefront epilogue (plugin ``wavefront_chief_ray``).  Declines (reference code runs instead):
gradients wanted, CPU tensors, a sample count that does not match the pupil grid's unit-disk mask.
"""
from __future__ import annotations

from collections import OrderedDict

_cell_cache: OrderedDict = OrderedDict()


def _cell_map(be, num_rays: int, like):
    """int32 (num_rays^2) tensor: index of the wavefront sample of every pupil-grid cell, -1 outside the unit disk, and
    the number of in-disk cells.  The mask is formed with the SAME backend ops and precision as the reference's
    (``be.linspace(-1, 1, n)``, ``meshgrid``, ``x**2 + y**2 <= 1``: fft.py:141-145 and distribution.py:182-186), so edge
    cells fall on the same side; memoised per (num_rays, dtype, device)."""
    import torch

    key = (int(num_rays), str(like.dtype), str(like.device))
    hit = _cell_cache.get(key)
    if hit is not None:
        _cell_cache.move_to_end(key)
        return hit
    x = be.linspace(-1, 1, num_rays)
    x, y = be.meshgrid(x, x)
    mask = (x.ravel() ** 2 + y.ravel() ** 2 <= 1).to(like.device)
    idx = torch.cumsum(mask.to(torch.int32), 0, dtype=torch.int32) - 1
    cell = torch.where(mask, idx, torch.full_like(idx, -1)).contiguous()
    out = (cell, int(mask.sum()))
    _cell_cache[key] = out
    while len(_cell_cache) > 8:
        _cell_cache.popitem(last=False)
    return out


def install(P, registry, be):
    """Wrap the three methods; returns the originals for ``uninstall``."""
    from optiland.psf.fft import ScalarFFTPSF

    orig_generate = ScalarFFTPSF._generate_pupils
    orig_pad = ScalarFFTPSF._pad_pupils
    orig_compute = ScalarFFTPSF._compute_psf

    def _engine():
        eng = P._state.get("engine")
        if eng is None or not hasattr(eng, "fft_pupil") or not P._state.get("fuse_fft_psf", True):
            return None
        return eng

    def generate_pupils(self):
        import torch

        eng = _engine()
        backend = registry.get(be.get_backend())
        self._olb_fft = None
        # (only under the plugin's own backend: with the NumPy backend active the reference's code runs)
        if eng is None or not hasattr(backend, "trace_optic") or bool(backend.grad_mode.requires_grad):
            return orig_generate(self)
        field = self.fields[0]          # PSF contains a single field (fft.py:147)
        n, g = int(self.num_rays), int(self.grid_size)
        pad = (g - n) // 2
        padded, pupils = [], []
        for wl in self.wavelengths:
            data = self.get_data(field, wl)
            opd, inten = data.opd, data.intensity
            if not (torch.is_tensor(opd) and torch.is_tensor(inten)) or opd.requires_grad or inten.requires_grad:
                return orig_generate(self)
            cell, count = _cell_map(be, n, opd)
            if opd.ndim != 1 or opd.numel() != count or inten.numel() != count:
                return orig_generate(self)      # the reference's masked assignment would raise / broadcast: its call
            buf = eng.fft_pupil(opd, inten, cell, n, g)
            if buf is None:
                return orig_generate(self)
            padded.append(buf)
            pupils.append(buf[pad:pad + n, pad:pad + n])
        self._olb_fft = (pupils, padded)
        return pupils

    def _ours(self):
        st = getattr(self, "_olb_fft", None)
        if st is None or self.pupils is not st[0] or len(self.pupils) != len(st[1]):
            return None
        return st[1]

    def pad_pupils(self):
        padded = _ours(self)
        return list(padded) if padded is not None else orig_pad(self)

    def compute_psf(self):
        import torch

        eng = _engine()
        padded = _ours(self) if eng is not None else None
        if not padded:
            return orig_compute(self)
        norm = float(self._get_normalization())
        g = padded[0].shape[-1]
        psf = torch.empty((g, g), dtype=padded[0].real.dtype, device=padded[0].device)
        last = len(padded) - 1
        for j, pupil in enumerate(padded):
            amp = torch.fft.fft2(pupil)          # library FFT (cuFFT), as the reference's be.fft.fft2
            eng.fft_psf_accumulate(amp, psf, j == 0, j == last, norm, 100.0)
        return psf

    ScalarFFTPSF._generate_pupils = generate_pupils
    ScalarFFTPSF._pad_pupils = pad_pupils
    ScalarFFTPSF._compute_psf = compute_psf
    return {"generate": orig_generate, "pad": orig_pad, "compute": orig_compute}


def uninstall(saved):
    from optiland.psf.fft import ScalarFFTPSF

    ScalarFFTPSF._generate_pupils = saved["generate"]
    ScalarFFTPSF._pad_pupils = saved["pad"]
    ScalarFFTPSF._compute_psf = saved["compute"]
