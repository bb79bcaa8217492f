// olb_fftpsf.cuh -- per-cell arithmetic of the FFT-PSF gridding kernels (second half).
//
// Reference: ScalarFFTPSF._generate_pupils / _pad_pupils / _compute_psf (optiland/psf/fft.py:123-227):
//   P = zeros(num_rays^2);  P[R2 <= 1] = sqrt(intensity) * exp(-1j * 2 pi * opd);  reshape;  zero-pad to grid_size
//   psf = sum_wavelengths |fftshift(fft2(P_padded))|^2 / norm * 100
// Seven element-wise passes over num_rays^2 / grid_size^2 arrays per wavelength around ONE library FFT.  Here each side
// of the FFT is one pass: `fft_pupil_cell` produces a cell of the PADDED pupil function directly (the masked scatter is
// a gather through a cell -> sample map), `fft_psf_cell` folds |.|^2, the fftshift, the sum over wavelengths and the
// normalisation into the read of the spectrum.
//
// Shared by olb_psf.cu (device) and tests/hostcheck (g++; TEST INFRASTRUCTURE) so that the index arithmetic can be
// checked in the GPU-less build container.
#ifndef OLB_FFTPSF_CUH_
#define OLB_FFTPSF_CUH_
#include <math.h>
#include <stdint.h>

#if defined(__CUDACC__)
#define OLB_FFT_HD __host__ __device__ __forceinline__
#else
#define OLB_FFT_HD inline
#endif

namespace olb {

// cos / sin of 2 pi * turns with the argument reduced EXACTLY to [-1/2, 1/2] turns first (opd is in waves: tens of
// turns on an aberrated system; the reduction is exact in floating point, the reference's 2 * pi * opd is not).
OLB_FFT_HD void cis_turns(double turns, double& cs, double& sn) {
  const double fr = turns - rint(turns);
#if defined(__CUDA_ARCH__)
  sincospi(2.0 * fr, &sn, &cs);
#else
  const double a = 6.283185307179586476925286766559 * fr;
  cs = cos(a); sn = sin(a);
#endif
}
OLB_FFT_HD void cis_turns(float turns, float& cs, float& sn) {
  const float fr = turns - rintf(turns);
#if defined(__CUDA_ARCH__)
  sincospif(2.0f * fr, &sn, &cs);
#else
  const float a = 6.2831853071795864769f * fr;
  cs = cosf(a); sn = sinf(a);
#endif
}
OLB_FFT_HD double fft_sqrt(double v) { return sqrt(v); }
OLB_FFT_HD float fft_sqrt(float v) { return sqrtf(v); }

// Cell (r, c) of the grid_size x grid_size padded pupil function.  The num_rays x num_rays pupil grid sits at rows /
// columns [pad, pad + num_rays) (pad = (grid_size - num_rays) / 2, be.pad's `pad_before`, fft.py:216-225);
// cell_ray[pr * num_rays + pc] is the index of the wavefront sample that belongs to pupil cell (pr, pc) -- the k-th cell
// with x^2 + y^2 <= 1 in row-major order holds sample k (distribution.py:175-186 builds the samples the same way) -- or
// -1 outside the unit disk.  NaN / negative intensity and non-finite OPD propagate as in the reference (sqrt, exp).
template <typename T>
OLB_FFT_HD void fft_pupil_cell(int32_t r, int32_t c, int32_t num_rays, int32_t pad, const int32_t* cell_ray, const T* opd,
                               const T* intensity, T& re, T& im) {
  re = 0; im = 0;
  const int32_t pr = r - pad, pc = c - pad;
  if (pr < 0 || pc < 0 || pr >= num_rays || pc >= num_rays) return;
  const int32_t k = cell_ray[(int64_t)pr * num_rays + pc];
  if (k < 0) return;
  const T a = fft_sqrt(intensity[k]);
  T cs, sn;
  cis_turns(-opd[k], cs, sn);            // exp(-i 2 pi opd)
  re = a * cs; im = a * sn;
}

// numpy / torch fftshift: out[(i + n / 2) % n] = in[i] along each axis (n / 2 rounded down).
OLB_FFT_HD int32_t fftshift_index(int32_t i, int32_t n) {
  const int32_t j = i + n / 2;
  return j >= n ? j - n : j;
}

// psf cell update: `first` -> acc = v, else acc += v; `last` -> acc / div * mul  (the reference's `/ norm * 100`)
template <typename T>
OLB_FFT_HD T fft_psf_cell(T acc, T re, T im, bool first, bool last, T div, T mul) {
  const T v = re * re + im * im;
  T s = first ? v : acc + v;
  if (last) s = s / div * mul;
  return s;
}

}  // namespace olb
#endif  // OLB_FFTPSF_CUH_
