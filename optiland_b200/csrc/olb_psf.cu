// olb_psf.cu -- Huygens-Fresnel PSF summation, sm_100a.
//
// Reference: NumbaSummation._huygens_fresnel_summation (optiland/psf/huygens_fresnel_strategies.py:97-160,
// a prange double loop) and TorchSummation.compute (:183-273, batched eager ops).  For every image point
// P and every pupil point Q:
//     field(P) += amp_Q * exp(-i k opd_Q) * exp(i k R) / R * 0.5 (1 + (P - Q).n_Q / R),   n_Q = Q / Rp
// psf = |field|^2.  O(N_image x N_pupil) transcendental work, no HBM traffic to speak of: the bound is
// the fp64 pipe.  k R is ~1e6 rad, so the phase must be formed in fp64; it is reduced EXACTLY to a
// fraction of a turn before the sine / cosine ((R - opd)/lambda - rint(.)), which keeps sincospi on
// its fast path (sincos(1e6) would take the Payne-Hanek slow path on every pair).
//
// One thread per image point; pupil points are staged through shared memory in tiles, each thread
// streams the tile with all operands in registers.  No tensor cores: the phase is not separable into
// a GEMM without the Fresnel approximation, which the reference does not make.
#include <cuda_runtime.h>
#include <stdint.h>

#include <string>

#include "../../include/olb.h"
#include "olb_fftpsf.cuh"

namespace olb {
int fail_psf(int code, const char* msg);   // olb_trace.cu
void count_launch();

static constexpr int PSF_BLOCK = 128;
static constexpr int PSF_TILE = 512;

struct PsfArgs {
  const double* ix; const double* iy; const double* iz;          // image points
  const double* px; const double* py; const double* pz;          // pupil points
  const double* pamp_re; const double* pamp_im; const double* popd;  // amplitude (complex allowed), opd [mm]
  double* psf; double* field;                                    // |field|^2 and (optional) complex field (re, im interleaved)
  int64_t n_img; int32_t n_pupil;
  double inv_lambda, inv_Rp;
  int32_t pupil_per_split;   // gridDim.y > 1: each y-slice of the grid sums its own pupil range into `field`
};

__global__ void __launch_bounds__(PSF_BLOCK) huygens_kernel(const __grid_constant__ PsfArgs a) {
  __shared__ double s_x[PSF_TILE], s_y[PSF_TILE], s_z[PSF_TILE], s_ar[PSF_TILE], s_ai[PSF_TILE], s_opd[PSF_TILE];
  const int64_t i = (int64_t)blockIdx.x * PSF_BLOCK + threadIdx.x;
  const bool valid = i < a.n_img;
  const double x = valid ? a.ix[i] : 0.0, y = valid ? a.iy[i] : 0.0, z = valid ? a.iz[i] : 0.0;
  double re = 0.0, im = 0.0;
  const int p_lo = blockIdx.y * a.pupil_per_split;
  const int p_hi = min(a.n_pupil, p_lo + a.pupil_per_split);
  for (int base = p_lo; base < p_hi; base += PSF_TILE) {
    const int m = min(PSF_TILE, p_hi - base);
    __syncthreads();
    for (int j = threadIdx.x; j < m; j += PSF_BLOCK) {
      s_x[j] = a.px[base + j]; s_y[j] = a.py[base + j]; s_z[j] = a.pz[base + j];
      s_ar[j] = a.pamp_re[base + j]; s_ai[j] = a.pamp_im ? a.pamp_im[base + j] : 0.0;
      s_opd[j] = a.popd[base + j];
    }
    __syncthreads();
    if (!valid) continue;
#pragma unroll 4
    for (int j = 0; j < m; ++j) {
      const double u = s_x[j], v = s_y[j], w = s_z[j];
      const double dx = x - u, dy = y - v, dz = z - w;
      const double R2 = fma(dx, dx, fma(dy, dy, dz * dz));
      const double rinv = rsqrt(R2);
      const double R = R2 * rinv;
      // phase / 2 pi = (R - opd) / lambda, reduced to [-1/2, 1/2] turns
      const double turns = (R - s_opd[j]) * a.inv_lambda;
      const double fr = turns - rint(turns);
      double sn, cs;
      sincospi(2.0 * fr, &sn, &cs);
      const double dot = fma(dx, u, fma(dy, v, dz * w)) * a.inv_Rp;      // (P - Q) . n_Q
      const double q = 0.5 * (1.0 + dot * rinv) * rinv;                   // obliquity / R
      const double ar = s_ar[j] * q, ai = s_ai[j] * q;
      re = fma(ar, cs, fma(-ai, sn, re));
      im = fma(ar, sn, fma(ai, cs, im));
    }
  }
  if (valid) {
    if (gridDim.y > 1) {   // partial sums of this pupil slice; |.|^2 is taken by finish_kernel
      atomicAdd(&a.field[2 * i], re);
      atomicAdd(&a.field[2 * i + 1], im);
    } else {
      a.psf[i] = re * re + im * im;
      if (a.field) { a.field[2 * i] = re; a.field[2 * i + 1] = im; }
    }
  }
}

__global__ void finish_kernel(const double* __restrict__ field, double* __restrict__ psf, int64_t n) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) psf[i] = field[2 * i] * field[2 * i] + field[2 * i + 1] * field[2 * i + 1];
}

}  // namespace olb

extern "C" int olb_huygens_psf_f64(const double* image_x, const double* image_y, const double* image_z, int64_t n_image,
                                   const double* pupil_x, const double* pupil_y, const double* pupil_z,
                                   const double* pupil_amp_re, const double* pupil_amp_im, const double* pupil_opd,
                                   int32_t n_pupil, double wavelength_mm, double Rp, double* psf, double* field,
                                   void* stream) {
  using namespace olb;
  if (!image_x || !image_y || !image_z || !pupil_x || !pupil_y || !pupil_z || !pupil_amp_re || !pupil_opd || !psf)
    return fail_psf(OLB_ERR_INVALID_ARG, "olb_huygens_psf_f64: NULL argument");
  if (n_image < 0 || n_pupil < 0 || !(wavelength_mm > 0) || Rp == 0)
    return fail_psf(OLB_ERR_INVALID_ARG, "olb_huygens_psf_f64: bad size / wavelength / Rp");
  if (n_image == 0) return OLB_OK;
  PsfArgs a{image_x, image_y, image_z, pupil_x, pupil_y, pupil_z, pupil_amp_re, pupil_amp_im, pupil_opd, psf, field,
            n_image, n_pupil, 1.0 / wavelength_mm, 1.0 / Rp, n_pupil};
  const int64_t grid = (n_image + PSF_BLOCK - 1) / PSF_BLOCK;
  // Few image points (a 128 x 128 PSF is 128 CTAs on 148 SMs): split the pupil sum over gridDim.y so that
  // ~8 CTAs per SM are in flight; needs the caller's `field` buffer for the partial sums.
  int splits = 1;
  if (field && grid < 8 * 148) {
    splits = (int)((8 * 148 + grid - 1) / grid);
    const int max_splits = (n_pupil + PSF_TILE - 1) / PSF_TILE;
    if (splits > max_splits) splits = max_splits;
    if (splits > 64) splits = 64;
    if (splits < 1) splits = 1;
  }
  cudaStream_t st = (cudaStream_t)stream;
  if (splits > 1) {
    a.pupil_per_split = ((n_pupil + splits - 1) / splits + PSF_TILE - 1) / PSF_TILE * PSF_TILE;
    splits = (n_pupil + a.pupil_per_split - 1) / a.pupil_per_split;
    cudaMemsetAsync(field, 0, (size_t)n_image * 2 * sizeof(double), st);
  }
  huygens_kernel<<<dim3((unsigned)grid, (unsigned)splits), PSF_BLOCK, 0, st>>>(a);
  if (splits > 1) finish_kernel<<<(unsigned)((n_image + 255) / 256), 256, 0, st>>>(field, psf, n_image);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return fail_psf(OLB_ERR_CUDA, cudaGetErrorString(e));
  count_launch();
  return OLB_OK;
}


// =============================================================================================================
// FFT-PSF gridding (second half; per-cell arithmetic and the reference citations: olb_fftpsf.cuh).
// Both kernels are one streaming pass: HBM-bound, 16 B (fp64) / 8 B (fp32) per cell written (pupil) resp. read +
// one real written (psf); grid-stride over the cells with 8 CTAs of 256 threads per SM.
// =============================================================================================================
namespace olb {

template <typename T> struct Cplx;
template <> struct Cplx<double> { using type = double2; };
template <> struct Cplx<float> { using type = float2; };

template <typename T>
__global__ void __launch_bounds__(256) fft_pupil_kernel(const T* __restrict__ opd, const T* __restrict__ intensity,
                                                        const int32_t* __restrict__ cell_ray, int32_t num_rays,
                                                        int32_t grid_size, int32_t pad, typename Cplx<T>::type* __restrict__ pupil) {
  const int64_t cells = (int64_t)grid_size * grid_size;
  for (int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; k < cells; k += (int64_t)gridDim.x * blockDim.x) {
    const int32_t r = (int32_t)(k / grid_size), c = (int32_t)(k - (int64_t)r * grid_size);
    T re, im;
    fft_pupil_cell<T>(r, c, num_rays, pad, cell_ray, opd, intensity, re, im);
    typename Cplx<T>::type v;
    v.x = re; v.y = im;
    pupil[k] = v;
  }
}

template <typename T>
__global__ void __launch_bounds__(256) fft_psf_kernel(const typename Cplx<T>::type* __restrict__ amp, int32_t grid_size,
                                                      int first, int last, T div, T mul, T* __restrict__ psf) {
  const int64_t cells = (int64_t)grid_size * grid_size;
  for (int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; k < cells; k += (int64_t)gridDim.x * blockDim.x) {
    const int32_t r = (int32_t)(k / grid_size), c = (int32_t)(k - (int64_t)r * grid_size);
    const typename Cplx<T>::type a = amp[k];                      // coalesced read of the spectrum ...
    const int64_t o = (int64_t)fftshift_index(r, grid_size) * grid_size + fftshift_index(c, grid_size);
    psf[o] = fft_psf_cell<T>(first ? (T)0 : psf[o], a.x, a.y, first != 0, last != 0, div, mul);  // ... shifted (still row-contiguous) write
  }
}

static unsigned stream_grid(int64_t cells) {
  const int64_t want = (cells + 255) / 256;
  const int64_t cap = 148 * 8;
  return (unsigned)(want < cap ? (want < 1 ? 1 : want) : cap);
}

template <typename T>
static int fft_pupil_impl(const T* opd, const T* intensity, int64_t n_samples, const int32_t* cell_ray, int32_t num_rays,
                          int32_t grid_size, T* pupil, void* stream, const char* name) {
  if (!opd || !intensity || !cell_ray || !pupil) return fail_psf(OLB_ERR_INVALID_ARG, (std::string(name) + ": NULL argument").c_str());
  if (num_rays <= 0 || grid_size < num_rays || n_samples < 0 || n_samples > (int64_t)num_rays * num_rays)
    return fail_psf(OLB_ERR_INVALID_ARG, (std::string(name) + ": need 0 < num_rays <= grid_size and n_samples <= num_rays^2").c_str());
  if ((reinterpret_cast<uintptr_t>(pupil) % (2 * sizeof(T))) != 0)
    return fail_psf(OLB_ERR_ALIGNMENT, (std::string(name) + ": pupil must be aligned to one complex element").c_str());
  const int64_t cells = (int64_t)grid_size * grid_size;
  fft_pupil_kernel<T><<<stream_grid(cells), 256, 0, (cudaStream_t)stream>>>(
      opd, intensity, cell_ray, num_rays, grid_size, (grid_size - num_rays) / 2, reinterpret_cast<typename Cplx<T>::type*>(pupil));
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return fail_psf(OLB_ERR_CUDA, cudaGetErrorString(e));
  count_launch();
  return OLB_OK;
}

template <typename T>
static int fft_psf_impl(const T* amp, int32_t grid_size, int first, int last, double div, double mul, T* psf, void* stream,
                        const char* name) {
  if (!amp || !psf) return fail_psf(OLB_ERR_INVALID_ARG, (std::string(name) + ": NULL argument").c_str());
  if (grid_size <= 0) return fail_psf(OLB_ERR_INVALID_ARG, (std::string(name) + ": grid_size must be positive").c_str());
  if ((reinterpret_cast<uintptr_t>(amp) % (2 * sizeof(T))) != 0)
    return fail_psf(OLB_ERR_ALIGNMENT, (std::string(name) + ": amp must be aligned to one complex element").c_str());
  const int64_t cells = (int64_t)grid_size * grid_size;
  fft_psf_kernel<T><<<stream_grid(cells), 256, 0, (cudaStream_t)stream>>>(
      reinterpret_cast<const typename Cplx<T>::type*>(amp), grid_size, first, last, (T)div, (T)mul, psf);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return fail_psf(OLB_ERR_CUDA, cudaGetErrorString(e));
  count_launch();
  return OLB_OK;
}

}  // namespace olb

extern "C" int olb_fft_pupil_f64(const double* opd_waves, const double* intensity, int64_t n_samples, const int32_t* cell_ray,
                                 int32_t num_rays, int32_t grid_size, double* pupil, void* stream) {
  return olb::fft_pupil_impl<double>(opd_waves, intensity, n_samples, cell_ray, num_rays, grid_size, pupil, stream, "olb_fft_pupil_f64");
}
extern "C" int olb_fft_pupil_f32(const float* opd_waves, const float* intensity, int64_t n_samples, const int32_t* cell_ray,
                                 int32_t num_rays, int32_t grid_size, float* pupil, void* stream) {
  return olb::fft_pupil_impl<float>(opd_waves, intensity, n_samples, cell_ray, num_rays, grid_size, pupil, stream, "olb_fft_pupil_f32");
}
extern "C" int olb_fft_psf_accumulate_f64(const double* amp, int32_t grid_size, int32_t first, int32_t last, double div,
                                          double mul, double* psf, void* stream) {
  return olb::fft_psf_impl<double>(amp, grid_size, first, last, div, mul, psf, stream, "olb_fft_psf_accumulate_f64");
}
extern "C" int olb_fft_psf_accumulate_f32(const float* amp, int32_t grid_size, int32_t first, int32_t last, double div,
                                          double mul, float* psf, void* stream) {
  return olb::fft_psf_impl<float>(amp, grid_size, first, last, div, mul, psf, stream, "olb_fft_psf_accumulate_f32");
}
