// Microbenchmark: the trace kernel's HBM access pattern without its arithmetic.
// Reads 8 arrays of N floats, writes 8 x S record rows of N floats each (float4, streaming),
// persistent grid.  Variants: tile order (strided / blocked), cache operator, tiles per CTA step.
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <vector>

template <int MODE>  // 0: st.cs strided tiles, 1: plain st strided, 2: st.cs blocked (contiguous range per CTA)
__global__ void __launch_bounds__(256, 2) pattern(const float* __restrict__ in, float* __restrict__ rec, long n, int S,
                                                   int work) {
  const long per_tile = 256 * 4;
  const long n_tiles = (n + per_tile - 1) / per_tile;
  long t0 = blockIdx.x, step = gridDim.x, t1 = n_tiles;
  if (MODE == 2) {
    long chunk = (n_tiles + gridDim.x - 1) / gridDim.x;
    t0 = blockIdx.x * chunk; t1 = min(n_tiles, t0 + chunk); step = 1;
  }
  for (long tile = t0; tile < t1; tile += step) {
    long base = (tile * 256 + threadIdx.x) * 4;
    if (base + 4 > n) continue;
    float4 v[8];
#pragma unroll
    for (int q = 0; q < 8; ++q) v[q] = __ldcs(reinterpret_cast<const float4*>(in + q * n + base));
    for (int s = 0; s < S; ++s) {
      // `work` dependent FMAs per surface per value: emulates arithmetic between the stores
      for (int w = 0; w < work; ++w) {
#pragma unroll
        for (int q = 0; q < 8; ++q) { v[q].x = fmaf(v[q].x, 1.0000001f, 1e-9f); v[q].y = fmaf(v[q].y, 1.0000001f, 1e-9f);
                                      v[q].z = fmaf(v[q].z, 1.0000001f, 1e-9f); v[q].w = fmaf(v[q].w, 1.0000001f, 1e-9f); }
      }
#pragma unroll
      for (int q = 0; q < 8; ++q) {
        float4* dst = reinterpret_cast<float4*>(rec + ((long)q * S + s) * n + base);
        if (MODE == 1) *dst = v[q]; else __stcs(dst, v[q]);
      }
    }
  }
}

int main(int argc, char** argv) {
  long n = 10000000; int S = 13;
  float *in, *rec;
  cudaMalloc(&in, 8 * n * 4); cudaMalloc(&rec, 8L * S * n * 4);
  cudaMemset(in, 0, 8 * n * 4);
  cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
  auto run = [&](int mode, int grid, int work) {
    float best = 1e9;
    for (int it = 0; it < 8; ++it) {
      cudaEventRecord(a);
      if (mode == 0) pattern<0><<<grid, 256>>>(in, rec, n, S, work);
      if (mode == 1) pattern<1><<<grid, 256>>>(in, rec, n, S, work);
      if (mode == 2) pattern<2><<<grid, 256>>>(in, rec, n, S, work);
      cudaEventRecord(b); cudaEventSynchronize(b);
      float ms; cudaEventElapsedTime(&ms, a, b); if (it > 1 && ms < best) best = ms;
    }
    double gb = 4.0 * n * (8 + 8 * S) / 1e9;
    printf("{\"mode\": %d, \"grid\": %d, \"work\": %d, \"ms\": %.4f, \"GBps\": %.1f}\n", mode, grid, work, best, gb / best * 1e3);
  };
  for (int mode = 0; mode < 3; ++mode)
    for (int grid : {148, 296, 592, 1184, 9766})
      run(mode, grid, 0);
  for (int work : {1, 2, 4, 8, 16}) run(0, 296, work);
  return 0;
}
