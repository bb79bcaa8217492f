// Store-pattern microbenchmark, part 2: does a different tile shape beat 6.05 TB/s?
//   V float4 per thread and row, either adjacent (32/64 B per thread) or interleaved (V separate
//   512 B warp segments); block size 128 / 256 / 512; one tile per CTA.
#include <cuda_runtime.h>
#include <cstdio>

template <int V, bool ADJ, int BS>
__global__ void __launch_bounds__(BS) pattern(const float* __restrict__ in, float* __restrict__ rec, long n, int S) {
  const long per_tile = (long)BS * 4 * V;
  const long tile = blockIdx.x;
  float4 v[8][V];
  long off[V];
#pragma unroll
  for (int j = 0; j < V; ++j)
    off[j] = ADJ ? tile * per_tile + ((long)threadIdx.x * V + j) * 4 : tile * per_tile + ((long)j * BS + threadIdx.x) * 4;
  if (off[V - 1] + 4 > n) return;
#pragma unroll
  for (int q = 0; q < 8; ++q)
#pragma unroll
    for (int j = 0; j < V; ++j) v[q][j] = __ldcs(reinterpret_cast<const float4*>(in + q * n + off[j]));
  for (int s = 0; s < S; ++s) {
#pragma unroll
    for (int q = 0; q < 8; ++q)
#pragma unroll
      for (int j = 0; j < V; ++j) {
        v[q][j].x = fmaf(v[q][j].x, 1.0000001f, 1e-9f);
        __stcs(reinterpret_cast<float4*>(rec + ((long)q * S + s) * n + off[j]), v[q][j]);
      }
  }
}

template <int V, bool ADJ, int BS>
void run(const float* in, float* rec, long n, int S, cudaEvent_t a, cudaEvent_t b) {
  long per_tile = (long)BS * 4 * V;
  int grid = (int)((n + per_tile - 1) / per_tile);
  float best = 1e9;
  for (int it = 0; it < 8; ++it) {
    cudaEventRecord(a);
    pattern<V, ADJ, BS><<<grid, BS>>>(in, rec, n, S);
    cudaEventRecord(b); cudaEventSynchronize(b);
    float ms; cudaEventElapsedTime(&ms, a, b); if (it > 1 && ms < best) best = ms;
  }
  double gb = 4.0 * n * (8 + 8 * S) / 1e9;
  printf("{\"V\": %d, \"adjacent\": %d, \"block\": %d, \"grid\": %d, \"ms\": %.4f, \"GBps\": %.1f}\n", V, (int)ADJ, BS, grid, best, gb / best * 1e3);
}

int main() {
  long n = 10000000 / 4096 * 4096; int S = 13;
  float *in, *rec;
  cudaMalloc(&in, 8 * n * 4); cudaMalloc(&rec, 8L * S * n * 4);
  cudaMemset(in, 0, 8 * n * 4);
  cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
  run<1, true, 128>(in, rec, n, S, a, b);
  run<1, true, 256>(in, rec, n, S, a, b);
  run<1, true, 512>(in, rec, n, S, a, b);
  run<2, true, 128>(in, rec, n, S, a, b);
  run<2, true, 256>(in, rec, n, S, a, b);
  run<2, false, 128>(in, rec, n, S, a, b);
  run<2, false, 256>(in, rec, n, S, a, b);
  run<4, false, 128>(in, rec, n, S, a, b);
  run<4, true, 128>(in, rec, n, S, a, b);
  return 0;
}
