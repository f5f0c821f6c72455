// Micro-benchmark: cost of the object-gradient scatter (one (P+1)x(P+1) window of complex values per
// pattern, heavily overlapping windows) with different reduction instructions.
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o tools/atomics_bench tools/atomics_bench.cu
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cuda_runtime.h>

template <int MODE>
__global__ void __launch_bounds__(512) scatter(float2* obj, const int2* org, int npat, int n, int W) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int pat = blockIdx.x; pat < npat; pat += gridDim.x) {
    const int R = org[pat].x, C = org[pat].y;
    if (MODE == 3) {  // float4: two pixels per lane, window start aligned down to even column
      const int C0 = C & ~1;
      for (int i = warp; i < W; i += 16)
        for (int j = 2 * lane; j < W + 1; j += 64) {
          float4 v = make_float4(1.f, 2.f, 3.f, 4.f);
          atomicAdd(reinterpret_cast<float4*>(obj + (size_t)(R + i) * n + C0 + j), v);
        }
    } else {
      for (int i = warp; i < W; i += 16)
        for (int j = lane; j < W; j += 32) {
          float2* p = obj + (size_t)(R + i) * n + C + j;
          if (MODE == 0) { atomicAdd(&p->x, 1.f); atomicAdd(&p->y, 2.f); }
          if (MODE == 1) atomicAdd(p, make_float2(1.f, 2.f));
          if (MODE == 2) *p = make_float2(1.f, 2.f);
        }
    }
  }
}

int main() {
  const int n = 512, W = 129, side = 32, npat = side * side * 8;
  float2* obj; int2* org;
  cudaMalloc(&obj, sizeof(float2) * n * (n + 2));
  cudaMemset(obj, 0, sizeof(float2) * n * (n + 2));
  std::vector<int2> h(npat);
  for (int k = 0; k < npat; ++k) {
    int s = k % (side * side);
    h[k].x = (s / side) * 12 + rand() % 6; h[k].y = (s % side) * 12 + rand() % 6;
    if (h[k].x > n - W - 2) h[k].x = n - W - 2;
    if (h[k].y > n - W - 2) h[k].y = n - W - 2;
  }
  cudaMalloc(&org, sizeof(int2) * npat);
  cudaMemcpy(org, h.data(), sizeof(int2) * npat, cudaMemcpyHostToDevice);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  const char* names[4] = {"2x scalar atomicAdd(float)", "atomicAdd(float2)  [red.v2.f32]",
                          "plain float2 store (bound)", "atomicAdd(float4)  [red.v4.f32]"};
  for (int grid : {148, 296}) for (int mode = 0; mode < 4; ++mode) {
    float best = 1e30f;
    for (int rep = 0; rep < 4; ++rep) {
      cudaEventRecord(e0);
      if (mode == 0) scatter<0><<<grid, 512>>>(obj, org, npat, n, W);
      if (mode == 1) scatter<1><<<grid, 512>>>(obj, org, npat, n, W);
      if (mode == 2) scatter<2><<<grid, 512>>>(obj, org, npat, n, W);
      if (mode == 3) scatter<3><<<grid, 512>>>(obj, org, npat, n, W);
      cudaEventRecord(e1); cudaEventSynchronize(e1);
      float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms;
    }
    cudaError_t err = cudaGetLastError();
    printf("grid %3d  %-34s %8.3f ms  %7.2f M patterns/s  %7.1f G pixel-updates/s  (%s)\n", grid,
           names[mode], best, npat / best * 1e-3, (double)npat * W * W / best * 1e-6,
           cudaGetErrorString(err));
  }
  return 0;
}
