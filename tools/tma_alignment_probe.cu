// TMA probe written with the libcu++ wrappers of the CUDA programming guide (known-good form).
#include <cuda.h>
#include <cuda/barrier>
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>
#include <vector>
using barrier = cuda::barrier<cuda::thread_scope_block>;
namespace cde = cuda::device::experimental;
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*,
                                  CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion,
                                  CUtensorMapFloatOOBfill);
constexpr int W = 32, H = 8;
__global__ void probe(const __grid_constant__ CUtensorMap tm, int* out, int x, int y) {
#ifdef DYN
  extern __shared__ __align__(128) unsigned char dsm[];
  int (&smem_buffer)[H][W] = *reinterpret_cast<int (*)[H][W]>(dsm);
#else
  __shared__ alignas(128) int smem_buffer[H][W];
#endif
#pragma nv_diag_suppress static_var_with_dynamic_init
  __shared__ barrier bar;
  if (threadIdx.x == 0) {
    init(&bar, blockDim.x);
    cde::fence_proxy_async_shared_cta();
  }
  __syncthreads();
#ifdef MYBAR
  __shared__ unsigned long long rawbar;
  unsigned urb = (unsigned)__cvta_generic_to_shared(&rawbar);
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(urb));
    asm volatile("fence.mbarrier_init.release.cluster;");
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    asm volatile("fence.proxy.async.shared::cta;");
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(urb), "r"((unsigned)sizeof(smem_buffer)));
    unsigned utile = (unsigned)__cvta_generic_to_shared(&smem_buffer);
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
                 ::"r"(utile), "l"(&tm), "r"(x), "r"(y), "r"(urb) : "memory");
  }
  {
    unsigned done = 0;
    while (!done)
      asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}"
                   : "=r"(done) : "r"(urb), "r"(0u) : "memory");
  }
  for (int i = threadIdx.x; i < W * H; i += blockDim.x) out[i] = smem_buffer[i / W][i % W];
  return;
#endif
  barrier::arrival_token token;
  if (threadIdx.x == 0) {
#ifdef ASM
    unsigned ubar = (unsigned)__cvta_generic_to_shared(cuda::device::barrier_native_handle(bar));
    unsigned utile = (unsigned)__cvta_generic_to_shared(&smem_buffer);
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
                 ::"r"(utile), "l"(&tm), "r"(x), "r"(y), "r"(ubar) : "memory");
#else
    cde::cp_async_bulk_tensor_2d_global_to_shared(&smem_buffer, &tm, x, y, bar);
#endif
    token = cuda::device::barrier_arrive_tx(bar, 1, sizeof(smem_buffer));
  } else {
    token = bar.arrive();
  }
  bar.wait(std::move(token));
  for (int i = threadIdx.x; i < W * H; i += blockDim.x) out[i] = smem_buffer[i / W][i % W];
}
int main(int argc, char** argv) {
  const int X0 = argc > 1 ? atoi(argv[1]) : 64;
  void* f = nullptr;
  cudaDriverEntryPointQueryResult q;
  cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f, cudaEnableDefault, &q);
  EncodeTiledFn enc = (EncodeTiledFn)f;
  const int n = 1024, nz = 256;
  std::vector<int> h((size_t)nz * n);
  for (size_t i = 0; i < h.size(); ++i) h[i] = (int)i + 1;
  int *d, *o;
  cudaMalloc(&d, h.size() * 4);
  cudaMemcpy(d, h.data(), h.size() * 4, cudaMemcpyHostToDevice);
  cudaMalloc(&o, W * H * 4);
  alignas(64) CUtensorMap tm;
  cuuint64_t dims[2] = {n, nz};
  cuuint64_t strides[1] = {n * 4ull};
  cuuint32_t box[2] = {W, H};
  cuuint32_t es[2] = {1, 1};
  CUresult r = enc(&tm, CU_TENSOR_MAP_DATA_TYPE_INT32, 2, d, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   CU_TENSOR_MAP_SWIZZLE_NONE,
#ifdef L2P
 CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
#else
 CU_TENSOR_MAP_L2_PROMOTION_NONE,
#endif
 CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  #ifdef DYN
  probe<<<1, 128, W * H * 4>>>(tm, o, X0, 10);
#else
  probe<<<1, 128>>>(tm, o, X0, 10);
#endif
  cudaError_t e = cudaDeviceSynchronize();
  std::vector<int> out(W * H);
  int bad = 0;
  if (e == cudaSuccess) {
    cudaMemcpy(out.data(), o, out.size() * 4, cudaMemcpyDeviceToHost);
    for (int y = 0; y < H; ++y)
      for (int x = 0; x < W; ++x)
        if (out[y * W + x] != h[(size_t)(10 + y) * n + X0 + x]) ++bad;
  }
  printf("libcu++ 2d int32 32x8: encode=%d run=%s mismatches=%d\n", (int)r, cudaGetErrorString(e), bad);
  return 0;
}
