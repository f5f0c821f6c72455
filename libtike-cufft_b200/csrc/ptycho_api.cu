// C ABI of the B200-native ptychography library (include/ptychofft_b200.h): plan object, launch
// glue for the fused kernels of ptycho_passes.cuh, and the small vector kernels of the CG loop.
//
// Replaces /root/reference/src/cuda/{ptychofft.cu,kernels.cu} (cuFFT plan + muloperator) and the
// CuPy elementwise/reduction code of src/libtike/cufft/ptycho.py:283-488.  Built in-tree by
// __graft_entry__.build() with
//   nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -shared -Xcompiler -fPIC
// There is no CPU or library fallback: unsupported sizes return PTX_EUNSUPPORTED.
#include <cuda_runtime.h>
#include <stdarg.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <atomic>
#include <new>
#include <vector>

#include "../../include/ptychofft_b200.h"
#include "ptycho_ops.h"

namespace ptx {

// ------------------------------------------------------------------------------------------
// small vector kernels on object / probe sized arrays
// ------------------------------------------------------------------------------------------
__global__ void k_dy_reduce(const float2* __restrict__ gr, const float2* __restrict__ g0,
                            const float2* __restrict__ d, size_t n, double* out) {
  __shared__ double red[8 * 3];
  double acc[3] = {0.0, 0.0, 0.0};
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n;
       i += (size_t)gridDim.x * blockDim.x) {
    const float2 a = gr[i], b = g0[i], c = d[i];
    const float2 df = make_float2(a.x - b.x, a.y - b.y);
    acc[0] += (double)(a.x * a.x + a.y * a.y);
    acc[1] += (double)(c.x * df.x + c.y * df.y);  // conj(d) * (g - g0)
    acc[2] += (double)(c.x * df.y - c.y * df.x);
  }
  block_reduce_add<3, 8>(acc, red, out, threadIdx.x);
}

__global__ void k_dy_update(float2* __restrict__ gr, float2* __restrict__ g0,
                            float2* __restrict__ d, size_t n, const double* red, int flags) {
  const int first = flags & 1;
  const bool zero_g = (flags & 2) != 0;  // the gradient is consumed here: leave it zeroed for the next pass
  float2 beta = make_float2(0.f, 0.f);
  if (!first) {
    // beta = ||g||^2 / (sum conj(d)(g-g0)) : a real divided by a complex (ptycho.py:369-371)
    const double nr = red[0], re = red[1], im = red[2];
    const double den = re * re + im * im;
    beta = make_float2((float)(nr * re / den), (float)(-nr * im / den));
  }
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n;
       i += (size_t)gridDim.x * blockDim.x) {
    const float2 a = gr[i];
    float2 r = make_float2(-a.x, -a.y);
    if (!first) {
      const float2 c = d[i];
      r.x += beta.x * c.x - beta.y * c.y;
      r.y += beta.x * c.y + beta.y * c.x;
    }
    d[i] = r;
    g0[i] = a;
    if (zero_g) gr[i] = make_float2(0.f, 0.f);
  }
}

// out = y + alpha * x   (psi + gamma dpsi of the position-correction block, ptycho.py:400)
__global__ void k_axpy_out(float2* __restrict__ out, const float2* __restrict__ y, const float2* __restrict__ x,
                           size_t n, float al) {
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n;
       i += (size_t)gridDim.x * blockDim.x) {
    const float2 a = y[i], b = x[i];
    out[i] = make_float2(a.x + al * b.x, a.y + al * b.y);
  }
}

// scan[0, :] += shifts: float32 += float64 the way CuPy casts it (ptycho.py:403)
__global__ void k_apply_shifts(float* __restrict__ scan, const double* __restrict__ shifts, size_t n) {
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n;
       i += (size_t)gridDim.x * blockDim.x)
    scan[i] = (float)((double)scan[i] + shifts[i]);
}

// dst[k] = src[idx[k]], k < 3: hands the a, b, cost of an accepted line-search candidate to the next
// iteration without a host round trip
__global__ void k_pick3(double* __restrict__ dst, const double* __restrict__ src, int i0, int i1, int i2) {
  if (threadIdx.x == 0) {
    dst[0] = src[i0];
    dst[1] = src[i1];
    dst[2] = src[i2];
  }
}

// Device-side decision of one fused line-search pass (ptycho.py:272-281): the first of the `kdec`
// candidates 2^-c0, 2^-(c0+1), ... whose cost row[1+j] is not above row[0] = f(p1) is accepted.  The kernel
// leaves HALF that step (the update the solver applies, ptycho.py:393, 461) in *gam -- 0 when none was
// accepted, which turns the updates queued behind it into no-ops -- the accepted index (or -1) in row[15]
// for the host to read whenever it gets there, and, when `carry` is given, the a, b, cost of the intensity
// at the half step (row[5+j+2], row[10+j+2], row[j+2]) for the next iteration ({1, 1, 0} when nothing was
// accepted: the probe rescaling that reads it is then a multiplication by one).
__global__ void k_ls_decide(double* __restrict__ row, int c0, int kdec, float* __restrict__ gam,
                            double* __restrict__ carry) {
  if (threadIdx.x != 0) return;
  int jj = -1;
  for (int j = 0; j < kdec; j++)
    if (!(row[1 + j] > row[0])) {
      jj = j;
      break;
    }
  row[15] = (double)jj;
  *gam = jj >= 0 ? ldexpf(0.5f, -(c0 + jj)) : 0.f;
  if (carry) {
    carry[0] = jj >= 0 ? row[5 + jj + 2] : 1.0;
    carry[1] = jj >= 0 ? row[10 + jj + 2] : 1.0;
    carry[2] = jj >= 0 ? row[jj + 2] : 0.0;
  }
}

__global__ void k_axpy_out_d(float2* __restrict__ out, const float2* __restrict__ y, const float2* __restrict__ x,
                             size_t n, const float* alpha) {
  const float al = *alpha;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n;
       i += (size_t)gridDim.x * blockDim.x) {
    const float2 a = y[i], b = x[i];
    out[i] = make_float2(a.x + al * b.x, a.y + al * b.y);
  }
}

__global__ void k_axpy(float2* __restrict__ y, const float2* __restrict__ x, size_t n,
                       const float* alpha) {
  const float al = *alpha;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n;
       i += (size_t)gridDim.x * blockDim.x) {
    float2 a = y[i];
    const float2 b = x[i];
    a.x += al * b.x;
    a.y += al * b.y;
    y[i] = a;
  }
}

__global__ void k_scale(float2* __restrict__ x, size_t n, const float* s) {
  const float sc = *s;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n;
       i += (size_t)gridDim.x * blockDim.x) {
    float2 a = x[i];
    a.x *= sc;
    a.y *= sc;
    x[i] = a;
  }
}

// CG scalars computed where their inputs are (no host round trip), with the reference's float32
// arithmetic: a, b are CuPy float32 sums; s = a/b; fpsi * (b/a); absfpsi * (a/b)^2 (ptycho.py:342-351)
__global__ void k_prep_scale(const double* red, int model, float* s_out, float* sc) {
  const float a = (float)red[0], b = (float)red[1];
  const float s = a / b;
  *s_out = s;
  sc[0] = model == PTX_MODEL_GAUSSIAN ? b / a : 1.f;
  sc[1] = (float)((double)s * (double)s);
}
// sc[2] = k / absmax^2  (ptycho.py:356: / max|probe_k|^2 ; 435, 441: / max|psi|^2 / nscan [* nmodes])
__global__ void k_prep_gscale(const float* absmax, double k, float* sc) {
  const double m = (double)*absmax;
  sc[2] = (float)(k / (m * m));
}
__global__ void k_axpy_s(float2* __restrict__ y, const float2* __restrict__ x, size_t n, float al) {
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n;
       i += (size_t)gridDim.x * blockDim.x) {
    float2 a = y[i];
    const float2 b = x[i];
    a.x += al * b.x;
    a.y += al * b.y;
    y[i] = a;
  }
}

// I += g^2 p2 + g p3 : the intensity after a step g along the searched direction, from the (p2, p3)
// the line search left behind (|t1 + g t2|^2 = |t1|^2 + g^2 |t2|^2 + 2 g Re(t1 conj t2))
__global__ void k_inten_update(float* __restrict__ inten, const float2* __restrict__ p23, size_t n, float g) {
  const float g2 = g * g;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n;
       i += (size_t)gridDim.x * blockDim.x) {
    const float2 v = __ldcs(p23 + i);
    inten[i] += g2 * v.x + g * v.y;
  }
}
__global__ void k_inten_update_d(float* __restrict__ inten, const float2* __restrict__ p23, size_t n,
                                 const float* gam) {
  const float g = *gam, g2 = g * g;
  if (g == 0.f) return;  // nothing accepted (yet): leave the map exactly as it is
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n;
       i += (size_t)gridDim.x * blockDim.x) {
    const float2 v = __ldcs(p23 + i);
    inten[i] += g2 * v.x + g * v.y;
  }
}

__global__ void k_absmax(const float2* __restrict__ x, size_t n, float* out) {
  float m = 0.f;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n;
       i += (size_t)gridDim.x * blockDim.x) {
    const float2 a = x[i];
    m = fmaxf(m, a.x * a.x + a.y * a.y);
  }
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) m = fmaxf(m, __shfl_down_sync(0xffffffffu, m, off));
  // non-negative floats order like their bit patterns
  if ((threadIdx.x & 31) == 0) atomicMax(reinterpret_cast<int*>(out), __float_as_int(sqrtf(m)));
}


// out[s][y][x] = in[ids[s]][(y + N/2) % N][(x + N/2) % N] / den : frame selection + fftshift +
// normalisation of the measured data in one pass (tests/catalyst/test_rec_script.py:44-46, 98-100, 209)
__global__ void k_prepare_data(const float* __restrict__ in, const long long* __restrict__ ids, size_t nsel,
                               int N, float den, int shift, float* __restrict__ out) {
  const int h = shift ? N / 2 : 0;
  const size_t total = nsel * (size_t)N * (size_t)(N / 4);
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total;
       i += (size_t)gridDim.x * blockDim.x) {
    const int x4 = (int)(i % (N / 4)) * 4;
    const int y = (int)((i / (N / 4)) % N);
    const size_t s = i / ((size_t)N * (N / 4));
    const size_t f = ids ? (size_t)ids[s] : s;
    int ys = y + h, xs = x4 + h;
    if (ys >= N) ys -= N;
    if (xs >= N) xs -= N;  // N/2 is a multiple of 4: the four source pixels stay contiguous
    const float4 v = __ldg(reinterpret_cast<const float4*>(in + (f * N + ys) * N + xs));
    *reinterpret_cast<float4*>(out + (s * N + y) * N + x4) = make_float4(v.x / den, v.y / den, v.z / den, v.w / den);
  }
}

}  // namespace ptx

// ==========================================================================================
// host side: plan object and C ABI
// ==========================================================================================
using namespace ptx;

static thread_local char g_err[512] = "";
static std::atomic<unsigned long long> g_launches{0};

static int fail(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
  return code;
}
#define CUDA_TRY(x)                                                                       \
  do {                                                                                    \
    cudaError_t e_ = (x);                                                                 \
    if (e_ != cudaSuccess) return fail(PTX_ECUDA, "%s: %s", #x, cudaGetErrorString(e_)); \
  } while (0)

struct ptx_plan {
  size_t ptheta, nz, n, nscan, ndet, nprb;
  const PlanOps* ops;
  bool freed;
  int device, num_sms, grid;
  int grid_k[K_COUNT];  // per kernel: CTAs that are resident at once, at most `grid` (0 = not asked yet)
  float2* tw;
  // per-CTA scratch, one contiguous [grid][...] array per kind; stash and accp are allocated the
  // first time a kernel that uses them is launched (ensure_scratch)
  float2* frame;
  float2* stash;
  float* accp;
  double* slots;
  // tensor maps of the object arrays seen lately (the CG loop alternates between a handful of pointers;
  // encoding costs a driver call per launch otherwise)
  struct TmEntry {
    const void* ptr;
    bool ok;
    CUtensorMap tm;
  } tmc[6];
  int tmc_next;
  size_t l2_window;  // bytes of `frame` covered by the persisting-L2 access-policy window (0: none)
  float l2_hit;      // hit ratio of that window (set-aside / window)
  Geo geo;
  // position correction (lazy): tables of the last upsampling factor
  double2* reg_E;
  double* reg_AT;  // [REG_NQ][N] Bessel table followed by [REG_JROWS][REG_NQ] Chebyshev table
  int reg_uf;
};

static const PlanOps* ops_for(int L) {
  switch (L) {
    case 6: return ops_l6();
    case 7: return ops_l7();
    case 8: return ops_l8();
    case 9: return ops_l9();
  }
  return nullptr;
}

static int plan_init(ptx_plan* p) {
  const PlanOps* ops = p->ops;
  std::vector<float2> tw(ops->tw_total + 1);
  ops->fill_tw(tw.data());
  CUDA_TRY(cudaMalloc(&p->tw, tw.size() * sizeof(float2)));
  CUDA_TRY(cudaMemcpy(p->tw, tw.data(), tw.size() * sizeof(float2), cudaMemcpyHostToDevice));
  // opt in to the large dynamic shared-memory carve-out for every kernel of this size class
  for (int k = 0; k < K_COUNT; ++k) {
    if (!ops->kernels[k]) continue;
    const bool reg = k == K_REG_OBJ || k == K_REG_FOURIER || k == K_REG_REAL;
    const bool pipe = k >= K_PIPE_GAUSS && k <= K_PIPEMC_POIS;
    const size_t bytes = pipe ? ops->smem_bytes_pipe
                              : reg && ops->smem_bytes_reg > ops->smem_bytes ? ops->smem_bytes_reg : ops->smem_bytes;
    CUDA_TRY(cudaFuncSetAttribute(ops->kernels[k], cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
  }
  // persistent grid: as many CTAs as fit at once
  int per_sm = 0;
  CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, ops->kernels[K_GRAD_GAUSS_OBJ],
                                                         ops->NT, ops->smem_bytes));
  if (per_sm < 1)
    return fail(PTX_ECUDA, "kernel does not fit on an SM (smem %zu B)", ops->smem_bytes);
  p->grid = p->num_sms * per_sm;
  if (const char* e = getenv("PTX_GRID")) {  // experiment: cap the persistent grid (L2 footprint of the frames)
    const int cap = atoi(e);
    if (cap > 0 && cap < p->grid) p->grid = cap;
  }
  const size_t npat = p->ptheta * p->nscan;
  if ((size_t)p->grid > npat) p->grid = (int)npat;
  CUDA_TRY(cudaMalloc(&p->slots, ops->slots_per_cta * p->grid * sizeof(double)));
  if (ops->frame_per_cta) {
    // N > 128: the per-CTA staging frames (8 N^2 bytes each) are written and read back twice per
    // pattern; one contiguous array.  PTX_L2_PERSIST=1 pins it in a persisting-L2 access-policy window
    // (experiment, OFF by default).  Measured on B200 (profiles/r02a_l2_window.txt): frame READS already
    // hit L2 (85 %), every frame WRITE is written back to DRAM with or without the window (the L2
    // cleans dirty lines eagerly; DRAM is 17 % busy, not the limiter), kernel times move by +-4 % at
    // 256^2 and get 15-60 % WORSE at 512^2, and the 83 MB set-aside slows every other kernel of the
    // process that lives on L2 reuse.
    const size_t bytes = ops->frame_per_cta * p->grid * sizeof(float2);
    CUDA_TRY(cudaMalloc(&p->frame, bytes));
    const char* e = getenv("PTX_L2_PERSIST");
    if (e && !strcmp(e, "1")) {
      int max_persist = 0, max_window = 0;
      cudaDeviceGetAttribute(&max_persist, cudaDevAttrMaxPersistingL2CacheSize, p->device);
      cudaDeviceGetAttribute(&max_window, cudaDevAttrMaxAccessPolicyWindowSize, p->device);
      size_t want = bytes < (size_t)max_persist ? bytes : (size_t)max_persist, cur = 0;
      cudaDeviceGetLimit(&cur, cudaLimitPersistingL2CacheSize);
      if (want > cur && cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, want) == cudaSuccess) cur = want;
      cudaGetLastError();
      p->l2_window = bytes < (size_t)max_window ? bytes : (size_t)max_window;
      p->l2_hit = p->l2_window ? (float)((double)cur / (double)p->l2_window) : 0.f;
      if (p->l2_hit > 1.f) p->l2_hit = 1.f;
      if (cur == 0) p->l2_window = 0;
    }
  }
  return PTX_OK;
}

// stash / accp arrays on first use (a registration-only or operators-only plan never pays for them)
static int ensure_scratch(ptx_plan* p, bool stash, bool accp) {
  if (stash && !p->stash)
    CUDA_TRY(cudaMalloc(&p->stash, p->ops->stash_per_cta * p->grid * sizeof(float2)));
  if (accp && !p->accp)
    CUDA_TRY(cudaMalloc(&p->accp, p->ops->accp_per_cta * p->grid * sizeof(float)));
  return PTX_OK;
}

static bool debug_sync() {
  static const bool on = getenv("PTX_DEBUG_SYNC") != nullptr;
  return on;
}

// Tensor map of an object array [T, nz, n] complex64 (8-byte elements) with the plan's patch box.
// TMA needs a 16-byte aligned base and row pitch: odd object widths fall back to plain loads.
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*,
                                  CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion,
                                  CUtensorMapFloatOOBfill);
static EncodeTiledFn encode_fn() {
  static EncodeTiledFn fn = []() -> EncodeTiledFn {
    void* f = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f, cudaEnableDefault, &q) != cudaSuccess ||
        q != cudaDriverEntryPointSuccess)
      return nullptr;
    return (EncodeTiledFn)f;
  }();
  return fn;
}
static bool encode_object_map(const ptx_plan* p, const void* psi, CUtensorMap* tm);
static bool make_object_map(ptx_plan* p, const void* psi, CUtensorMap* tm) {
  for (auto& e : p->tmc)
    if (e.ptr == psi && psi) {
      if (e.ok) *tm = e.tm;
      return e.ok;
    }
  auto& e = p->tmc[p->tmc_next];
  p->tmc_next = (p->tmc_next + 1) % 6;
  e.ptr = psi;
  e.ok = encode_object_map(p, psi, &e.tm);
  if (e.ok) *tm = e.tm;
  return e.ok;
}
static bool encode_object_map(const ptx_plan* p, const void* psi, CUtensorMap* tm) {
  const PlanOps* ops = p->ops;
  if (!ops->patch_w || !psi || ((uintptr_t)psi & 15) || (p->n & 1) || !encode_fn()) return false;
  const cuuint64_t dims[3] = {p->n, p->nz, p->ptheta};
  const cuuint64_t strides[2] = {p->n * 8, p->n * p->nz * 8};
  const cuuint32_t box[3] = {(cuuint32_t)ops->patch_w, (cuuint32_t)ops->patch_h, 1};
  const cuuint32_t estr[3] = {1, 1, 1};
  return encode_fn()(tm, CU_TENSOR_MAP_DATA_TYPE_UINT64, 3, const_cast<void*>(psi), dims, strides, box,
                     estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                     CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

static bool reg_kernel(int kid) { return kid == K_REG_OBJ || kid == K_REG_FOURIER || kid == K_REG_REAL; }

static int launch(ptx_plan* p, int kid, PassArgs& a, cudaStream_t st) {
  const PlanOps* ops = p->ops;
  const int npat = a.g.T * a.g.S;
  int grid = npat < p->grid ? npat : p->grid;
  alignas(64) CUtensorMap tm_a, tm_b;
  memset(&tm_a, 0, sizeof(tm_a));
  memset(&tm_b, 0, sizeof(tm_b));
  a.use_tma = 0;
  // column-strip gather through the shared tile (ptycho_device.cuh: gather_strip).  Measured on B200
  // (profiles/r02l_strip_gather.txt): 64^2 fwd +43 %, fused gradient +6 % (object) / +13 % (probe); 128^2
  // -4 % ... +1 % (there the extra pass through the 128 KB tile eats what the cheaper loads save), so by
  // default only the 64^2 plan uses it.  PTX_STRIP=0 / 1 force it off / on for every single-tile plan.
  static const int strip = []() {
    const char* e = getenv("PTX_STRIP");
    return !e ? -1 : !strcmp(e, "0") ? 0 : 1;
  }();
  a.strip = strip < 0 ? (ops->N == 64) : strip;
  // PTX_TMA_GATHER = sel (default: where the prefetched tensor copy measured faster than strided
  // loads -- intensity and line-search passes, and the forward operator of the multi-block plans;
  // profiles/) | all | off
  static const int policy = []() {
    const char* e = getenv("PTX_TMA_GATHER");
    return !e ? 1 : !strcmp(e, "off") ? 0 : !strcmp(e, "all") ? 2 : 1;
  }();
  const bool ls = (kid == K_LS_GAUSS || kid == K_LS_POIS || kid == K_LSAB_GAUSS || kid == K_LSAB_POIS ||
                   kid == K_LSC_GAUSS || kid == K_LSC_POIS || kid == K_LSCAB_GAUSS || kid == K_LSCAB_POIS);
  const bool inten = (kid == K_INT_GAUSS || kid == K_INT_POIS);
  // the object-gradient pass of the 64^2 plan gathers the NEXT pattern inside its scatter loop
  // (scatter_gather_impl) with plain loads: no tensor map under any policy
  const bool grad_obj1 = ops->N == 64 && (kid == K_GRAD_GAUSS_OBJ || kid == K_GRAD_POIS_OBJ ||
                                          kid == K_GRADC_GAUSS_OBJ || kid == K_GRADC_POIS_OBJ);
  // (round 2, profiles/r02h_tma256.txt: at 256^2 the probe-gradient pass gains 22 % and the object pass 1 %
  // from the prefetched tensor copy, so the multi-tile plans use it for every gradient pass as well)
  const bool grad_any = kid == K_GRAD_GAUSS_OBJ || kid == K_GRAD_GAUSS_PRB || kid == K_GRAD_POIS_OBJ ||
                        kid == K_GRAD_POIS_PRB || kid == K_GRADC_GAUSS_OBJ || kid == K_GRADC_GAUSS_PRB ||
                        kid == K_GRADC_POIS_OBJ || kid == K_GRADC_POIS_PRB;
  const bool want = !grad_obj1 &&
                    (policy == 2 || (policy == 1 && (ls || inten || (ops->RC > 1 && (kid == K_FWD || grad_any)))));
  // position correction: object patches of both images by tensor copy (profiles/r02t_reg_probe.txt: 64^2
  // 0.364 -> 0.311 ms, 128^2 0.416 -> 0.403 ms, 256^2 2.061 -> 1.812 ms per 1024 patterns; identical shifts)
  const bool reg_tma = kid == K_REG_OBJ && policy != 0;
  if ((want || reg_tma) && kid != K_NEAR && kid != K_ADJ_OBJ && kid != K_ADJ_PRB &&
      (!reg_kernel(kid) || reg_tma) && a.psi) {
    const bool two = ls || reg_tma;
    if (make_object_map(p, a.psi, &tm_a) && (!two || make_object_map(p, a.psi_b, &tm_b))) a.use_tma = 1;
  }
  void* params[] = {&a, &tm_a, &tm_b};
  const bool nodata = kid == K_FWD || kid == K_NEAR || kid == K_ADJ_OBJ || kid == K_ADJ_PRB;
  const bool reg = kid == K_REG_OBJ || kid == K_REG_FOURIER || kid == K_REG_REAL;
  // scratch kinds this launch touches
  const bool grad_prb = kid == K_GRAD_GAUSS_PRB || kid == K_GRAD_POIS_PRB || kid == K_GRADC_GAUSS_PRB ||
                        kid == K_GRADC_POIS_PRB;
  const bool need_stash = kid == K_ADJ_PRB || grad_prb || (ls && !a.far_in) || kid == K_REG_OBJ ||
                          kid == K_REG_REAL;
  const bool need_accp = reg || (inten && a.nmodes > 1) || (ls && a.npairs > 1);
  int rc = ensure_scratch(p, need_stash, need_accp);
  if (rc) return rc;
  a.frame = p->frame;
  a.slots = p->slots;
  a.stash = need_stash ? p->stash : nullptr;
  a.accp = need_accp ? p->accp : nullptr;
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  const bool pipe = kid >= K_PIPE_GAUSS && kid <= K_PIPEMC_POIS;
  cfg.blockDim = dim3(pipe ? ops->NT_pipe : ops->NT);
  cfg.dynamicSmemBytes = pipe ? ops->smem_bytes_pipe
                              : reg ? ops->smem_bytes_reg : nodata ? ops->smem_bytes_nodata : ops->smem_bytes;
  // A persistent grid larger than what is resident at once only queues CTAs behind the first wave (the
  // position-correction kernel of the 64^2 plan needs 166-250 registers: 2 CTAs per SM, not the plan's 4)
  if (!p->grid_k[kid]) {
    int occ = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, ops->kernels[kid], (int)cfg.blockDim.x,
                                                      cfg.dynamicSmemBytes) != cudaSuccess || occ < 1) {
      cudaGetLastError();
      occ = 1;
    }
    const int fit = p->num_sms * occ;
    p->grid_k[kid] = fit < p->grid ? fit : p->grid;
  }
  if (grid > p->grid_k[kid]) grid = p->grid_k[kid];
  cfg.gridDim = dim3(grid);
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  if (p->l2_window) {  // keep the staging frames in the persisting part of L2; everything else streams
    attr[0].id = cudaLaunchAttributeAccessPolicyWindow;
    attr[0].val.accessPolicyWindow.base_ptr = p->frame;
    attr[0].val.accessPolicyWindow.num_bytes = p->l2_window;
    attr[0].val.accessPolicyWindow.hitRatio = p->l2_hit;
    attr[0].val.accessPolicyWindow.hitProp = cudaAccessPropertyPersisting;
    attr[0].val.accessPolicyWindow.missProp = cudaAccessPropertyStreaming;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
  }
  cudaError_t e = cudaLaunchKernelExC(&cfg, ops->kernels[kid], params);
  g_launches.fetch_add(1);
  if (e != cudaSuccess) return fail(PTX_ECUDA, "launch %s: %s", ops->names[kid], cudaGetErrorString(e));
  CUDA_TRY(cudaGetLastError());
  if (debug_sync()) {  // PTX_DEBUG_SYNC=1: attribute asynchronous faults to the kernel that raised them
    e = cudaStreamSynchronize(st);
    if (e != cudaSuccess)
      return fail(PTX_ECUDA, "%s (grid %d): %s", ops->names[kid], grid, cudaGetErrorString(e));
  }
  return PTX_OK;
}

static int check_plan(const ptx_plan* p) {
  if (!p) return fail(PTX_EINVAL, "null plan");
  if (p->freed) return fail(PTX_EFREED, "plan used after free()");
  return PTX_OK;
}

static PassArgs base_args(const ptx_plan* p) {
  PassArgs a;
  memset(&a, 0, sizeof(a));
  a.g = p->geo;
  a.tw = p->tw;
  return a;
}

extern "C" {

const char* ptx_last_error(void) { return g_err; }

unsigned long long ptx_launch_count(void) { return g_launches.load(); }

int ptx_device_ok(void) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess || n < 1) {
    cudaGetLastError();
    return 0;
  }
  int dev = 0, major = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return 0;
  if (cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev) != cudaSuccess) return 0;
  return major == 10 ? 1 : 0;
}

int ptx_create(ptx_plan** out, size_t ptheta, size_t nz, size_t n, size_t nscan, size_t ndet,
               size_t nprb) {
  if (!out) return fail(PTX_EINVAL, "null out pointer");
  *out = nullptr;
  if (!ptheta || !nz || !n || !nscan || !ndet || !nprb)
    return fail(PTX_EINVAL, "all sizes must be positive");
  if (nprb > ndet) return fail(PTX_EINVAL, "probe_shape %zu exceeds detector_shape %zu", nprb, ndet);
  if (nprb + 1 > nz || nprb + 1 > n)
    return fail(PTX_EINVAL, "object %zux%zu too small for a %zu probe", nz, n, nprb);
  int L = 0;
  while (((size_t)1 << L) < ndet) ++L;
  if (((size_t)1 << L) != ndet || !ops_for(L))
    return fail(PTX_EUNSUPPORTED,
                "detector_shape=%zu: this build has sm_100a kernels for 64, 128, 256 and 512 only "
                "(no CPU or cuFFT fallback)", ndet);
  if (ptheta * nscan > 0x7fffffffull / 2) return fail(PTX_EINVAL, "too many patterns per call");
  ptx_plan* p = new (std::nothrow) ptx_plan();
  if (!p) return fail(PTX_EINVAL, "out of host memory");
  p->ptheta = ptheta; p->nz = nz; p->n = n; p->nscan = nscan; p->ndet = ndet; p->nprb = nprb;
  p->ops = ops_for(L);
  p->freed = false;
  p->tw = nullptr;
  p->frame = nullptr;
  p->stash = nullptr;
  p->accp = nullptr;
  p->slots = nullptr;
  p->l2_window = 0;
  p->l2_hit = 0.f;
  memset(p->tmc, 0, sizeof(p->tmc));
  p->tmc_next = 0;
  p->reg_E = nullptr;
  p->reg_AT = nullptr;
  p->reg_uf = 0;
  p->geo.T = (int)ptheta; p->geo.nz = (int)nz; p->geo.n = (int)n; p->geo.S = (int)nscan;
  p->geo.P = (int)nprb; p->geo.N = (int)ndet; p->geo.o = (int)((ndet - nprb) / 2);
  p->geo.kappa = 1.0f / (float)ndet;
  cudaError_t e = cudaGetDevice(&p->device);
  if (e == cudaSuccess) e = cudaDeviceGetAttribute(&p->num_sms, cudaDevAttrMultiProcessorCount, p->device);
  int major = 0;
  if (e == cudaSuccess) e = cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, p->device);
  if (e != cudaSuccess) {
    delete p;
    return fail(PTX_ECUDA, "no usable CUDA device: %s", cudaGetErrorString(e));
  }
  if (major != 10) {
    delete p;
    return fail(PTX_EUNSUPPORTED, "device compute capability %d.x: this library is sm_100a only", major);
  }
  int rc = plan_init(p);
  if (rc) {
    if (p->tw) cudaFree(p->tw);
    if (p->frame) cudaFree(p->frame);
    if (p->slots) cudaFree(p->slots);
    delete p;
    return rc;
  }
  *out = p;
  return PTX_OK;
}

int ptx_free(ptx_plan* p) {
  if (!p) return fail(PTX_EINVAL, "null plan");
  if (!p->freed) {
    cudaFree(p->tw);
    if (p->frame) cudaFree(p->frame);
    if (p->stash) cudaFree(p->stash);
    if (p->accp) cudaFree(p->accp);
    if (p->slots) cudaFree(p->slots);
    if (p->reg_E) cudaFree(p->reg_E);
    if (p->reg_AT) cudaFree(p->reg_AT);
    p->tw = nullptr;
    p->frame = nullptr;
    p->stash = nullptr;
    p->accp = nullptr;
    p->slots = nullptr;
    p->reg_E = nullptr;
    p->reg_AT = nullptr;
    p->freed = true;
  }
  return PTX_OK;
}

int ptx_destroy(ptx_plan* p) {
  if (!p) return PTX_OK;
  ptx_free(p);
  delete p;
  return PTX_OK;
}

size_t ptx_dim(const ptx_plan* p, int which) {
  if (!p) return 0;
  switch (which) {
    case PTX_DIM_PTHETA: return p->ptheta;
    case PTX_DIM_NZ: return p->nz;
    case PTX_DIM_N: return p->n;
    case PTX_DIM_NSCAN: return p->nscan;
    case PTX_DIM_NDET: return p->ndet;
    case PTX_DIM_NPRB: return p->nprb;
  }
  return 0;
}

int ptx_fwd(ptx_plan* p, void* g, const void* f, const void* scan, const void* prb,
            size_t prb_angle_stride, void* stream) {
  int rc = check_plan(p);
  if (rc) return rc;
  if (!g || !f || !scan || !prb) return fail(PTX_EINVAL, "ptx_fwd: null array");
  PassArgs a = base_args(p);
  a.far = (float2*)g;
  a.psi = (const float2*)f;
  a.scan = (const float2*)scan;
  a.prb = (const float2*)prb;
  a.prb_ts = prb_angle_stride ? prb_angle_stride : p->nprb * p->nprb;
  return launch(p, K_FWD, a, (cudaStream_t)stream);
}

int ptx_debug_nearplane(ptx_plan* p, void* near, const void* f, const void* scan, const void* prb,
                        size_t prb_angle_stride, void* stream) {
  int rc = check_plan(p);
  if (rc) return rc;
  if (!near || !f || !scan || !prb) return fail(PTX_EINVAL, "ptx_debug_nearplane: null array");
  PassArgs a = base_args(p);
  a.far = (float2*)near;
  a.psi = (const float2*)f;
  a.scan = (const float2*)scan;
  a.prb = (const float2*)prb;
  a.prb_ts = prb_angle_stride ? prb_angle_stride : p->nprb * p->nprb;
  return launch(p, K_NEAR, a, (cudaStream_t)stream);
}

int ptx_adj(ptx_plan* p, void* f, const void* g, const void* scan, void* prb,
            size_t prb_angle_stride, int flg, void* stream) {
  int rc = check_plan(p);
  if (rc) return rc;
  if (!g || !f || !scan || !prb) return fail(PTX_EINVAL, "ptx_adj: null array");
  if (flg != 0 && flg != 1) return fail(PTX_EINVAL, "ptx_adj: flg must be 0 (object) or 1 (probe)");
  PassArgs a = base_args(p);
  a.far_in = (const float2*)g;
  a.scan = (const float2*)scan;
  const size_t ts = prb_angle_stride ? prb_angle_stride : p->nprb * p->nprb;
  if (flg == 0) {
    a.grad = (float2*)f;
    a.prb = (const float2*)prb;
    a.prb_ts = ts;
    return launch(p, K_ADJ_OBJ, a, (cudaStream_t)stream);
  } else {
    a.psi = (const float2*)f;
    a.grad = (float2*)prb;
    a.grad_ts = ts;
    return launch(p, K_ADJ_PRB, a, (cudaStream_t)stream);
  }
}

int ptx_cg_intensity(ptx_plan* p, const void* psi, const void* scan, const void* probe, int nmodes,
                     const float* data, float* inten_out, const float* iscale_dev, int model,
                     double* red, void* stream) {
  int rc = check_plan(p);
  if (rc) return rc;
  if (!psi || !scan || !probe || !data || !red || nmodes < 1)
    return fail(PTX_EINVAL, "ptx_cg_intensity: bad argument");
  PassArgs a = base_args(p);
  a.psi = (const float2*)psi;
  a.scan = (const float2*)scan;
  a.prb = (const float2*)probe;
  a.prb_ms = p->nprb * p->nprb;
  a.prb_ts = a.prb_ms * nmodes;
  a.nmodes = nmodes;
  a.data = data;
  a.inten_out = inten_out;
  a.sc = iscale_dev;
  a.red = red;
  if (model == PTX_MODEL_GAUSSIAN) return launch(p, K_INT_GAUSS, a, (cudaStream_t)stream);
  if (model == PTX_MODEL_POISSON) return launch(p, K_INT_POIS, a, (cudaStream_t)stream);
  return fail(PTX_EINVAL, "unknown model %d", model);
}

int ptx_cg_grad(ptx_plan* p, int what, const void* psi, const void* scan, const void* probe,
                int nmodes, int mode, const float* data, const float* inten_in, const float* sc,
                int model, void* grad_out, size_t grad_angle_stride, void* far_out, void* stream) {
  int rc = check_plan(p);
  if (rc) return rc;
  if (!psi || !scan || !probe || !data || !sc || !grad_out || nmodes < 1 || mode < 0 ||
      mode >= nmodes || (what != 0 && what != 1))
    return fail(PTX_EINVAL, "ptx_cg_grad: bad argument");
  PassArgs a = base_args(p);
  const size_t pp = p->nprb * p->nprb;
  a.psi = (const float2*)psi;
  a.scan = (const float2*)scan;
  a.prb = (const float2*)probe + (size_t)mode * pp;
  a.prb_ts = pp * nmodes;
  a.data = data;
  a.inten_in = inten_in;
  a.sc = sc;
  a.grad = (float2*)grad_out;
  a.grad_ts = grad_angle_stride ? grad_angle_stride : pp;
  a.far = (float2*)far_out;
  cudaStream_t st = (cudaStream_t)stream;
  // PTX_PIPE=1: object gradient of the 128^2 plan through the warp-specialised, pipelined kernel
  // (ptycho_pipe.cuh).  Parity-green but not yet faster than the single-role kernel on B200 (44.3 k vs
  // 40.0 k clk per pattern and SM, profiles/r02f_pipe_ncu.txt: both roles ~84 % busy, issue slots and the
  // probe's L2 latency bind), hence opt-in.
  static const bool use_pipe = []() {
    const char* e = getenv("PTX_PIPE");
    return e && !strcmp(e, "1");
  }();
  if (what == 0 && use_pipe && p->ops->NT_pipe && p->nprb == p->ndet &&
      (model == PTX_MODEL_GAUSSIAN || model == PTX_MODEL_POISSON)) {
    return launch(p, (model == PTX_MODEL_GAUSSIAN ? (far_out ? K_PIPEC_GAUSS : K_PIPE_GAUSS)
                                                  : (far_out ? K_PIPEC_POIS : K_PIPE_POIS)) +
                         (inten_in ? K_PIPEM_GAUSS - K_PIPE_GAUSS : 0), a, st);
  }
  if (model == PTX_MODEL_GAUSSIAN)
    return launch(p, far_out ? (what == 0 ? K_GRADC_GAUSS_OBJ : K_GRADC_GAUSS_PRB)
                             : (what == 0 ? K_GRAD_GAUSS_OBJ : K_GRAD_GAUSS_PRB), a, st);
  if (model == PTX_MODEL_POISSON)
    return launch(p, far_out ? (what == 0 ? K_GRADC_POIS_OBJ : K_GRADC_POIS_PRB)
                             : (what == 0 ? K_GRAD_POIS_OBJ : K_GRAD_POIS_PRB), a, st);
  return fail(PTX_EINVAL, "unknown model %d", model);
}

int ptx_cg_linesearch(ptx_plan* p, const void* obj_a, const void* prb_a, int nmodes_a, int mode_a0,
                      const void* obj_b, const void* prb_b, int nmodes_b, int mode_b0, int npairs,
                      const void* scan, const float* data, const float* p1_in, const void* far_a,
                      int model, int c0, int ncand, int want_ab, void* p23_out, double* cost,
                      void* stream) {
  int rc = check_plan(p);
  if (rc) return rc;
  if (!obj_a || !prb_a || !obj_b || !prb_b || !scan || !data || !cost || npairs < 1 || ncand < 1 ||
      ncand > 4 || c0 < 0 || mode_a0 < 0 || mode_b0 < 0 || mode_a0 + npairs > nmodes_a ||
      mode_b0 + npairs > nmodes_b)
    return fail(PTX_EINVAL, "ptx_cg_linesearch: bad argument");
  PassArgs a = base_args(p);
  const size_t pp = p->nprb * p->nprb;
  a.psi = (const float2*)obj_a;
  a.psi_b = (const float2*)obj_b;
  a.prb = (const float2*)prb_a + (size_t)mode_a0 * pp;
  a.prb_b = (const float2*)prb_b + (size_t)mode_b0 * pp;
  a.prb_ts = pp * nmodes_a;
  a.prb_b_ts = pp * nmodes_b;
  a.prb_ms = pp;
  a.prb_b_ms = pp;
  a.scan = (const float2*)scan;
  a.data = data;
  a.inten_in = p1_in;
  a.far_in = (const float2*)far_a;
  a.far_ms = p->ptheta * p->nscan * p->ndet * p->ndet;
  a.p23 = (float2*)p23_out;
  a.npairs = npairs;
  a.c0 = c0;
  a.ncand = ncand;
  a.red = cost;
  cudaStream_t st = (cudaStream_t)stream;
  if (model == PTX_MODEL_GAUSSIAN)
    return launch(p, far_a ? (want_ab ? K_LSCAB_GAUSS : K_LSC_GAUSS) : (want_ab ? K_LSAB_GAUSS : K_LS_GAUSS), a, st);
  if (model == PTX_MODEL_POISSON)
    return launch(p, far_a ? (want_ab ? K_LSCAB_POIS : K_LSC_POIS) : (want_ab ? K_LSAB_POIS : K_LS_POIS), a, st);
  return fail(PTX_EINVAL, "unknown model %d", model);
}

// ------------------------------------------------------------------------------------------
// position correction (ptycho_register.cuh)
// ------------------------------------------------------------------------------------------
// E[j][k] = exp(2 pi i j k_signed / (uf N)), j < U; zero rows up to REG_EROWS.  Rebuilt (stream
// ordered) only when the upsampling factor changes.
// PTX_REG_ALGO = lowrank (default: Jacobi-Anger factorisation of the window kernel on the FP64
// tensor cores) | dmma (the reference's two matrix products, on the tensor cores) | dfma (same, on
// the scalar FP64 pipe).  All three give the same shifts (tests/test_gpu_register.py).
static int reg_algo() {
  static const int algo = []() {
    const char* e = getenv("PTX_REG_ALGO");
    return !e ? 2 : !strcmp(e, "dfma") ? 0 : !strcmp(e, "dmma") ? 1 : 2;
  }();
  return algo;
}

// J_0..J_{nmax-1}(x) by Miller's backward recurrence, normalised with J_0 + 2 sum J_2k = 1
static void bessel_j(long double x, int nmax, long double* out) {
  const long double ax = fabsl(x);
  if (ax < 1e-30L) {
    for (int n = 0; n < nmax; ++n) out[n] = n == 0 ? 1.0L : 0.0L;
    return;
  }
  const int start = nmax + 40;  // x <= 0.75 pi: J_n decays like (x/2)^n / n!, far below 1e-40 here
  long double jp1 = 0.0L, j = 1e-300L, sum = 0.0L;
  std::vector<long double> v(start + 1, 0.0L);
  for (int n = start; n >= 1; --n) {  // J_{n-1} = (2n/x) J_n - J_{n+1}
    const long double jm1 = (2.0L * n / ax) * j - jp1;
    jp1 = j;
    j = jm1;
    v[n - 1] = j;
    if (((n - 1) & 1) == 0 && n - 1 > 0) sum += 2.0L * j;
  }
  sum += v[0];
  for (int n = 0; n < nmax; ++n) {
    long double r = v[n] / sum;
    if (x < 0 && (n & 1)) r = -r;  // J_n(-x) = (-1)^n J_n(x)
    out[n] = r;
  }
}

static int reg_prepare(ptx_plan* p, int uf, int* U_out, cudaStream_t st) {
  if (uf < 1) return fail(PTX_EINVAL, "upsample_factor must be a positive integer");
  const int U = (3 * uf + 1) / 2;  // ceil(1.5 uf)
  if (U > REG_UMAX)
    return fail(PTX_EUNSUPPORTED, "upsample_factor %d: this build supports factors up to 100", uf);
  *U_out = U;
  if (uf == 1 || p->reg_uf == uf) return PTX_OK;
  const size_t N = p->ndet, count = (size_t)REG_EROWS * N;
  if (!p->reg_E) CUDA_TRY(cudaMalloc(&p->reg_E, count * sizeof(double2)));
  std::vector<double2> h(count, make_double2(0.0, 0.0));
  const long long period = (long long)uf * (long long)N;
  const long double PI2 = 6.283185307179586476925286766559L;
  for (int j = 0; j < U; ++j)
    for (size_t k = 0; k < N; ++k) {
      const long long ks = k < N / 2 ? (long long)k : (long long)k - (long long)N;
      long long q = ((long long)j * ks) % period;
      if (q < 0) q += period;
      const long double ang = PI2 * (long double)q / (long double)period;
      h[(size_t)j * N + k] = make_double2((double)cosl(ang), (double)sinl(ang));
    }
  // low-rank form: a[n][k] = eps_n J_n(theta_k), theta_k = 2 pi dftshift k_signed / (uf N);
  // T[j][n] = T_n(x_j), x_j = (j - dftshift) / dftshift, zero rows beyond U
  const int dftshift = U / 2;
  const size_t na = (size_t)REG_NQ * N, nt = (size_t)REG_JROWS * REG_NQ;
  if (!p->reg_AT) CUDA_TRY(cudaMalloc(&p->reg_AT, (na + nt) * sizeof(double)));
  std::vector<double> hat(na + nt, 0.0);
  for (size_t k = 0; k < N; ++k) {
    const long long ks = k < N / 2 ? (long long)k : (long long)k - (long long)N;
    long double jn[REG_NQ];
    bessel_j(PI2 * (long double)dftshift * (long double)ks / (long double)period, REG_NQ, jn);
    for (int n = 0; n < REG_NQ; ++n) hat[(size_t)n * N + k] = (double)((n ? 2.0L : 1.0L) * jn[n]);
  }
  for (int j = 0; j < U; ++j) {
    const long double x = dftshift ? (long double)(j - dftshift) / (long double)dftshift : 0.0L;
    long double t0 = 1.0L, t1 = x;
    for (int n = 0; n < REG_NQ; ++n) {  // T_{n+1} = 2 x T_n - T_{n-1}
      hat[na + (size_t)j * REG_NQ + n] = (double)t0;
      const long double t2 = 2.0L * x * t1 - t0;
      t0 = t1;
      t1 = t2;
    }
  }
  CUDA_TRY(cudaStreamSynchronize(st));  // a previous launch may still read the old tables
  CUDA_TRY(cudaMemcpy(p->reg_E, h.data(), count * sizeof(double2), cudaMemcpyHostToDevice));
  CUDA_TRY(cudaMemcpy(p->reg_AT, hat.data(), hat.size() * sizeof(double), cudaMemcpyHostToDevice));
  p->reg_uf = uf;
  return PTX_OK;
}

int ptx_register_translation(ptx_plan* p, const void* src, const void* target, size_t nimg,
                             int fourier_space, int upsample_factor, double* shifts, void* stream) {
  int rc = check_plan(p);
  if (rc) return rc;
  if (!src || !target || !shifts || !nimg || nimg > 0x3fffffffull)
    return fail(PTX_EINVAL, "ptx_register_translation: bad argument");
  cudaStream_t st = (cudaStream_t)stream;
  int U = 0;
  rc = reg_prepare(p, upsample_factor, &U, st);
  if (rc) return rc;
  PassArgs a = base_args(p);
  a.g.T = 1;
  a.g.S = (int)nimg;
  a.far_in = (const float2*)src;
  a.psi_b = (const float2*)target;
  a.reg_E = p->reg_E;
  a.reg_U = U;
  a.reg_uf = upsample_factor;
  a.reg_out = shifts;
  a.reg_A = p->reg_AT;
  a.reg_T = p->reg_AT ? p->reg_AT + (size_t)REG_NQ * p->ndet : nullptr;
  a.reg_algo = reg_algo();
  return launch(p, fourier_space ? K_REG_FOURIER : K_REG_REAL, a, st);
}

int ptx_cg_position_shifts(ptx_plan* p, const void* psi_a, const void* psi_b, const void* scan,
                           int upsample_factor, double* shifts, void* stream) {
  int rc = check_plan(p);
  if (rc) return rc;
  if (!psi_a || !psi_b || !scan || !shifts) return fail(PTX_EINVAL, "ptx_cg_position_shifts: null array");
  cudaStream_t st = (cudaStream_t)stream;
  int U = 0;
  rc = reg_prepare(p, upsample_factor, &U, st);
  if (rc) return rc;
  PassArgs a = base_args(p);
  a.g.T = 1;  // angle 0 of the chunk only, like the reference (ptycho.py:399-403)
  a.psi = (const float2*)psi_a;
  a.psi_b = (const float2*)psi_b;
  a.scan = (const float2*)scan;
  a.reg_E = p->reg_E;
  a.reg_U = U;
  a.reg_uf = upsample_factor;
  a.reg_out = shifts;
  a.reg_A = p->reg_AT;
  a.reg_T = p->reg_AT ? p->reg_AT + (size_t)REG_NQ * p->ndet : nullptr;
  a.reg_algo = reg_algo();
  return launch(p, K_REG_OBJ, a, st);
}

int ptx_prepare_data(const float* raw, const long long* ids, size_t nsel, size_t n, float denominator,
                     int fftshift, float* out, void* stream) {
  if (!raw || !out || !nsel) return fail(PTX_EINVAL, "ptx_prepare_data: bad argument");
  if (n < 8 || n % 8) return fail(PTX_EINVAL, "ptx_prepare_data: frame size must be a multiple of 8");
  if (((uintptr_t)raw | (uintptr_t)out) & 15) return fail(PTX_EINVAL, "ptx_prepare_data: arrays must be 16-byte aligned");
  const size_t total = nsel * n * (n / 4);
  size_t b = (total + 255) / 256;
  static const size_t cap = []() {
    int dev = 0, sms = 0;
    if (cudaGetDevice(&dev) != cudaSuccess ||
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || sms < 1)
      sms = 148;
    return (size_t)sms * 16;
  }();
  const int grid = (int)(b > cap ? cap : b);
  k_prepare_data<<<grid, 256, 0, (cudaStream_t)stream>>>(raw, ids, nsel, (int)n, denominator, fftshift, out);
  g_launches.fetch_add(1);
  CUDA_TRY(cudaGetLastError());
  return PTX_OK;
}

static int vec_grid(size_t n) {
  size_t b = (n + 255) / 256;
  return (int)(b < 1 ? 1 : (b > 1184 ? 1184 : b));
}

int ptx_cg_intensity_step(float* inten, const void* p23, size_t n, float step, void* stream) {
  if (!inten || !p23) return fail(PTX_EINVAL, "ptx_cg_intensity_step: null array");
  k_inten_update<<<vec_grid(n), 256, 0, (cudaStream_t)stream>>>(inten, (const float2*)p23, n, step);
  g_launches.fetch_add(1);
  CUDA_TRY(cudaGetLastError());
  return PTX_OK;
}

int ptx_cg_intensity_step_dev(float* inten, const void* p23, size_t n, const float* step_dev, void* stream) {
  if (!inten || !p23 || !step_dev) return fail(PTX_EINVAL, "ptx_cg_intensity_step_dev: null array");
  k_inten_update_d<<<vec_grid(n), 256, 0, (cudaStream_t)stream>>>(inten, (const float2*)p23, n, step_dev);
  g_launches.fetch_add(1);
  CUDA_TRY(cudaGetLastError());
  return PTX_OK;
}

int ptx_cg_ls_decide(double* cost_row, int c0, int kdec, float* half_step_dev, double* carry, void* stream) {
  if (!cost_row || !half_step_dev || c0 < 0 || kdec < 0 || kdec > 4)
    return fail(PTX_EINVAL, "ptx_cg_ls_decide: bad argument");
  k_ls_decide<<<1, 32, 0, (cudaStream_t)stream>>>(cost_row, c0, kdec, half_step_dev, carry);
  g_launches.fetch_add(1);
  CUDA_TRY(cudaGetLastError());
  return PTX_OK;
}

int ptx_vec_axpy_out_dev(void* out, const void* y, const void* x, size_t n, const float* alpha_dev, void* stream) {
  if (!out || !y || !x || !alpha_dev) return fail(PTX_EINVAL, "ptx_vec_axpy_out_dev: null array");
  k_axpy_out_d<<<vec_grid(n), 256, 0, (cudaStream_t)stream>>>((float2*)out, (const float2*)y, (const float2*)x, n,
                                                              alpha_dev);
  g_launches.fetch_add(1);
  CUDA_TRY(cudaGetLastError());
  return PTX_OK;
}

int ptx_vec_dai_yuan_reduce(const void* g, const void* g0, const void* d, size_t n, double* red,
                            void* stream) {
  if (!g || !g0 || !d || !red) return fail(PTX_EINVAL, "ptx_vec_dai_yuan_reduce: null array");
  k_dy_reduce<<<vec_grid(n), 256, 0, (cudaStream_t)stream>>>((const float2*)g, (const float2*)g0,
                                                             (const float2*)d, n, red);
  g_launches.fetch_add(1);
  CUDA_TRY(cudaGetLastError());
  return PTX_OK;
}

int ptx_vec_dai_yuan_update(void* g, void* g0, void* d, size_t n, const double* red, int first,
                            void* stream) {
  if (!g || !g0 || !d || !red) return fail(PTX_EINVAL, "ptx_vec_dai_yuan_update: null array");
  k_dy_update<<<vec_grid(n), 256, 0, (cudaStream_t)stream>>>((float2*)g, (float2*)g0,
                                                             (float2*)d, n, red, first);
  g_launches.fetch_add(1);
  CUDA_TRY(cudaGetLastError());
  return PTX_OK;
}

int ptx_vec_axpy(void* y, const void* x, size_t n, const float* alpha_dev, void* stream) {
  if (!y || !x || !alpha_dev) return fail(PTX_EINVAL, "ptx_vec_axpy: null array");
  k_axpy<<<vec_grid(n), 256, 0, (cudaStream_t)stream>>>((float2*)y, (const float2*)x, n, alpha_dev);
  g_launches.fetch_add(1);
  CUDA_TRY(cudaGetLastError());
  return PTX_OK;
}

int ptx_vec_axpy_s(void* y, const void* x, size_t n, float alpha, void* stream) {
  if (!y || !x) return fail(PTX_EINVAL, "ptx_vec_axpy_s: null array");
  k_axpy_s<<<vec_grid(n), 256, 0, (cudaStream_t)stream>>>((float2*)y, (const float2*)x, n, alpha);
  g_launches.fetch_add(1);
  CUDA_TRY(cudaGetLastError());
  return PTX_OK;
}

int ptx_vec_axpy_out(void* out, const void* y, const void* x, size_t n, float alpha, void* stream) {
  if (!out || !y || !x) return fail(PTX_EINVAL, "ptx_vec_axpy_out: null array");
  k_axpy_out<<<vec_grid(n), 256, 0, (cudaStream_t)stream>>>((float2*)out, (const float2*)y, (const float2*)x, n, alpha);
  g_launches.fetch_add(1);
  CUDA_TRY(cudaGetLastError());
  return PTX_OK;
}

int ptx_vec_zero(void* x, size_t nbytes, void* stream) {
  if (!x) return fail(PTX_EINVAL, "ptx_vec_zero: null array");
  CUDA_TRY(cudaMemsetAsync(x, 0, nbytes, (cudaStream_t)stream));
  return PTX_OK;
}

int ptx_cg_apply_shifts(float* scan, const double* shifts, size_t nscan, void* stream) {
  if (!scan || !shifts) return fail(PTX_EINVAL, "ptx_cg_apply_shifts: null array");
  k_apply_shifts<<<vec_grid(2 * nscan), 256, 0, (cudaStream_t)stream>>>(scan, shifts, 2 * nscan);
  g_launches.fetch_add(1);
  CUDA_TRY(cudaGetLastError());
  return PTX_OK;
}

int ptx_cg_pick3(double* dst, const double* src, int i0, int i1, int i2, void* stream) {
  if (!dst || !src || i0 < 0 || i1 < 0 || i2 < 0) return fail(PTX_EINVAL, "ptx_cg_pick3: bad argument");
  k_pick3<<<1, 32, 0, (cudaStream_t)stream>>>(dst, src, i0, i1, i2);
  g_launches.fetch_add(1);
  CUDA_TRY(cudaGetLastError());
  return PTX_OK;
}

int ptx_cg_prep_scale(const double* red, int model, float* s_out, float* sc, void* stream) {
  if (!red || !s_out || !sc) return fail(PTX_EINVAL, "ptx_cg_prep_scale: null array");
  k_prep_scale<<<1, 1, 0, (cudaStream_t)stream>>>(red, model, s_out, sc);
  g_launches.fetch_add(1);
  CUDA_TRY(cudaGetLastError());
  return PTX_OK;
}

int ptx_cg_prep_gscale(const float* absmax, double k, float* sc, void* stream) {
  if (!absmax || !sc) return fail(PTX_EINVAL, "ptx_cg_prep_gscale: null array");
  k_prep_gscale<<<1, 1, 0, (cudaStream_t)stream>>>(absmax, k, sc);
  g_launches.fetch_add(1);
  CUDA_TRY(cudaGetLastError());
  return PTX_OK;
}

int ptx_vec_scale(void* x, size_t n, const float* s_dev, void* stream) {
  if (!x || !s_dev) return fail(PTX_EINVAL, "ptx_vec_scale: null array");
  k_scale<<<vec_grid(n), 256, 0, (cudaStream_t)stream>>>((float2*)x, n, s_dev);
  g_launches.fetch_add(1);
  CUDA_TRY(cudaGetLastError());
  return PTX_OK;
}

int ptx_vec_absmax(const void* x, size_t n, float* out, void* stream) {
  if (!x || !out) return fail(PTX_EINVAL, "ptx_vec_absmax: null array");
  k_absmax<<<vec_grid(n), 256, 0, (cudaStream_t)stream>>>((const float2*)x, n, out);
  g_launches.fetch_add(1);
  CUDA_TRY(cudaGetLastError());
  return PTX_OK;
}

}  // extern "C"
