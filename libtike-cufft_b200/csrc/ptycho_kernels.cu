// Fused sm_100a ptychography kernels and their C ABI (include/ptychofft_b200.h).
//
// Replaces /root/reference/src/cuda/{ptychofft.cu,kernels.cu} (cuFFT plan + muloperator) and the
// CuPy elementwise/reduction code of src/libtike/cufft/ptycho.py:283-488.  Built in-tree by
// __graft_entry__.build() with
//   nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -shared -Xcompiler -fPIC
// There is no CPU or library fallback: unsupported sizes return PTX_EUNSUPPORTED.
#include <cuda_runtime.h>
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <atomic>
#include <new>
#include <vector>

#include "../../include/ptychofft_b200.h"
#include "ptycho_device.cuh"

namespace ptx {

// ------------------------------------------------------------------------------------------
// kernel parameter block
// ------------------------------------------------------------------------------------------
struct PassArgs {
  Geo g;
  const float2* tw;       // twiddle tables (global, copied to smem by every CTA)
  float2* scratch;        // per-CTA thread-private scratch (L2 resident)
  size_t scratch_per_cta; // in float2
  // arrays
  const float2* psi;   // [T,nz,n]
  const float2* psi_b; // second object (line search)
  const float2* prb;   // probe base of the mode to use, angle stride prb_ts
  const float2* prb_b;
  size_t prb_ts, prb_b_ts; // complex elements between angles
  size_t prb_ms, prb_b_ms; // complex elements between modes (pair loop)
  const float2* scan;      // [T,S]
  const float* data;       // [T,S,N,N]
  const float* inten_in;   // [T,S,N,N] or null
  float* inten_out;        // [T,S,N,N] or null
  float2* far;             // [T,S,N,N]
  const float2* far_in;
  float2* grad;            // object gradient [T,nz,n] or probe gradient base (angle stride grad_ts)
  size_t grad_ts;
  const float* sc;         // device scalars
  double* red;
  int nmodes, npairs, c0, ncand;
};

template <class P>
struct Smem {
  static constexpr int TILE = TileGeom<P::L>::WORDS;       // float2
  static constexpr int TW = TwLayout<P>::TOTAL;            // float2
  static constexpr int RED = (P::NT / 32) * 12;            // doubles
  static constexpr size_t BYTES = (size_t)(TILE + TW) * sizeof(float2) + RED * sizeof(double);
};

template <class P>
__device__ __forceinline__ void smem_setup(unsigned char* raw, const float2* tw_g, float2*& tile,
                                           float2*& tw, double*& red, int tid) {
  tile = reinterpret_cast<float2*>(raw);
  tw = tile + Smem<P>::TILE;
  red = reinterpret_cast<double*>(tw + Smem<P>::TW);
  for (int i = tid; i < Smem<P>::TW; i += P::NT) tw[i] = tw_g[i];
  __syncthreads();
}

// Square root / division without the IEEE slow-path subroutine calls: MUFU.RSQ / MUFU.RCP plus one
// Newton step, accurate to ~1 ulp for the non-negative, normal-range inputs of this path.
__device__ __forceinline__ float fsqrt(float x) {
  if (x < 1e-35f) return 0.f;
  const float r = rsqrtf(x);
  float s = x * r;                       // ~sqrt(x)
  s = fmaf(fmaf(-s, s, x), 0.5f * r, s); // one Newton step
  return s;
}
__device__ __forceinline__ float fdiv(float a, float b) {
  const float r = __frcp_rn(b);
  const float q = a * r;
  return fmaf(fmaf(-q, b, a), r, q);
}

// minimisation functional per pixel (ptycho.py:308-314), x = intensity estimate, d = data
template <int MODEL>
__device__ __forceinline__ float minf_px(float x, float d, float sqd) {
  if (MODEL == PTX_MODEL_GAUSSIAN) {
    const float r = fsqrt(fabsf(x)) - sqd;
    return r * r;
  } else {
    const float ax = fabsf(x);
    return ax - d * logf(ax + 1e-32f);
  }
}

// ------------------------------------------------------------------------------------------
// API forward: g = FFT2(pad(kappa * prb * patch))                      (ptychofft.cu:60-73)
// ------------------------------------------------------------------------------------------
template <class P>
__global__ void __launch_bounds__(P::NT) k_fwd(const PassArgs a) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int tid = threadIdx.x;
  float2 *tile, *tw;
  double* red;
  smem_setup<P>(smem_raw, a.tw, tile, tw, red, tid);
  const Geo g = a.g;
  int xf0, yf0, xf2, yf2;
  fixed_coords<typename P::S0, P::WBITS>(tid, xf0, yf0);
  fixed_coords<typename P::S2, P::WBITS>(tid, xf2, yf2);
  const int npat = g.T * g.S;
  for (int pat = blockIdx.x; pat < npat; pat += gridDim.x) {
    const int t = pat / g.S;
    const Pat p = make_pat(a.scan, pat, g);
    float2* out = a.far + (size_t)pat * P::N * P::N;
    float2 v[P::E];
    if (p.skip) {
#pragma unroll
      for (int e = 0; e < P::E; ++e) out[spec_index<P>(e, xf2, yf2)] = make_float2(0.f, 0.f);
      continue;
    }
    gather_s0<P>(v, a.psi + (size_t)t * g.nz * g.n, a.prb + (size_t)t * a.prb_ts, g, p, xf0, yf0);
    fft_forward<P>(v, tile, tw, tid);
#pragma unroll
    for (int e = 0; e < P::E; ++e) out[spec_index<P>(e, xf2, yf2)] = v[e];
    __syncthreads();  // tile is reused by the next pattern's stage-0 store
  }
}

// Parity hook: the zero-padded near-plane frame (kernels.cu:95-107 output, before the FFT), natural order.
// Integer work of the path (patch origin, window offset, skip rule) is checked bit-exactly through it.
template <class P>
__global__ void __launch_bounds__(P::NT) k_nearplane(const PassArgs a) {
  const int tid = threadIdx.x;
  const Geo g = a.g;
  int xf0, yf0;
  fixed_coords<typename P::S0, P::WBITS>(tid, xf0, yf0);
  const int npat = g.T * g.S;
  for (int pat = blockIdx.x; pat < npat; pat += gridDim.x) {
    const int t = pat / g.S;
    const Pat p = make_pat(a.scan, pat, g);
    float2* out = a.far + (size_t)pat * P::N * P::N;
    float2 v[P::E];
    if (!p.skip) {
      gather_s0<P>(v, a.psi + (size_t)t * g.nz * g.n, a.prb + (size_t)t * a.prb_ts, g, p, xf0, yf0);
    } else {
#pragma unroll
      for (int e = 0; e < P::E; ++e) v[e] = make_float2(0.f, 0.f);
    }
#pragma unroll
    for (int e = 0; e < P::E; ++e) {
      int dx, dy;
      elem_offset<typename P::S0>(e, dx, dy);
      out[(yf0 | dy) * P::N + (xf0 | dx)] = v[e];
    }
  }
}

// thread-private probe-gradient accumulator kept in L2-resident scratch, flushed per angle
template <class P>
__device__ __forceinline__ void pacc_flush(float2* acc, float2* __restrict__ gp, const Geo& g,
                                           int xf0, int yf0, int tid) {
  using ST = typename P::S0;
#pragma unroll
  for (int e = 0; e < P::E; ++e) {
    int dx, dy;
    elem_offset<ST>(e, dx, dy);
    const int iy = (yf0 | dy) - g.o, ix = (xf0 | dx) - g.o;
    if ((unsigned)iy < (unsigned)g.P && (unsigned)ix < (unsigned)g.P) {
      const float2 s = acc[e * P::NT + tid];
      atomicAdd(gp + iy * g.P + ix, s);
      acc[e * P::NT + tid] = make_float2(0.f, 0.f);
    }
  }
}

// acc += scale * near * conj(patch)                                    (kernels.cu:82-94)
template <class P>
__device__ __forceinline__ void pacc_add(float2 (&v)[P::E], float2* acc,
                                         const float2* __restrict__ psi_t, float scale,
                                         const Geo& g, const Pat& p, int xf0, int yf0, int tid) {
  using ST = typename P::S0;
#pragma unroll
  for (int e = 0; e < P::E; ++e) {
    int dx, dy;
    elem_offset<ST>(e, dx, dy);
    const int iy = (yf0 | dy) - g.o, ix = (xf0 | dx) - g.o;
    if ((unsigned)iy < (unsigned)g.P && (unsigned)ix < (unsigned)g.P) {
      const float2 f = patch_at(psi_t, g, p, iy, ix);
      float2 s = acc[e * P::NT + tid];
      s.x += scale * (v[e].x * f.x + v[e].y * f.y);
      s.y += scale * (v[e].y * f.x - v[e].x * f.y);
      acc[e * P::NT + tid] = s;
    }
  }
}

// ------------------------------------------------------------------------------------------
// API adjoints: inverse FFT + object scatter (FLG 0) or probe reduction (FLG 1)  (ptychofft.cu:76-88)
// ------------------------------------------------------------------------------------------
template <class P, int FLG>
__global__ void __launch_bounds__(P::NT) k_adj(const PassArgs a) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int tid = threadIdx.x;
  float2 *tile, *tw;
  double* red;
  smem_setup<P>(smem_raw, a.tw, tile, tw, red, tid);
  const Geo g = a.g;
  int xf0, yf0, xf2, yf2;
  fixed_coords<typename P::S0, P::WBITS>(tid, xf0, yf0);
  fixed_coords<typename P::S2, P::WBITS>(tid, xf2, yf2);
  float2* acc = a.scratch + (size_t)blockIdx.x * a.scratch_per_cta;
  if (FLG == 1) {
#pragma unroll
    for (int e = 0; e < P::E; ++e) acc[e * P::NT + tid] = make_float2(0.f, 0.f);
  }
  int t_cur = -1;
  const int npat = g.T * g.S;
  for (int pat = blockIdx.x; pat < npat; pat += gridDim.x) {
    const int t = pat / g.S;
    if (FLG == 1 && t != t_cur) {
      if (t_cur >= 0) pacc_flush<P>(acc, a.grad + (size_t)t_cur * a.grad_ts, g, xf0, yf0, tid);
      t_cur = t;
    }
    const Pat p = make_pat(a.scan, pat, g);
    if (p.skip) continue;
    const float2* in = a.far_in + (size_t)pat * P::N * P::N;
    float2 v[P::E];
#pragma unroll
    for (int e = 0; e < P::E; ++e) v[e] = __ldg(in + spec_index<P>(e, xf2, yf2));
    fft_inverse<P>(v, tile, tw, tid);
    if (FLG == 0) {
      scatter_obj<P>(v, tile, a.prb + (size_t)t * a.prb_ts, g.kappa, a.grad + (size_t)t * g.nz * g.n,
                     g, p, tid);
    } else {
      pacc_add<P>(v, acc, a.psi + (size_t)t * g.nz * g.n, g.kappa, g, p, xf0, yf0, tid);
    }
    __syncthreads();
  }
  if (FLG == 1 && t_cur >= 0)
    pacc_flush<P>(acc, a.grad + (size_t)t_cur * a.grad_ts, g, xf0, yf0, tid);
}

// ------------------------------------------------------------------------------------------
// CG pass A: I = sum_k |F_k|^2, reductions a = sum sqrt(I d), b = sum I, cost   (ptycho.py:330-343)
// ------------------------------------------------------------------------------------------
template <class P, int MODEL>
__global__ void __launch_bounds__(P::NT) k_intensity(const PassArgs a) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int tid = threadIdx.x;
  float2 *tile, *tw;
  double* red;
  smem_setup<P>(smem_raw, a.tw, tile, tw, red, tid);
  const Geo g = a.g;
  int xf0, yf0, xf2, yf2;
  fixed_coords<typename P::S0, P::WBITS>(tid, xf0, yf0);
  fixed_coords<typename P::S2, P::WBITS>(tid, xf2, yf2);
  const float iscale = a.sc ? a.sc[0] : 1.f;
  double acc[3] = {0.0, 0.0, 0.0};
  const int npat = g.T * g.S;
  for (int pat = blockIdx.x; pat < npat; pat += gridDim.x) {
    const int t = pat / g.S;
    const Pat p = make_pat(a.scan, pat, g);
    float I[P::E];
#pragma unroll
    for (int e = 0; e < P::E; ++e) I[e] = 0.f;
    if (!p.skip) {
      for (int k = 0; k < a.nmodes; ++k) {
        float2 v[P::E];
        gather_s0<P>(v, a.psi + (size_t)t * g.nz * g.n, a.prb + (size_t)t * a.prb_ts + k * a.prb_ms,
                     g, p, xf0, yf0);
        fft_forward<P>(v, tile, tw, tid);
#pragma unroll
        for (int e = 0; e < P::E; ++e) I[e] += v[e].x * v[e].x + v[e].y * v[e].y;
        __syncthreads();
      }
    }
    const float* d = a.data + (size_t)pat * P::N * P::N;
    float* io = a.inten_out ? a.inten_out + (size_t)pat * P::N * P::N : nullptr;
    float sa = 0.f, sb = 0.f, scost = 0.f;
#pragma unroll
    for (int e = 0; e < P::E; ++e) {
      const int k = spec_index<P>(e, xf2, yf2);
      const float dd = __ldg(d + k);
      sa += fsqrt(I[e] * dd);
      sb += I[e];
      scost += minf_px<MODEL>(I[e] * iscale, dd, fsqrt(dd));
      if (io) io[k] = I[e];
    }
    acc[0] += (double)sa;
    acc[1] += (double)sb;
    acc[2] += (double)scost;
  }
  block_reduce_add<3, P::NT / 32>(acc, red, a.red, tid);
}

// ------------------------------------------------------------------------------------------
// CG pass B/D: fused fwd -> residual -> inverse -> object scatter (WHAT 0) / probe reduction (WHAT 1)
//   ptycho.py:347-363 (object), 421-441 (probe)
// sc = {fscale, iscale, gscale}
// ------------------------------------------------------------------------------------------
template <class P, int MODEL, int WHAT>
__global__ void __launch_bounds__(P::NT) k_grad(const PassArgs a) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int tid = threadIdx.x;
  float2 *tile, *tw;
  double* red;
  smem_setup<P>(smem_raw, a.tw, tile, tw, red, tid);
  const Geo g = a.g;
  int xf0, yf0, xf2, yf2;
  fixed_coords<typename P::S0, P::WBITS>(tid, xf0, yf0);
  fixed_coords<typename P::S2, P::WBITS>(tid, xf2, yf2);
  const float fscale = a.sc[0], iscale = a.sc[1], gscale = a.sc[2] * g.kappa;
  float2* acc = a.scratch + (size_t)blockIdx.x * a.scratch_per_cta;
  if (WHAT == 1) {
#pragma unroll
    for (int e = 0; e < P::E; ++e) acc[e * P::NT + tid] = make_float2(0.f, 0.f);
  }
  int t_cur = -1;
  const int npat = g.T * g.S;
  for (int pat = blockIdx.x; pat < npat; pat += gridDim.x) {
    const int t = pat / g.S;
    if (WHAT == 1 && t != t_cur) {
      if (t_cur >= 0) pacc_flush<P>(acc, a.grad + (size_t)t_cur * a.grad_ts, g, xf0, yf0, tid);
      t_cur = t;
    }
    const Pat p = make_pat(a.scan, pat, g);
    if (p.skip) continue;  // F = 0 -> residual 0 -> no contribution
    const float2* psi_t = a.psi + (size_t)t * g.nz * g.n;
    const float2* prb_t = a.prb + (size_t)t * a.prb_ts;
    float2 v[P::E];
    gather_s0<P>(v, psi_t, prb_t, g, p, xf0, yf0);
    fft_forward<P>(v, tile, tw, tid);
    const float* d = a.data + (size_t)pat * P::N * P::N;
    const float* ii = a.inten_in ? a.inten_in + (size_t)pat * P::N * P::N : nullptr;
#pragma unroll
    for (int e = 0; e < P::E; ++e) {
      const int k = spec_index<P>(e, xf2, yf2);
      const float dd = __ldg(d + k);
      const float I = ii ? __ldg(ii + k) * iscale : (v[e].x * v[e].x + v[e].y * v[e].y);
      float f;
      if (MODEL == PTX_MODEL_GAUSSIAN)
        f = fscale * (1.f - fdiv(fsqrt(dd), fsqrt(I) + 1e-32f));
      else
        f = fscale * (1.f - fdiv(dd, I + 1e-32f));
      v[e].x *= f;
      v[e].y *= f;
    }
    fft_inverse<P>(v, tile, tw, tid);
    if (WHAT == 0)
      scatter_obj<P>(v, tile, prb_t, gscale, a.grad + (size_t)t * g.nz * g.n, g, p, tid);
    else
      pacc_add<P>(v, acc, psi_t, gscale, g, p, xf0, yf0, tid);
    __syncthreads();
  }
  if (WHAT == 1 && t_cur >= 0)
    pacc_flush<P>(acc, a.grad + (size_t)t_cur * a.grad_ts, g, xf0, yf0, tid);
}

// ------------------------------------------------------------------------------------------
// CG pass C/E: line-search costs for up to 8 step candidates at once
//   ptycho.py:383-393 (object), 451-461 (probe), 253-281 (line_search_sqr)
// The first far field of a pair is parked in thread-private scratch while the second is transformed.
// ------------------------------------------------------------------------------------------
template <class P, int MODEL>
__global__ void __launch_bounds__(P::NT) k_linesearch(const PassArgs a) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int tid = threadIdx.x;
  float2 *tile, *tw;
  double* red;
  smem_setup<P>(smem_raw, a.tw, tile, tw, red, tid);
  const Geo g = a.g;
  int xf0, yf0, xf2, yf2;
  fixed_coords<typename P::S0, P::WBITS>(tid, xf0, yf0);
  fixed_coords<typename P::S2, P::WBITS>(tid, xf2, yf2);
  float2* stash = a.scratch + (size_t)blockIdx.x * a.scratch_per_cta;  // E*NT float2
  float* accp = reinterpret_cast<float*>(stash + P::E * P::NT);        // 3*E*NT floats
  double acc[9];
#pragma unroll
  for (int c = 0; c < 9; ++c) acc[c] = 0.0;
  const bool multi = a.npairs > 1;
  const int npat = g.T * g.S;
  for (int pat = blockIdx.x; pat < npat; pat += gridDim.x) {
    const int t = pat / g.S;
    const Pat p = make_pat(a.scan, pat, g);
    const float* d = a.data + (size_t)pat * P::N * P::N;
    const float* p1in = a.inten_in ? a.inten_in + (size_t)pat * P::N * P::N : nullptr;
    float cost[9];
#pragma unroll
    for (int c = 0; c < 9; ++c) cost[c] = 0.f;
    if (multi) {
#pragma unroll
      for (int e = 0; e < P::E; ++e) {
        accp[(0 * P::E + e) * P::NT + tid] = 0.f;
        accp[(1 * P::E + e) * P::NT + tid] = 0.f;
        accp[(2 * P::E + e) * P::NT + tid] = 0.f;
      }
    }
    for (int j = 0; j < a.npairs; ++j) {
      float2 v[P::E];
      if (!p.skip) {
        gather_s0<P>(v, a.psi + (size_t)t * g.nz * g.n, a.prb + (size_t)t * a.prb_ts + j * a.prb_ms,
                     g, p, xf0, yf0);
        fft_forward<P>(v, tile, tw, tid);
      } else {
#pragma unroll
        for (int e = 0; e < P::E; ++e) v[e] = make_float2(0.f, 0.f);
      }
#pragma unroll
      for (int e = 0; e < P::E; ++e) stash[e * P::NT + tid] = v[e];
      __syncthreads();
      if (!p.skip) {
        gather_s0<P>(v, a.psi_b + (size_t)t * g.nz * g.n,
                     a.prb_b + (size_t)t * a.prb_b_ts + j * a.prb_b_ms, g, p, xf0, yf0);
        fft_forward<P>(v, tile, tw, tid);
      }
#pragma unroll
      for (int e = 0; e < P::E; ++e) {
        const float2 t1 = stash[e * P::NT + tid];
        const float2 t2 = v[e];
        float q1 = t1.x * t1.x + t1.y * t1.y;
        float q2 = t2.x * t2.x + t2.y * t2.y;
        float q3 = 2.f * (t1.x * t2.x + t1.y * t2.y);
        if (multi) {
          q1 += accp[(0 * P::E + e) * P::NT + tid];
          q2 += accp[(1 * P::E + e) * P::NT + tid];
          q3 += accp[(2 * P::E + e) * P::NT + tid];
          if (j + 1 < a.npairs) {
            accp[(0 * P::E + e) * P::NT + tid] = q1;
            accp[(1 * P::E + e) * P::NT + tid] = q2;
            accp[(2 * P::E + e) * P::NT + tid] = q3;
          }
        }
        if (j + 1 == a.npairs) {
          const int k = spec_index<P>(e, xf2, yf2);
          const float dd = __ldg(d + k);
          const float sqd = fsqrt(dd);
          if (p1in) q1 = __ldg(p1in + k);
          cost[0] += minf_px<MODEL>(q1, dd, sqd);
          float gam = exp2f(-(float)a.c0);
          for (int c = 0; c < a.ncand; ++c) {
            cost[1 + c] += minf_px<MODEL>(q1 + gam * gam * q2 + gam * q3, dd, sqd);
            gam *= 0.5f;
          }
        }
      }
      __syncthreads();
    }
#pragma unroll
    for (int c = 0; c < 9; ++c) acc[c] += (double)cost[c];
  }
  block_reduce_add<9, P::NT / 32>(acc, red, a.red, tid);
}

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

__global__ void k_dy_update(const float2* __restrict__ gr, float2* __restrict__ g0,
                            float2* __restrict__ d, size_t n, const double* red, int first) {
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
  int L;
  bool freed;
  int device, num_sms, grid;
  float2* tw;
  float2* scratch;
  size_t scratch_per_cta;  // float2
  Geo geo;
};

template <class P>
static int plan_init(ptx_plan* p) {
  std::vector<float2> tw(TwLayout<P>::TOTAL);
  fill_twiddles<P>(tw.data());
  CUDA_TRY(cudaMalloc(&p->tw, tw.size() * sizeof(float2)));
  CUDA_TRY(cudaMemcpy(p->tw, tw.data(), tw.size() * sizeof(float2), cudaMemcpyHostToDevice));
  // opt in to the large dynamic shared-memory carve-out for every kernel of this size class
  void (*kernels[])(const PassArgs) = {
      k_fwd<P>,           k_adj<P, 0>,        k_adj<P, 1>,        k_intensity<P, 0>,
      k_intensity<P, 1>,  k_grad<P, 0, 0>,    k_grad<P, 0, 1>,    k_grad<P, 1, 0>,
      k_grad<P, 1, 1>,    k_linesearch<P, 0>, k_linesearch<P, 1>};
  for (auto k : kernels)
    CUDA_TRY(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  (int)Smem<P>::BYTES));
  // persistent grid: as many CTAs as fit at once
  int per_sm = 0;
  CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_grad<P, 0, 0>, P::NT,
                                                         Smem<P>::BYTES));
  if (per_sm < 1) return fail(PTX_ECUDA, "kernel does not fit on an SM (smem %zu B)", Smem<P>::BYTES);
  p->grid = p->num_sms * per_sm;
  // scratch: E*NT float2 (stash / probe accumulators) + 3*E*NT floats (p1,p2,p3 accumulators)
  p->scratch_per_cta = (size_t)P::E * P::NT + (size_t)3 * P::E * P::NT / 2;
  CUDA_TRY(cudaMalloc(&p->scratch, p->scratch_per_cta * p->grid * sizeof(float2)));
  return PTX_OK;
}

static bool debug_sync() {
  static const bool on = getenv("PTX_DEBUG_SYNC") != nullptr;
  return on;
}

template <class P, class K>
static int launch_named(ptx_plan* p, K kernel, const char* name, const PassArgs& a, cudaStream_t st) {
  const int npat = a.g.T * a.g.S;
  const int grid = npat < p->grid ? npat : p->grid;
  kernel<<<grid, P::NT, Smem<P>::BYTES, st>>>(a);
  g_launches.fetch_add(1);
  CUDA_TRY(cudaGetLastError());
  if (debug_sync()) {  // PTX_DEBUG_SYNC=1: attribute asynchronous faults to the kernel that raised them
    cudaError_t e = cudaStreamSynchronize(st);
    if (e != cudaSuccess) return fail(PTX_ECUDA, "%s (grid %d): %s", name, grid, cudaGetErrorString(e));
  }
  return PTX_OK;
}
#define LAUNCH(st, a, ...) launch_named<PL>(p, __VA_ARGS__, #__VA_ARGS__, a, st)

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
  a.scratch = p->scratch;
  a.scratch_per_cta = p->scratch_per_cta;
  return a;
}

#define DISPATCH_L(p, STMT)                                             \
  switch ((p)->L) {                                                     \
    case 6: {                                                           \
      using PL = Plan<6>;                                               \
      STMT;                                                             \
    } break;                                                            \
    case 7: {                                                           \
      using PL = Plan<7>;                                               \
      STMT;                                                             \
    } break;                                                            \
    default:                                                            \
      return fail(PTX_EUNSUPPORTED, "ndet=%zu not built", (p)->ndet);   \
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
  if (((size_t)1 << L) != ndet || (L != 6 && L != 7))
    return fail(PTX_EUNSUPPORTED,
                "detector_shape=%zu: this build has sm_100a kernels for 64 and 128 only "
                "(no CPU or cuFFT fallback)", ndet);
  if (ptheta * nscan > 0x7fffffffull / 2) return fail(PTX_EINVAL, "too many patterns per call");
  ptx_plan* p = new (std::nothrow) ptx_plan();
  if (!p) return fail(PTX_EINVAL, "out of host memory");
  p->ptheta = ptheta; p->nz = nz; p->n = n; p->nscan = nscan; p->ndet = ndet; p->nprb = nprb;
  p->L = L;
  p->freed = false;
  p->tw = nullptr;
  p->scratch = nullptr;
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
  int rc;
  if (L == 6) rc = plan_init<Plan<6>>(p); else rc = plan_init<Plan<7>>(p);
  if (rc) {
    if (p->tw) cudaFree(p->tw);
    if (p->scratch) cudaFree(p->scratch);
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
    cudaFree(p->scratch);
    p->tw = nullptr;
    p->scratch = nullptr;
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
  DISPATCH_L(p, return (LAUNCH((cudaStream_t)stream, a, k_fwd<PL>)));
  return PTX_OK;
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
  const int npat = a.g.T * a.g.S;
  const int grid = npat < p->grid ? npat : p->grid;
  switch (p->L) {
    case 6: k_nearplane<Plan<6>><<<grid, Plan<6>::NT, 0, (cudaStream_t)stream>>>(a); break;
    case 7: k_nearplane<Plan<7>><<<grid, Plan<7>::NT, 0, (cudaStream_t)stream>>>(a); break;
    default: return fail(PTX_EUNSUPPORTED, "ndet=%zu not built", p->ndet);
  }
  g_launches.fetch_add(1);
  CUDA_TRY(cudaGetLastError());
  return PTX_OK;
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
    DISPATCH_L(p, return (LAUNCH((cudaStream_t)stream, a, k_adj<PL, 0>)));
  } else {
    a.psi = (const float2*)f;
    a.grad = (float2*)prb;
    a.grad_ts = ts;
    DISPATCH_L(p, return (LAUNCH((cudaStream_t)stream, a, k_adj<PL, 1>)));
  }
  return PTX_OK;
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
  if (model == PTX_MODEL_GAUSSIAN) {
    DISPATCH_L(p, return (LAUNCH((cudaStream_t)stream, a, k_intensity<PL, 0>)));
  } else if (model == PTX_MODEL_POISSON) {
    DISPATCH_L(p, return (LAUNCH((cudaStream_t)stream, a, k_intensity<PL, 1>)));
  }
  return fail(PTX_EINVAL, "unknown model %d", model);
}

int ptx_cg_grad(ptx_plan* p, int what, const void* psi, const void* scan, const void* probe,
                int nmodes, int mode, const float* data, const float* inten_in, const float* sc,
                int model, void* grad_out, size_t grad_angle_stride, void* stream) {
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
  cudaStream_t st = (cudaStream_t)stream;
  if (model == PTX_MODEL_GAUSSIAN) {
    if (what == 0) { DISPATCH_L(p, return (LAUNCH(st, a, k_grad<PL, 0, 0>))); }
    else { DISPATCH_L(p, return (LAUNCH(st, a, k_grad<PL, 0, 1>))); }
  } else if (model == PTX_MODEL_POISSON) {
    if (what == 0) { DISPATCH_L(p, return (LAUNCH(st, a, k_grad<PL, 1, 0>))); }
    else { DISPATCH_L(p, return (LAUNCH(st, a, k_grad<PL, 1, 1>))); }
  }
  return fail(PTX_EINVAL, "unknown model %d", model);
}

int ptx_cg_linesearch(ptx_plan* p, const void* obj_a, const void* prb_a, int nmodes_a, int mode_a0,
                      const void* obj_b, const void* prb_b, int nmodes_b, int mode_b0, int npairs,
                      const void* scan, const float* data, const float* p1_in, int model, int c0,
                      int ncand, double* cost, void* stream) {
  int rc = check_plan(p);
  if (rc) return rc;
  if (!obj_a || !prb_a || !obj_b || !prb_b || !scan || !data || !cost || npairs < 1 || ncand < 1 ||
      ncand > 8 || c0 < 0 || mode_a0 < 0 || mode_b0 < 0 || mode_a0 + npairs > nmodes_a ||
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
  a.npairs = npairs;
  a.c0 = c0;
  a.ncand = ncand;
  a.red = cost;
  cudaStream_t st = (cudaStream_t)stream;
  if (model == PTX_MODEL_GAUSSIAN) {
    DISPATCH_L(p, return (LAUNCH(st, a, k_linesearch<PL, 0>)));
  } else if (model == PTX_MODEL_POISSON) {
    DISPATCH_L(p, return (LAUNCH(st, a, k_linesearch<PL, 1>)));
  }
  return fail(PTX_EINVAL, "unknown model %d", model);
}

static int vec_grid(size_t n) {
  size_t b = (n + 255) / 256;
  return (int)(b < 1 ? 1 : (b > 1184 ? 1184 : b));
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

int ptx_vec_dai_yuan_update(const void* g, void* g0, void* d, size_t n, const double* red, int first,
                            void* stream) {
  if (!g || !g0 || !d || !red) return fail(PTX_EINVAL, "ptx_vec_dai_yuan_update: null array");
  k_dy_update<<<vec_grid(n), 256, 0, (cudaStream_t)stream>>>((const float2*)g, (float2*)g0,
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
