// The fused sm_100a kernels, templated on the detector-size plan (fft_tile.cuh).
//
// Replaces /root/reference/src/cuda/{ptychofft.cu,kernels.cu} (cuFFT plan + muloperator) and the
// CuPy elementwise/reduction code of src/libtike/cufft/ptycho.py:283-488.  Instantiated once per
// detector size by plan_l6.cu .. plan_l9.cu; launched by ptycho_api.cu through PlanOps.
#pragma once

#include "../../include/ptychofft_b200.h"
#include "ptycho_ops.h"

namespace ptx {

template <class P>
struct Smem {
  // float2: the FFT tile; the TMA-landed object patch reuses the region and may be a bit larger
  static constexpr int TILE = (Patch<P>::TMA && Patch<P>::WORDS > TileGeom<P>::WORDS)
                                  ? ((Patch<P>::WORDS + 15) / 16 * 16)
                                  : TileGeom<P>::WORDS;
  static constexpr int TW = TwLayout<P>::TOTAL;    // float2
  static constexpr int RED = (P::NT / 32) * 16;    // doubles: cross-warp reduction scratch
  static constexpr int DBUF = P::NX * P::NY;       // floats: measured-data tile (TMA bulk copy)
  static constexpr size_t OFF_TW = (size_t)TILE * sizeof(float2);
  // [tile | twiddles | reduction scratch | mbarrier | data tile]: kernels that never read measured
  // data are launched without the last part
  static constexpr size_t OFF_RED = (OFF_TW + (size_t)TW * sizeof(float2) + 15) / 16 * 16;
  static constexpr size_t OFF_BAR = OFF_RED + (size_t)RED * sizeof(double);
  static constexpr size_t OFF_DBUF = (OFF_BAR + 16 + 127) / 128 * 128;
  static constexpr size_t BYTES_NODATA = OFF_DBUF;
  static constexpr size_t BYTES = OFF_DBUF + (size_t)DBUF * sizeof(float);
};

template <class P>
struct Scratch {  // per-CTA global scratch; every kind is its own contiguous [grid][...] array
  static constexpr size_t FRAME = P::RC > 1 ? (size_t)P::N * P::N : 0;  // float2
  static constexpr size_t STASH = (size_t)P::N * P::N;                  // float2
  static constexpr size_t ACCP = (size_t)3 * P::N * P::N;               // floats
  static constexpr size_t SLOTS = (size_t)16 * P::NT;                   // doubles
};

template <class P>
__device__ __forceinline__ void cta_setup(Cta<P>& c, unsigned char* raw, const PassArgs& a) {
  c.tid = threadIdx.x;
  c.strip = a.strip != 0;
  c.tile = reinterpret_cast<float2*>(raw);
  float2* tw = reinterpret_cast<float2*>(raw + Smem<P>::OFF_TW);
  c.dbuf = reinterpret_cast<float*>(raw + Smem<P>::OFF_DBUF);
  c.red = reinterpret_cast<double*>(raw + Smem<P>::OFF_RED);
  c.bar = reinterpret_cast<unsigned long long*>(raw + Smem<P>::OFF_BAR);
  for (int i = c.tid; i < Smem<P>::TW; i += P::NT) tw[i] = a.tw[i];
  c.tw = tw;
  // kernels are only handed the scratch kinds they use (ptycho_api.cu: launch()); the others are null
  c.frame = a.frame + (size_t)blockIdx.x * Scratch<P>::FRAME;
  c.stash = a.stash + (size_t)blockIdx.x * Scratch<P>::STASH;
  c.accp = a.accp + (size_t)blockIdx.x * Scratch<P>::ACCP;
  c.slots = a.slots + (size_t)blockIdx.x * Scratch<P>::SLOTS + c.tid;
  if (a.slots) {
#pragma unroll
    for (int k = 0; k < 16; ++k) c.slots[k * P::NT] = 0.0;
  }
  fixed_coords<typename P::S0, P::WBITS>(c.tid, c.xf0, c.yf0);
  fixed_coords<typename P::S2, P::WBITS>(c.tid, c.xf2, c.yf2);
  c.sbase = spec_base<P>(c.xf2, c.yf2);
  c.lbase = pos_to_freq_y<P>(c.yf2) * P::N + pos_to_freq_x<P>(c.xf2);
  dp_init<P>(c);
  __syncthreads();
}

// PTX_MATH selects how the per-pixel square roots, reciprocals and logarithms of the residual and of
// the cost are evaluated:
//   0 (default)  SFU approximations (MUFU.RSQ / MUFU.RCP, <= 2 ulp, relative error 2.4e-7 -- far below
//                the 1e-5 operator bar): a handful of instructions per pixel (logf is CUDA's 1-ulp one);
//   1            IEEE-rounded sqrtf / division and the reference's literal expression order
//                (ptycho.py:353, 360, 308-314).  Built as a second library by `__graft_entry__.build(
//                variants=True)` for the CG-parity A/B of tests/tools/cg_parity_probe.py.
#ifndef PTX_MATH
#define PTX_MATH 0
#endif

__device__ __forceinline__ float frsq(float x) {
#if PTX_MATH == 1
  return 1.0f / sqrtf(fmaxf(x, 1e-35f));
#else
  float r;
  asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(fmaxf(x, 1e-35f)));
  return r;
#endif
}
__device__ __forceinline__ float fsqrt(float x) {  // 0 -> 0
#if PTX_MATH == 1
  return sqrtf(x);
#else
  return x * frsq(x);
#endif
}
__device__ __forceinline__ float frcp(float x) {
#if PTX_MATH == 1
  return 1.0f / x;
#else
  float r;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
#endif
}

// residual factor of one far-field pixel (ptycho.py:353, 360): the residual is F * factor
//   gaussian: fscale * (1 - sqrt(d) / (sqrt(I) + 1e-32)) = fscale - fscale * d * rsqrt(d * I)
//   poisson:  fscale * (1 - d / (I + 1e-32))
// The one-MUFU Gaussian form needs d * I to stay a normal number: both factors are scaled by 2^32
// first (exact), which keeps the product representable from d * I = 5e-55 (data normalised far below
// 1 next to vanishing intensities) up to 1.8e19 -- no branch, two extra multiplies.
template <int MODEL>
__device__ __forceinline__ float residual_factor(float d, float I, float fscale) {
  if (MODEL == PTX_MODEL_GAUSSIAN) {
#if PTX_MATH == 1
    return fscale - fscale * sqrtf(d) / (sqrtf(I) + 1e-32f);
#else
    const float a = d * 4294967296.f;
    return fmaf(-(fscale * a), frsq(a * (I * 4294967296.f)), fscale);
#endif
  } else {
#if PTX_MATH == 1
    return fscale - fscale * d / (I + 1e-32f);
#else
    return fmaf(-(fscale * d), frcp(I + 1e-32f), fscale);
#endif
  }
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

// near plane of column block cb: object patch by TMA tensor copy when the tensor map is usable,
// strided read-only loads otherwise (odd object width, unaligned base, 512^2 detectors).
// The copy is pipelined: block 0 of a pattern is normally already in flight (patch_prefetch, issued
// when the previous tile user released the tile), and the copy of block cb+1 is issued as soon as
// block cb's taps have been read, overlapping the cross butterfly and the frame stores.
template <class P>
__device__ __forceinline__ void gather_any(float2 (&v)[P::E], Cta<P>& c, int cb, bool use_tma,
                                           const CUtensorMap* tm, int key, int t,
                                           const float2* __restrict__ psi_t,
                                           const float2* __restrict__ prb, const Geo& g,
                                           const Pat& p) {
  if (Patch<P>::TMA && use_tma) {
    if (cb == 0) {
      if (c.pp_key != key) {
        __syncthreads();  // every thread is done with the tile (e.g. the last inverse stage's loads)
        patch_issue<P>(c, tm, g, p, t, 0);
      }
      c.pp_key = -1;
    }
    const int shift = c.pshift;
    gather_tma<P>(v, c, cb, shift, prb, g, p);  // ends with a block barrier: the tile is free again
    if (cb + 1 < P::RC) patch_issue<P>(c, tm, g, p, t, cb + 1);
  } else {
    // the plain-load path is the exception where tensor copies are the rule: keep its predicate
    // arithmetic inside this branch (opaque_zero, ptycho_device.cuh); 512^2 has no tensor-copy path
    const int zero = Patch<P>::TMA ? opaque_zero() : 0;
    const Geo g2 = tied(g, zero);
    const Pat p2 = tied(p, zero);
    gather_nat<P>(v, c, cb, psi_t, prb, g2, p2);
  }
}
// Issue block 0 of pattern `pat`'s patch ahead of time.  To be called by every thread right after a
// block barrier that released the tile; nothing may touch the tile until that pattern's gather.
template <class P>
__device__ __forceinline__ void patch_prefetch(Cta<P>& c, bool use_tma, const CUtensorMap* tm, int key,
                                               int pat, int npat, const float2* __restrict__ scan,
                                               const Geo& g) {
  if (!(Patch<P>::TMA && use_tma) || pat >= npat) return;
  const Pat p = make_pat(scan, pat, g);
  if (p.skip) return;
  patch_issue<P>(c, tm, g, p, pat / g.S, 0);
  c.pp_key = key;
}

// re-arm the data pipe with the tile that follows (pat, k1) in this CTA's schedule
template <class P>
__device__ __forceinline__ void dp_next(const Cta<P>& c, const float* data, int pat, int k1, int npat) {
  int nk = k1 + 1;
  if (nk == P::RC) {
    nk = 0;
    pat += gridDim.x;
  }
  if (pat < npat) dp_issue<P>(c, data + (size_t)pat * P::N * P::N, nk);
}

// ------------------------------------------------------------------------------------------
// API forward: g = FFT2(pad(kappa * prb * patch))                      (ptychofft.cu:60-73)
// ------------------------------------------------------------------------------------------
template <class P>
__global__ void __launch_bounds__(P::NT, P::MINB) k_fwd(const PassArgs a, const __grid_constant__ CUtensorMap tm_a,
                                               const __grid_constant__ CUtensorMap tm_b) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  Cta<P> c;
  cta_setup<P>(c, smem_raw, a);
  const Geo g = a.g;
  const int npat = g.T * g.S;
  for (int pat = blockIdx.x; pat < npat; pat += gridDim.x) {
    const int t = pat / g.S;
    const Pat p = make_pat(a.scan, pat, g);
    float2* out = a.far + (size_t)pat * P::N * P::N;
    const float2* psi_t = a.psi + (size_t)t * g.nz * g.n;
    const float2* prb_t = a.prb + (size_t)t * a.prb_ts;
    spectrum_pass<P>(
        c, p.skip, [&](int cb, float2(&v)[P::E]) {
          gather_any<P>(v, c, cb, a.use_tma, &tm_a, 2 * pat, t, psi_t, prb_t, g, p);
        },
        [&](int k1, float2(&v)[P::E]) {
#pragma unroll
          for (int e = 0; e < P::E; ++e) out[spec_index<P>(c, k1, e)] = v[e];
        },
        [&](int k1) {
          if (k1 == P::RC - 1)
            patch_prefetch<P>(c, a.use_tma, &tm_a, 2 * (pat + gridDim.x), pat + gridDim.x, npat, a.scan, g);
        });
  }
}

// Parity hook: the zero-padded near-plane frame (kernels.cu:95-107 output, before the FFT), natural
// order.  Integer work of the path (patch origin, window offset, skip rule) is checked bit-exactly
// through it.
template <class P>
__global__ void __launch_bounds__(P::NT, P::MINB) k_nearplane(const PassArgs a,
                                                     const __grid_constant__ CUtensorMap tm_a,
                                                     const __grid_constant__ CUtensorMap tm_b) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  Cta<P> c;
  cta_setup<P>(c, smem_raw, a);
  const Geo g = a.g;
  const int npat = g.T * g.S;
  for (int pat = blockIdx.x; pat < npat; pat += gridDim.x) {
    const int t = pat / g.S;
    const Pat p = make_pat(a.scan, pat, g);
    float2* out = a.far + (size_t)pat * P::N * P::N;
    for (int cb = 0; cb < P::RC; ++cb) {
      float2 v[P::E];
      if (!p.skip) {
        gather_nat<P>(v, c, cb, a.psi + (size_t)t * g.nz * g.n, a.prb + (size_t)t * a.prb_ts, g, p);
      } else {
#pragma unroll
        for (int e = 0; e < P::E; ++e) v[e] = make_float2(0.f, 0.f);
      }
#pragma unroll
      for (int e = 0; e < P::E; ++e) {
        int y, x;
        nat_coord<P>(c, cb, e, y, x);
        out[y * P::N + x] = v[e];
      }
    }
  }
}

// ------------------------------------------------------------------------------------------
// API adjoints: inverse FFT + object scatter (FLG 0) or probe reduction (FLG 1)  (ptychofft.cu:76-88)
// ------------------------------------------------------------------------------------------
template <class P, int FLG>
__global__ void __launch_bounds__(P::NT, P::MINB) k_adj(const PassArgs a, const __grid_constant__ CUtensorMap tm_a,
                                               const __grid_constant__ CUtensorMap tm_b) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  Cta<P> c;
  cta_setup<P>(c, smem_raw, a);
  const Geo g = a.g;
  if (FLG == 1) pacc_zero<P>(c);
  int t_cur = -1;
  const int npat = g.T * g.S;
  for (int pat = blockIdx.x; pat < npat; pat += gridDim.x) {
    const int t = pat / g.S;
    if (FLG == 1 && t != t_cur) {
      if (t_cur >= 0) pacc_flush<P>(c, a.grad + (size_t)t_cur * a.grad_ts, g);
      t_cur = t;
    }
    const Pat p = make_pat(a.scan, pat, g);
    if (p.skip) continue;
    const float2* in = a.far_in + (size_t)pat * P::N * P::N;
    const float2* psi_t = a.psi + (size_t)t * g.nz * g.n;
    const float2* prb_t = a.prb + (size_t)t * a.prb_ts;
    float2* grad_t = a.grad + (size_t)t * g.nz * g.n;
    inverse_pass<P>(
        c,
        [&](int k1, float2(&v)[P::E]) {
#pragma unroll
          for (int e = 0; e < P::E; ++e) v[e] = __ldg(in + spec_index<P>(c, k1, e));
        },
        [&](int cb, float2(&v)[P::E]) {
          if (FLG == 0)
            scatter_block<P>(v, c, cb, prb_t, g.kappa, grad_t, g, p, []() {});
          else
            pacc_add<P>(v, c, cb, psi_t, g.kappa, g, p);
        });
  }
  if (FLG == 1 && t_cur >= 0) pacc_flush<P>(c, a.grad + (size_t)t_cur * a.grad_ts, g);
}

// minf_px plus the two sums of the NEXT iteration's intensity pass for the same intensity x
// (ptycho.py:342-343: a = sum sqrt(I d), b = sum I): the probe line search evaluates exactly the
// intensity the next iteration starts from, so that pass need not run again (M = 1).
template <int MODEL>
__device__ __forceinline__ float minf_ab_px(float x, float d, float sqd, float& sa, float& sb) {
  const float ax = fabsf(x);
  const float sx = fsqrt(ax);
  sa += sx * sqd;
  sb += x;
  if (MODEL == PTX_MODEL_GAUSSIAN) {
    const float r = sx - sqd;
    return r * r;
  } else {
    return ax - d * logf(ax + 1e-32f);
  }
}

// ------------------------------------------------------------------------------------------
// CG pass A: I = sum_k |F_k|^2, reductions a = sum sqrt(I d), b = sum I, cost   (ptycho.py:330-343)
// With several modes the running sum is parked in thread-private scratch between modes.
// ------------------------------------------------------------------------------------------
template <class P, int MODEL, bool MULTI>
__device__ __forceinline__ void intensity_body(Cta<P>& c, const PassArgs& a, const CUtensorMap* tm) {
  const Geo g = a.g;
  const float iscale = a.sc ? a.sc[0] : 1.f;
  const int npat = g.T * g.S;
  if ((int)blockIdx.x < npat) dp_issue<P>(c, a.data + (size_t)blockIdx.x * P::N * P::N, 0);
  for (int pat = blockIdx.x; pat < npat; pat += gridDim.x) {
    const int t = pat / g.S;
    const Pat p = make_pat(a.scan, pat, g);
    const float2* psi_t = a.psi + (size_t)t * g.nz * g.n;
    float* io = a.inten_out ? a.inten_out + (size_t)pat * P::N * P::N : nullptr;
    const int kfirst = (MULTI && !p.skip) ? 0 : a.nmodes - 1;  // a skipped pattern: one zero pass
    for (int k = kfirst; k < a.nmodes; ++k) {
      const float2* prb_k = a.prb + (size_t)t * a.prb_ts + (size_t)k * a.prb_ms;
      const bool first = !MULTI || (k == kfirst), last = !MULTI || (k + 1 == a.nmodes);
      spectrum_pass<P>(
          c, p.skip, [&](int cb, float2(&v)[P::E]) {
            gather_any<P>(v, c, cb, a.use_tma, tm, 2 * pat, t, psi_t, prb_k, g, p);
          },
          [&](int k1, float2(&v)[P::E]) {
            float* ia = c.accp + (size_t)k1 * P::E * P::NT + c.tid;
            if (!last) {
#pragma unroll
              for (int e = 0; e < P::E; ++e) {
                float I = v[e].x * v[e].x + v[e].y * v[e].y;
                if (!first) I += ia[e * P::NT];
                ia[e * P::NT] = I;
              }
              return;
            }
            float sa = 0.f, sb = 0.f, scost = 0.f;  // fp32 over 32 pixels, double across tiles
            dp_wait<P>(c);
#pragma unroll
            for (int e = 0; e < P::E; ++e) {
              float I = v[e].x * v[e].x + v[e].y * v[e].y;
              if (!first) I += ia[e * P::NT];
              const float dd = c.dbuf[data_index<P>(c, e)];
              sa += fsqrt(I * dd);
              sb += I;
              scost += minf_px<MODEL>(I * iscale, dd, fsqrt(dd));
              if (io) io[spec_index<P>(c, k1, e)] = I;
            }
            c.slots[0 * P::NT] += (double)sa;
            c.slots[1 * P::NT] += (double)sb;
            c.slots[2 * P::NT] += (double)scost;
          },
          [&](int k1) {
            if (last) dp_next<P>(c, a.data, pat, k1, npat);
            if (k1 == P::RC - 1) {  // the tile is free: next mode of this pattern, or the next pattern
              const int np = last ? pat + (int)gridDim.x : pat;
              patch_prefetch<P>(c, a.use_tma, tm, 2 * np, np, npat, a.scan, g);
            }
          });
    }
  }
  double acc[3];
#pragma unroll
  for (int k = 0; k < 3; ++k) acc[k] = c.slots[k * P::NT];
  block_reduce_add<3, P::NT / 32>(acc, c.red, a.red, c.tid);
}
template <class P, int MODEL>
__global__ void __launch_bounds__(P::NT, P::MINB) k_intensity(const PassArgs a,
                                                     const __grid_constant__ CUtensorMap tm_a,
                                                     const __grid_constant__ CUtensorMap tm_b) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  Cta<P> c;
  cta_setup<P>(c, smem_raw, a);
  if (a.nmodes == 1)
    intensity_body<P, MODEL, false>(c, a, &tm_a);
  else
    intensity_body<P, MODEL, true>(c, a, &tm_a);
}

// ------------------------------------------------------------------------------------------
// CG pass B/D: fused fwd -> residual -> inverse -> object scatter (WHAT 0) / probe reduction (WHAT 1)
//   ptycho.py:347-363 (object), 421-441 (probe)
// sc = {fscale, iscale, gscale}
// ------------------------------------------------------------------------------------------
template <class P, int MODEL, int WHAT, bool CACHE>
__global__ void __launch_bounds__(P::NT, P::MINB) k_grad(const PassArgs a, const __grid_constant__ CUtensorMap tm_a,
                                                const __grid_constant__ CUtensorMap tm_b) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  Cta<P> c;
  cta_setup<P>(c, smem_raw, a);
  const Geo g = a.g;
  const float fscale = a.sc[0], iscale = a.sc[1], gscale = a.sc[2] * g.kappa;
  if (WHAT == 1) pacc_zero<P>(c);
  int t_cur = -1;
  const int npat = g.T * g.S;
  if ((int)blockIdx.x < npat) dp_issue<P>(c, a.data + (size_t)blockIdx.x * P::N * P::N, 0);
  float2 vv[P::E];    // near plane / spectrum of the pattern in flight (single-tile object pass)
  bool have = false;  // vv already holds the gathered near plane of `pat`
  for (int pat = blockIdx.x; pat < npat; pat += gridDim.x) {
    const int t = pat / g.S;
    if (WHAT == 1 && t != t_cur) {
      if (t_cur >= 0) pacc_flush<P>(c, a.grad + (size_t)t_cur * a.grad_ts, g);
      t_cur = t;
    }
    const Pat p = make_pat(a.scan, pat, g);
    if (p.skip) {  // F = 0 -> residual 0 -> no contribution; drain the data pipe all the same
      for (int k1 = 0; k1 < P::RC; ++k1) {
        dp_wait<P>(c);
        __syncthreads();
        dp_next<P>(c, a.data, pat, k1, npat);
      }
      continue;
    }
    const float2* psi_t = a.psi + (size_t)t * g.nz * g.n;
    const float2* prb_t = a.prb + (size_t)t * a.prb_ts;
    float2* grad_t = a.grad + (size_t)t * g.nz * g.n;
    const float* ii = a.inten_in ? a.inten_in + (size_t)pat * P::N * P::N : nullptr;
    float2* fc = CACHE ? a.far + (size_t)pat * P::N * P::N : nullptr;
    auto residual = [&](int k1, float2(&v)[P::E]) {
      if (CACHE) {  // keep F(psi, probe) for the line search that follows (ptycho.py:385, 457): 8 N^2 B
#pragma unroll
        for (int e = 0; e < P::E; ++e) __stcs(fc + spec_index<P>(c, k1, e), v[e]);  // streaming: read back a pass later
      }
      dp_wait<P>(c);
#pragma unroll
      for (int e = 0; e < P::E; ++e) {
        const float dd = c.dbuf[data_index<P>(c, e)];
        const float I = ii ? __ldg(ii + spec_index<P>(c, k1, e)) * iscale
                           : (v[e].x * v[e].x + v[e].y * v[e].y);
        const float f = residual_factor<MODEL>(dd, I, fscale);
        v[e].x *= f;
        v[e].y *= f;
      }
    };
    if constexpr (WHAT == 0 && P::N == 64) {
      // object pass of the 64^2 plan: the next pattern's gather rides in this pattern's scatter
      // loop (scatter_gather_impl) whenever both are plain full-window interior patterns.  Measured
      // (profiles/r01j_scatter_gather.txt): +6 % at 64^2; at 128^2 the same fusion LOSES 6 % -- there
      // both phases are bound by load/store-unit wavefronts, not by L2 latency, so nothing overlaps
      // and the extra live registers cost spills -- hence 64^2 only.
      if (!have) gather_nat<P>(vv, c, 0, psi_t, prb_t, g, p);
      fft_forward<P>(vv, c.tile, c.tw, c.tid);
      residual(0, vv);
      fft_inverse<P>(vv, c.tile, c.tw, c.tid);
      dp_next<P>(c, a.data, pat, 0, npat);
      const int np = pat + (int)gridDim.x;
      Pat pn;
      pn.skip = true;
      pn.inside = false;
      if (np < npat) pn = make_pat(a.scan, np, g);
      have = g.P == P::N && p.inside && np < npat && !pn.skip && pn.inside && (np / g.S) * (size_t)a.prb_ts == (size_t)t * a.prb_ts;
      if (have)
        scatter_gather_impl<P>(vv, c, prb_t, gscale, grad_t, a.psi + (size_t)(np / g.S) * g.nz * g.n, g, p, pn,
                               []() {});
      else
        scatter_block<P>(vv, c, 0, prb_t, gscale, grad_t, g, p, []() {});
      continue;
    }
    fused_pass<P>(
        c, [&](int cb, float2(&v)[P::E]) {
          gather_any<P>(v, c, cb, a.use_tma, &tm_a, 2 * pat, t, psi_t, prb_t, g, p);
        },
        residual,
        [&](int k1) {
          dp_next<P>(c, a.data, pat, k1, npat);
          if (WHAT == 1 && P::RC == 1 && Patch<P>::TMA && a.use_tma) {  // probe pass: the tile is idle after the inverse
            __syncthreads();
            patch_prefetch<P>(c, a.use_tma, &tm_a, 2 * (pat + gridDim.x), pat + gridDim.x, npat, a.scan, g);
          }
        },
        [&](int cb, float2(&v)[P::E]) {
          if (WHAT == 0)
            scatter_block<P>(v, c, cb, prb_t, gscale, grad_t, g, p, [&]() {
              if (cb == P::RC - 1)
                patch_prefetch<P>(c, a.use_tma, &tm_a, 2 * (pat + gridDim.x), pat + gridDim.x, npat,
                                  a.scan, g);
            });
          else
            pacc_add<P>(v, c, cb, psi_t, gscale, g, p);
        });
  }
  if (WHAT == 1 && t_cur >= 0) pacc_flush<P>(c, a.grad + (size_t)t_cur * a.grad_ts, g);
}

// ------------------------------------------------------------------------------------------
// CG pass C/E: line-search costs for 4 step candidates 2^-c0 .. 2^-(c0+3) at once
//   ptycho.py:383-393 (object), 451-461 (probe), 253-281 (line_search_sqr)
// The first far field of a pair is parked in thread-private scratch while the second is transformed.
// ------------------------------------------------------------------------------------------
template <class P, int MODEL, bool AB, bool CACHED>
__global__ void __launch_bounds__(P::NT, P::MINB) k_linesearch(const PassArgs a,
                                                      const __grid_constant__ CUtensorMap tm_a,
                                                      const __grid_constant__ CUtensorMap tm_b) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  Cta<P> c;
  cta_setup<P>(c, smem_raw, a);
  const Geo g = a.g;
  constexpr size_t NN = (size_t)P::N * P::N;
  const bool multi = a.npairs > 1;
  const int npat = g.T * g.S;
  const float gam0 = exp2f(-(float)a.c0);
  if ((int)blockIdx.x < npat) dp_issue<P>(c, a.data + (size_t)blockIdx.x * NN, 0);
  for (int pat = blockIdx.x; pat < npat; pat += gridDim.x) {
    const int t = pat / g.S;
    const Pat p = make_pat(a.scan, pat, g);
    const float* p1in = a.inten_in ? a.inten_in + (size_t)pat * NN : nullptr;
    float2* p23o = a.p23 ? a.p23 + (size_t)pat * NN : nullptr;  // (p2, p3): intensity at any step later
    const float2* psi_a = a.psi + (size_t)t * g.nz * g.n;
    const float2* psi_b = a.psi_b + (size_t)t * g.nz * g.n;
    for (int j = 0; j < a.npairs; ++j) {
      const float2* prb_a = a.prb + (size_t)t * a.prb_ts + (size_t)j * a.prb_ms;
      const float2* prb_b = a.prb_b + (size_t)t * a.prb_b_ts + (size_t)j * a.prb_b_ms;
      const bool first = (j == 0), last = (j + 1 == a.npairs);
      // the pair's first far field was left in HBM by the gradient pass that preceded this search
      // (reading 8 N^2 bytes costs a quarter of recomputing gather + transform)
      const float2* t1c = CACHED ? a.far_in + (size_t)j * a.far_ms + (size_t)pat * NN : nullptr;
      if constexpr (!CACHED)
        spectrum_pass<P>(
            c, p.skip, [&](int cb, float2(&v)[P::E]) {
              gather_any<P>(v, c, cb, a.use_tma, &tm_a, 2 * pat, t, psi_a, prb_a, g, p);
            },
            [&](int k1, float2(&v)[P::E]) {
              float2* st = c.stash + (size_t)k1 * P::E * P::NT + c.tid;
#pragma unroll
              for (int e = 0; e < P::E; ++e) st[e * P::NT] = v[e];
            },
            [&](int k1) {  // the second object's patch of the same pattern comes next
              if (k1 == P::RC - 1) patch_prefetch<P>(c, a.use_tma, &tm_b, 2 * pat + 1, pat, npat, a.scan, g);
            });
      spectrum_pass<P>(
          c, p.skip, [&](int cb, float2(&v)[P::E]) {
            if (CACHED && cb == 0 && !p.skip && (c.tid & 15) == 0) {
              // pull the cached far field from HBM into L2 meanwhile: a warp's register e covers two
              // 128-byte lines (32 lanes x 8 B), one prefetch per line
              for (int k1 = 0; k1 < P::RC; ++k1)
#pragma unroll
                for (int e = 0; e < P::E; ++e)
                  asm volatile("prefetch.global.L2 [%0];" ::"l"(t1c + spec_index<P>(c, k1, e)));
            }
            gather_any<P>(v, c, cb, a.use_tma, &tm_b, 2 * pat + 1, t, psi_b, prb_b, g, p);
          },
          [&](int k1, float2(&v)[P::E]) {
            const float2* st = c.stash + (size_t)k1 * P::E * P::NT + c.tid;
            float* ap = c.accp + (size_t)k1 * P::E * P::NT + c.tid;
            float cost[5];  // fp32 over 32 pixels, double across tiles
            float sab[AB ? 10 : 1];  // a and b of every candidate (AB)
#pragma unroll
            for (int q = 0; q < 5; ++q) cost[q] = 0.f;
#pragma unroll
            for (int q = 0; q < (AB ? 10 : 1); ++q) sab[q] = 0.f;
            const float2* t1p = (CACHED && !p.skip) ? t1c + c.sbase + k1 * P::N : nullptr;
            if (last) dp_wait<P>(c);
            // the first far field comes from L2 / HBM: its loads are issued CH at a time ahead of
            // the arithmetic that consumes them (one exposed round trip per CH pixels, not per pixel)
            constexpr int CH = 8;
#pragma unroll
            for (int e0 = 0; e0 < P::E; e0 += CH) {
              float2 t1v[CH];
#pragma unroll
              for (int j = 0; j < CH; ++j) {
                const int e = e0 + j;
                if (CACHED) {
                  int dx, dy;
                  elem_offset<typename P::S2>(e, dx, dy);
                  // volatile: ptxas otherwise batches all 32 loads ahead of the arithmetic (64 live
                  // registers) and spills the spectrum -- 576 B of stack per thread in round 1, the
                  // "unexplained" DRAM writes of profiles/r01m_all128_ncu.txt were that local traffic
                  t1v[j] = make_float2(0.f, 0.f);
                  if (t1p)
                    asm volatile("ld.volatile.global.v2.f32 {%0, %1}, [%2];"
                                 : "=f"(t1v[j].x), "=f"(t1v[j].y)
                                 : "l"(t1p + (P::RC * pos_to_freq_y<P>(dy) * P::N + pos_to_freq_x<P>(dx)))
                                 : "memory");
                } else {
                  t1v[j] = st[e * P::NT];
                }
              }
#pragma unroll
              for (int j = 0; j < CH; ++j) {
                const int e = e0 + j;
                const float2 t1 = t1v[j];
                const float2 t2 = v[e];
                float q1 = t1.x * t1.x + t1.y * t1.y;
                float q2 = t2.x * t2.x + t2.y * t2.y;
                float q3 = 2.f * (t1.x * t2.x + t1.y * t2.y);
                if (multi) {
                  if (!first) {
                    q1 += ap[e * P::NT];
                    q2 += ap[NN + e * P::NT];
                    q3 += ap[2 * NN + e * P::NT];
                  }
                  if (!last) {
                    ap[e * P::NT] = q1;
                    ap[NN + e * P::NT] = q2;
                    ap[2 * NN + e * P::NT] = q3;
                  }
                }
                if (last) {
                  const float dd = c.dbuf[data_index<P>(c, e)];
                  const float sqd = fsqrt(dd);
                  if (p1in) q1 = __ldg(p1in + spec_index<P>(c, k1, e));
                  if (p23o) __stcs(p23o + spec_index<P>(c, k1, e), make_float2(q2, q3));
                  float gam = gam0;
                  if (AB) {
                    cost[0] += minf_ab_px<MODEL>(q1, dd, sqd, sab[0], sab[5]);
#pragma unroll
                    for (int q = 0; q < 4; ++q) {
                      cost[1 + q] += minf_ab_px<MODEL>(q1 + gam * gam * q2 + gam * q3, dd, sqd, sab[1 + q],
                                                       sab[6 + q]);
                      gam *= 0.5f;
                    }
                  } else {
                    cost[0] += minf_px<MODEL>(q1, dd, sqd);
#pragma unroll
                    for (int q = 0; q < 4; ++q) {
                      cost[1 + q] += minf_px<MODEL>(q1 + gam * gam * q2 + gam * q3, dd, sqd);
                      gam *= 0.5f;
                    }
                  }
                }
              }
            }
            if (last) {
#pragma unroll
              for (int q = 0; q < 5; ++q) c.slots[q * P::NT] += (double)cost[q];
              if (AB) {
#pragma unroll
                for (int q = 0; q < 10; ++q) c.slots[(5 + q) * P::NT] += (double)sab[q];
              }
            }
          },
          [&](int k1) {
            if (last) dp_next<P>(c, a.data, pat, k1, npat);
            if (k1 == P::RC - 1) {  // next pair of this pattern, or the next pattern
              const int np = last ? pat + (int)gridDim.x : pat;
              if (CACHED)  // only second objects are gathered
                patch_prefetch<P>(c, a.use_tma, &tm_b, 2 * np + 1, np, npat, a.scan, g);
              else
                patch_prefetch<P>(c, a.use_tma, &tm_a, 2 * np, np, npat, a.scan, g);
            }
          });
    }
  }
  double acc[AB ? 15 : 5];
#pragma unroll
  for (int q = 0; q < (AB ? 15 : 5); ++q) acc[q] = c.slots[q * P::NT];
  block_reduce_add<(AB ? 15 : 5), P::NT / 32>(acc, c.red, a.red, c.tid);
}

}  // namespace ptx
