// Device building blocks of the fused ptychography kernels (sm_100a).
//
// One CTA owns one diffraction pattern at a time (persistent loop over patterns).  The pattern's
// N x N complex tile is staged through shared memory by the register-resident FFT of
// fft_tile.cuh; everything else -- patch gather with bilinear sub-pixel weights, probe multiply,
// zero padding, residual against the measured data, conjugate-probe multiply, bilinear
// scatter-add -- is done on the registers / the shared tile of the same CTA, so the far field
// never visits HBM in the fused passes.
//
// Reference semantics reproduced here (paths relative to /root/reference):
//   scan split / skip rule / indices ... src/cuda/kernels.cu:19-63
//   forward multiply .................... src/cuda/kernels.cu:95-107
//   object adjoint ...................... src/cuda/kernels.cu:69-81
//   probe adjoint ....................... src/cuda/kernels.cu:82-94
#pragma once

#include <cuda_runtime.h>

#include "fft_tile.cuh"

namespace ptx {

struct Geo {
  int T, nz, n, S, P, N, o;  // o = (N - P) / 2, window offset of the probe in the padded frame
  float kappa;               // 1 / N, the reference's "fft constant" (kernels.cu:65)
};

// Per-pattern, block-uniform context.
struct Pat {
  int R, C;                  // integer patch origin (row, col): trunc toward zero like modff
  float w00, w01, w10, w11;  // bilinear weights, kernels.cu:97-100
  bool skip;                 // integer part negative -> pattern skipped (kernels.cu:39)
  bool inside;               // the (P+1) x (P+1) window lies fully inside the object
};

__device__ __forceinline__ Pat make_pat(const float2* __restrict__ scan, int idx, const Geo& g) {
  const float2 sc = __ldg(scan + idx);  // .x = row (vertical), .y = column (horizontal)
  const float rI = truncf(sc.x), cI = truncf(sc.y);
  const float rho = sc.x - rI, gam = sc.y - cI;
  Pat p;
  p.skip = (rI < 0.f) || (cI < 0.f);  // -0.0 is not < 0: (-1,0) is NOT skipped, as in the reference
  p.R = (int)rI;
  p.C = (int)cI;
  p.w00 = (1.f - gam) * (1.f - rho);
  p.w01 = gam * (1.f - rho);
  p.w10 = (1.f - gam) * rho;
  p.w11 = gam * rho;
  p.inside = (p.R + g.P + 1 <= g.nz) && (p.C + g.P + 1 <= g.n);
  return p;
}

// Bilinear object patch value at probe pixel (iy, ix).  Outside the object the field is taken as 0
// (the reference reads out of bounds there).
__device__ __forceinline__ float2 patch_at(const float2* __restrict__ psi_t, const Geo& g,
                                           const Pat& p, int iy, int ix) {
  const int r = p.R + iy, c = p.C + ix;
  const float2* q = psi_t + (size_t)r * g.n + c;
  float2 f00, f01, f10, f11;
  if (p.inside) {
    f00 = __ldg(q);
    f01 = __ldg(q + 1);
    f10 = __ldg(q + g.n);
    f11 = __ldg(q + g.n + 1);
  } else {
    const float2 z = make_float2(0.f, 0.f);
    const bool r0 = r < g.nz, r1 = r + 1 < g.nz, c0 = c < g.n, c1 = c + 1 < g.n;
    f00 = (r0 && c0) ? __ldg(q) : z;
    f01 = (r0 && c1) ? __ldg(q + 1) : z;
    f10 = (r1 && c0) ? __ldg(q + g.n) : z;
    f11 = (r1 && c1) ? __ldg(q + g.n + 1) : z;
  }
  float2 t;
  t.x = f00.x * p.w00 + f01.x * p.w01 + f10.x * p.w10 + f11.x * p.w11;
  t.y = f00.y * p.w00 + f01.y * p.w01 + f10.y * p.w10 + f11.y * p.w11;
  return t;
}

// near[o+iy, o+ix] = kappa * prb[iy,ix] * patch[iy,ix], zero elsewhere; in stage-0 ownership.
template <class P>
__device__ __forceinline__ void gather_s0(float2 (&v)[P::E], const float2* __restrict__ psi_t,
                                          const float2* __restrict__ prb, const Geo& g,
                                          const Pat& p, int xf, int yf) {
  using ST = typename P::S0;
#pragma unroll
  for (int e = 0; e < P::E; ++e) {
    int dx, dy;
    elem_offset<ST>(e, dx, dy);
    const int iy = (yf | dy) - g.o, ix = (xf | dx) - g.o;
    float2 r = make_float2(0.f, 0.f);
    if ((unsigned)iy < (unsigned)g.P && (unsigned)ix < (unsigned)g.P) {
      const float2 t = patch_at(psi_t, g, p, iy, ix);
      const float2 pr = __ldg(prb + iy * g.P + ix);
      r.x = g.kappa * (pr.x * t.x - pr.y * t.y);
      r.y = g.kappa * (pr.x * t.y + pr.y * t.x);
    }
    v[e] = r;
  }
}

// ---------------------------------------------------------------- CTA-wide transforms
// forward: v (stage-0 ownership, natural order) -> v (stage-2 ownership, digit-reversed spectrum)
template <class P>
__device__ __forceinline__ void fft_forward(float2 (&v)[P::E], float2* tile, const float2* tw,
                                            int tid) {
  using TL = TwLayout<P>;
  constexpr int L = P::L;
  int xf, yf;
  fixed_coords<typename P::S0, P::WBITS>(tid, xf, yf);
  stage_compute<typename P::S0, false>(v, xf, yf, tw + TL::X0, tw + TL::Y0);
  stage_store<typename P::S0, L>(v, tile, xf, yf);
  __syncthreads();
  fixed_coords<typename P::S1, P::WBITS>(tid, xf, yf);
  stage_load<typename P::S1, L>(v, tile, xf, yf);
  stage_compute<typename P::S1, false>(v, xf, yf, tw + TL::X1, tw + TL::Y1);
  stage_store<typename P::S1, L>(v, tile, xf, yf);
  __syncthreads();
  fixed_coords<typename P::S2, P::WBITS>(tid, xf, yf);
  stage_load<typename P::S2, L>(v, tile, xf, yf);
  stage_compute<typename P::S2, false>(v, xf, yf, tw + TL::X2, tw + TL::Y2);
}

// inverse (unnormalised): v (stage-2 ownership spectrum) -> v (stage-0 ownership, natural order)
// Every stage writes exactly the tile positions its own thread read last, so no barrier is needed
// between a forward transform and the inverse that follows it.
template <class P>
__device__ __forceinline__ void fft_inverse(float2 (&v)[P::E], float2* tile, const float2* tw,
                                            int tid) {
  using TL = TwLayout<P>;
  constexpr int L = P::L;
  int xf, yf;
  fixed_coords<typename P::S2, P::WBITS>(tid, xf, yf);
  stage_compute<typename P::S2, true>(v, xf, yf, tw + TL::X2, tw + TL::Y2);
  stage_store<typename P::S2, L>(v, tile, xf, yf);
  __syncthreads();
  fixed_coords<typename P::S1, P::WBITS>(tid, xf, yf);
  stage_load<typename P::S1, L>(v, tile, xf, yf);
  stage_compute<typename P::S1, true>(v, xf, yf, tw + TL::X1, tw + TL::Y1);
  stage_store<typename P::S1, L>(v, tile, xf, yf);
  __syncthreads();
  fixed_coords<typename P::S0, P::WBITS>(tid, xf, yf);
  stage_load<typename P::S0, L>(v, tile, xf, yf);
  stage_compute<typename P::S0, true>(v, xf, yf, tw + TL::X0, tw + TL::Y0);
}

// frequency (ky, kx) of spectrum register e of this thread (stage-2 ownership)
template <class P>
__device__ __forceinline__ int spec_index(int e, int xf2, int yf2) {
  int dx, dy;
  elem_offset<typename P::S2>(e, dx, dy);
  return pos_to_freq_y<P>(yf2 | dy) * P::N + pos_to_freq_x<P>(xf2 | dx);
}

// ---------------------------------------------------------------- object adjoint: scatter-add
// v holds the near field in stage-0 ownership.  t = scale * conj(prb) * near is parked in the shared
// tile, then every output pixel of the (P+1) x (P+1) window combines its four bilinear taps and
// issues ONE vector reduction (red.global.add.v2.f32) instead of the reference's 8 scalar atomics
// per input pixel (kernels.cu:73-80).
template <class P>
__device__ __forceinline__ void scatter_obj(float2 (&v)[P::E], float2* tile,
                                            const float2* __restrict__ prb, float scale,
                                            float2* __restrict__ grad_t, const Geo& g, const Pat& p,
                                            int tid) {
  using ST = typename P::S0;
  using G = TileGeom<P::L>;
  int xf, yf;
  fixed_coords<ST, P::WBITS>(tid, xf, yf);
#pragma unroll
  for (int e = 0; e < P::E; ++e) {
    int dx, dy;
    elem_offset<ST>(e, dx, dy);
    const int y = yf | dy, x = xf | dx;
    const int iy = y - g.o, ix = x - g.o;
    if ((unsigned)iy < (unsigned)g.P && (unsigned)ix < (unsigned)g.P) {
      const float2 pr = __ldg(prb + iy * g.P + ix);
      float2 t;  // conj(prb) * near
      t.x = scale * (pr.x * v[e].x + pr.y * v[e].y);
      t.y = scale * (pr.x * v[e].y - pr.y * v[e].x);
      tile[G::idx(y, x)] = t;
    }
  }
  __syncthreads();
  const int lane = tid & 31, warp = tid >> 5;
  constexpr int NW = P::NT / 32;
  const int W = g.P + 1;
  const float2 z = make_float2(0.f, 0.f);
  for (int i = warp; i < W; i += NW) {
    const int r = p.R + i;
    if (r >= g.nz) break;
    for (int j = lane; j < W; j += 32) {
      const int c = p.C + j;
      const bool up = i > 0, lo = i < g.P, lf = j > 0, rt = j < g.P;
      const float2 a = (lo && rt) ? tile[G::idx(g.o + i, g.o + j)] : z;
      const float2 b = (lo && lf) ? tile[G::idx(g.o + i, g.o + j - 1)] : z;
      const float2 cc = (up && rt) ? tile[G::idx(g.o + i - 1, g.o + j)] : z;
      const float2 d = (up && lf) ? tile[G::idx(g.o + i - 1, g.o + j - 1)] : z;
      float2 out;
      out.x = a.x * p.w00 + b.x * p.w01 + cc.x * p.w10 + d.x * p.w11;
      out.y = a.y * p.w00 + b.y * p.w01 + cc.y * p.w10 + d.y * p.w11;
      if (c < g.n) atomicAdd(grad_t + (size_t)r * g.n + c, out);
    }
  }
}

// ---------------------------------------------------------------- block reductions (double)
template <int K, int NW>
__device__ __forceinline__ void block_reduce_add(double (&val)[K], double* red_smem, double* out,
                                                 int tid) {
  const int lane = tid & 31, warp = tid >> 5;
#pragma unroll
  for (int k = 0; k < K; ++k) {
    double x = val[k];
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) x += __shfl_down_sync(0xffffffffu, x, off);
    if (lane == 0) red_smem[warp * K + k] = x;
  }
  __syncthreads();
  if (tid < K) {
    double s = 0.0;
    for (int w = 0; w < NW; ++w) s += red_smem[w * K + tid];
    atomicAdd(out + tid, s);
  }
  __syncthreads();
}

}  // namespace ptx
