// Warp-specialised, software-pipelined object-gradient kernel for the 128^2 detector (sm_100a).
//
// k_grad (ptycho_passes.cuh) runs gather -> FFT -> residual -> IFFT -> scatter back to back in one
// 512-thread CTA: the load/store-unit-bound phases (gather: L2 latency + wavefronts, scatter: L2
// reduction throughput) never overlap the FMA / shared-memory-bound transforms, and ncu shows no
// pipe above 60 % (profiles/r01l_bench_grad_ncu.txt).  Here one CTA of 640 threads splits the roles:
//
//   warps 0-15 ("FFT warps", 512 threads, 32 complex registers each) transform: they pull the object
//       patch of pattern n from a shared buffer B, multiply by the probe, run FFT2 -> residual against
//       the measured data -> IFFT2 in registers, multiply by the conjugate probe and push the result
//       back into B;
//   warps 16-19 ("helpers", 128 threads, one frame column each) meanwhile scatter pattern n-1 out of
//       B into the object gradient (separable bilinear spread, one red.global.add.v2.f32 per object
//       pixel) and gather the patch of pattern n+1 into B (one coalesced load per object pixel, the
//       right tap by shuffle, the row below carried in a register) -- kernels.cu:69-81 and 95-107.
//
// B changes hands by two named barriers (bar.arrive / bar.sync, producer-consumer): the swap
// "result(n-1) out, patch(n) in" is one in-place pass of the FFT warps over their own 32 positions.
// Which side does the probe multiplies was decided by ncu's stall samples (profiles/r02f_pipe_ncu.txt):
// with them in the helper loops the FFT warps sat 61 % of their time at the hand-over barrier.
// Registers: 640 threads x 96 = 61440 of the SM's 65536 (the transform core needs < 96, measured with
// -Xptxas -v).  Shared memory: B (128 x 130 complex = 133 KB) + the exchange tile + twiddles.  To
// make room for B the exchange tile holds ONE float per element: real and imaginary parts change
// ownership one after the other (three 512-thread barriers per exchange instead of one; the same
// 128 B/clk of shared-memory traffic).  With 32-bit accesses a wavefront is 32 lanes, so the tile
// geometry differs from Plan<7>: row pitch 132 = 4 mod 32, skew(x) = (x >> 5) & 3, and stage 0 takes
// y2 (not y0) as its fifth lane bit -- every lane bit then moves the bank index by a distinct power
// of two in all three stages (audited on the CPU by tests/emu_fft.cpp, plan "7P").
#pragma once

#include "ptycho_passes.cuh"

namespace ptx {

struct Pipe {
  using P = Plan7P;
  static constexpr int NFFT = 512, NHELP = 128, NTHREADS = NFFT + NHELP;
  static constexpr int PB = 130;  // pitch of B in float2: guard column + 128 + 1
  static constexpr int TILE_F = P::NY * P::RS;            // floats
  static constexpr int B_WORDS = P::N * PB;               // float2
  static constexpr size_t OFF_B = ((size_t)TILE_F * 4 + 15) / 16 * 16;
  static constexpr size_t OFF_TW = OFF_B + (size_t)B_WORDS * 8;
  static constexpr size_t BYTES = OFF_TW + (size_t)TwLayout<P>::TOTAL * 8 + 16;
  // named barriers (0 is __syncthreads)
  static constexpr int BAR_FFT = 1, BAR_FULL = 2, BAR_SWAPPED = 3;
};

__device__ __forceinline__ void nbar_sync(int id, int count) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(count) : "memory");
}
__device__ __forceinline__ void nbar_arrive(int id, int count) {
  asm volatile("bar.arrive %0, %1;" ::"r"(id), "r"(count) : "memory");
}

// ---------------------------------------------------------------- split-plane exchange
// Ownership SA -> SB through the float tile: real parts, then imaginary parts.
template <class SA, class SB>
__device__ __forceinline__ void pipe_exchange(float2 (&v)[32], float* tile, int tid) {
  using P = Plan7P;
  using G = TileGeom<P>;
  int xa, ya, xb, yb;
  fixed_coords<SA, P::WBITS>(tid, xa, ya);
  fixed_coords<SB, P::WBITS>(tid, xb, yb);
  float* pa = tile + G::idx(ya, xa);
  const float* pb = tile + G::idx(yb, xb);
#pragma unroll
  for (int e = 0; e < 32; ++e) {
    int dx, dy;
    elem_offset<SA>(e, dx, dy);
    pa[G::idx(dy, dx)] = v[e].x;
  }
  nbar_sync(Pipe::BAR_FFT, Pipe::NFFT);
  float nx[32];
#pragma unroll
  for (int e = 0; e < 32; ++e) {
    int dx, dy;
    elem_offset<SB>(e, dx, dy);
    nx[e] = pb[G::idx(dy, dx)];
  }
  nbar_sync(Pipe::BAR_FFT, Pipe::NFFT);
#pragma unroll
  for (int e = 0; e < 32; ++e) {
    int dx, dy;
    elem_offset<SA>(e, dx, dy);
    pa[G::idx(dy, dx)] = v[e].y;
  }
  nbar_sync(Pipe::BAR_FFT, Pipe::NFFT);
#pragma unroll
  for (int e = 0; e < 32; ++e) {
    int dx, dy;
    elem_offset<SB>(e, dx, dy);
    v[e] = make_float2(nx[e], pb[G::idx(dy, dx)]);
  }
}

__device__ __forceinline__ void pipe_fft_forward(float2 (&v)[32], float* tile, const float2* tw, int tid) {
  using P = Plan7P;
  using TL = TwLayout<P>;
  int xf, yf;
  fixed_coords<P::S0, P::WBITS>(tid, xf, yf);
  stage_compute<P::S0, false>(v, xf, yf, tw + TL::X0, tw + TL::Y0);
  pipe_exchange<P::S0, P::S1>(v, tile, tid);
  fixed_coords<P::S1, P::WBITS>(tid, xf, yf);
  stage_compute<P::S1, false>(v, xf, yf, tw + TL::X1, tw + TL::Y1);
  pipe_exchange<P::S1, P::S2>(v, tile, tid);
  fixed_coords<P::S2, P::WBITS>(tid, xf, yf);
  stage_compute<P::S2, false>(v, xf, yf, tw + TL::X2, tw + TL::Y2);
}
__device__ __forceinline__ void pipe_fft_inverse(float2 (&v)[32], float* tile, const float2* tw, int tid) {
  using P = Plan7P;
  using TL = TwLayout<P>;
  int xf, yf;
  fixed_coords<P::S2, P::WBITS>(tid, xf, yf);
  stage_compute<P::S2, true>(v, xf, yf, tw + TL::X2, tw + TL::Y2);
  pipe_exchange<P::S2, P::S1>(v, tile, tid);
  fixed_coords<P::S1, P::WBITS>(tid, xf, yf);
  stage_compute<P::S1, true>(v, xf, yf, tw + TL::X1, tw + TL::Y1);
  pipe_exchange<P::S1, P::S0>(v, tile, tid);
  fixed_coords<P::S0, P::WBITS>(tid, xf, yf);
  stage_compute<P::S0, true>(v, xf, yf, tw + TL::X0, tw + TL::Y0);
}

// ---------------------------------------------------------------- helpers: gather into B
// Thread h owns frame column x = h.  Walks the rows top to bottom: one object pixel per row is
// loaded (a warp reads 32 adjacent columns of one object row), the right tap comes from the next
// lane, the horizontally interpolated value of the row below is carried to the next iteration.
//   B[y][x] = kappa * ((1-rho) hq[iy][ix] + rho hq[iy+1][ix]),   (iy, ix) = (y - o, x - o)
//   hq[r][c] = (1-gam) psi[R+r][C+c] + gam psi[R+r][C+c+1]        (kernels.cu:95-107; zero outside)
// (the probe factor is applied by the FFT warps when they pick the patch up)
__device__ __forceinline__ void pipe_gather_any(float2* __restrict__ B, int h, const float2* __restrict__ psi_t,
                                                const Geo& g, const Pat& p) {
  constexpr int PB = Pipe::PB, N = Plan7P::N, CH = 4;
  const int lane = h & 31;
  const int ix = h - g.o;
  const int oc = p.C + ix;
  const bool colv = (unsigned)ix < (unsigned)g.P;                   // a probe column
  const bool tap0 = ix >= 0 && ix <= g.P && oc < g.n;               // this lane's own object pixel is needed
  const bool tap1 = lane == 31 && ix + 1 <= g.P && ix + 1 >= 0 && oc + 1 < g.n;  // right tap no lane provides
  const float2 z = make_float2(0.f, 0.f);
  const float a0 = 1.f - p.gam, a1 = p.gam, kb0 = g.kappa * (1.f - p.rho), kb1 = g.kappa * p.rho;
  const float2* col = psi_t + (ptrdiff_t)(p.R - g.o) * g.n + oc;    // + y * n: object pixel of frame row y
  float2* bcol = B + 1 + h;
  auto hrow = [&](float2 f0, float2 fx) {
    float2 f1 = make_float2(__shfl_down_sync(0xffffffffu, f0.x, 1), __shfl_down_sync(0xffffffffu, f0.y, 1));
    if (lane == 31) f1 = fx;
    return make_float2(a0 * f0.x + a1 * f1.x, a0 * f0.y + a1 * f1.y);
  };
  // rows of the window [o, o+P] (the last one only as the lower tap) that lie inside the object
  auto rowok = [&](int y) { return y >= g.o && y <= g.o + g.P && p.R + y - g.o < g.nz; };
  float2 hc;
  {
    const bool ok = rowok(0);
    const float2 f0 = (ok && tap0) ? __ldg(col) : z;
    const float2 fx = (ok && tap1) ? __ldg(col + 1) : z;
    hc = hrow(f0, fx);
  }
#pragma unroll 1
  for (int y0 = 0; y0 < N; y0 += CH) {
    float2 f0[CH], fx[CH];
#pragma unroll
    for (int j = 0; j < CH; ++j) {
      const int y = y0 + 1 + j;  // tap row of output row y0 + j
      const bool ok = rowok(y);
      f0[j] = (ok && tap0) ? __ldg(col + (ptrdiff_t)y * g.n) : z;
      fx[j] = (ok && tap1) ? __ldg(col + (ptrdiff_t)y * g.n + 1) : z;
    }
#pragma unroll
    for (int j = 0; j < CH; ++j) {
      const int yo = y0 + j;
      const float2 hn = hrow(f0[j], fx[j]);
      const bool inw = colv && yo >= g.o && yo < g.o + g.P;
      bcol[yo * PB] = inw ? make_float2(kb0 * hc.x + kb1 * hn.x, kb0 * hc.y + kb1 * hn.y) : z;
      hc = hn;
    }
  }
}

// ---------------------------------------------------------------- helpers: scatter out of B
// B holds t = gscale * conj(prb) * IFFT2(residual) (frame order, zero outside the probe window).  The
// bilinear spread is separable:
//   hq[y][x] = (1-gam) t[y][x] + gam t[y][x-1],  out[y][x] = (1-rho) hq[y][x] + rho hq[y-1][x],
// one vector reduction per object pixel (kernels.cu:69-81 issues 8 scalar atomics per probe pixel).
__device__ __forceinline__ void pipe_scatter_any(const float2* __restrict__ B, int h, float2* __restrict__ grad_t,
                                                 const Geo& g, const Pat& p) {
  constexpr int PB = Pipe::PB, N = Plan7P::N, CH = 4;
  const int ix = h - g.o;
  const int oc = p.C + ix;
  const bool outc = ix >= 0 && ix <= g.P && oc < g.n;                // this thread's object column exists
  const float2 z = make_float2(0.f, 0.f);
  const float a0 = 1.f - p.gam, a1 = p.gam, b0 = 1.f - p.rho, b1 = p.rho;
  const float2* bcol = B + 1 + h;
  float2* dst = grad_t + (ptrdiff_t)(p.R - g.o) * g.n + oc;          // + y * n
  float2 hp = z;
#pragma unroll 1
  for (int y0 = 0; y0 < N; y0 += CH) {
    float2 tc[CH], tl[CH];
#pragma unroll
    for (int j = 0; j < CH; ++j) {
      tc[j] = bcol[(y0 + j) * PB];
      tl[j] = bcol[(y0 + j) * PB - 1];  // column 0 of B is a guard column of zeros
    }
#pragma unroll
    for (int j = 0; j < CH; ++j) {
      const int y = y0 + j;
      const float2 hq = make_float2(a0 * tc[j].x + a1 * tl[j].x, a0 * tc[j].y + a1 * tl[j].y);
      if (outc && y >= g.o && y <= g.o + g.P && p.R + y - g.o < g.nz)
        atomicAdd(dst + (ptrdiff_t)y * g.n, make_float2(b0 * hq.x + b1 * hp.x, b0 * hq.y + b1 * hp.y));
      hp = hq;
    }
  }
  if (g.o + g.P == N) {  // full window: the row below and the column right of the frame
    if (outc && p.R + g.P < g.nz) atomicAdd(dst + (ptrdiff_t)N * g.n, make_float2(b1 * hp.x, b1 * hp.y));
    if (h < 32 && p.C + g.P < g.n) {  // object column C + P: only the gam * t[.][P-1] share
      const float2* bl = B + N;         // B[y][1 + (N-1)]
      float2* de = grad_t + (ptrdiff_t)p.R * g.n + p.C + g.P;
      for (int y = h; y <= N; y += 32) {
        const float2 t0 = y < N ? bl[y * PB] : z;
        const float2 tu = y > 0 ? bl[(y - 1) * PB] : z;
        if (p.R + y < g.nz)
          atomicAdd(de + (ptrdiff_t)y * g.n, make_float2(a1 * (b0 * t0.x + b1 * tu.x), a1 * (b0 * t0.y + b1 * tu.y)));
      }
    }
  }
}

// ---------------------------------------------------------------- helpers: fast paths
// Full probe window (P == N) lying inside the object: no predicates, pointers advanced row by row.
__device__ __forceinline__ void pipe_gather_fast(float2* __restrict__ B, int h, const float2* __restrict__ psi_t,
                                                 const Geo& g, const Pat& p) {
  constexpr int PB = Pipe::PB, N = Plan7P::N, CH = 8;
  const bool last = (h & 31) == 31;
  const float2 z = make_float2(0.f, 0.f);
  const float a0 = 1.f - p.gam, a1 = p.gam, kb0 = g.kappa * (1.f - p.rho), kb1 = g.kappa * p.rho;
  const int n = g.n;
  const float2* src = psi_t + (size_t)p.R * n + p.C + h;
  float2* bc = B + 1 + h;
  auto hval = [&](float2 f0, float2 fx) {
    float2 f1 = make_float2(__shfl_down_sync(0xffffffffu, f0.x, 1), __shfl_down_sync(0xffffffffu, f0.y, 1));
    if (last) f1 = fx;
    return make_float2(a0 * f0.x + a1 * f1.x, a0 * f0.y + a1 * f1.y);
  };
  float2 hc;
  {
    const float2 f0 = __ldg(src);
    float2 fx = z;
    if (last) fx = __ldg(src + 1);
    hc = hval(f0, fx);
  }
  src += n;
  // two chunks of loads in flight: the taps of rows y0+CH .. are requested before rows y0 .. are used
  float2 f0[CH], fx[CH];
#pragma unroll
  for (int j = 0; j < CH; ++j) {
    f0[j] = __ldg(src + j * n);
    fx[j] = z;
    if (last) fx[j] = __ldg(src + j * n + 1);
  }
  src += CH * n;
#pragma unroll 1
  for (int y0 = 0; y0 < N; y0 += CH) {
    float2 g0[CH], gx[CH];
    if (y0 + CH < N) {
#pragma unroll
      for (int j = 0; j < CH; ++j) {
        g0[j] = __ldg(src + j * n);
        gx[j] = z;
        if (last) gx[j] = __ldg(src + j * n + 1);
      }
      src += CH * n;
    }
#pragma unroll
    for (int j = 0; j < CH; ++j) {
      const float2 hn = hval(f0[j], fx[j]);
      bc[(y0 + j) * PB] = make_float2(kb0 * hc.x + kb1 * hn.x, kb0 * hc.y + kb1 * hn.y);
      hc = hn;
    }
#pragma unroll
    for (int j = 0; j < CH; ++j) {
      f0[j] = g0[j];
      fx[j] = gx[j];
    }
  }
}

__device__ __forceinline__ void pipe_scatter_fast(const float2* __restrict__ B, int h, float2* __restrict__ grad_t,
                                                  const Geo& g, const Pat& p) {
  constexpr int PB = Pipe::PB, N = Plan7P::N, CH = 8;
  const float2 z = make_float2(0.f, 0.f);
  const float a0 = 1.f - p.gam, a1 = p.gam, b0 = 1.f - p.rho, b1 = p.rho;
  const int n = g.n;
  const float2* bc = B + 1 + h;
  float2* dst = grad_t + (size_t)p.R * n + p.C + h;
  float2 hp = z;
#pragma unroll 1
  for (int y0 = 0; y0 < N; y0 += CH) {
    float2 tc[CH], tl[CH];
#pragma unroll
    for (int j = 0; j < CH; ++j) {
      tc[j] = bc[(y0 + j) * PB];
      tl[j] = bc[(y0 + j) * PB - 1];  // column 0 of B is a guard column of zeros
    }
#pragma unroll
    for (int j = 0; j < CH; ++j) {
      const float2 hq = make_float2(a0 * tc[j].x + a1 * tl[j].x, a0 * tc[j].y + a1 * tl[j].y);
      atomicAdd(dst + j * n, make_float2(b0 * hq.x + b1 * hp.x, b0 * hq.y + b1 * hp.y));
      hp = hq;
    }
    dst += CH * n;
  }
  atomicAdd(dst, make_float2(b1 * hp.x, b1 * hp.y));  // the row below the frame
  if (h < 32) {  // object column C + P: only the gam * t[.][P-1] share
    const float2* bl = B + N;  // B[y][1 + (N-1)]
    float2* de = grad_t + (size_t)p.R * n + p.C + N;
    for (int y = h; y <= N; y += 32) {
      const float2 t0 = y < N ? bl[y * PB] : z;
      const float2 tu = y > 0 ? bl[(y - 1) * PB] : z;
      atomicAdd(de + (size_t)y * n, make_float2(a1 * (b0 * t0.x + b1 * tu.x), a1 * (b0 * t0.y + b1 * tu.y)));
    }
  }
}

__device__ __forceinline__ void pipe_gather(float2* __restrict__ B, int h, const float2* __restrict__ psi_t,
                                            const Geo& g, const Pat& p) {
  if (g.P == Plan7P::N && p.inside)
    pipe_gather_fast(B, h, psi_t, g, p);
  else
    pipe_gather_any(B, h, psi_t, g, p);
}
__device__ __forceinline__ void pipe_scatter(const float2* __restrict__ B, int h, float2* __restrict__ grad_t,
                                             const Geo& g, const Pat& p) {
  if (g.P == Plan7P::N && p.inside)
    pipe_scatter_fast(B, h, grad_t, g, p);
  else
    pipe_scatter_any(B, h, grad_t, g, p);
}

// The hand-over pass of the FFT warps over their own 32 positions of B, in place:
//   out: B <- v   (the previous pattern's gscale * conj(prb) * IFFT2(residual));   in: v <- B (kappa * patch)
// Eight positions at a time through volatile accesses: an unrestricted schedule hoists all 32 loads and
// needs old + new values live (128 registers), which puts the whole array into local memory.
template <bool LOAD>
__device__ __forceinline__ void pipe_handover(float2 (&v)[32], float2* bp) {
  using P = Plan7P;
  const unsigned sb = smem_u32(bp);
#pragma unroll
  for (int e0 = 0; e0 < 32; e0 += 8) {
    float2 t[8];
    if (LOAD) {
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        int dx, dy;
        elem_offset<P::S0>(e0 + j, dx, dy);
        asm volatile("ld.volatile.shared.v2.f32 {%0, %1}, [%2];"
                     : "=f"(t[j].x), "=f"(t[j].y)
                     : "r"(sb + (unsigned)(dy * Pipe::PB + dx) * 8u)
                     : "memory");
      }
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      int dx, dy;
      elem_offset<P::S0>(e0 + j, dx, dy);
      asm volatile("st.volatile.shared.v2.f32 [%0], {%1, %2};" ::"r"(sb + (unsigned)(dy * Pipe::PB + dx) * 8u),
                   "f"(v[e0 + j].x), "f"(v[e0 + j].y)
                   : "memory");
      if (LOAD) v[e0 + j] = t[j];
    }
  }
}
// v <- s * prb * v (CONJ = false, kernels.cu:105-106) or s * conj(prb) * v (CONJ = true, kernels.cu:71-72) on
// the thread's stage-0 positions (the pipelined kernel is only launched for P == N: the probe covers the
// frame).  Both multiplies of a pattern use that pattern's own probe, read from L2 each time, so that
// nothing depends on "same angle as the previous pattern?".  The loads are pinned behind a barrier of
// the FFT warps, eight at a time: ptxas otherwise batches all 32 (64 registers) and spills the spectrum.
template <bool CONJ>
__device__ __forceinline__ void pipe_probe_mul(float2 (&v)[32], const float2* __restrict__ pp, float s) {
  using P = Plan7P;
#pragma unroll
  for (int e0 = 0; e0 < 32; e0 += 8) {
    float2 pr[8];
    __syncwarp();  // volatile loads stay in program order (ptxas batches .nc loads even across barriers)
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      int dx, dy;
      elem_offset<P::S0>(e0 + j, dx, dy);
      asm volatile("ld.volatile.global.v2.f32 {%0, %1}, [%2];"
                   : "=f"(pr[j].x), "=f"(pr[j].y)
                   : "l"(pp + dy * P::N + dx)
                   : "memory");
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const float2 w = v[e0 + j];
      if (CONJ)
        v[e0 + j] = make_float2(s * (pr[j].x * w.x + pr[j].y * w.y), s * (pr[j].x * w.y - pr[j].y * w.x));
      else
        v[e0 + j] = make_float2(pr[j].x * w.x - pr[j].y * w.y, pr[j].x * w.y + pr[j].y * w.x);
    }
  }
}

// ------------------------------------------------------------------------------------------
// CG pass B, object gradient, pipelined (same contract as k_grad<P, MODEL, 0, CACHE>):
//   grad += gscale * adj(F * (1 - sqrt(d)/sqrt(I)), scan, probe)       (ptycho.py:347-363)
// ------------------------------------------------------------------------------------------
template <int MODEL, bool CACHE, bool MULTI>
__global__ void __launch_bounds__(Pipe::NTHREADS) k_grad_pipe(const PassArgs a,
                                                              const __grid_constant__ CUtensorMap tm_a,
                                                              const __grid_constant__ CUtensorMap tm_b) {
  using P = Plan7P;
  extern __shared__ __align__(128) unsigned char smem_raw[];
  float* tile = reinterpret_cast<float*>(smem_raw);
  float2* B = reinterpret_cast<float2*>(smem_raw + Pipe::OFF_B);
  float2* tw = reinterpret_cast<float2*>(smem_raw + Pipe::OFF_TW);
  const int tid = threadIdx.x;
  for (int i = tid; i < TwLayout<P>::TOTAL; i += Pipe::NTHREADS) tw[i] = a.tw[i];
  for (int i = tid; i < P::N; i += Pipe::NTHREADS) B[i * Pipe::PB] = make_float2(0.f, 0.f);  // guard column
  __syncthreads();
  const Geo g = a.g;
  const int npat = g.T * g.S;
  constexpr size_t NN = (size_t)P::N * P::N;

  if (tid < Pipe::NFFT) {
    // ================================================================ FFT warps
    const float fscale = a.sc[0], iscale = a.sc[1], gscale = a.sc[2] * g.kappa;
    int xf0, yf0, xf2, yf2;
    fixed_coords<P::S0, P::WBITS>(tid, xf0, yf0);
    fixed_coords<P::S2, P::WBITS>(tid, xf2, yf2);
    const int lbase = pos_to_freq_y<P>(yf2) * P::N + pos_to_freq_x<P>(xf2);
    float2* bp = B + yf0 * Pipe::PB + 1 + xf0;
    float2 v[32];
#pragma unroll
    for (int e = 0; e < 32; ++e) v[e] = make_float2(0.f, 0.f);
    bool have = false;
    for (int pat = blockIdx.x; pat < npat; pat += gridDim.x) {
      const Pat p = make_pat(a.scan, pat, g);
      if (p.skip) continue;  // F = 0 -> residual 0 -> no contribution
      nbar_sync(Pipe::BAR_FULL, Pipe::NTHREADS);
      const float2* prb_in = a.prb + (size_t)(pat / g.S) * a.prb_ts;
      pipe_handover<true>(v, bp);  // (before the first pattern v is zero)
      __threadfence_block();
      nbar_arrive(Pipe::BAR_SWAPPED, Pipe::NTHREADS);
      pipe_probe_mul<false>(v, prb_in + yf0 * P::N + xf0, 1.f);
      pipe_fft_forward(v, tile, tw, tid);
      {
        const float* dpat = a.data + (size_t)pat * NN + lbase;
        const float* ipat = MULTI ? a.inten_in + (size_t)pat * NN + lbase : nullptr;  // sum_k |F_k|^2
        float2* fc = CACHE ? a.far + (size_t)pat * NN + lbase : nullptr;
        constexpr int CH = 8;
#pragma unroll
        for (int e0 = 0; e0 < 32; e0 += CH) {
          float dd[CH], iv[MULTI ? CH : 1];
#pragma unroll
          for (int j = 0; j < CH; ++j) {
            int dx, dy;
            elem_offset<P::S2>(e0 + j, dx, dy);
            const int off = pos_to_freq_y<P>(dy) * P::N + pos_to_freq_x<P>(dx);
            dd[j] = __ldcs(dpat + off);
            if (MULTI) iv[j] = __ldg(ipat + off);
            if (CACHE) __stcs(fc + off, v[e0 + j]);  // F(psi, probe) for the line search that follows
          }
#pragma unroll
          for (int j = 0; j < CH; ++j) {
            const int e = e0 + j;
            const float I = MULTI ? iv[MULTI ? j : 0] * iscale : (v[e].x * v[e].x + v[e].y * v[e].y);
            const float f = residual_factor<MODEL>(dd[j], I, fscale);
            v[e].x *= f;
            v[e].y *= f;
          }
        }
      }
      pipe_fft_inverse(v, tile, tw, tid);
      pipe_probe_mul<true>(v, prb_in + yf0 * P::N + xf0, gscale);
      have = true;
    }
    if (have) {  // hand the last result over
      nbar_sync(Pipe::BAR_FULL, Pipe::NTHREADS);
      pipe_handover<false>(v, bp);
      __threadfence_block();
      nbar_arrive(Pipe::BAR_SWAPPED, Pipe::NTHREADS);
    }
  } else {
    // ================================================================ helper warps
    const int h = tid - Pipe::NFFT;
    auto next_pattern = [&](int pat) {  // first non-skipped pattern of this CTA at or after `pat`
      for (; pat < npat; pat += gridDim.x) {
        const float2 sc = __ldg(a.scan + pat);
        const float rI = truncf(sc.x), cI = truncf(sc.y);
        if (!((rI < 0.f) || (cI < 0.f) || !(rI < (float)g.nz) || !(cI < (float)g.n))) return pat;
      }
      return -1;
    };
    auto l2_prefetch = [&](int pat) {  // measured data (and intensity map) of `pat` towards L2
      if (h == 0) {
        asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(a.data + (size_t)pat * NN),
                     "r"((unsigned)(NN * 4))
                     : "memory");
        if (a.inten_in)
          asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(a.inten_in + (size_t)pat * NN),
                       "r"((unsigned)(NN * 4))
                       : "memory");
      }
    };
    int prev = -1, cur = next_pattern(blockIdx.x);
    if (cur >= 0) {
      const int t = cur / g.S;
      l2_prefetch(cur);
      pipe_gather(B, h, a.psi + (size_t)t * g.nz * g.n, g, make_pat(a.scan, cur, g));
      __threadfence_block();
      nbar_arrive(Pipe::BAR_FULL, Pipe::NTHREADS);
    }
    while (cur >= 0) {
      const int nxt = next_pattern(cur + gridDim.x);
      if (nxt >= 0) l2_prefetch(nxt);
      nbar_sync(Pipe::BAR_SWAPPED, Pipe::NTHREADS);  // B: result of `prev`; the FFT warps hold near(cur)
      if (prev >= 0) {
        const int t = prev / g.S;
        pipe_scatter(B, h, a.grad + (size_t)t * g.nz * g.n, g, make_pat(a.scan, prev, g));
      }
      if (nxt >= 0) {
        const int t = nxt / g.S;
        nbar_sync(4, Pipe::NHELP);  // every helper is done reading B (scatter reads neighbouring columns)
        pipe_gather(B, h, a.psi + (size_t)t * g.nz * g.n, g, make_pat(a.scan, nxt, g));
      }
      __threadfence_block();
      nbar_arrive(Pipe::BAR_FULL, Pipe::NTHREADS);
      prev = cur;
      cur = nxt;
    }
    if (prev >= 0) {
      nbar_sync(Pipe::BAR_SWAPPED, Pipe::NTHREADS);
      const int t = prev / g.S;
      pipe_scatter(B, h, a.grad + (size_t)t * g.nz * g.n, g, make_pat(a.scan, prev, g));
    }
  }
}

}  // namespace ptx
