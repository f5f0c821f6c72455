// Device building blocks of the fused ptychography kernels (sm_100a).
//
// One CTA owns one diffraction pattern at a time (persistent loop over patterns).  The pattern's
// frame is staged through shared memory by the register-resident FFT of fft_tile.cuh; everything
// else -- patch gather with bilinear sub-pixel weights, probe multiply, zero padding, residual
// against the measured data, conjugate-probe multiply, bilinear scatter-add -- is done on the
// registers / the shared tile of the same CTA, so the far field never visits HBM in the fused
// passes.  Detectors larger than 128^2 do not fit one SM: their frame is cut into RC = N^2/16384
// local tiles by a radix-RC "cross" butterfly and staged through a per-CTA scratch frame that
// stays L2-resident (4 x 8 N^2 bytes of L2 traffic per pattern, no HBM round trip).
//
// "Natural ownership" of column block cb (cb < RC): which near-plane pixels (y, x) the 32 registers
// of a thread hold before the forward / after the inverse transform -- stage-0 ownership of the
// whole tile for RC = 1, the cross-stage ownership of fft_tile.cuh (N rows x N/RC columns) else.
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
  float rho, gam;            // fractional parts (row, col)
  float w00, w01, w10, w11;  // bilinear weights, kernels.cu:97-100
  bool skip;                 // integer part negative -> pattern skipped (kernels.cu:39)
  bool inside;               // the (P+1) x (P+1) window lies fully inside the object
};

__device__ __forceinline__ Pat make_pat(const float2* __restrict__ scan, int idx, const Geo& g) {
  const float2 sc = __ldg(scan + idx);  // .x = row (vertical), .y = column (horizontal)
  const float rI = truncf(sc.x), cI = truncf(sc.y);
  Pat p;
  p.rho = sc.x - rI;
  p.gam = sc.y - cI;
  p.skip = (rI < 0.f) || (cI < 0.f);  // -0.0 is not < 0: (-1,0) is NOT skipped, as in the reference
  // a window that starts at or beyond the object's last row / column (or a NaN / inf position, for
  // which the int conversion below would saturate and the bounds arithmetic overflow) only sees the
  // zero extension of Q11: far field 0, no adjoint contribution -- the same thing as a skip
  p.skip = p.skip || !(rI < (float)g.nz) || !(cI < (float)g.n);
  p.R = p.skip ? 0 : (int)rI;
  p.C = p.skip ? 0 : (int)cI;
  p.w00 = (1.f - p.gam) * (1.f - p.rho);
  p.w01 = p.gam * (1.f - p.rho);
  p.w10 = (1.f - p.gam) * p.rho;
  p.w11 = p.gam * p.rho;
  p.inside = (p.R + g.P + 1 <= g.nz) && (p.C + g.P + 1 <= g.n);
  return p;
}

// A zero the compiler cannot see through.  ptxas hoists the (pure) bounds / window predicate arithmetic of a
// rarely taken generic branch above the branch that selects it, where it costs the common path
// instructions and registers (k_linesearch at 128^2: 18 k warp instructions per pattern and 400-700 B of
// spills, profiles/r02v_ls128_regions.txt).  Adding this to a quantity every predicate of the slow
// branch depends on (the window offset, the patch origin) ties that arithmetic to the branch.
__device__ __forceinline__ int opaque_zero() {
  int z;
  asm volatile("mov.u32 %0, 0;" : "=r"(z));
  return z;
}
__device__ __forceinline__ Geo tied(const Geo& g, int z) {
  Geo r = g;
  r.o += z;
  return r;
}
__device__ __forceinline__ Pat tied(const Pat& p, int z) {
  Pat r = p;
  r.R += z;
  r.C += z;
  return r;
}

// Bilinear object patch value at probe pixel (iy, ix).  Outside the object the field is taken as 0
// (the reference reads out of bounds there).
template <bool INSIDE>
__device__ __forceinline__ float2 patch_at(const float2* __restrict__ psi_t, const Geo& g,
                                           const Pat& p, int iy, int ix) {
  const int r = p.R + iy, c = p.C + ix;
  const float2* q = psi_t + (size_t)r * g.n + c;
  float2 f00, f01, f10, f11;
  if (INSIDE) {
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

// ---------------------------------------------------------------- per-CTA context
template <class P>
struct Cta {
  float2* tile;        // shared: NY x (NX+4) complex
  const float2* tw;    // shared twiddle tables
  double* red;         // shared cross-warp reduction scratch
  float* dbuf;         // shared: the measured-data tile of the current (sub-)spectrum, NY x NX f32
  unsigned long long* bar;  // shared mbarriers: [0] bulk copy into dbuf, [1] object-patch tensor copy
  unsigned phase;      // parity of the next dbuf completion to wait for
  unsigned phase2;     // ... of the next patch completion
  int pshift;          // 0/1: column offset of the patch inside its (even-aligned) TMA box
  int pp_key;          // key (2*pattern + object) of the block-0 patch in flight / landed, or -1
  double* slots;       // global thread-private running sums: slots[k * NT], k < 9
  float2* frame;       // global scratch frame [RC][NY][NX] (RC > 1 only)
  float2* stash;       // global thread-private scratch, N*N complex: [(k*E + e)*NT + tid]
  float* accp;         // global thread-private scratch, 3*N*N floats
  int tid, xf0, yf0, xf2, yf2;
  bool strip;          // the tile is free whenever a gather starts: the column-strip gather may use it
  int sbase;           // spec_index of the thread's stage-2 coordinates (element part is immediate)
  int lbase;           // same within one sub-tile's data buffer: fy_local * N + fx
};

// ---------------------------------------------------------------- measured-data pipe (TMA bulk copy)
// The N^2 (or, N > 128, the NY rows ky = k1 mod RC of the) measured intensities a pattern's pointwise
// step needs are pulled into shared memory by cp.async.bulk (UBLKCP) one (sub-)tile ahead: issued
// right after the previous tile has been consumed, waited for just before use, so the HBM latency
// hides behind a whole inverse + forward transform.  One buffer, one mbarrier, at most one copy in
// flight.  dp_issue must be called by all threads, after a block barrier that follows the last read
// of the buffer.
__device__ __forceinline__ unsigned smem_u32(const void* p) {
  return (unsigned)__cvta_generic_to_shared(p);
}
template <class P>
__device__ __forceinline__ void dp_init(Cta<P>& c) {
  if (c.tid == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(c.bar)));
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(c.bar + 1)));
    asm volatile("fence.mbarrier_init.release.cluster;");
  }
  c.phase = 0;
  c.phase2 = 0;
  c.pp_key = -1;
}
__device__ __forceinline__ void mbar_wait(unsigned bar, unsigned phase) {
  unsigned done = 0;
  while (!done) {
    asm volatile(
        "{\n.reg .pred p;\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
        "selp.u32 %0, 1, 0, p;\n}"
        : "=r"(done)
        : "r"(bar), "r"(phase)
        : "memory");
  }
}
template <class P>
__device__ __forceinline__ void dp_issue(const Cta<P>& c, const float* d_pat, int k1) {
  if (c.tid >= 32) return;
  const unsigned bar = smem_u32(c.bar);
  constexpr unsigned TOTAL = P::NX * P::NY * 4;
  if (c.tid == 0) {
    asm volatile("fence.proxy.async.shared::cta;");
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(TOTAL));
  }
  __syncwarp();
  // the measured data are read exactly once per pass: evict-first in L2, so that the stream does
  // not push out what IS reused there (object, probe, and the scratch frames of the N > 128 plans)
  unsigned long long pol;
  asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
  if (P::RC == 1) {
    if (c.tid == 0)
      asm volatile(
          "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;" ::
              "r"(smem_u32(c.dbuf)), "l"(d_pat), "r"(TOTAL), "r"(bar), "l"(pol)
          : "memory");
  } else {  // rows ky = k1 + RC * j of the N x N data frame, one bulk copy per row
    for (int j = c.tid; j < P::NY; j += 32)
      asm volatile(
          "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;" ::
              "r"(smem_u32(c.dbuf + j * P::N)), "l"(d_pat + (size_t)(k1 + P::RC * j) * P::N),
              "r"((unsigned)(P::N * 4)), "r"(bar), "l"(pol)
          : "memory");
  }
}
template <class P>
__device__ __forceinline__ void dp_wait(Cta<P>& c) {
  mbar_wait(smem_u32(c.bar), c.phase);
  c.phase ^= 1;
}
// index into dbuf of spectrum register e (stage-2 ownership)
template <class P>
__device__ __forceinline__ int data_index(const Cta<P>& c, int e) {
  int dx, dy;
  elem_offset<typename P::S2>(e, dx, dy);
  return c.lbase + (pos_to_freq_y<P>(dy) * P::N + pos_to_freq_x<P>(dx));
}

// frame coordinates (y, x) of natural-ownership register e of column block cb
template <class P>
__device__ __forceinline__ void nat_coord(const Cta<P>& c, int cb, int e, int& y, int& x) {
  if (P::RC == 1) {
    int dx, dy;
    elem_offset<typename P::S0>(e, dx, dy);
    y = c.yf0 | dy;
    x = c.xf0 | dx;
  } else {
    int ylow, xc;
    Cross<P>::pair(c.tid, e / P::RC, ylow, xc);
    y = (e % P::RC) * P::NY + ylow;
    x = cb * Cross<P>::CW + xc;
  }
}

// natural frequency index ky*N + kx of spectrum register e (stage-2 ownership) of sub-tile k1
// (pos_to_freq is a bit permutation, hence additive over the disjoint thread / element bit fields)
template <class P>
__device__ __forceinline__ int spec_base(int xf2, int yf2) {
  return P::RC * pos_to_freq_y<P>(yf2) * P::N + pos_to_freq_x<P>(xf2);
}
template <class P>
__device__ __forceinline__ int spec_index(const Cta<P>& c, int k1, int e) {
  int dx, dy;
  elem_offset<typename P::S2>(e, dx, dy);
  return c.sbase + k1 * P::N + (P::RC * pos_to_freq_y<P>(dy) * P::N + pos_to_freq_x<P>(dx));
}

// near[o+iy, o+ix] = kappa * prb[iy,ix] * patch[iy,ix], zero elsewhere; natural ownership of block cb.
// FULL: probe window == frame (P == N, o == 0), no window test; INSIDE: no object-bounds tests.
template <class P, bool FULL, bool INSIDE>
__device__ __forceinline__ void gather_impl(float2 (&v)[P::E], const Cta<P>& c, int cb,
                                            const float2* __restrict__ psi_t,
                                            const float2* __restrict__ prb, const Geo& g,
                                            const Pat& p) {
#pragma unroll
  for (int e = 0; e < P::E; ++e) {
    int y, x;
    nat_coord<P>(c, cb, e, y, x);
    const int iy = FULL ? y : y - g.o, ix = FULL ? x : x - g.o;
    float2 r = make_float2(0.f, 0.f);
    if (FULL || ((unsigned)iy < (unsigned)g.P && (unsigned)ix < (unsigned)g.P)) {
      const float2 pr = __ldg(prb + iy * g.P + ix);
      const float2 t = patch_at<INSIDE>(psi_t, g, p, iy, ix);
      r.x = g.kappa * (pr.x * t.x - pr.y * t.y);
      r.y = g.kappa * (pr.x * t.y + pr.y * t.x);
    }
    v[e] = r;
  }
}
// Single-tile plans, full probe window inside the object: the near plane is produced by COLUMN
// STRIPS and handed to stage-0 ownership through the (free) shared tile.  Thread t owns frame column
// x = t % N over a run of N*N/NT rows: one coalesced load per object pixel (a warp reads 32 adjacent
// columns of one object row), the right tap comes from the next lane, the horizontally interpolated
// value of the row below is carried in a register -- 1.03 global loads per pixel instead of four
// unaligned 8-byte taps (3.5 wavefronts each; ncu, profiles/r02k_grad128_ncu.txt: the tap gather was
// 20 % of the fused kernel, bound by L2 latency and LSU queueing).  Costs one extra pass through the
// tile (store by strips, barrier, load by stage-0 ownership): 2 k wavefronts against 5.5 k saved.
template <class P>
__device__ __forceinline__ void gather_strip(float2 (&v)[P::E], const Cta<P>& c,
                                             const float2* __restrict__ psi_t,
                                             const float2* __restrict__ prb, const Geo& g,
                                             const Pat& p) {
  static_assert(P::RC == 1, "single-tile plans only");
  constexpr int N = P::N, RUN = N * N / P::NT, CH = 8;
  static_assert(RUN % CH == 0 && N % 32 == 0, "strip geometry");
  const int x = c.tid % N, r0 = (c.tid / N) * RUN;
  const bool last = (c.tid & 31) == 31;
  const float2 z = make_float2(0.f, 0.f);
  const float a0 = 1.f - p.gam, a1 = p.gam, kb0 = g.kappa * (1.f - p.rho), kb1 = g.kappa * p.rho;
  const int n = g.n;
  const float2* src = psi_t + (size_t)(p.R + r0) * n + p.C + x;
  const float2* pp = prb + r0 * N + x;
  float2* tp = c.tile + TileGeom<P>::idx(r0, x);  // + j * RS per row
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
  __syncthreads();  // the previous user of the tile (an inverse stage, a probe pass) is done reading it
#pragma unroll 1
  for (int j0 = 0; j0 < RUN; j0 += CH) {
    float2 f0[CH], fx[CH], pr[CH];
#pragma unroll
    for (int j = 0; j < CH; ++j) {
      f0[j] = __ldg(src + j * n);
      fx[j] = z;
      if (last) fx[j] = __ldg(src + j * n + 1);
      pr[j] = __ldg(pp + j * N);
    }
    src += CH * n;
    pp += CH * N;
#pragma unroll
    for (int j = 0; j < CH; ++j) {
      const float2 hn = hval(f0[j], fx[j]);
      const float2 t = make_float2(kb0 * hc.x + kb1 * hn.x, kb0 * hc.y + kb1 * hn.y);
      tp[(j0 + j) * TileGeom<P>::RS] = make_float2(pr[j].x * t.x - pr[j].y * t.y, pr[j].x * t.y + pr[j].y * t.x);
      hc = hn;
    }
  }
  __syncthreads();
  stage_load<typename P::S0, P>(v, c.tile, c.xf0, c.yf0);
  // stage 0 stores to the positions this thread just loaded: no barrier needed before fft_forward
}
template <class P>
__device__ __forceinline__ void gather_nat(float2 (&v)[P::E], const Cta<P>& c, int cb,
                                           const float2* __restrict__ psi_t,
                                           const float2* __restrict__ prb, const Geo& g,
                                           const Pat& p) {
  if (g.P == P::N) {  // block-uniform
    if constexpr (P::RC == 1) {
      if (p.inside && c.strip) {
        gather_strip<P>(v, c, psi_t, prb, g, p);
        return;
      }
    }
    if (p.inside) {
      gather_impl<P, true, true>(v, c, cb, psi_t, prb, g, p);
    } else {
      const int z = opaque_zero();
      gather_impl<P, true, false>(v, c, cb, psi_t, prb, tied(g, z), tied(p, z));
    }
  } else {
    const int z = opaque_zero();
    gather_impl<P, false, false>(v, c, cb, psi_t, prb, tied(g, z), tied(p, z));
  }
}

// ---------------------------------------------------------------- object patch by TMA tensor copy
// The (N+1) x (CW+1) object window under column block cb is pulled into the (free) shared tile by
// cp.async.bulk.tensor (UTMALDG): one instruction instead of 4 x 32 strided global loads per thread,
// every object pixel crosses L2 -> SM once, and pixels outside the object arrive as zeros (the zero
// extension of patch_at<false>).  The box starts at frame pixel (0, cb*CW), i.e. at
// object pixel (R - o, C - o + cb*CW), so frame pixel (y, x) has its taps at [y][x - cb*CW] + {0,1}.
template <class P>
struct Patch {
  static constexpr bool TMA = P::N <= 256;
  // box width: CW + 1 columns are needed; TMA wants the box to START on a 16-byte boundary of the
  // global row (measured: an odd complex64 column coordinate raises "illegal instruction"), so the
  // box starts at the even column at or left of the patch and is one column wider; even width.
  static constexpr int W = Cross<P>::CW + 4;
  static constexpr int H = P::N >= 256 ? 129 : P::N + 1;        // box height (<= 256)
  static constexpr int NLOAD = P::N >= 256 ? P::N / 128 : 1;    // boxes stacked every 128 rows
  static constexpr unsigned BYTES = (unsigned)NLOAD * H * W * 8;
  static constexpr int WORDS = ((NLOAD - 1) * 128 + H) * W;     // float2, lands in the tile region
};
template <class P>
__device__ __forceinline__ void patch_issue(Cta<P>& c, const CUtensorMap* tm, const Geo& g,
                                            const Pat& p, int t, int cb) {
  const int cx0 = p.C - g.o + cb * Cross<P>::CW;
  c.pshift = cx0 & 1;  // column of frame pixel x inside the landed box: (x - cb*CW) + pshift
  if (c.tid != 0) return;
  const unsigned bar = smem_u32(c.bar + 1);
  asm volatile("fence.proxy.async.shared::cta;");
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(Patch<P>::BYTES));
  const int cx = cx0 - (cx0 & 1), cy = p.R - g.o;
#pragma unroll
  for (int j = 0; j < Patch<P>::NLOAD; ++j)
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes "
        "[%0], [%1, {%2, %3, %4}], [%5];" ::"r"(smem_u32(c.tile + j * 128 * Patch<P>::W)),
        "l"(tm), "r"(cx), "r"(cy + 128 * j), "r"(t), "r"(bar)
        : "memory");
}
// near plane of block cb from the landed patch; ends with a block barrier (the tile is about to be
// overwritten by the transform / the next patch)
template <class P, bool FULL, bool ONES = false>
__device__ __forceinline__ void gather_tma_impl(float2 (&v)[P::E], Cta<P>& c, int cb, int shift,
                                                const float2* __restrict__ prb, const Geo& g,
                                                const Pat& p) {
  constexpr int W = Patch<P>::W;
  // probe values first: their L2 latency overlaps the wait for the tensor copy
#pragma unroll
  for (int e = 0; e < P::E; ++e) {
    int y, x;
    nat_coord<P>(c, cb, e, y, x);
    const int iy = FULL ? y : y - g.o, ix = FULL ? x : x - g.o;
    v[e] = make_float2(0.f, 0.f);
    if (FULL || ((unsigned)iy < (unsigned)g.P && (unsigned)ix < (unsigned)g.P))
      v[e] = ONES ? make_float2(1.f, 0.f) : __ldg(prb + iy * g.P + ix);
  }
  mbar_wait(smem_u32(c.bar + 1), c.phase2);
  c.phase2 ^= 1;
  const float k00 = g.kappa * p.w00, k01 = g.kappa * p.w01, k10 = g.kappa * p.w10, k11 = g.kappa * p.w11;
#pragma unroll
  for (int e = 0; e < P::E; ++e) {
    int y, x;
    nat_coord<P>(c, cb, e, y, x);
    const float2* q = c.tile + y * W + (x - cb * Cross<P>::CW) + shift;
    const float2 f00 = q[0], f01 = q[1], f10 = q[W], f11 = q[W + 1];
    float2 t;
    t.x = f00.x * k00 + f01.x * k01 + f10.x * k10 + f11.x * k11;
    t.y = f00.y * k00 + f01.y * k01 + f10.y * k10 + f11.y * k11;
    const float2 pr = v[e];
    v[e] = make_float2(pr.x * t.x - pr.y * t.y, pr.x * t.y + pr.y * t.x);
  }
  __syncthreads();
}
template <class P, bool ONES = false>
__device__ __forceinline__ void gather_tma(float2 (&v)[P::E], Cta<P>& c, int cb, int shift,
                                           const float2* __restrict__ prb, const Geo& g,
                                           const Pat& p) {
  if (g.P == P::N)
    gather_tma_impl<P, true, ONES>(v, c, cb, shift, prb, g, p);
  else
    gather_tma_impl<P, false, ONES>(v, c, cb, shift, prb, tied(g, opaque_zero()), p);
}

// barrier of the S1 <-> S2 exchange: only the XG2 threads that actually trade data (fft_tile.cuh)
template <class P>
__device__ __forceinline__ void xchg2_barrier(int tid) {
  if (P::XG2 == 32) {
    __syncwarp();
  } else if (P::XG2 > 0) {
    asm volatile("bar.sync %0, %1;" ::"r"(1 + tid / (P::XG2 > 0 ? P::XG2 : 1)), "n"(P::XG2 > 0 ? P::XG2 : 32) : "memory");
  } else {
    __syncthreads();
  }
}

// ---------------------------------------------------------------- CTA-wide local transforms
// forward: v (stage-0 ownership, natural order) -> v (stage-2 ownership, digit-reversed spectrum)
template <class P>
__device__ __forceinline__ void fft_forward(float2 (&v)[P::E], float2* tile, const float2* tw,
                                            int tid) {
  using TL = TwLayout<P>;
  int xf, yf;
  fixed_coords<typename P::S0, P::WBITS>(tid, xf, yf);
  stage_compute<typename P::S0, false>(v, xf, yf, tw + TL::X0, tw + TL::Y0);
  stage_store<typename P::S0, P>(v, tile, xf, yf);
  __syncthreads();
  fixed_coords<typename P::S1, P::WBITS>(tid, xf, yf);
  stage_load<typename P::S1, P>(v, tile, xf, yf);
  stage_compute<typename P::S1, false>(v, xf, yf, tw + TL::X1, tw + TL::Y1);
  stage_store<typename P::S1, P>(v, tile, xf, yf);
  xchg2_barrier<P>(tid);
  fixed_coords<typename P::S2, P::WBITS>(tid, xf, yf);
  stage_load<typename P::S2, P>(v, tile, xf, yf);
  stage_compute<typename P::S2, false>(v, xf, yf, tw + TL::X2, tw + TL::Y2);
}

// inverse (unnormalised): v (stage-2 ownership spectrum) -> v (stage-0 ownership, natural order)
// Every stage writes exactly the tile positions its own thread read last, so no barrier is needed
// between a forward transform and the inverse that follows it.
template <class P>
__device__ __forceinline__ void fft_inverse(float2 (&v)[P::E], float2* tile, const float2* tw,
                                            int tid) {
  using TL = TwLayout<P>;
  int xf, yf;
  fixed_coords<typename P::S2, P::WBITS>(tid, xf, yf);
  stage_compute<typename P::S2, true>(v, xf, yf, tw + TL::X2, tw + TL::Y2);
  stage_store<typename P::S2, P>(v, tile, xf, yf);
  xchg2_barrier<P>(tid);
  fixed_coords<typename P::S1, P::WBITS>(tid, xf, yf);
  stage_load<typename P::S1, P>(v, tile, xf, yf);
  stage_compute<typename P::S1, true>(v, xf, yf, tw + TL::X1, tw + TL::Y1);
  stage_store<typename P::S1, P>(v, tile, xf, yf);
  __syncthreads();
  fixed_coords<typename P::S0, P::WBITS>(tid, xf, yf);
  stage_load<typename P::S0, P>(v, tile, xf, yf);
  stage_compute<typename P::S0, true>(v, xf, yf, tw + TL::X0, tw + TL::Y0);
}

// Frame accesses: L2-only (.cg); -DPTX_FRAME_HINT=1 adds an L2 evict_last cache hint (measured: no effect on DRAM bytes
// or time at full c4 size, profiles/r02p_frame_hint.txt, so off) -- the frame is
// the one thing in this kernel that IS re-used from L2 (written and read back twice per pattern) while
// 260 KB of measured data stream past it per pattern (evict_first, dp_issue).
#ifndef PTX_FRAME_HINT
#define PTX_FRAME_HINT 0
#endif
__device__ __forceinline__ unsigned long long frame_policy() {
  unsigned long long pol;
  asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(pol));
  return pol;
}
__device__ __forceinline__ float2 frame_ld(const float2* p, unsigned long long pol) {
#if PTX_FRAME_HINT
  float2 r;
  asm volatile("ld.global.cg.L2::cache_hint.v2.f32 {%0, %1}, [%2], %3;" : "=f"(r.x), "=f"(r.y) : "l"(p), "l"(pol));
  return r;
#else
  (void)pol;
  return __ldcg(p);
#endif
}
__device__ __forceinline__ void frame_st(float2* p, float2 v, unsigned long long pol) {
#if PTX_FRAME_HINT
  asm volatile("st.global.cg.L2::cache_hint.v2.f32 [%0], {%1, %2}, %3;" ::"l"(p), "f"(v.x), "f"(v.y), "l"(pol) : "memory");
#else
  (void)pol;
  __stcg(p, v);
#endif
}

// stage-0 ownership <-> scratch frame sub-tile k1 (L2-only accesses: written and read by
// different threads of the same CTA with a block barrier in between)
template <class P>
__device__ __forceinline__ void frame_load_s0(float2 (&v)[P::E], const Cta<P>& c, int k1) {
  const float2* f = c.frame + (size_t)k1 * P::NX * P::NY;
  const unsigned long long pol = frame_policy();
#pragma unroll
  for (int e = 0; e < P::E; ++e) {
    int dx, dy;
    elem_offset<typename P::S0>(e, dx, dy);
    v[e] = frame_ld(f + (c.yf0 | dy) * P::NX + (c.xf0 | dx), pol);
  }
}
template <class P>
__device__ __forceinline__ void frame_store_s0(const float2 (&v)[P::E], const Cta<P>& c, int k1) {
  float2* f = c.frame + (size_t)k1 * P::NX * P::NY;
  const unsigned long long pol = frame_policy();
#pragma unroll
  for (int e = 0; e < P::E; ++e) {
    int dx, dy;
    elem_offset<typename P::S0>(e, dx, dy);
    frame_st(f + (c.yf0 | dy) * P::NX + (c.xf0 | dx), v[e], pol);
  }
}
// cross (natural) ownership of block cb <-> scratch frame: register j + RC*b <-> sub-tile j
template <class P>
__device__ __forceinline__ void frame_load_cross(float2 (&v)[P::E], const Cta<P>& c, int cb) {
  const unsigned long long pol = frame_policy();
#pragma unroll
  for (int e = 0; e < P::E; ++e) {
    int ylow, xc;
    Cross<P>::pair(c.tid, e / P::RC, ylow, xc);
    v[e] = frame_ld(c.frame + scratch_index<P>(e % P::RC, ylow, cb * Cross<P>::CW + xc), pol);
  }
}
template <class P>
__device__ __forceinline__ void frame_store_cross(const float2 (&v)[P::E], const Cta<P>& c, int cb) {
  const unsigned long long pol = frame_policy();
#pragma unroll
  for (int e = 0; e < P::E; ++e) {
    int ylow, xc;
    Cross<P>::pair(c.tid, e / P::RC, ylow, xc);
    frame_st(c.frame + scratch_index<P>(e % P::RC, ylow, cb * Cross<P>::CW + xc), v[e], pol);
  }
}

// ---------------------------------------------------------------- pattern-level passes
// gather(cb, v): fill v with the near plane of column block cb in natural ownership.
// point(k1, v):  consume / modify the spectrum of sub-tile k1 (stage-2 ownership, spec_index).
// after(k1):     runs once a block barrier has passed since point(k1): the place to re-arm the
//                measured-data pipe (dp_issue) for the next tile.
// near(cb, v):   consume the near plane of column block cb after the inverse transform.
// Every pass ends with the shared tile free for reuse (trailing block barrier).

// forward only: fwd operator, intensities, line-search costs.  `zero`: the pattern is skipped,
// its far field is identically 0 and point() is evaluated on zeros without any transform.
template <class P, class Gather, class Point, class After>
__device__ __forceinline__ void spectrum_pass(Cta<P>& c, bool zero, Gather gather, Point point,
                                              After after) {
  float2 v[P::E];
  if (zero) {
    for (int k1 = 0; k1 < P::RC; ++k1) {
#pragma unroll
      for (int e = 0; e < P::E; ++e) v[e] = make_float2(0.f, 0.f);
      point(k1, v);
      __syncthreads();
      after(k1);
    }
    return;
  }
  if (P::RC == 1) {
    gather(0, v);
    fft_forward<P>(v, c.tile, c.tw, c.tid);
    point(0, v);
    __syncthreads();
    after(0);
  } else {
    for (int cb = 0; cb < P::RC; ++cb) {
      gather(cb, v);
      cross_compute<P, false>(v, c.tid, c.tw + TwLayout<P>::CROSS);
      frame_store_cross<P>(v, c, cb);
    }
    __syncthreads();
    for (int k1 = 0; k1 < P::RC; ++k1) {
      frame_load_s0<P>(v, c, k1);
      fft_forward<P>(v, c.tile, c.tw, c.tid);
      point(k1, v);
      __syncthreads();
      after(k1);
    }
  }
}

// forward -> pointwise -> inverse with the spectrum kept in registers (the fused gradient)
template <class P, class Gather, class Point, class After, class Near>
__device__ __forceinline__ void fused_pass(Cta<P>& c, Gather gather, Point point, After after,
                                           Near near) {
  float2 v[P::E];
  if (P::RC == 1) {
    gather(0, v);
    fft_forward<P>(v, c.tile, c.tw, c.tid);
    point(0, v);
    fft_inverse<P>(v, c.tile, c.tw, c.tid);
    after(0);
    near(0, v);
  } else {
    for (int cb = 0; cb < P::RC; ++cb) {
      gather(cb, v);
      cross_compute<P, false>(v, c.tid, c.tw + TwLayout<P>::CROSS);
      frame_store_cross<P>(v, c, cb);
    }
    __syncthreads();
    for (int k1 = 0; k1 < P::RC; ++k1) {
      frame_load_s0<P>(v, c, k1);
      fft_forward<P>(v, c.tile, c.tw, c.tid);
      point(k1, v);
      fft_inverse<P>(v, c.tile, c.tw, c.tid);
      after(k1);
      frame_store_s0<P>(v, c, k1);  // the positions this thread loaded: no hazard
      __syncthreads();              // tile reuse by the next sub-tile; frame complete after the last
    }
    for (int cb = 0; cb < P::RC; ++cb) {
      frame_load_cross<P>(v, c, cb);
      cross_compute<P, true>(v, c.tid, c.tw + TwLayout<P>::CROSS);
      near(cb, v);
    }
    __syncthreads();  // frame reuse by the next pattern
  }
}

// inverse only: the API adjoints.  load(k1, v) fills the spectrum registers of sub-tile k1.
template <class P, class Load, class Near>
__device__ __forceinline__ void inverse_pass(Cta<P>& c, Load load, Near near) {
  float2 v[P::E];
  if (P::RC == 1) {
    load(0, v);
    fft_inverse<P>(v, c.tile, c.tw, c.tid);
    near(0, v);
    __syncthreads();  // the next inverse starts by writing stage-2 positions other threads just read
  } else {
    for (int k1 = 0; k1 < P::RC; ++k1) {
      load(k1, v);
      fft_inverse<P>(v, c.tile, c.tw, c.tid);
      frame_store_s0<P>(v, c, k1);
      __syncthreads();
    }
    for (int cb = 0; cb < P::RC; ++cb) {
      frame_load_cross<P>(v, c, cb);
      cross_compute<P, true>(v, c.tid, c.tw + TwLayout<P>::CROSS);
      near(cb, v);
    }
    __syncthreads();
  }
}

// ---------------------------------------------------------------- object adjoint: scatter-add
// v holds the near plane of column block cb (natural ownership).  t = scale * conj(prb) * near is
// parked in the shared tile as an N-row x CW-column block; then every thread walks one column of
// the block over a run of 32 rows, forms the bilinear spread separably
//     h[y][x] = (1-gam) t[y][x] + gam t[y][x-1],   out[y][x] = (1-rho) h[y][x] + rho h[y-1][x]
// in registers (2 shared loads per output pixel) and issues ONE vector reduction
// (red.global.add.v2.f32) per object pixel instead of the reference's 8 scalar atomics per probe
// pixel (kernels.cu:73-80).  The column to the right of the block only receives the gam * t part
// (the neighbouring block adds its own share: the adds commute).  Leaves the tile free.
template <class P, bool FULL, class TileFree>
__device__ __forceinline__ void scatter_impl(float2 (&v)[P::E], const Cta<P>& c, int cb,
                                             const float2* __restrict__ prb, float scale,
                                             float2* __restrict__ grad_t, const Geo& g,
                                             const Pat& p, TileFree tile_free) {
  // FULL: probe window == frame and the (P+1)^2 footprint lies inside the object: no predicates.
  constexpr int ROWS = P::N, COLS = Cross<P>::CW;
  constexpr int PITCH = TileGeom<P>::WORDS / ROWS;
  constexpr int RUN = ROWS * COLS / P::NT;  // 32 output rows per thread
  static_assert(PITCH >= COLS + 1, "scatter block (+ one zero column) must fit the tile");
  static_assert(ROWS <= P::NT, "one thread per row zeroes the left guard column");
  float2* tile = c.tile;
  const float2 z = make_float2(0.f, 0.f);
  __syncthreads();  // other threads may still be reading the tile (last inverse stage)
  // block layout: row y at tile[y*PITCH], column 0 = zero guard (left neighbour of the block's first
  // column, owned by the previous block), column 1 + xl = t[y][xl]
#pragma unroll
  for (int e = 0; e < P::E; ++e) {
    int y, x;
    nat_coord<P>(c, cb, e, y, x);
    const int iy = FULL ? y : y - g.o, ix = FULL ? x : x - g.o;
    float2 t = z;
    if (FULL || ((unsigned)iy < (unsigned)g.P && (unsigned)ix < (unsigned)g.P)) {
      const float2 pr = __ldg(prb + iy * g.P + ix);
      t.x = scale * (pr.x * v[e].x + pr.y * v[e].y);  // conj(prb) * near
      t.y = scale * (pr.x * v[e].y - pr.y * v[e].x);
    }
    tile[y * PITCH + 1 + (x - cb * COLS)] = t;
  }
  if (c.tid < ROWS) tile[c.tid * PITCH] = z;
  __syncthreads();
  const int xl = c.tid % COLS, r0 = (c.tid / COLS) * RUN;
  const int x = cb * COLS + xl;                    // frame column of this thread's outputs
  const int oc = p.C + x - (FULL ? 0 : g.o);       // object column
  const float a0 = 1.f - p.gam, a1 = p.gam, b0 = 1.f - p.rho, b1 = p.rho;
  {
    const bool colok = FULL || ((x >= g.o) && (x <= g.o + g.P) && (oc < g.n));
    const float2* tp = tile + r0 * PITCH + xl;  // tp[0] = left tap, tp[1] = centre tap of row r0
    float2 hp = z;
    if (r0 > 0) {
      const float2 tl = tp[-PITCH], tc = tp[1 - PITCH];
      hp = make_float2(a0 * tc.x + a1 * tl.x, a0 * tc.y + a1 * tl.y);
    }
    const int orow0 = p.R + r0 - (FULL ? 0 : g.o);
    float2* dst = grad_t + (ptrdiff_t)orow0 * g.n + oc;
#pragma unroll 8
    for (int i = 0; i < RUN; ++i) {
      const float2 tl = tp[i * PITCH], tc = tp[i * PITCH + 1];
      const float2 hc = make_float2(a0 * tc.x + a1 * tl.x, a0 * tc.y + a1 * tl.y);
      const float2 out = make_float2(b0 * hc.x + b1 * hp.x, b0 * hc.y + b1 * hp.y);
      const int y = r0 + i;
      if (FULL || (colok && (y >= g.o) && (y <= g.o + g.P) && (orow0 + i < g.nz)))
        atomicAdd(dst + (ptrdiff_t)i * g.n, out);
      hp = hc;
    }
    if (r0 + RUN == ROWS) {  // the last run also emits row ROWS (only the rho * h[ROWS-1] share)
      if (FULL || (colok && (ROWS <= g.o + g.P) && (orow0 + RUN < g.nz)))
        atomicAdd(dst + (ptrdiff_t)RUN * g.n, make_float2(b1 * hp.x, b1 * hp.y));
    }
  }
  // the column right of the block, x = (cb+1)*COLS, receives only the gam * t[.][COLS-1] share
  if (c.tid < 32) {
    const int xe = (cb + 1) * COLS;
    const int oce = p.C + xe - (FULL ? 0 : g.o);
    const bool colok = FULL || ((xe >= g.o) && (xe <= g.o + g.P) && (oce < g.n));
    for (int y = c.tid; y <= ROWS; y += 32) {
      const float2 tc = y < ROWS ? tile[y * PITCH + COLS] : z;
      const float2 tu = y > 0 ? tile[(y - 1) * PITCH + COLS] : z;
      const int orow = p.R + y - (FULL ? 0 : g.o);
      if (FULL || (colok && (y >= g.o) && (y <= g.o + g.P) && (orow < g.nz)))
        atomicAdd(grad_t + (ptrdiff_t)orow * g.n + oce,
                  make_float2(a1 * (b0 * tc.x + b1 * tu.x), a1 * (b0 * tc.y + b1 * tu.y)));
    }
  }
  __syncthreads();
  tile_free();
}
// Scatter of the current pattern with the GATHER OF THE NEXT ONE folded into its loop (P == N, both
// footprints inside the object).  Once t is parked in the tile the 32 registers of v are dead, and
// the scatter loop -- two shared loads, a few FMAs and one fire-and-forget vector reduction per
// output pixel -- leaves the load/store unit's global path idle: each of its 32 iterations also
// fetches the four bilinear taps and the probe value of one near-plane pixel of pattern `pn`, whose
// L2 latency (the whole cost of the stand-alone gather phase) hides under the scatter.  On return v
// holds kappa * prb * patch(pn), ready for fft_forward.
template <class P, class TileFree>
__device__ __forceinline__ void scatter_gather_impl(float2 (&v)[P::E], const Cta<P>& c,
                                                    const float2* __restrict__ prb, float scale,
                                                    float2* __restrict__ grad_t,
                                                    const float2* __restrict__ psi_next, const Geo& g,
                                                    const Pat& p, const Pat& pn, TileFree tile_free) {
  static_assert(P::RC == 1, "single-tile plans only");
  constexpr int ROWS = P::N, COLS = P::N;
  constexpr int PITCH = TileGeom<P>::WORDS / ROWS;
  constexpr int RUN = ROWS * COLS / P::NT;
  static_assert(RUN == P::E, "one gathered pixel per scatter iteration");
  float2* tile = c.tile;
  const float2 z = make_float2(0.f, 0.f);
  __syncthreads();
#pragma unroll
  for (int e = 0; e < P::E; ++e) {
    int y, x;
    nat_coord<P>(c, 0, e, y, x);
    const float2 pr = __ldg(prb + y * g.P + x);
    float2 t;
    t.x = scale * (pr.x * v[e].x + pr.y * v[e].y);
    t.y = scale * (pr.x * v[e].y - pr.y * v[e].x);
    tile[y * PITCH + 1 + x] = t;
  }
  if (c.tid < ROWS) tile[c.tid * PITCH] = z;
  __syncthreads();
  const int xl = c.tid % COLS, r0 = (c.tid / COLS) * RUN;
  const int oc = p.C + xl;
  const float a0 = 1.f - p.gam, a1 = p.gam, b0 = 1.f - p.rho, b1 = p.rho;
  const float k00 = g.kappa * pn.w00, k01 = g.kappa * pn.w01, k10 = g.kappa * pn.w10, k11 = g.kappa * pn.w11;
  {
    const float2* tp = tile + r0 * PITCH + xl;
    float2 hp = z;
    if (r0 > 0) {
      const float2 tl = tp[-PITCH], tc = tp[1 - PITCH];
      hp = make_float2(a0 * tc.x + a1 * tl.x, a0 * tc.y + a1 * tl.y);
    }
    float2* dst = grad_t + (ptrdiff_t)(p.R + r0) * g.n + oc;
    constexpr int CH = 4;  // gathered pixels in flight per thread (5 loads each): bounded register use
#pragma unroll
    for (int i0 = 0; i0 < RUN; i0 += CH) {
      float2 f[CH][4], pr[CH];
#pragma unroll
      for (int j = 0; j < CH; ++j) {  // next pattern, near-plane pixels of registers i0 .. i0 + CH - 1
        int y, x;
        nat_coord<P>(c, 0, i0 + j, y, x);
        const float2* q = psi_next + (size_t)(pn.R + y) * g.n + (pn.C + x);
        f[j][0] = __ldg(q);
        f[j][1] = __ldg(q + 1);
        f[j][2] = __ldg(q + g.n);
        f[j][3] = __ldg(q + g.n + 1);
        pr[j] = __ldg(prb + y * g.P + x);
      }
#pragma unroll
      for (int j = 0; j < CH; ++j) {  // this pattern, output rows r0 + i0 .. r0 + i0 + CH - 1
        const int i = i0 + j;
        const float2 tl = tp[i * PITCH], tc = tp[i * PITCH + 1];
        const float2 hc = make_float2(a0 * tc.x + a1 * tl.x, a0 * tc.y + a1 * tl.y);
        atomicAdd(dst + (ptrdiff_t)i * g.n, make_float2(b0 * hc.x + b1 * hp.x, b0 * hc.y + b1 * hp.y));
        hp = hc;
      }
#pragma unroll
      for (int j = 0; j < CH; ++j) {
        float2 t;
        t.x = f[j][0].x * k00 + f[j][1].x * k01 + f[j][2].x * k10 + f[j][3].x * k11;
        t.y = f[j][0].y * k00 + f[j][1].y * k01 + f[j][2].y * k10 + f[j][3].y * k11;
        v[i0 + j] = make_float2(pr[j].x * t.x - pr[j].y * t.y, pr[j].x * t.y + pr[j].y * t.x);
      }
      asm volatile("" ::: "memory");  // keep the next chunk's loads behind this chunk's arithmetic
    }
    if (r0 + RUN == ROWS)
      atomicAdd(dst + (ptrdiff_t)RUN * g.n, make_float2(b1 * hp.x, b1 * hp.y));
  }
  if (c.tid < 32) {
    const int oce = p.C + COLS;
    for (int y = c.tid; y <= ROWS; y += 32) {
      const float2 tc = y < ROWS ? tile[y * PITCH + COLS] : z;
      const float2 tu = y > 0 ? tile[(y - 1) * PITCH + COLS] : z;
      atomicAdd(grad_t + (ptrdiff_t)(p.R + y) * g.n + oce,
                make_float2(a1 * (b0 * tc.x + b1 * tu.x), a1 * (b0 * tc.y + b1 * tu.y)));
    }
  }
  __syncthreads();
  tile_free();
}
// tile_free(): called by every thread once the tile is no longer needed (after a block barrier)
template <class P, class TileFree>
__device__ __forceinline__ void scatter_block(float2 (&v)[P::E], const Cta<P>& c, int cb,
                                              const float2* __restrict__ prb, float scale,
                                              float2* __restrict__ grad_t, const Geo& g,
                                              const Pat& p, TileFree tile_free) {
  if (g.P == P::N && p.inside) {
    scatter_impl<P, true>(v, c, cb, prb, scale, grad_t, g, p, tile_free);
  } else {
    const int z = opaque_zero();
    scatter_impl<P, false>(v, c, cb, prb, scale, grad_t, tied(g, z), tied(p, z), tile_free);
  }
}

// ---------------------------------------------------------------- probe adjoint
// thread-private accumulators in L2-resident scratch: acc[(cb*E + e)*NT + tid], natural ownership.
// acc += scale * near * conj(patch)                                    (kernels.cu:82-94)
template <class P, bool FULL, bool INSIDE>
__device__ __forceinline__ void pacc_impl(float2 (&v)[P::E], const Cta<P>& c, int cb,
                                          const float2* __restrict__ psi_t, float scale,
                                          const Geo& g, const Pat& p) {
  float2* acc = c.stash + (size_t)cb * P::E * P::NT + c.tid;
#pragma unroll
  for (int e = 0; e < P::E; ++e) {
    int y, x;
    nat_coord<P>(c, cb, e, y, x);
    const int iy = FULL ? y : y - g.o, ix = FULL ? x : x - g.o;
    if (FULL || ((unsigned)iy < (unsigned)g.P && (unsigned)ix < (unsigned)g.P)) {
      const float2 f = patch_at<INSIDE>(psi_t, g, p, iy, ix);
      float2 s = acc[e * P::NT];
      s.x += scale * (v[e].x * f.x + v[e].y * f.y);
      s.y += scale * (v[e].y * f.x - v[e].x * f.y);
      acc[e * P::NT] = s;
    }
  }
}
template <class P>
__device__ __forceinline__ void pacc_add(float2 (&v)[P::E], const Cta<P>& c, int cb,
                                         const float2* __restrict__ psi_t, float scale,
                                         const Geo& g, const Pat& p) {
  if (g.P == P::N && p.inside) {
    pacc_impl<P, true, true>(v, c, cb, psi_t, scale, g, p);
  } else {
    const int z = opaque_zero();
    pacc_impl<P, false, false>(v, c, cb, psi_t, scale, tied(g, z), tied(p, z));
  }
}
template <class P>
__device__ __forceinline__ void pacc_zero(const Cta<P>& c) {
  for (int i = 0; i < P::RC * P::E; ++i) c.stash[(size_t)i * P::NT + c.tid] = make_float2(0.f, 0.f);
}
// one vector atomic per probe pixel per CTA and angle (reference: 2 scalar atomics per pattern)
template <class P>
__device__ __forceinline__ void pacc_flush(const Cta<P>& c, float2* __restrict__ gp, const Geo& g) {
  for (int cb = 0; cb < P::RC; ++cb) {
    float2* acc = c.stash + (size_t)cb * P::E * P::NT + c.tid;
#pragma unroll
    for (int e = 0; e < P::E; ++e) {
      int y, x;
      nat_coord<P>(c, cb, e, y, x);
      const int iy = y - g.o, ix = x - g.o;
      if ((unsigned)iy < (unsigned)g.P && (unsigned)ix < (unsigned)g.P) {
        atomicAdd(gp + iy * g.P + ix, acc[e * P::NT]);
        acc[e * P::NT] = make_float2(0.f, 0.f);
      }
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
