// In-place, register-staged 2-D FFT of one N x N complex64 tile owned by one CTA.
//
// The tile lives in shared memory BETWEEN stages and in registers INSIDE a stage:
// every thread owns E = 32 elements per stage, runs a 2-D radix (2^bx along x,
// 2^by along y) butterfly on them in registers and puts them back at the same
// tile positions (decimation in frequency, digits taken from the top).  A stage's
// read set equals its write set, so the only block barriers are the ones between
// stages.  After the last forward stage the spectrum sits in registers at
// digit-reversed tile positions; the inverse runs the conjugate stages backwards
// from exactly that layout, so a forward -> pointwise -> inverse chain never
// permutes anything and touches shared memory 2 x (stages-1) times per transform.
//
// Replaces the cuFFT batched plan of the reference (src/cuda/ptychofft.cu:13-20,
// 72, 85): forward = unnormalised exp(-2 pi i ..), inverse = unnormalised
// exp(+2 pi i ..), DC at [0,0].
//
// The header is host+device: tests/emu_fft.cpp compiles it with g++ and runs all
// "threads" of a CTA sequentially to check the index algebra on the CPU.
#pragma once

#include <math.h>
#include <stdint.h>

#if defined(__CUDACC__)
#define PTX_HD __host__ __device__ __forceinline__
#else
#define PTX_HD inline
#ifndef PTX_HOST_FLOAT2
#define PTX_HOST_FLOAT2
struct float2 {
  float x, y;
};
static inline float2 make_float2(float a, float b) {
  float2 r;
  r.x = a;
  r.y = b;
  return r;
}
#endif
#endif

namespace ptx {

// ---------------------------------------------------------------- complex helpers
// PTX_F32X2 = 1 (device code, default off): complex add / subtract as ONE packed instruction on the
// (re, im) register pair (sm_100 FADD2 / FFMA2: __fadd2_rn, __ffma2_rn), bit-identical to the scalar
// pair (a - b = fma(b, -1, a) is exact).  It removes 20 % of the FP32 instructions of a transform
// (SASS count) but measured no faster on B200 -- 128^2 fused gradient 0.593 vs 0.582 ms, 256^2 1.590 vs
// 1.594 ms (profiles/r02i_f32x2.txt): the packed forms buy no issue slots and pin register pairs, which
// costs spills in the 128-register kernels.  Kept as a switch for the record.
#ifndef PTX_F32X2
#define PTX_F32X2 0
#endif
#if defined(__CUDA_ARCH__) && PTX_F32X2
PTX_HD float2 cadd(float2 a, float2 b) { return __fadd2_rn(a, b); }
PTX_HD float2 csub(float2 a, float2 b) { return __ffma2_rn(b, make_float2(-1.f, -1.f), a); }
#else
PTX_HD float2 cadd(float2 a, float2 b) { return make_float2(a.x + b.x, a.y + b.y); }
PTX_HD float2 csub(float2 a, float2 b) { return make_float2(a.x - b.x, a.y - b.y); }
#endif
// a + rot90(d), a - rot90(d) without materialising the rotated pair (it would cost two moves)
template <bool INV>
PTX_HD float2 cadd_rot(float2 a, float2 d) {
  return INV ? make_float2(a.x - d.y, a.y + d.x) : make_float2(a.x + d.y, a.y - d.x);
}
template <bool INV>
PTX_HD float2 csub_rot(float2 a, float2 d) {
  return INV ? make_float2(a.x + d.y, a.y - d.x) : make_float2(a.x - d.y, a.y + d.x);
}
PTX_HD float2 cmul(float2 a, float2 b) {
  return make_float2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x);
}
// a * conj(b)
PTX_HD float2 cmulc(float2 a, float2 b) {
  return make_float2(a.x * b.x + a.y * b.y, a.y * b.x - a.x * b.y);
}
// multiply by -i (forward quarter turn) or +i (inverse)
template <bool INV>
PTX_HD float2 rot90(float2 a) {
  return INV ? make_float2(-a.y, a.x) : make_float2(a.y, -a.x);
}

// ---------------------------------------------------------------- radix butterflies
// dft<R, INV>(v, base, stride): in-place R-point DFT of v[base + j*stride], natural order out.
template <int R, bool INV, int STRIDE, int EN>
struct Dft;

template <bool INV, int STRIDE, int EN>
struct Dft<1, INV, STRIDE, EN> {
  static PTX_HD void run(float2 (&v)[EN], int base) { (void)v; (void)base; }
};

template <bool INV, int STRIDE, int EN>
struct Dft<2, INV, STRIDE, EN> {
  static PTX_HD void run(float2 (&v)[EN], int base) {
    float2 a = v[base], b = v[base + STRIDE];
    v[base] = cadd(a, b);
    v[base + STRIDE] = csub(a, b);
  }
};

template <bool INV, int STRIDE, int EN>
struct Dft<4, INV, STRIDE, EN> {
  static PTX_HD void run(float2 (&v)[EN], int base) {
    float2 a0 = v[base], a1 = v[base + STRIDE], a2 = v[base + 2 * STRIDE], a3 = v[base + 3 * STRIDE];
    float2 t0 = cadd(a0, a2), t1 = csub(a0, a2), t2 = cadd(a1, a3), d3 = csub(a1, a3);
    v[base] = cadd(t0, t2);
    v[base + STRIDE] = cadd_rot<INV>(t1, d3);
    v[base + 2 * STRIDE] = csub(t0, t2);
    v[base + 3 * STRIDE] = csub_rot<INV>(t1, d3);
  }
};

template <bool INV, int STRIDE, int EN>
struct Dft<8, INV, STRIDE, EN> {
  static PTX_HD void run(float2 (&v)[EN], int base) {
    const float h = 0.70710678118654752440f;
    float2 a0 = v[base], a1 = v[base + STRIDE], a2 = v[base + 2 * STRIDE], a3 = v[base + 3 * STRIDE];
    float2 a4 = v[base + 4 * STRIDE], a5 = v[base + 5 * STRIDE], a6 = v[base + 6 * STRIDE],
           a7 = v[base + 7 * STRIDE];
    // even 4-point: a0 a2 a4 a6 ; odd 4-point: a1 a3 a5 a7
    float2 e0 = cadd(a0, a4), e1 = csub(a0, a4), e2 = cadd(a2, a6), e3 = csub(a2, a6);
    float2 E0 = cadd(e0, e2), E1 = cadd_rot<INV>(e1, e3), E2 = csub(e0, e2), E3 = csub_rot<INV>(e1, e3);
    float2 o0 = cadd(a1, a5), o1 = csub(a1, a5), o2 = cadd(a3, a7), o3 = csub(a3, a7);
    float2 O0 = cadd(o0, o2), O1 = cadd_rot<INV>(o1, o3), O2 = csub(o0, o2), O3 = csub_rot<INV>(o1, o3);
    // twiddles W8^k (forward: exp(-i pi k/4); inverse: conjugate)
    // W8^1 = (1 -/+ i) h ; W8^2 = -/+ i ; W8^3 = (-1 -/+ i) h
    float2 T1 = INV ? make_float2((O1.x - O1.y) * h, (O1.x + O1.y) * h)
                    : make_float2((O1.x + O1.y) * h, (O1.y - O1.x) * h);
    float2 T3 = INV ? make_float2((-O3.x - O3.y) * h, (O3.x - O3.y) * h)
                    : make_float2((O3.y - O3.x) * h, (-O3.x - O3.y) * h);
    v[base] = cadd(E0, O0);
    v[base + 4 * STRIDE] = csub(E0, O0);
    v[base + STRIDE] = cadd(E1, T1);
    v[base + 5 * STRIDE] = csub(E1, T1);
    v[base + 2 * STRIDE] = cadd_rot<INV>(E2, O2);
    v[base + 6 * STRIDE] = csub_rot<INV>(E2, O2);
    v[base + 3 * STRIDE] = cadd(E3, T3);
    v[base + 7 * STRIDE] = csub(E3, T3);
  }
};

template <bool INV, int STRIDE, int EN>
struct Dft<16, INV, STRIDE, EN> {
  // 16 = 4 x 4 : four 4-point DFTs over stride-4 subsequences, twiddle W16^(m k), four 4-point DFTs.
  static PTX_HD void run(float2 (&v)[EN], int base) {
    const float c1 = 0.92387953251128675613f, s1 = 0.38268343236508977173f;
    const float h = 0.70710678118654752440f;
    float2 t[16];
#pragma unroll
    for (int j = 0; j < 16; ++j) t[j] = v[base + j * STRIDE];
    // DIF: n = n1*4 + m ; A[k1][m] = sum_n1 x[n1*4+m] W4^(n1 k1) ; times W16^(m k1) ; 4-pt over m.
    float2 A[16];
#pragma unroll
    for (int m = 0; m < 4; ++m) {
      float2 a0 = t[m], a1 = t[4 + m], a2 = t[8 + m], a3 = t[12 + m];
      float2 u0 = cadd(a0, a2), u1 = csub(a0, a2), u2 = cadd(a1, a3), u3 = csub(a1, a3);
      A[0 * 4 + m] = cadd(u0, u2);
      A[1 * 4 + m] = cadd_rot<INV>(u1, u3);
      A[2 * 4 + m] = csub(u0, u2);
      A[3 * 4 + m] = csub_rot<INV>(u1, u3);
    }
    // twiddles W16^(m*k1), m,k1 in 1..3 : exponents 1,2,3,2,4,6,3,6,9
    const float2 w1 = make_float2(c1, INV ? s1 : -s1);
    const float2 w2 = make_float2(h, INV ? h : -h);
    const float2 w3 = make_float2(s1, INV ? c1 : -c1);
    const float2 w6 = make_float2(-h, INV ? h : -h);
    const float2 w9 = make_float2(-c1, INV ? -s1 : s1);
    A[1 * 4 + 1] = cmul(A[1 * 4 + 1], w1);
    A[1 * 4 + 2] = cmul(A[1 * 4 + 2], w2);
    A[1 * 4 + 3] = cmul(A[1 * 4 + 3], w3);
    A[2 * 4 + 1] = cmul(A[2 * 4 + 1], w2);
    A[2 * 4 + 2] = rot90<INV>(A[2 * 4 + 2]);
    A[2 * 4 + 3] = cmul(A[2 * 4 + 3], w6);
    A[3 * 4 + 1] = cmul(A[3 * 4 + 1], w3);
    A[3 * 4 + 2] = cmul(A[3 * 4 + 2], w6);
    A[3 * 4 + 3] = cmul(A[3 * 4 + 3], w9);
    // X[k1 + 4 q] = sum_m A[k1][m] W4^(m q)
#pragma unroll
    for (int k1 = 0; k1 < 4; ++k1) {
      float2 a0 = A[k1 * 4 + 0], a1 = A[k1 * 4 + 1], a2 = A[k1 * 4 + 2], a3 = A[k1 * 4 + 3];
      float2 u0 = cadd(a0, a2), u1 = csub(a0, a2), u2 = cadd(a1, a3), u3 = csub(a1, a3);
      v[base + (k1 + 0) * STRIDE] = cadd(u0, u2);
      v[base + (k1 + 4) * STRIDE] = cadd_rot<INV>(u1, u3);
      v[base + (k1 + 8) * STRIDE] = csub(u0, u2);
      v[base + (k1 + 12) * STRIDE] = csub_rot<INV>(u1, u3);
    }
  }
};

// ---------------------------------------------------------------- stage description
// A stage owns x bits [XLO, XLO+BX) and y bits [YLO, YLO+BY) as radix digits and, when
// BX+BY < log2(E), BB "batch" bits (independent butterflies): BMAP::bit(i), i < BB.
// WMAP::bit(j) tells which tile-coordinate bit work-item (thread) bit j drives.
// Bit codes: axis*16 + bit, axis 0 = x, 1 = y.  Lane bits come first (j = 0..4).
// Batch bits must lie above the radix digits of their axis (twiddles ignore them).
template <int XLO_, int BX_, int YLO_, int BY_, int BB_, class BMAP_, class WMAP_>
struct Stage {
  static constexpr int XLO = XLO_, BX = BX_, YLO = YLO_, BY = BY_, BB = BB_;
  static constexpr int RX = 1 << BX, RY = 1 << BY, NB = 1 << BB;
  static constexpr int E = RX * RY * NB;
  using WMAP = WMAP_;
  using BMAP = BMAP_;
};
struct NoBatch {
  static constexpr int bit(int) { return 0; }
};

// ---------------------------------------------------------------- plans
// Plan<L>, L = log2(detector size N).  A CTA of NT threads works on one LOCAL tile of NY x NX
// complex values (always 16384 for N >= 128) with E = 32 registers per thread and three 2-D radix
// stages S0..S2.  For N > 128 the N x N frame is RC = N^2/16384 local tiles: a "cross" radix-RC
// butterfly along y (digit y[L-1 : L-LC], natural-order thread ownership of N x N/RC column blocks)
// splits the frame into RC independent NY x NX sub-transforms (sub-tile k1 holds the rows of output
// frequency ky = k1 mod RC), which are staged through an L2-resident scratch frame.
//
// Shared-memory geometry of a tile: element (y, x) lives at y*RS + x + skew(x), where skew(x) is a
// carry-free function of the HIGH column bits (bit-linear: skew(a|b) = skew(a) + skew(b) for
// disjoint a, b) and RS = 4 mod 16.  With 64-bit accesses (16 lanes per wavefront) every stage of the
// plans below is bank-conflict free: bank pair = (4*y + x + skew(x)) mod 16 (audited on the CPU by
// tests/emu_fft.cpp).  Because both terms are additive in disjoint bit fields, a thread's 32
// element addresses are ONE per-thread base plus compile-time immediates: no per-access index math.
template <int L>
struct Plan;

template <class P>
struct TileGeom {
  static constexpr int RS = P::RS;
  static constexpr int WORDS = P::NY * RS;  // float2 elements
  static PTX_HD int idx(int y, int x) { return y * RS + x + P::skew(x); }
};

// N = 128: 512 threads x 32 elements; x = 3+2+2 bits, y = 2+3+2 bits (+1 batch bit on y).
struct W128_1 {  // free bits: x[3:0], y[4:0]
  static constexpr int bit(int j) {
    return j < 4 ? (0 * 16 + j) : (1 * 16 + (j - 4));
  }
};
struct W128_2 {  // free bits: x[1:0], y[1:0], x[6:4], y[6:5]
  static constexpr int bit(int j) {
    return j < 2 ? (0 * 16 + j) : j < 4 ? (1 * 16 + (j - 2)) : j < 7 ? (0 * 16 + (j - 4 + 4))
                                                                      : (1 * 16 + (j - 7 + 5));
  }
};
struct B128_3 {  // batch bit: y[2]
  static constexpr int bit(int) { return 16 + 2; }
};
struct W128_3 {  // free bits: x[6:2], y[6:3]; lanes = x4,x5,x2,x3,x6 (coalesced spectrum order)
  static constexpr int bit(int j) {
    return j == 0 ? 4 : j == 1 ? 5 : j == 2 ? 2 : j == 3 ? 3 : j == 4 ? 6 : (1 * 16 + (j - 5 + 3));
  }
};
#ifndef PTX_MINB6
#define PTX_MINB6 3
#endif

template <>
struct Plan<7> {
  static constexpr int L = 7, N = 128, LX = 7, LY = 7, NX = 128, NY = 128, RC = 1, LC = 0;
  static constexpr int E = 32, NT = NX * NY / E, NSTAGE = 3, WBITS = 9, MINB = 1;
  using S0 = Stage<4, 3, 5, 2, 0, NoBatch, W128_1>;
  using S1 = Stage<2, 2, 2, 3, 0, NoBatch, W128_2>;
  using S2 = Stage<0, 2, 0, 2, 1, B128_3, W128_3>;
  static constexpr int RS = 148;  // 128 + max skew 7, = 4 mod 16
  static PTX_HD constexpr int skew(int x) { return x >> 4; }  // stage-2 lanes x4,x5 -> +1,+2
  // The S1 <-> S2 exchange only moves data inside groups of 4 consecutive warps (S1 warps are
  // (x5,x6,y5,y6), S2 warps (y3,y4,y5,y6): both have (y5,y6) on top; tests/emu_fft.cpp audits it
  // when XG2 = 128), so a 128-thread named barrier could replace the block barrier there.  Measured
  // on B200 (profiles/r01j_group_barrier.txt): fwd / adj +2.5 %, fused gradient +0 %, line search
  // -8 % (the TMA-fed passes lose more from the drift than the butterflies gain) -> block barrier.
  static constexpr int XG2 = 0;
};

// N = 128, geometry of the pipelined kernel (ptycho_pipe.cuh): same stages as Plan<7>, but the exchange
// tile holds ONE float per element (real and imaginary planes change ownership one after the other),
// i.e. 32-bit accesses with 32 lanes per wavefront: row pitch 132 = 4 mod 32, skew(x) = (x >> 5) & 3,
// and stage 0 takes y2 instead of y0 as its fifth lane bit, so that in every stage the five lane bits
// move the bank index by five distinct powers of two (S0: 1,2,4,8,16; S1: 1,2,4,8,16; S2: 16,1,4,8,2).
struct WP128_1 {  // stage 0: free bits x[3:0], y[4:0]; lanes x0..x3, y2; then y0, y1, y3, y4
  static constexpr int bit(int j) {
    return j < 4 ? j : j == 4 ? (16 + 2) : j == 5 ? (16 + 0) : j == 6 ? (16 + 1) : (16 + (j - 7 + 3));
  }
};
struct Plan7P {
  static constexpr int L = 7, N = 128, LX = 7, LY = 7, NX = 128, NY = 128, RC = 1, LC = 0;
  static constexpr int E = 32, NT = NX * NY / E, NSTAGE = 3, WBITS = 9, MINB = 1;
  using S0 = Stage<4, 3, 5, 2, 0, NoBatch, WP128_1>;
  using S1 = Stage<2, 2, 2, 3, 0, NoBatch, W128_2>;
  using S2 = Stage<0, 2, 0, 2, 1, B128_3, W128_3>;
  static constexpr int RS = 132;  // floats; 4 mod 32
  static PTX_HD constexpr int skew(int x) { return (x >> 5) & 3; }
  static constexpr int XG2 = 0;
};


// N = 64: 128 threads x 32 elements; x = 3+3 bits, y = 2+2+2 (last stage: radix-4 on y, 8 batches).
struct W64_1 {  // active x[5:3], y[5:4]; free: x[2:0], y[3:0]; lanes x0,x1,y0,y1,x2
  static constexpr int bit(int j) {
    return j < 2 ? j : j < 4 ? (16 + (j - 2)) : j == 4 ? 2 : (16 + (j - 5 + 2));
  }
};
struct W64_2 {  // active x[2:0], y[3:2]; free: y[1:0], x[5:3], y[5:4]; lanes y0,y1,x4,x5,x3
  static constexpr int bit(int j) {
    return j < 2 ? (16 + j) : j == 2 ? 4 : j == 3 ? 5 : j == 4 ? 3 : (16 + (j - 5 + 4));
  }
};
struct B64_3 {  // batch bits: x0, x1, y2
  static constexpr int bit(int i) { return i < 2 ? i : (16 + 2); }
};
struct W64_3 {  // active y[1:0]; free: x[5:2], y[5:3]; lanes x4,x5,x2,x3,y3
  static constexpr int bit(int j) {
    return j == 0 ? 4 : j == 1 ? 5 : j == 2 ? 2 : j == 3 ? 3 : (16 + (j - 4 + 3));
  }
};
template <>
struct Plan<6> {
  static constexpr int L = 6, N = 64, LX = 6, LY = 6, NX = 64, NY = 64, RC = 1, LC = 0;
  static constexpr int E = 32, NT = NX * NY / E, NSTAGE = 3, WBITS = 7;
  // resident CTAs per SM the kernels are compiled for.  Without a bound ptxas takes 190 registers for the fused
  // gradient and only 2 CTAs (8 warps) fit.  Measured on B200 (profiles/r02z_grad64_ncu.txt), 8192 patterns: 3 CTAs
  // (170 registers) +9 ... +20 % on every pass; 4 CTAs (128 registers, spills) the same again on large batches but
  // 1-2 patterns per CTA at the 1024 patterns of one CG angle, where 3 is the most even (CG 1440-1490 it/s).
  static constexpr int MINB = PTX_MINB6;
  using S0 = Stage<3, 3, 4, 2, 0, NoBatch, W64_1>;
  using S1 = Stage<0, 3, 2, 2, 0, NoBatch, W64_2>;
  using S2 = Stage<0, 0, 0, 2, 3, B64_3, W64_3>;
  static constexpr int RS = 68;
  static PTX_HD constexpr int skew(int x) { return (x >> 4) & 3; }
  static constexpr int XG2 = 32;  // the S1 <-> S2 exchange stays inside a warp
};

// N = 256: cross radix 4 on y[7:6]; local tile 64 (y) x 256 (x): x = 3+3+2 bits, y = 2+2+2 bits.
struct WBIG_0 {  // free bits: x[4:0], y[3:0]; lanes x0..x4 (coalesced scratch rows)
  static constexpr int bit(int j) { return j < 5 ? j : (16 + (j - 5)); }
};
struct W256_1 {  // active x[4:2], y[3:2]; free: x[1:0], y[1:0], x[7:5], y[5:4]; lanes x0,x1,y0,y1,x5
  static constexpr int bit(int j) {
    return j < 2 ? j : j < 4 ? (16 + (j - 2)) : j < 7 ? (5 + (j - 4)) : (16 + 4 + (j - 7));
  }
};
struct B256_2 {  // batch bit: y[5]
  static constexpr int bit(int) { return 16 + 5; }
};
struct W256_2 {  // active x[1:0], y[1:0]; lanes x5,x6,x7,x2,x3 (= frequency bits 0..4), then x4, y[4:2]
  static constexpr int bit(int j) {
    return j < 3 ? (5 + j) : j < 6 ? (2 + (j - 3)) : (16 + 2 + (j - 6));
  }
};
template <>
struct Plan<8> {
  static constexpr int L = 8, N = 256, LX = 8, LY = 6, NX = 256, NY = 64, RC = 4, LC = 2;
  static constexpr int E = 32, NT = NX * NY / E, NSTAGE = 3, WBITS = 9, MINB = 1;
  using S0 = Stage<5, 3, 4, 2, 0, NoBatch, WBIG_0>;
  using S1 = Stage<2, 3, 2, 2, 0, NoBatch, W256_1>;
  using S2 = Stage<0, 2, 0, 2, 1, B256_2, W256_2>;
  static constexpr int XG2 = 0;   // S1 <-> S2 exchange spans the CTA: block barrier
  static constexpr int RS = 276;  // 256 + max skew 11, = 4 mod 16
  static PTX_HD constexpr int skew(int x) {  // stage-2 lanes x5,x6,x7 -> +1,+2,+8 (x2 gives +4)
    return ((x >> 5) & 3) + (((x >> 7) & 1) << 3);
  }
};

// N = 512: cross radix 16 on y[8:5]; local tile 32 (y) x 512 (x): x = 4+3+2 bits, y = 1+2+2 bits.
struct W512_1 {  // active x[4:2], y[3:2]; free: x[1:0], y[1:0], x[8:5], y[4]; lanes x0,x1,y0,y1,x5
  static constexpr int bit(int j) {
    return j < 2 ? j : j < 4 ? (16 + (j - 2)) : j < 8 ? (5 + (j - 4)) : (16 + 4);
  }
};
struct B512_2 {  // batch bit: y[4]
  static constexpr int bit(int) { return 16 + 4; }
};
struct W512_2 {  // active x[1:0], y[1:0]; lanes x5..x8,x2 (= frequency bits 0..4), then x3,x4,y2,y3
  static constexpr int bit(int j) {
    return j < 4 ? (5 + j) : j < 7 ? (2 + (j - 4)) : (16 + 2 + (j - 7));
  }
};
template <>
struct Plan<9> {
  static constexpr int L = 9, N = 512, LX = 9, LY = 5, NX = 512, NY = 32, RC = 16, LC = 4;
  static constexpr int E = 32, NT = NX * NY / E, NSTAGE = 3, WBITS = 9, MINB = 1;
  using S0 = Stage<5, 4, 4, 1, 0, NoBatch, WBIG_0>;
  using S1 = Stage<2, 3, 2, 2, 0, NoBatch, W512_1>;
  using S2 = Stage<0, 2, 0, 2, 1, B512_2, W512_2>;
  static constexpr int XG2 = 0;
  static constexpr int RS = 532;  // 512 + max skew 15, = 4 mod 16
  static PTX_HD constexpr int skew(int x) { return (x >> 5) & 15; }  // stage-2 lanes x5..x8
};

// ---------------------------------------------------------------- coordinates of a thread's elements
template <class ST, int WBITS>
PTX_HD void fixed_coords(int w, int& xf, int& yf) {
  xf = 0;
  yf = 0;
#pragma unroll
  for (int j = 0; j < WBITS; ++j) {
    const int code = ST::WMAP::bit(j);
    const int b = (w >> j) & 1;
    if (code >= 16)
      yf |= b << (code - 16);
    else
      xf |= b << code;
  }
}

// element e = ex + RX*(ey + RY*eb)
template <class ST>
PTX_HD void elem_offset(int e, int& dx, int& dy) {
  const int ex = e & (ST::RX - 1);
  const int ey = (e >> ST::BX) & (ST::RY - 1);
  const int eb = e >> (ST::BX + ST::BY);
  dx = ex << ST::XLO;
  dy = ey << ST::YLO;
#pragma unroll
  for (int i = 0; i < ST::BB; ++i) {
    const int code = ST::BMAP::bit(i);
    const int b = (eb >> i) & 1;
    if (code >= 16)
      dy |= b << (code - 16);
    else
      dx |= b << code;
  }
}

// ---------------------------------------------------------------- twiddle tables
// For stage S and axis A the table holds W_{2^(LO+B)}^(m k), k = 1..R-1, m = 0..2^LO-1, laid out [k-1][m].
template <class ST>
struct TwSize {
  static constexpr int X = ST::XLO > 0 ? (ST::RX - 1) * (1 << ST::XLO) : 0;
  static constexpr int Y = ST::YLO > 0 ? (ST::RY - 1) * (1 << ST::YLO) : 0;
};
template <class P>
struct TwLayout {
  static constexpr int X0 = 0;
  static constexpr int Y0 = X0 + TwSize<typename P::S0>::X;
  static constexpr int X1 = Y0 + TwSize<typename P::S0>::Y;
  static constexpr int Y1 = X1 + TwSize<typename P::S1>::X;
  static constexpr int X2 = Y1 + TwSize<typename P::S1>::Y;
  static constexpr int Y2 = X2 + TwSize<typename P::S2>::X;
  static constexpr int CROSS = Y2 + TwSize<typename P::S2>::Y;  // W_N^(ylow k1), [k1-1][ylow]
  static constexpr int TOTAL = CROSS + (P::RC > 1 ? (P::RC - 1) * P::NY : 0);
};

// host-side fill (double precision -> float)
template <class ST>
inline void fill_twiddles_stage(float2* tx, float2* ty) {
  const double PI2 = 6.283185307179586476925286766559;
  if (ST::XLO > 0) {
    const int M = 1 << ST::XLO, S = M * ST::RX;
    for (int k = 1; k < ST::RX; ++k)
      for (int m = 0; m < M; ++m) {
        double a = -PI2 * (double)(m * k) / (double)S;
        tx[(k - 1) * M + m] = make_float2((float)cos(a), (float)sin(a));
      }
  }
  if (ST::YLO > 0) {
    const int M = 1 << ST::YLO, S = M * ST::RY;
    for (int k = 1; k < ST::RY; ++k)
      for (int m = 0; m < M; ++m) {
        double a = -PI2 * (double)(m * k) / (double)S;
        ty[(k - 1) * M + m] = make_float2((float)cos(a), (float)sin(a));
      }
  }
}
template <class P>
inline void fill_twiddles(float2* tw) {
  using TL = TwLayout<P>;
  fill_twiddles_stage<typename P::S0>(tw + TL::X0, tw + TL::Y0);
  fill_twiddles_stage<typename P::S1>(tw + TL::X1, tw + TL::Y1);
  fill_twiddles_stage<typename P::S2>(tw + TL::X2, tw + TL::Y2);
  if (P::RC > 1) {
    const double PI2 = 6.283185307179586476925286766559;
    for (int k = 1; k < P::RC; ++k)
      for (int m = 0; m < P::NY; ++m) {
        double a = -PI2 * (double)(m * k) / (double)P::N;
        tw[TL::CROSS + (k - 1) * P::NY + m] = make_float2((float)cos(a), (float)sin(a));
      }
  }
}

// ---------------------------------------------------------------- one stage, in registers
// Forward: butterflies along x, x twiddles, butterflies along y, y twiddles (order commutes).
// Inverse: conjugate twiddles first, then conjugate butterflies.
template <class ST, bool INV>
PTX_HD void stage_compute(float2 (&v)[ST::E], int xf, int yf, const float2* twx, const float2* twy) {
  constexpr int E = ST::E;
  constexpr int MX = 1 << ST::XLO, MY = 1 << ST::YLO;
  const int mx = xf & (MX - 1), my = yf & (MY - 1);
  if (INV) {
    if (ST::XLO > 0 && ST::BX > 0) {
#pragma unroll
      for (int k = 1; k < ST::RX; ++k) {
        const float2 w = twx[(k - 1) * MX + mx];
#pragma unroll
        for (int o = 0; o < E / ST::RX; ++o) v[o * ST::RX + k] = cmulc(v[o * ST::RX + k], w);
      }
    }
    if (ST::YLO > 0 && ST::BY > 0) {
#pragma unroll
      for (int k = 1; k < ST::RY; ++k) {
        const float2 w = twy[(k - 1) * MY + my];
#pragma unroll
        for (int b = 0; b < ST::NB; ++b)
#pragma unroll
          for (int i = 0; i < ST::RX; ++i) {
            const int e = i + ST::RX * (k + ST::RY * b);
            v[e] = cmulc(v[e], w);
          }
      }
    }
  }
  // x butterflies: elements e = ex + RX*o, stride 1
#pragma unroll
  for (int o = 0; o < E / ST::RX; ++o) Dft<ST::RX, INV, 1, E>::run(v, o * ST::RX);
  if (!INV && ST::XLO > 0 && ST::BX > 0) {
#pragma unroll
    for (int k = 1; k < ST::RX; ++k) {
      const float2 w = twx[(k - 1) * MX + mx];
#pragma unroll
      for (int o = 0; o < E / ST::RX; ++o) v[o * ST::RX + k] = cmul(v[o * ST::RX + k], w);
    }
  }
  // y butterflies: elements e = i + RX*(ey + RY*b), stride RX
#pragma unroll
  for (int b = 0; b < ST::NB; ++b)
#pragma unroll
    for (int i = 0; i < ST::RX; ++i) Dft<ST::RY, INV, ST::RX, E>::run(v, i + ST::RX * ST::RY * b);
  if (!INV && ST::YLO > 0 && ST::BY > 0) {
#pragma unroll
    for (int k = 1; k < ST::RY; ++k) {
      const float2 w = twy[(k - 1) * MY + my];
#pragma unroll
      for (int b = 0; b < ST::NB; ++b)
#pragma unroll
        for (int i = 0; i < ST::RX; ++i) {
          const int e = i + ST::RX * (k + ST::RY * b);
          v[e] = cmul(v[e], w);
        }
    }
  }
}

// idx() is additive over the disjoint thread / element bit fields: one base per thread and stage,
// compile-time immediates per element.
template <class ST, class P>
PTX_HD void stage_load(float2 (&v)[ST::E], const float2* tile, int xf, int yf) {
  const float2* base = tile + TileGeom<P>::idx(yf, xf);
#pragma unroll
  for (int e = 0; e < ST::E; ++e) {
    int dx, dy;
    elem_offset<ST>(e, dx, dy);
    v[e] = base[TileGeom<P>::idx(dy, dx)];
  }
}
template <class ST, class P>
PTX_HD void stage_store(const float2 (&v)[ST::E], float2* tile, int xf, int yf) {
  float2* base = tile + TileGeom<P>::idx(yf, xf);
#pragma unroll
  for (int e = 0; e < ST::E; ++e) {
    int dx, dy;
    elem_offset<ST>(e, dx, dy);
    base[TileGeom<P>::idx(dy, dx)] = v[e];
  }
}

// ---------------------------------------------------------------- spectrum position <-> frequency
// After the forward transform the value at tile position p (per axis) is X[k] with the digits of
// k in reversed order: position digits (top..bottom) are k1, k2, k3 and k = k1 + R1*k2 + R1*R2*k3.
template <class P>
PTX_HD int pos_to_freq_x(int p) {
  using A = typename P::S0;
  using B = typename P::S1;
  using C = typename P::S2;
  const int k1 = (p >> A::XLO) & (A::RX - 1);
  const int k2 = (p >> B::XLO) & (B::RX - 1);
  const int k3 = (p >> C::XLO) & (C::RX - 1);
  return k1 + A::RX * (k2 + B::RX * k3);
}
template <class P>
PTX_HD int pos_to_freq_y(int p) {
  using A = typename P::S0;
  using B = typename P::S1;
  using C = typename P::S2;
  const int k1 = (p >> A::YLO) & (A::RY - 1);
  const int k2 = (p >> B::YLO) & (B::RY - 1);
  const int k3 = (p >> C::YLO) & (C::RY - 1);
  return k1 + A::RY * (k2 + B::RY * k3);
}

// ---------------------------------------------------------------- cross stage (N > 128)
// Natural ("cross") ownership of column block c: thread tid owns, for b < E/RC, the RC elements
// (y = j*NY + ylow, x = c*CW + xc), j < RC, with pair index p = tid + NT*b, xc = p % CW, ylow = p / CW
// (CW = NX/RC = NY).  Register e = j + RC*b.  For RC = 1 the natural ownership is stage-0 ownership.
template <class P>
struct Cross {
  static constexpr int RC = P::RC, CW = P::NX / P::RC, NB = P::E / P::RC;
  static constexpr int LCW = P::LX - P::LC;  // log2(CW)
  static PTX_HD void pair(int tid, int b, int& ylow, int& xc) {
    const int p = tid + P::NT * b;
    xc = p & (CW - 1);
    ylow = p >> LCW;
  }
};

// forward: radix-RC butterfly over j, then twiddle W_N^(ylow*k1); inverse: conjugates, reversed.
template <class P, bool INV>
PTX_HD void cross_compute(float2 (&v)[P::E], int tid, const float2* twc) {
  using C = Cross<P>;
  if (INV) {
#pragma unroll
    for (int b = 0; b < C::NB; ++b) {
      int ylow, xc;
      C::pair(tid, b, ylow, xc);
#pragma unroll
      for (int k = 1; k < P::RC; ++k)
        v[k + P::RC * b] = cmulc(v[k + P::RC * b], twc[(k - 1) * P::NY + ylow]);
    }
  }
#pragma unroll
  for (int b = 0; b < C::NB; ++b) Dft<P::RC, INV, 1, P::E>::run(v, P::RC * b);
  if (!INV) {
#pragma unroll
    for (int b = 0; b < C::NB; ++b) {
      int ylow, xc;
      C::pair(tid, b, ylow, xc);
#pragma unroll
      for (int k = 1; k < P::RC; ++k)
        v[k + P::RC * b] = cmul(v[k + P::RC * b], twc[(k - 1) * P::NY + ylow]);
    }
  }
}

// scratch frame layout: sub-tile k1 is a contiguous NY x NX row-major block
template <class P>
PTX_HD int scratch_index(int k1, int ylow, int x) {
  return (k1 * P::NY + ylow) * P::NX + x;
}

}  // namespace ptx
