// Position correction: batched phase correlation with an upsampled matrix DFT around the peak.
//
// Replaces register_translation_batch + _upsampled_dft_batch of the reference
// (/root/reference/src/libtike/cufft/ptycho.py:163-248) and, in its fused form, the whole
// position-correction block of the CG loop (ptycho.py:398-403): two extra fwd() calls with an
// all-ones probe, an element-wise product, cupy.fft.ifft2, two argmax passes and two complex128
// einsums ([S,U,N] x [S,N,N] and [S,U,N] x [S,U,N], U = 150) -- about ten full-array kernels and
// 0.4 GB of complex128 temporaries per 1000 patterns at 128^2 -- become ONE kernel in which a CTA
// keeps a pattern's whole chain on chip:
//
//   F_a, F_b (far fields of the two objects / the two given images)     fft_tile.cuh transforms
//   P = F_a conj(F_b)                        complex64, like the reference (ptycho.py:207)
//   c = IFFT2(P), (y*, x*) = first argmax |c|, wrapped to signed shifts  (ptycho.py:208-219)
//   G[jr, jc] = sum_{r,c} P[r,c] W^((jr-off_r) k_r + (jc-off_c) k_c),  W = exp(2 pi i / (uf N)),
//       off = dftshift - shift * uf, k = signed frequency index          (ptycho.py:221-233, 163-190)
//   (jr*, jc*) = first argmax |G|;  shift += (j* - dftshift) / uf        (ptycho.py:234-240)
//
// The upsampled DFT is evaluated exactly as the reference does it -- as two matrix products in
// float64 -- because near the peak neighbouring samples of |G| differ by far less than float32
// resolution.  Since `off` is an integer, W^((j-off) k) = E[j][k] * ph[k] with a PATTERN-INDEPENDENT
// table E[j][k] = W^(j k) (one [U, N] complex128 table per plan and upsampling factor, L2 resident)
// and two per-pattern phase vectors ph_r, ph_c (N sincospi each).  Per jc-chunk of JC columns:
//   stage 1   T[jc][r]  = ph_r[r] * sum_c E[jc][c] (P[r][c] ph_c[c])     register tile 2 x 4 / thread
//   stage 2   G[jr][jc] = sum_r E[jr][r] T[jc][r]                        register tile TJR x TJC / thread
// with A/B tiles staged through shared memory (pitch KCH+1 complex128 = conflict-free 128-bit
// loads for lanes on consecutive rows) and T kept in shared memory: the float64 pipe (DFMA) is the
// roof, 4 DFMA per complex multiply-add, (U' N^2 + U'^2 N) of them per pattern (U' = U rounded up).
#pragma once

#include "ptycho_passes.cuh"

namespace ptx {

template <class P>
struct RegCfg;  // JC: window columns per chunk; KCH: contraction chunk; TJR x TJC: stage-2 thread tile
template <>
struct RegCfg<Plan<6>> {
  static constexpr int JC = 32, KCH = 16, TJR = 5, TJC = 2;
};
template <>
struct RegCfg<Plan<7>> {
  static constexpr int JC = 32, KCH = 16, TJR = 5, TJC = 2;
};
template <>
struct RegCfg<Plan<8>> {
  static constexpr int JC = 16, KCH = 8, TJR = 5, TJC = 1;
};
template <>
struct RegCfg<Plan<9>> {
  static constexpr int JC = 8, KCH = 4, TJR = 3, TJC = 1;
};

// Geometry of the tensor-core (DMMA, mma.sync.m8n8k4.f64) form of the same two products.  A warp owns
// 8 window columns x 32 rows (stage 1: 4 C tiles) and 8 columns x MT2 row tiles (stage 2); a complex
// tile product is 4 real DMMAs on fragments that ONE 128-bit shared load per lane delivers (re, im).
// Operand traffic is ~0.3 B per FMA instead of the 3 B of the scalar-DFMA register tiles, which ncu
// showed to be bound by the shared-memory data pipe (profiles/r01i_register128_ncu.txt).
template <class P>
struct RegMma {
  using C = RegCfg<P>;
  static constexpr int N = P::N, NT = P::NT, NW = NT / 32;
  static constexpr int JC = C::JC, KCH = C::KCH;
  static constexpr int KP = (KCH % 8 == 4) ? KCH : KCH + 4;  // pitch = 4 mod 8 complex128: the two
  static constexpr int TP = N + 4;                           // rows of a quarter warp hit disjoint banks
  static constexpr int W1J = JC / 8, W1R = NW / W1J, RCH = 32 * W1R;
  static constexpr int W2C = JC / 8, W2R = NW / W2C;
  static constexpr int NTILE = (REG_UMAX + 7) / 8;                       // 19 row tiles of 8 cover U <= 150
  static constexpr int MT2 = W2R >= NTILE ? 1 : (W2R == 1 ? 5 : (NTILE + W2R - 1) / W2R);
  static constexpr int JRCH = 8 * MT2 * W2R;                             // window rows per stage-2 pass
  static constexpr int SZ_A = JC * KP, SZ_B = RCH * KP, SZ_A2 = JRCH * KP;
  static constexpr int SZ_AB = (SZ_A + SZ_B) > SZ_A2 ? (SZ_A + SZ_B) : SZ_A2;
  static constexpr int OFF_T = SZ_AB, OFF_PH = OFF_T + JC * TP, TOTAL = OFF_PH + 2 * N;
  static constexpr size_t BYTES = (size_t)TOTAL * 16;
  static_assert(W1J * W1R == NW && W2C * W2R == NW && N % RCH == 0 && N % KCH == 0 && KCH % 4 == 0, "tiling");
  static_assert(REG_EROWS % JC == 0 && REG_EROWS >= ((REG_UMAX + JRCH - 1) / JRCH) * JRCH, "E rows");
};

// Geometry of the low-rank (Jacobi-Anger) form.  The matrix-DFT kernel over the 1.5-pixel window,
//   W^((j - dftshift) k) = exp(i theta_k x_j),  theta_k = 2 pi dftshift k / (uf N) in [-0.75 pi, 0.75 pi],
//   x_j = (j - dftshift) / dftshift in [-1, 1],
// has numerical rank ~20:  exp(i theta x) = sum_n eps_n i^n J_n(theta) T_n(x)  (eps_0 = 1, eps_n = 2),
// and J_n(0.75 pi) < 1e-20 for n >= 24.  With a[n][k] = eps_n J_n(theta_k) (real) and T[j][n] = T_n(x_j)
//   G = T (i^(n+m) D) T^t,   D[n][m] = sum_r a[n][r] ph_r[r] (sum_c a[m][c] ph_c[c] P[r][c]),
// i.e. 24 N^2 + 24^2 N + 24^2 U + 24 U^2 real-times-complex multiply-adds instead of U N^2 + U^2 N
// complex ones: 12x fewer FP64 operations at N = 128, 20x at N = 256, equal to the direct products
// to rounding (2.7e-15 of the peak, tests/test_register_oracle.py).  All four products run on the
// FP64 tensor cores (2 DMMAs per real-times-complex tile step).
template <class P>
struct RegCheb {
  static constexpr int N = P::N, NT = P::NT, NW = NT / 32;
  static constexpr int NQ = REG_NQ, QT = NQ / 8;     // expansion terms, in tiles of 8
  static constexpr int KCH = N >= 512 ? 8 : 16;      // contraction chunk of stage 1
  static constexpr int KP = KCH + 4;                 // pitch of the staged A (real) / B (complex) tiles
  static constexpr int RCH = N < 128 ? N : 128;      // image rows per pass of stages 1-2
  static constexpr int NTW = (RCH / 8) / NW;         // stage-1 row tiles per warp
  static constexpr int DP = RCH + 4;                 // pitch of a2s / D1s
  static constexpr int QP = NQ + 4;                  // pitch of Tsm / Ds / Hs
  static constexpr int PPW = (QT * QT + NW - 1) / NW;  // stage-2 tile pairs per warp
  static constexpr int JROWS = 160;                  // window rows/cols rounded up to tiles (U <= 150)
  // phase A (bytes): ph | As | Bs | a2s | D1s        phase B: Tsm | Ds | Hs   (same region)
  static constexpr size_t A_PH = 0, A_AS = A_PH + (size_t)2 * N * 16, A_BS = A_AS + (size_t)NQ * KP * 8,
                          A_A2 = A_BS + (size_t)RCH * KP * 16, A_D1 = A_A2 + (size_t)NQ * DP * 8,
                          A_END = A_D1 + (size_t)NQ * DP * 16;
  static constexpr size_t B_T = 0, B_DS = B_T + (size_t)JROWS * QP * 8, B_HS = B_DS + (size_t)NQ * QP * 16,
                          B_END = B_HS + (size_t)JROWS * QP * 16;
  static constexpr size_t BYTES = A_END > B_END ? A_END : B_END;
  static_assert(NTW >= 1 && NTW * NW * 8 == RCH && N % RCH == 0 && N % KCH == 0, "tiling");
};

template <class P>
struct RegGeom {
  using C = RegCfg<P>;
  static constexpr int N = P::N, NT = P::NT, NW = NT / 32;
  static constexpr int JC = C::JC, KCH = C::KCH, TJR = C::TJR, TJC = C::TJC;
  static constexpr int KP = KCH + 1;  // tile pitch in complex128
  // stage 1: a warp is 4 (jc) x 8 (r) lanes, 2 x 4 outputs per lane -> 8 jc x 32 r per warp
  static constexpr int W1J = JC / 8, W1R = NW / W1J, RCH = 32 * W1R;
  // stage 2: a warp is 8 (jc) x 4 (jr) lanes, TJC x TJR outputs per lane
  static constexpr int W2C = JC / (8 * TJC), W2R = NW / W2C, JRCH = 4 * TJR * W2R;
  static constexpr int TP = N + 1;  // pitch of T
  // shared layout (complex128 units)
  static constexpr int SZ_A = JC * KP, SZ_B = RCH * KP, SZ_A2 = JRCH * KP;
  static constexpr int SZ_AB = (SZ_A + SZ_B) > SZ_A2 ? (SZ_A + SZ_B) : SZ_A2;
  static constexpr int OFF_T = SZ_AB, OFF_PH = OFF_T + JC * TP, TOTAL = OFF_PH + 2 * N;
  static constexpr size_t BYTES_FMA = (size_t)TOTAL * 16;
  static constexpr size_t BYTES_DIRECT = BYTES_FMA > RegMma<P>::BYTES ? BYTES_FMA : RegMma<P>::BYTES;
  static constexpr size_t BYTES = BYTES_DIRECT > RegCheb<P>::BYTES ? BYTES_DIRECT : RegCheb<P>::BYTES;
  // the GEMM region reuses the FFT tile when it fits, else it follows the mbarriers
  static constexpr bool IN_TILE = BYTES <= Smem<P>::OFF_TW;
  static constexpr size_t OFF = IN_TILE ? 0 : Smem<P>::OFF_DBUF;
  static constexpr size_t SMEM = IN_TILE ? Smem<P>::BYTES_NODATA : Smem<P>::OFF_DBUF + BYTES;
  static_assert(W1J >= 1 && W1R >= 1 && W1J * W1R == NW, "stage-1 warp grid");
  static_assert(W2C >= 1 && W2R >= 1 && W2C * W2R == NW, "stage-2 warp grid");
  static_assert(N % RCH == 0 && N % KCH == 0, "chunking");
  static_assert(REG_EROWS % JC == 0 && REG_EROWS >= ((REG_UMAX + JRCH - 1) / JRCH) * JRCH, "E rows");
};

struct Best {
  float v;
  int i;
};
__device__ __forceinline__ void best_take(Best& b, float v, int i) {
  if (v > b.v || (v == b.v && i < b.i)) {
    b.v = v;
    b.i = i;
  }
}
struct BestD {
  double v;
  int i;
};
__device__ __forceinline__ void bestd_take(BestD& b, double v, int i) {
  if (v > b.v || (v == b.v && i < b.i)) {
    b.v = v;
    b.i = i;
  }
}
// first-occurrence argmax over the CTA (numpy / cupy argmax: the smallest index among equal maxima)
template <int NW>
__device__ __forceinline__ int block_argmax(BestD b, double* red, int tid) {
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) {
    const double ov = __shfl_down_sync(0xffffffffu, b.v, off);
    const int oi = __shfl_down_sync(0xffffffffu, b.i, off);
    bestd_take(b, ov, oi);
  }
  int* ri = reinterpret_cast<int*>(red + NW);
  __syncthreads();
  if ((tid & 31) == 0) {
    red[tid >> 5] = b.v;
    ri[tid >> 5] = b.i;
  }
  __syncthreads();
  BestD r;
  r.v = red[0];
  r.i = ri[0];
  for (int w = 1; w < NW; ++w) bestd_take(r, red[w], ri[w]);
  __syncthreads();
  return r.i;
}

__device__ __forceinline__ void zfma(double2& acc, const double2 a, const double2 b) {
  acc.x = fma(a.x, b.x, acc.x);
  acc.x = fma(-a.y, b.y, acc.x);
  acc.y = fma(a.x, b.y, acc.y);
  acc.y = fma(a.y, b.x, acc.y);
}

// The matrix-DFT refinement of one pattern.  praw: [N][N] complex64 image product in natural
// frequency order (this CTA's scratch); (sy, sx): whole-pixel shifts.  Returns the flat index
// jr * U + jc of the first maximum of |G|.
template <class P>
__device__ __forceinline__ int register_refine(const float2* __restrict__ praw,
                                               const double2* __restrict__ E, double2* sm,
                                               double* red, int tid, int sy, int sx, int U, int uf,
                                               int dftshift) {
  using G = RegGeom<P>;
  constexpr int N = G::N, NT = G::NT, KP = G::KP, KCH = G::KCH, JC = G::JC;
  double2* As = sm;
  double2* Bs = sm + G::SZ_A;
  double2* A2s = sm;
  double2* Ts = sm + G::OFF_T;
  double2* phr = sm + G::OFF_PH;
  double2* phc = phr + N;
  const int lane = tid & 31, warp = tid >> 5;
  // per-pattern phase vectors: ph[k] = W^(-off k_signed), -off = shift * uf - dftshift
  const int period = uf * N;
  for (int k = tid; k < 2 * N; k += NT) {
    const int kk = k < N ? k : k - N;
    const int ks = kk < N / 2 ? kk : kk - N;
    const int m = (k < N ? sy : sx) * uf - dftshift;
    long long q = ((long long)m * ks) % period;
    if (q < 0) q += period;
    double s, c;
    sincospi(2.0 * (double)q / (double)period, &s, &c);
    (k < N ? phr : phc)[kk] = make_double2(c, s);
  }
  // stage-1 thread coordinates
  const int tjl = lane & 3, trl = lane >> 2;
  const int w1j = warp % G::W1J, w1r = warp / G::W1J;
  // stage-2 thread coordinates
  const int cl = lane & 7, rl = lane >> 3;
  const int w2c = warp % G::W2C, w2r = warp / G::W2C;
  BestD best;
  best.v = -1.0;
  best.i = 0;
  for (int jc0 = 0; jc0 < U; jc0 += JC) {
    // ---------------- stage 1: T[jc][r] = ph_r[r] sum_c E[jc0+jc][c] P[r][c] ph_c[c]
    for (int r0 = 0; r0 < N; r0 += G::RCH) {
      double2 acc[2][4];
#pragma unroll
      for (int i = 0; i < 2; ++i)
#pragma unroll
        for (int q = 0; q < 4; ++q) acc[i][q] = make_double2(0.0, 0.0);
      // the global loads of chunk k0 + KCH are issued before chunk k0 is consumed: their L2 latency
      // hides behind the DFMA work of a whole chunk
      constexpr int NA = (JC * KCH + NT - 1) / NT, NB = (G::RCH * KCH + NT - 1) / NT;
      double2 ra[NA];
      float2 rb[NB];
      auto fetch1 = [&](int k0) {
#pragma unroll
        for (int u = 0; u < NA; ++u) {
          const int t = tid + u * NT;
          if (JC * KCH % NT == 0 || t < JC * KCH)
            ra[u] = __ldg(E + (size_t)(jc0 + t / KCH) * N + k0 + t % KCH);
        }
#pragma unroll
        for (int u = 0; u < NB; ++u) {
          const int t = tid + u * NT;
          if (G::RCH * KCH % NT == 0 || t < G::RCH * KCH)
            rb[u] = __ldcg(praw + (size_t)(r0 + t / KCH) * N + k0 + t % KCH);
        }
      };
      fetch1(0);
      for (int k0 = 0; k0 < N; k0 += KCH) {
        __syncthreads();  // the previous tiles (and, first time round, the phase vectors) are settled
#pragma unroll
        for (int u = 0; u < NA; ++u) {
          const int t = tid + u * NT;
          if (JC * KCH % NT == 0 || t < JC * KCH) As[(t / KCH) * KP + t % KCH] = ra[u];
        }
#pragma unroll
        for (int u = 0; u < NB; ++u) {
          const int t = tid + u * NT;
          if (G::RCH * KCH % NT == 0 || t < G::RCH * KCH) {
            const double2 ph = phc[k0 + t % KCH];
            const float2 pv = rb[u];
            Bs[(t / KCH) * KP + t % KCH] = make_double2((double)pv.x * ph.x - (double)pv.y * ph.y,
                                                        (double)pv.x * ph.y + (double)pv.y * ph.x);
          }
        }
        __syncthreads();
        if (k0 + KCH < N) fetch1(k0 + KCH);
        const double2* ap = As + (tjl + 8 * w1j) * KP;
        const double2* bp = Bs + (trl + 32 * w1r) * KP;
#pragma unroll 4
        for (int kk = 0; kk < KCH; ++kk) {
          const double2 a0 = ap[kk], a1 = ap[4 * KP + kk];
          double2 b[4];
#pragma unroll
          for (int q = 0; q < 4; ++q) b[q] = bp[8 * q * KP + kk];
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            zfma(acc[0][q], a0, b[q]);
            zfma(acc[1][q], a1, b[q]);
          }
        }
      }
#pragma unroll
      for (int i = 0; i < 2; ++i)
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const int j = tjl + 4 * i + 8 * w1j, r = r0 + trl + 8 * q + 32 * w1r;
          const double2 ph = phr[r], a = acc[i][q];
          Ts[j * G::TP + r] = make_double2(a.x * ph.x - a.y * ph.y, a.x * ph.y + a.y * ph.x);
        }
    }
    // ---------------- stage 2: G[jr][jc0+jc] = sum_r E[jr][r] T[jc][r]
    for (int jr0 = 0; jr0 < U; jr0 += G::JRCH) {
      double2 acc[G::TJR][G::TJC];
#pragma unroll
      for (int i = 0; i < G::TJR; ++i)
#pragma unroll
        for (int q = 0; q < G::TJC; ++q) acc[i][q] = make_double2(0.0, 0.0);
      constexpr int NA2 = (G::JRCH * KCH + NT - 1) / NT;
      double2 ra2[NA2];
      auto fetch2 = [&](int k0) {
#pragma unroll
        for (int u = 0; u < NA2; ++u) {
          const int t = tid + u * NT;
          if (G::JRCH * KCH % NT == 0 || t < G::JRCH * KCH)
            ra2[u] = __ldg(E + (size_t)(jr0 + t / KCH) * N + k0 + t % KCH);
        }
      };
      fetch2(0);
      for (int k0 = 0; k0 < N; k0 += KCH) {
        __syncthreads();  // stage-1 tiles / previous A2 tile consumed, T complete
#pragma unroll
        for (int u = 0; u < NA2; ++u) {
          const int t = tid + u * NT;
          if (G::JRCH * KCH % NT == 0 || t < G::JRCH * KCH) A2s[(t / KCH) * KP + t % KCH] = ra2[u];
        }
        __syncthreads();
        if (k0 + KCH < N) fetch2(k0 + KCH);
        const double2* ap = A2s + (rl + 4 * G::TJR * w2r) * KP;
        const double2* bp = Ts + (cl + 8 * G::TJC * w2c) * G::TP + k0;
#pragma unroll 4
        for (int kk = 0; kk < KCH; ++kk) {
          double2 b[G::TJC];
#pragma unroll
          for (int q = 0; q < G::TJC; ++q) b[q] = bp[8 * q * G::TP + kk];
#pragma unroll
          for (int i = 0; i < G::TJR; ++i) {
            const double2 a = ap[4 * i * KP + kk];
#pragma unroll
            for (int q = 0; q < G::TJC; ++q) zfma(acc[i][q], a, b[q]);
          }
        }
      }
#pragma unroll
      for (int i = 0; i < G::TJR; ++i)
#pragma unroll
        for (int q = 0; q < G::TJC; ++q) {
          const int jr = jr0 + rl + 4 * i + 4 * G::TJR * w2r;
          const int jc = jc0 + cl + 8 * q + 8 * G::TJC * w2c;
          if (jr < U && jc < U) {
            const double2 a = acc[i][q];
            bestd_take(best, a.x * a.x + a.y * a.y, jr * U + jc);
          }
        }
    }
  }
  return block_argmax<G::NW>(best, red, tid);
}

__device__ __forceinline__ void dmma(double (&c)[2], double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
               : "+d"(c[0]), "+d"(c[1])
               : "d"(a), "d"(b));
}

// Tensor-core form of register_refine: same chunking, staging and results (up to summation order).
template <class P>
__device__ __forceinline__ int register_refine_mma(const float2* __restrict__ praw,
                                                   const double2* __restrict__ E, double2* sm,
                                                   double* red, int tid, int sy, int sx, int U, int uf,
                                                   int dftshift) {
  using G = RegMma<P>;
  constexpr int N = G::N, NT = G::NT, KP = G::KP, KCH = G::KCH, JC = G::JC, TP = G::TP;
  double2* As = sm;
  double2* Bs = sm + G::SZ_A;
  double2* A2s = sm;
  double2* Ts = sm + G::OFF_T;
  double2* phr = sm + G::OFF_PH;
  double2* phc = phr + N;
  const int lane = tid & 31, warp = tid >> 5;
  const int period = uf * N;
  for (int k = tid; k < 2 * N; k += NT) {
    const int kk = k < N ? k : k - N;
    const int ks = kk < N / 2 ? kk : kk - N;
    const int m = (k < N ? sy : sx) * uf - dftshift;
    long long q = ((long long)m * ks) % period;
    if (q < 0) q += period;
    double s, c;
    sincospi(2.0 * (double)q / (double)period, &s, &c);
    (k < N ? phr : phc)[kk] = make_double2(c, s);
  }
  const int fr = lane >> 2, fk = lane & 3;  // fragment row (A: m, B: n) and k index of this lane
  const int w1j = warp % G::W1J, w1r = warp / G::W1J;
  const int w2c = warp % G::W2C, w2r = warp / G::W2C;
  BestD best;
  best.v = -1.0;
  best.i = 0;
  for (int jc0 = 0; jc0 < U; jc0 += JC) {
    // ---------------- stage 1: T[jc][r] = ph_r[r] sum_c E[jc0+jc][c] P[r][c] ph_c[c]
    for (int r0 = 0; r0 < N; r0 += G::RCH) {
      double cre[4][2], cim[4][2];
#pragma unroll
      for (int q = 0; q < 4; ++q) cre[q][0] = cre[q][1] = cim[q][0] = cim[q][1] = 0.0;
      constexpr int NA = (JC * KCH + NT - 1) / NT, NB = (G::RCH * KCH + NT - 1) / NT;
      double2 ra[NA];
      float2 rb[NB];
      auto fetch1 = [&](int k0) {
#pragma unroll
        for (int u = 0; u < NA; ++u) {
          const int t = tid + u * NT;
          if (JC * KCH % NT == 0 || t < JC * KCH)
            ra[u] = __ldg(E + (size_t)(jc0 + t / KCH) * N + k0 + t % KCH);
        }
#pragma unroll
        for (int u = 0; u < NB; ++u) {
          const int t = tid + u * NT;
          if (G::RCH * KCH % NT == 0 || t < G::RCH * KCH)
            rb[u] = __ldcg(praw + (size_t)(r0 + t / KCH) * N + k0 + t % KCH);
        }
      };
      fetch1(0);
      for (int k0 = 0; k0 < N; k0 += KCH) {
        __syncthreads();
#pragma unroll
        for (int u = 0; u < NA; ++u) {
          const int t = tid + u * NT;
          if (JC * KCH % NT == 0 || t < JC * KCH) As[(t / KCH) * KP + t % KCH] = ra[u];
        }
#pragma unroll
        for (int u = 0; u < NB; ++u) {
          const int t = tid + u * NT;
          if (G::RCH * KCH % NT == 0 || t < G::RCH * KCH) {
            const double2 ph = phc[k0 + t % KCH];
            const float2 pv = rb[u];
            Bs[(t / KCH) * KP + t % KCH] = make_double2((double)pv.x * ph.x - (double)pv.y * ph.y,
                                                        (double)pv.x * ph.y + (double)pv.y * ph.x);
          }
        }
        __syncthreads();
        if (k0 + KCH < N) fetch1(k0 + KCH);
        const double2* ap = As + (8 * w1j + fr) * KP + fk;
        const double2* bp = Bs + (32 * w1r + fr) * KP + fk;
#pragma unroll
        for (int kk = 0; kk < KCH; kk += 4) {
          const double2 a = ap[kk];
          const double nai = -a.y;
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            const double2 b = bp[8 * q * KP + kk];
            dmma(cre[q], a.x, b.x);
            dmma(cre[q], nai, b.y);
            dmma(cim[q], a.x, b.y);
            dmma(cim[q], a.y, b.x);
          }
        }
      }
      // C fragment: row = fr (window column), columns 2 fk + {0, 1} (rows r of the n tile)
#pragma unroll
      for (int q = 0; q < 4; ++q)
#pragma unroll
        for (int u = 0; u < 2; ++u) {
          const int r = r0 + 32 * w1r + 8 * q + 2 * fk + u;
          const double2 ph = phr[r];
          const double ax = cre[q][u], ay = cim[q][u];
          Ts[(8 * w1j + fr) * TP + r] = make_double2(ax * ph.x - ay * ph.y, ax * ph.y + ay * ph.x);
        }
    }
    // ---------------- stage 2: G[jr][jc0+jc] = sum_r E[jr][r] T[jc][r]
    for (int jr0 = 0; jr0 < U; jr0 += G::JRCH) {
      double cre[G::MT2][2], cim[G::MT2][2];
#pragma unroll
      for (int i = 0; i < G::MT2; ++i) cre[i][0] = cre[i][1] = cim[i][0] = cim[i][1] = 0.0;
      constexpr int NA2 = (G::JRCH * KCH + NT - 1) / NT;
      double2 ra2[NA2];
      auto fetch2 = [&](int k0) {
#pragma unroll
        for (int u = 0; u < NA2; ++u) {
          const int t = tid + u * NT;
          if (G::JRCH * KCH % NT == 0 || t < G::JRCH * KCH)
            ra2[u] = __ldg(E + (size_t)(jr0 + t / KCH) * N + k0 + t % KCH);
        }
      };
      fetch2(0);
      for (int k0 = 0; k0 < N; k0 += KCH) {
        __syncthreads();
#pragma unroll
        for (int u = 0; u < NA2; ++u) {
          const int t = tid + u * NT;
          if (G::JRCH * KCH % NT == 0 || t < G::JRCH * KCH) A2s[(t / KCH) * KP + t % KCH] = ra2[u];
        }
        __syncthreads();
        if (k0 + KCH < N) fetch2(k0 + KCH);
        const double2* ap = A2s + (8 * G::MT2 * w2r + fr) * KP + fk;
        const double2* bp = Ts + (8 * w2c + fr) * TP + k0 + fk;
#pragma unroll
        for (int kk = 0; kk < KCH; kk += 4) {
          const double2 b = bp[kk];
          const double nbi = -b.y;
#pragma unroll
          for (int i = 0; i < G::MT2; ++i) {
            if (jr0 + 8 * (G::MT2 * w2r + i) < U) {  // warp-uniform: tiles beyond the window are skipped
              const double2 a = ap[8 * i * KP + kk];
              dmma(cre[i], a.x, b.x);
              dmma(cre[i], a.y, nbi);
              dmma(cim[i], a.x, b.y);
              dmma(cim[i], a.y, b.x);
            }
          }
        }
      }
#pragma unroll
      for (int i = 0; i < G::MT2; ++i)
#pragma unroll
        for (int u = 0; u < 2; ++u) {
          const int jr = jr0 + 8 * (G::MT2 * w2r + i) + fr;
          const int jc = jc0 + 8 * w2c + 2 * fk + u;
          if (jr < U && jc < U)
            bestd_take(best, cre[i][u] * cre[i][u] + cim[i][u] * cim[i][u], jr * U + jc);
        }
    }
  }
  return block_argmax<G::NW>(best, red, tid);
}

__device__ __forceinline__ double2 rot_i(double x, double y, int q) {  // (x + i y) * i^q
  switch (q & 3) {
    case 0: return make_double2(x, y);
    case 1: return make_double2(-y, x);
    case 2: return make_double2(-x, -y);
    default: return make_double2(y, -x);
  }
}

// Low-rank form of the refinement (see RegCheb).  atab: [NQ][N] doubles, ttab: [JROWS][NQ] doubles.
template <class P>
__device__ __forceinline__ int register_refine_cheb(const float2* __restrict__ praw,
                                                    const double* __restrict__ atab,
                                                    const double* __restrict__ ttab,
                                                    unsigned char* smraw, double* red, int tid, int sy,
                                                    int sx, int U) {
  using G = RegCheb<P>;
  constexpr int N = G::N, NT = G::NT, NW = G::NW, NQ = G::NQ, QT = G::QT, KCH = G::KCH, KP = G::KP;
  constexpr int RCH = G::RCH, NTW = G::NTW, DP = G::DP, QP = G::QP, PPW = G::PPW;
  double2* phr = reinterpret_cast<double2*>(smraw + G::A_PH);
  double2* phc = phr + N;
  double* As = reinterpret_cast<double*>(smraw + G::A_AS);
  double2* Bs = reinterpret_cast<double2*>(smraw + G::A_BS);
  double* a2s = reinterpret_cast<double*>(smraw + G::A_A2);
  double2* D1s = reinterpret_cast<double2*>(smraw + G::A_D1);
  double* Tsm = reinterpret_cast<double*>(smraw + G::B_T);
  double2* Ds = reinterpret_cast<double2*>(smraw + G::B_DS);
  double2* Hs = reinterpret_cast<double2*>(smraw + G::B_HS);
  const int lane = tid & 31, warp = tid >> 5, fr = lane >> 2, fk = lane & 3;
  // whole-pixel shift as a phase ramp: ph[k] = exp(2 pi i s k_signed / N)
  for (int k = tid; k < 2 * N; k += NT) {
    const int kk = k < N ? k : k - N;
    const int ks = kk < N / 2 ? kk : kk - N;
    int q = ((k < N ? sy : sx) * ks) % N;
    if (q < 0) q += N;
    double sn, cs;
    sincospi(2.0 * (double)q / (double)N, &sn, &cs);
    (k < N ? phr : phc)[kk] = make_double2(cs, sn);
  }
  double d2re[PPW][2], d2im[PPW][2];
#pragma unroll
  for (int s = 0; s < PPW; ++s) d2re[s][0] = d2re[s][1] = d2im[s][0] = d2im[s][1] = 0.0;
  for (int r0 = 0; r0 < N; r0 += RCH) {
    __syncthreads();  // stage 2 of the previous pass is done with a2s / D1s; ph is settled
    for (int t = tid; t < NQ * RCH; t += NT) a2s[(t / RCH) * DP + t % RCH] = __ldg(atab + (size_t)(t / RCH) * N + r0 + t % RCH);
    // ---------------- stage 1: D1[m][r] = ph_r[r] sum_c a[m][c] ph_c[c] P[r][c]
    double cre[QT][NTW][2], cim[QT][NTW][2];
#pragma unroll
    for (int i = 0; i < QT; ++i)
#pragma unroll
      for (int q = 0; q < NTW; ++q) cre[i][q][0] = cre[i][q][1] = cim[i][q][0] = cim[i][q][1] = 0.0;
    constexpr int NA = (NQ * KCH + NT - 1) / NT, NB = (RCH * KCH + NT - 1) / NT;
    double ra[NA];
    float2 rb[NB];
    auto fetch1 = [&](int k0) {
#pragma unroll
      for (int u = 0; u < NA; ++u) {
        const int t = tid + u * NT;
        if (t < NQ * KCH) ra[u] = __ldg(atab + (size_t)(t / KCH) * N + k0 + t % KCH);
      }
#pragma unroll
      for (int u = 0; u < NB; ++u) {
        const int t = tid + u * NT;
        if (RCH * KCH % NT == 0 || t < RCH * KCH)
          rb[u] = __ldcg(praw + (size_t)(r0 + t / KCH) * N + k0 + t % KCH);
      }
    };
    fetch1(0);
    for (int k0 = 0; k0 < N; k0 += KCH) {
      __syncthreads();
#pragma unroll
      for (int u = 0; u < NA; ++u) {
        const int t = tid + u * NT;
        if (t < NQ * KCH) As[(t / KCH) * KP + t % KCH] = ra[u];
      }
#pragma unroll
      for (int u = 0; u < NB; ++u) {
        const int t = tid + u * NT;
        if (RCH * KCH % NT == 0 || t < RCH * KCH) {
          const double2 ph = phc[k0 + t % KCH];
          const float2 pv = rb[u];
          Bs[(t / KCH) * KP + t % KCH] = make_double2((double)pv.x * ph.x - (double)pv.y * ph.y,
                                                      (double)pv.x * ph.y + (double)pv.y * ph.x);
        }
      }
      __syncthreads();
      if (k0 + KCH < N) fetch1(k0 + KCH);
      const double* ap = As + fr * KP + fk;
      const double2* bp = Bs + (8 * NTW * warp + fr) * KP + fk;
#pragma unroll
      for (int kk = 0; kk < KCH; kk += 4) {
        double av[QT];
#pragma unroll
        for (int i = 0; i < QT; ++i) av[i] = ap[8 * i * KP + kk];
#pragma unroll
        for (int q = 0; q < NTW; ++q) {
          const double2 b = bp[8 * q * KP + kk];
#pragma unroll
          for (int i = 0; i < QT; ++i) {
            dmma(cre[i][q], av[i], b.x);
            dmma(cim[i][q], av[i], b.y);
          }
        }
      }
    }
#pragma unroll
    for (int q = 0; q < NTW; ++q)
#pragma unroll
      for (int u = 0; u < 2; ++u) {
        const int rl = 8 * (NTW * warp + q) + 2 * fk + u;
        const double2 ph = phr[r0 + rl];
#pragma unroll
        for (int i = 0; i < QT; ++i) {
          const double ax = cre[i][q][u], ay = cim[i][q][u];
          D1s[(8 * i + fr) * DP + rl] = make_double2(ax * ph.x - ay * ph.y, ax * ph.y + ay * ph.x);
        }
      }
    __syncthreads();
    // ---------------- stage 2: D[n][m] += sum_r a[n][r] D1[m][r]
#pragma unroll
    for (int s = 0; s < PPW; ++s) {
      const int pi = warp + NW * s;
      if (pi < QT * QT) {
        const double* ap = a2s + (8 * (pi / QT) + fr) * DP + fk;
        const double2* bp = D1s + (8 * (pi % QT) + fr) * DP + fk;
#pragma unroll 8
        for (int kk = 0; kk < RCH; kk += 4) {
          const double av = ap[kk];
          const double2 b = bp[kk];
          dmma(d2re[s], av, b.x);
          dmma(d2im[s], av, b.y);
        }
      }
    }
  }
  __syncthreads();  // phase A is over: the region becomes Tsm | Ds | Hs
#pragma unroll
  for (int s = 0; s < PPW; ++s) {
    const int pi = warp + NW * s;
    if (pi < QT * QT) {
#pragma unroll
      for (int u = 0; u < 2; ++u) {
        const int n = 8 * (pi / QT) + fr, m = 8 * (pi % QT) + 2 * fk + u;
        Ds[n * QP + m] = rot_i(d2re[s][u], d2im[s][u], n + m);
      }
    }
  }
  for (int t = tid; t < G::JROWS * NQ; t += NT) Tsm[(t / NQ) * QP + t % NQ] = __ldg(ttab + t);
  __syncthreads();
  // ---------------- stage 3: H[n][jc] = sum_m D[n][m] T[jc][m], stored as Hs[jc][n]
  for (int jt = warp; jt < G::JROWS / 8; jt += NW) {
    double hre[QT][2], him[QT][2];
#pragma unroll
    for (int i = 0; i < QT; ++i) hre[i][0] = hre[i][1] = him[i][0] = him[i][1] = 0.0;
#pragma unroll
    for (int kk = 0; kk < NQ; kk += 4) {
      const double b = Tsm[(8 * jt + fr) * QP + kk + fk];
#pragma unroll
      for (int i = 0; i < QT; ++i) {
        const double2 a = Ds[(8 * i + fr) * QP + kk + fk];
        dmma(hre[i], a.x, b);
        dmma(him[i], a.y, b);
      }
    }
#pragma unroll
    for (int i = 0; i < QT; ++i)
#pragma unroll
      for (int u = 0; u < 2; ++u)
        Hs[(8 * jt + 2 * fk + u) * QP + 8 * i + fr] = make_double2(hre[i][u], him[i][u]);
  }
  __syncthreads();
  // ---------------- stage 4: G[jr][jc] = sum_n T[jr][n] H[n][jc]; running first-occurrence argmax
  BestD best;
  best.v = -1.0;
  best.i = 0;
  const int JT = (U + 7) / 8;
  constexpr int GW = 5;  // window-column tiles per work unit (independent accumulation chains)
  const int NG = (JT + GW - 1) / GW;
  for (int un = warp; un < JT * NG; un += NW) {
    const int jrt = un / NG, g0 = (un % NG) * GW;
    double av[NQ / 4];
#pragma unroll
    for (int s = 0; s < NQ / 4; ++s) av[s] = Tsm[(8 * jrt + fr) * QP + 4 * s + fk];
    double gre[GW][2], gim[GW][2];
#pragma unroll
    for (int w = 0; w < GW; ++w) gre[w][0] = gre[w][1] = gim[w][0] = gim[w][1] = 0.0;
#pragma unroll
    for (int s = 0; s < NQ / 4; ++s) {
#pragma unroll
      for (int w = 0; w < GW; ++w) {
        const double2 b = Hs[(8 * (g0 + w) + fr) * QP + 4 * s + fk];  // tiles up to JROWS/8 exist
        dmma(gre[w], av[s], b.x);
        dmma(gim[w], av[s], b.y);
      }
    }
#pragma unroll
    for (int w = 0; w < GW; ++w)
#pragma unroll
      for (int u = 0; u < 2; ++u) {
        const int jr = 8 * jrt + fr, jc = 8 * (g0 + w) + 2 * fk + u;
        if (jr < U && jc < U)
          bestd_take(best, gre[w][u] * gre[w][u] + gim[w][u] * gim[w][u], jr * U + jc);
      }
  }
  return block_argmax<NW>(best, red, tid);
}

// near plane under an all-ones probe (ptycho.py:399-400: probe[:, 0] * 0 + 1): kappa * patch inside
// the probe window, no probe loads
template <class P, bool INSIDE>
__device__ __forceinline__ void gather_ones_impl(float2 (&v)[P::E], const Cta<P>& c, int cb,
                                                 const float2* __restrict__ psi_t, const Geo& g,
                                                 const Pat& p) {
#pragma unroll
  for (int e = 0; e < P::E; ++e) {
    int y, x;
    nat_coord<P>(c, cb, e, y, x);
    const int iy = y - g.o, ix = x - g.o;
    float2 r = make_float2(0.f, 0.f);
    if ((unsigned)iy < (unsigned)g.P && (unsigned)ix < (unsigned)g.P) {
      const float2 t = patch_at<INSIDE>(psi_t, g, p, iy, ix);
      r = make_float2(g.kappa * t.x, g.kappa * t.y);
    }
    v[e] = r;
  }
}
template <class P>
__device__ __forceinline__ void gather_ones(float2 (&v)[P::E], const Cta<P>& c, int cb,
                                            const float2* __restrict__ psi_t, const Geo& g, const Pat& p) {
  if (p.inside) {
    gather_ones_impl<P, true>(v, c, cb, psi_t, g, p);
  } else {
    const int z = opaque_zero();
    gather_ones_impl<P, false>(v, c, cb, psi_t, tied(g, z), tied(p, z));
  }
}

// MODE 0: far fields of two OBJECTS under an all-ones probe at the scan positions (ptycho.py:398-401)
// MODE 1: src / target given in Fourier space [S,N,N] (register_translation_batch(space='fourier'))
// MODE 2: src / target given in real space [S,N,N]    (space='real': fft2 of both first)
template <class P, int MODE>
__global__ void __launch_bounds__(P::NT) k_register(const PassArgs a,
                                                    const __grid_constant__ CUtensorMap tm_a,
                                                    const __grid_constant__ CUtensorMap tm_b) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  using G = RegGeom<P>;
  Cta<P> c;
  cta_setup<P>(c, smem_raw, a);
  const Geo g = a.g;
  constexpr size_t NN = (size_t)P::N * P::N;
  double2* gemm = reinterpret_cast<double2*>(smem_raw + G::OFF);
  float2* praw = reinterpret_cast<float2*>(c.accp);  // [N][N] complex64, natural frequency order
  const int npat = g.T * g.S;
  const int U = a.reg_U, uf = a.reg_uf, dftshift = U / 2;
  for (int pat = blockIdx.x; pat < npat; pat += gridDim.x) {
    const int t = pat / g.S;
    Pat p;
    p.skip = false;
    if (MODE == 0) p = make_pat(a.scan, pat, g);
    double* out = a.reg_out + 2 * (size_t)pat;
    if (p.skip) {  // both far fields are identically 0: every argmax lands on index 0 (ptycho.py:210, 236)
      if (c.tid == 0) {
        const double s = uf > 1 ? -(double)dftshift / (double)uf : 0.0;
        out[0] = s;
        out[1] = s;
      }
      continue;
    }
    const float2* src = MODE == 0 ? a.psi + (size_t)t * g.nz * g.n : a.far_in + (size_t)pat * NN;
    const float2* tgt = MODE == 0 ? a.psi_b + (size_t)t * g.nz * g.n : a.psi_b + (size_t)pat * NN;
    Best bst;
    bst.v = -1.f;
    bst.i = 0;
    // v holds the target's spectrum; src_at(e) fetches the source's
    auto product = [&](int k1, float2(&v)[P::E], auto src_at) {
      constexpr int CH = 8;  // source loads in flight ahead of their use (one round trip per CH pixels)
#pragma unroll
      for (int e0 = 0; e0 < P::E; e0 += CH) {
        float2 sv[CH];
#pragma unroll
        for (int j = 0; j < CH; ++j) sv[j] = src_at(e0 + j);
#pragma unroll
        for (int j = 0; j < CH; ++j) {
          const int e = e0 + j;
          v[e] = cmulc(sv[j], v[e]);  // src * conj(target), complex64 (ptycho.py:207)
          __stcg(praw + spec_index<P>(c, k1, e), v[e]);
        }
      }
    };
    auto peak = [&](int cb, float2(&v)[P::E]) {
#pragma unroll
      for (int e = 0; e < P::E; ++e) {
        int y, x;
        nat_coord<P>(c, cb, e, y, x);
        best_take(bst, v[e].x * v[e].x + v[e].y * v[e].y, y * P::N + x);
      }
    };
    if (MODE == 1) {
      inverse_pass<P>(
          c,
          [&](int k1, float2(&v)[P::E]) {
#pragma unroll
            for (int e = 0; e < P::E; ++e) v[e] = __ldg(tgt + spec_index<P>(c, k1, e));
            product(k1, v, [&](int e) { return __ldg(src + spec_index<P>(c, k1, e)); });
          },
          peak);
    } else {
      auto load_nat = [&](const float2* img, int cb, float2(&v)[P::E]) {
        if (MODE == 0) {
          if (Patch<P>::TMA && a.use_tma) {
            // all-ones probe, object patch by tensor copy (block cb + 1 is in flight while block cb's
            // taps are read): the plain four-tap gather was 18 % of this kernel's stalls at 256^2
            const CUtensorMap* tm = img == src ? &tm_a : &tm_b;
            if (cb == 0) {
              __syncthreads();  // every thread is done with the tile
              patch_issue<P>(c, tm, g, p, t, 0);
            }
            const int shift = c.pshift;
            gather_tma<P, true>(v, c, cb, shift, nullptr, g, p);  // ends with a block barrier
            if (cb + 1 < P::RC) patch_issue<P>(c, tm, g, p, t, cb + 1);
          } else {
            gather_ones<P>(v, c, cb, img, g, p);
          }
        } else {
#pragma unroll
          for (int e = 0; e < P::E; ++e) {
            int y, x;
            nat_coord<P>(c, cb, e, y, x);
            v[e] = __ldg(img + y * P::N + x);
          }
        }
      };
      spectrum_pass<P>(
          c, false, [&](int cb, float2(&v)[P::E]) { load_nat(src, cb, v); },
          [&](int k1, float2(&v)[P::E]) {
            float2* st = c.stash + (size_t)k1 * P::E * P::NT + c.tid;
#pragma unroll
            for (int e = 0; e < P::E; ++e) st[e * P::NT] = v[e];
          },
          [&](int) {});
      fused_pass<P>(
          c, [&](int cb, float2(&v)[P::E]) { load_nat(tgt, cb, v); },
          [&](int k1, float2(&v)[P::E]) {
            const float2* st = c.stash + (size_t)k1 * P::E * P::NT + c.tid;
            product(k1, v, [&](int e) { return st[e * P::NT]; });
          },
          [&](int) {}, peak);
    }
    BestD bd;
    bd.v = (double)bst.v;
    bd.i = bst.i;
    const int imax = block_argmax<G::NW>(bd, c.red, c.tid);  // its barriers also publish praw
    int sy = imax / P::N, sx = imax % P::N;
    if (sy > P::N / 2) sy -= P::N;  // ptycho.py:213-219: strictly beyond the midpoint wraps
    if (sx > P::N / 2) sx -= P::N;
    double oy = (double)sy, ox = (double)sx;
    if (uf > 1) {
      const int j =
          a.reg_algo == 2
              ? register_refine_cheb<P>(praw, a.reg_A, a.reg_T, smem_raw + G::OFF, c.red, c.tid, sy, sx, U)
              : a.reg_algo == 1
                    ? register_refine_mma<P>(praw, a.reg_E, gemm, c.red, c.tid, sy, sx, U, uf, dftshift)
                    : register_refine<P>(praw, a.reg_E, gemm, c.red, c.tid, sy, sx, U, uf, dftshift);
      oy += (double)(j / U - dftshift) / (double)uf;
      ox += (double)(j % U - dftshift) / (double)uf;
    }
    if (c.tid == 0) {
      out[0] = oy;
      out[1] = ox;
    }
    __syncthreads();  // the GEMM region (= the FFT tile) is free again
  }
}

}  // namespace ptx
