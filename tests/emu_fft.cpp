// CPU emulation of the CTA-wide tile FFT (libtike-cufft_b200/csrc/fft_tile.cuh).
// Runs every "thread" of a CTA sequentially, stage by stage (a stage boundary == __syncthreads),
// including the cross stage + scratch frame of the N > 128 plans, and checks forward and inverse
// transforms against a double-precision separable DFT.  Also audits shared-memory bank conflicts.
// Built and run by tests/test_host_logic.py with g++ (no GPU needed).
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "fft_tile.cuh"

using namespace ptx;

template <class P>
static int check(bool word32 = false) {
  constexpr int L = P::L;
  using G = TileGeom<P>;
  using TL = TwLayout<P>;
  using C = Cross<P>;
  constexpr int N = P::N, NT = P::NT, E = P::E, RC = P::RC, NX = P::NX, NY = P::NY;
  static_assert(NX * NY == NT * E, "local tile = NT x E");
  static_assert(NX == N && NY * RC == N, "local tile rows x cross radix = N");
  std::vector<float2> tw(TL::TOTAL + 1);
  fill_twiddles<P>(tw.data());
  std::vector<float2> in(N * N), tile(G::WORDS), scratch(N * N);
  srand(1234 + L);
  for (auto& c : in) c = make_float2(rand() / (float)RAND_MAX - 0.5f, rand() / (float)RAND_MAX - 0.5f);
  // reference: separable DFT in double
  std::vector<double> wr(N), wi(N);
  for (int k = 0; k < N; ++k) {
    wr[k] = cos(-2.0 * M_PI * k / N);
    wi[k] = sin(-2.0 * M_PI * k / N);
  }
  std::vector<double> rr(N * N), ri(N * N), tr(N * N), ti(N * N);
  for (int y = 0; y < N; ++y)
    for (int k = 0; k < N; ++k) {
      double sr = 0, si = 0;
      for (int x = 0; x < N; ++x) {
        const int a = (x * k) % N;
        sr += in[y * N + x].x * wr[a] - in[y * N + x].y * wi[a];
        si += in[y * N + x].x * wi[a] + in[y * N + x].y * wr[a];
      }
      tr[y * N + k] = sr;
      ti[y * N + k] = si;
    }
  for (int x = 0; x < N; ++x)
    for (int k = 0; k < N; ++k) {
      double sr = 0, si = 0;
      for (int y = 0; y < N; ++y) {
        const int a = (y * k) % N;
        sr += tr[y * N + x] * wr[a] - ti[y * N + x] * wi[a];
        si += tr[y * N + x] * wi[a] + ti[y * N + x] * wr[a];
      }
      rr[k * N + x] = sr;
      ri[k * N + x] = si;
    }

  std::vector<char> seen(N * N, 0);
  // ---- forward cross stage (natural ownership of column blocks) -> scratch frame
  if (RC > 1) {
    for (int c = 0; c < RC; ++c)
      for (int t = 0; t < NT; ++t) {
        float2 v[E];
        for (int b = 0; b < C::NB; ++b) {
          int ylow, xc;
          C::pair(t, b, ylow, xc);
          for (int j = 0; j < RC; ++j) {
            v[j + RC * b] = in[(j * NY + ylow) * N + c * C::CW + xc];
            seen[(j * NY + ylow) * N + c * C::CW + xc]++;
          }
        }
        cross_compute<P, false>(v, t, tw.data() + TL::CROSS);
        for (int b = 0; b < C::NB; ++b) {
          int ylow, xc;
          C::pair(t, b, ylow, xc);
          for (int j = 0; j < RC; ++j) scratch[scratch_index<P>(j, ylow, c * C::CW + xc)] = v[j + RC * b];
        }
      }
    for (int i = 0; i < N * N; ++i)
      if (seen[i] != 1) {
        printf("L=%d cross coverage error at %d: %d\n", L, i, seen[i]);
        return 1;
      }
  }
  std::vector<std::vector<float2>> regs((size_t)NT * RC, std::vector<float2>(E));
  std::vector<char> seen2(N * N, 0);
  double num = 0, den = 0;
  for (int k1 = 0; k1 < RC; ++k1) {
    const float2* src = RC > 1 ? scratch.data() + (size_t)k1 * NX * NY : in.data();
    std::fill(seen.begin(), seen.end(), 0);
    for (int t = 0; t < NT; ++t) {  // stage 0: input straight from "global" (gather or scratch)
      using ST = typename P::S0;
      int xf, yf;
      fixed_coords<ST, P::WBITS>(t, xf, yf);
      float2 v[E];
      for (int e = 0; e < E; ++e) {
        int dx, dy;
        elem_offset<ST>(e, dx, dy);
        v[e] = src[(yf | dy) * NX + (xf | dx)];
        seen[(yf | dy) * NX + (xf | dx)]++;
      }
      stage_compute<ST, false>(v, xf, yf, tw.data() + TL::X0, tw.data() + TL::Y0);
      stage_store<ST, P>(v, tile.data(), xf, yf);
    }
    for (int i = 0; i < NX * NY; ++i)
      if (seen[i] != 1) {
        printf("L=%d stage0 coverage error at %d: %d\n", L, i, seen[i]);
        return 1;
      }
    for (int t = 0; t < NT; ++t) {
      using ST = typename P::S1;
      int xf, yf;
      fixed_coords<ST, P::WBITS>(t, xf, yf);
      float2 v[E];
      stage_load<ST, P>(v, tile.data(), xf, yf);
      stage_compute<ST, false>(v, xf, yf, tw.data() + TL::X1, tw.data() + TL::Y1);
      stage_store<ST, P>(v, tile.data(), xf, yf);
    }
    for (int t = 0; t < NT; ++t) {
      using ST = typename P::S2;
      int xf, yf;
      fixed_coords<ST, P::WBITS>(t, xf, yf);
      float2 v[E];
      stage_load<ST, P>(v, tile.data(), xf, yf);
      stage_compute<ST, false>(v, xf, yf, tw.data() + TL::X2, tw.data() + TL::Y2);
      for (int e = 0; e < E; ++e) {
        int dx, dy;
        elem_offset<ST>(e, dx, dy);
        const int kx = pos_to_freq_x<P>(xf | dx), ky = k1 + RC * pos_to_freq_y<P>(yf | dy);
        seen2[ky * N + kx]++;
        double er = v[e].x - rr[ky * N + kx], ei = v[e].y - ri[ky * N + kx];
        num += er * er + ei * ei;
        den += rr[ky * N + kx] * rr[ky * N + kx] + ri[ky * N + kx] * ri[ky * N + kx];
        regs[(size_t)k1 * NT + t][e] = v[e];
      }
    }
  }
  for (int i = 0; i < N * N; ++i)
    if (seen2[i] != 1) {
      printf("L=%d spectrum coverage error at %d: %d\n", L, i, seen2[i]);
      return 1;
    }
  const double efwd = sqrt(num / den);
  // ---- inverse, starting from the registers of the last forward stage
  std::vector<float2> out(N * N);
  for (int k1 = 0; k1 < RC; ++k1) {
    for (int t = 0; t < NT; ++t) {
      using ST = typename P::S2;
      int xf, yf;
      fixed_coords<ST, P::WBITS>(t, xf, yf);
      float2 v[E];
      for (int e = 0; e < E; ++e) v[e] = regs[(size_t)k1 * NT + t][e];
      stage_compute<ST, true>(v, xf, yf, tw.data() + TL::X2, tw.data() + TL::Y2);
      stage_store<ST, P>(v, tile.data(), xf, yf);
    }
    for (int t = 0; t < NT; ++t) {
      using ST = typename P::S1;
      int xf, yf;
      fixed_coords<ST, P::WBITS>(t, xf, yf);
      float2 v[E];
      stage_load<ST, P>(v, tile.data(), xf, yf);
      stage_compute<ST, true>(v, xf, yf, tw.data() + TL::X1, tw.data() + TL::Y1);
      stage_store<ST, P>(v, tile.data(), xf, yf);
    }
    float2* dst = RC > 1 ? scratch.data() + (size_t)k1 * NX * NY : out.data();
    for (int t = 0; t < NT; ++t) {
      using ST = typename P::S0;
      int xf, yf;
      fixed_coords<ST, P::WBITS>(t, xf, yf);
      float2 v[E];
      stage_load<ST, P>(v, tile.data(), xf, yf);
      stage_compute<ST, true>(v, xf, yf, tw.data() + TL::X0, tw.data() + TL::Y0);
      for (int e = 0; e < E; ++e) {
        int dx, dy;
        elem_offset<ST>(e, dx, dy);
        dst[(yf | dy) * NX + (xf | dx)] = v[e];
      }
    }
  }
  if (RC > 1) {
    for (int c = 0; c < RC; ++c)
      for (int t = 0; t < NT; ++t) {
        float2 v[E];
        for (int b = 0; b < C::NB; ++b) {
          int ylow, xc;
          C::pair(t, b, ylow, xc);
          for (int j = 0; j < RC; ++j) v[j + RC * b] = scratch[scratch_index<P>(j, ylow, c * C::CW + xc)];
        }
        cross_compute<P, true>(v, t, tw.data() + TL::CROSS);
        for (int b = 0; b < C::NB; ++b) {
          int ylow, xc;
          C::pair(t, b, ylow, xc);
          for (int j = 0; j < RC; ++j) out[(j * NY + ylow) * N + c * C::CW + xc] = v[j + RC * b];
        }
      }
  }
  num = den = 0;
  for (int i = 0; i < N * N; ++i) {
    double er = out[i].x / (double)(N * N) - in[i].x, ei = out[i].y / (double)(N * N) - in[i].y;
    num += er * er + ei * ei;
    den += in[i].x * in[i].x + in[i].y * in[i].y;
  }
  const double einv = sqrt(num / den);
  // ---- bank-conflict audit: 64-bit accesses, 16 lanes per wavefront, bank pair = idx mod 16
  // (word32: the split-plane exchange of the pipelined kernel -- one float per element, 32 lanes per
  // wavefront, bank = idx mod 32)
  int worst = 1;
  const int LW = word32 ? 32 : 16;
  auto audit = [&](auto st_tag) {
    using ST = decltype(st_tag);
    for (int w0 = 0; w0 < NT; w0 += LW)
      for (int e = 0; e < E; ++e) {
        int cnt[32] = {0};
        for (int l = 0; l < LW; ++l) {
          int xf, yf, dx, dy;
          fixed_coords<ST, P::WBITS>(w0 + l, xf, yf);
          elem_offset<ST>(e, dx, dy);
          cnt[G::idx(yf | dy, xf | dx) & (LW - 1)]++;
        }
        for (int b = 0; b < LW; ++b) worst = cnt[b] > worst ? cnt[b] : worst;
      }
  };
  audit(typename P::S0{});
  audit(typename P::S1{});
  audit(typename P::S2{});
  // ---- S1 <-> S2 exchange audit: every tile position is written and read inside one group of
  // P::XG2 consecutive threads (0 = whole CTA), which is what lets fft_forward / fft_inverse use a
  // group barrier there instead of a block barrier
  int closed = 1;
  if (P::XG2 > 0) {
    std::vector<int> owner1(G::WORDS, -1);
    for (int w = 0; w < NT; ++w)
      for (int e = 0; e < E; ++e) {
        int xf, yf, dx, dy;
        fixed_coords<typename P::S1, P::WBITS>(w, xf, yf);
        elem_offset<typename P::S1>(e, dx, dy);
        owner1[G::idx(yf | dy, xf | dx)] = w / P::XG2;
      }
    for (int w = 0; w < NT; ++w)
      for (int e = 0; e < E; ++e) {
        int xf, yf, dx, dy;
        fixed_coords<typename P::S2, P::WBITS>(w, xf, yf);
        elem_offset<typename P::S2>(e, dx, dy);
        if (owner1[G::idx(yf | dy, xf | dx)] != w / P::XG2) closed = 0;
      }
    if (!closed) printf("L=%d S1<->S2 exchange is NOT closed within groups of %d threads\n", L, P::XG2);
  }
  // ---- spectrum coalescing audit: the 32 lanes of a warp must own 32 consecutive kx of one ky
  int coalesced = 1;
  for (int w0 = 0; w0 < NT; w0 += 32)
    for (int e = 0; e < E; ++e) {
      int lo = 1 << 30, hi = -1;
      for (int l = 0; l < 32; ++l) {
        int xf, yf, dx, dy;
        fixed_coords<typename P::S2, P::WBITS>(w0 + l, xf, yf);
        elem_offset<typename P::S2>(e, dx, dy);
        const int k = pos_to_freq_y<P>(yf | dy) * N + pos_to_freq_x<P>(xf | dx);
        lo = k < lo ? k : lo;
        hi = k > hi ? k : hi;
      }
      if (hi - lo > 63) coalesced = 0;  // N = 64 plan: two 16-runs of adjacent rows are accepted
    }
  printf("L=%d%s N=%d fwd_rel_l2=%.3e inv_rel_l2=%.3e worst_bank_conflict=%d spectrum_coalesced=%d\n", L,
         word32 ? "P" : "", N, efwd, einv, worst, coalesced);
  return (efwd < 2e-6 && einv < 2e-6 && closed && (!word32 || (worst == 1 && coalesced))) ? 0 : 1;
}

int main() {
  int rc = 0;
  rc |= check<Plan<6>>();
  rc |= check<Plan<7>>();
  rc |= check<Plan<8>>();
  rc |= check<Plan<9>>();
  rc |= check<Plan7P>(true);  // geometry of the pipelined 128^2 kernel (ptycho_pipe.cuh)
  printf(rc ? "EMU FAILED\n" : "EMU OK\n");
  return rc;
}
