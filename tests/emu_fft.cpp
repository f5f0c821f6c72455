// CPU emulation of the CTA-wide tile FFT (libtike-cufft_b200/csrc/fft_tile.cuh).
// Runs every "thread" of a CTA sequentially, stage by stage (a stage boundary == __syncthreads),
// and checks forward and inverse transforms against a double-precision separable DFT.
// Built and run by tests/test_host_logic.py with g++ (no GPU needed).
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "fft_tile.cuh"

using namespace ptx;

template <int L>
static int check() {
  using P = Plan<L>;
  using G = TileGeom<L>;
  using TL = TwLayout<P>;
  constexpr int N = P::N, NT = P::NT, E = P::E;
  std::vector<float2> tw(TL::TOTAL + 1);
  fill_twiddles<P>(tw.data());
  std::vector<float2> in(N * N), tile(G::WORDS);
  srand(1234 + L);
  for (auto& c : in) c = make_float2(rand() / (float)RAND_MAX - 0.5f, rand() / (float)RAND_MAX - 0.5f);
  // reference: separable DFT in double
  std::vector<double> rr(N * N), ri(N * N), tr(N * N), ti(N * N);
  for (int y = 0; y < N; ++y)
    for (int k = 0; k < N; ++k) {
      double sr = 0, si = 0;
      for (int x = 0; x < N; ++x) {
        double a = -2.0 * M_PI * (double)((x * k) % N) / N;
        sr += in[y * N + x].x * cos(a) - in[y * N + x].y * sin(a);
        si += in[y * N + x].x * sin(a) + in[y * N + x].y * cos(a);
      }
      tr[y * N + k] = sr;
      ti[y * N + k] = si;
    }
  for (int x = 0; x < N; ++x)
    for (int k = 0; k < N; ++k) {
      double sr = 0, si = 0;
      for (int y = 0; y < N; ++y) {
        double a = -2.0 * M_PI * (double)((y * k) % N) / N;
        sr += tr[y * N + x] * cos(a) - ti[y * N + x] * sin(a);
        si += tr[y * N + x] * sin(a) + ti[y * N + x] * cos(a);
      }
      rr[k * N + x] = sr;
      ri[k * N + x] = si;
    }

  std::vector<std::vector<float2>> regs(NT, std::vector<float2>(E));
  std::vector<char> seen(N * N, 0);
  // ---- forward
  for (int t = 0; t < NT; ++t) {  // stage 0: input straight from "global" (the gather)
    using ST = typename P::S0;
    int xf, yf;
    fixed_coords<ST, P::WBITS>(t, xf, yf);
    float2 v[E];
    for (int e = 0; e < E; ++e) {
      int dx, dy;
      elem_offset<ST>(e, dx, dy);
      v[e] = in[(yf | dy) * N + (xf | dx)];
      seen[(yf | dy) * N + (xf | dx)]++;
    }
    stage_compute<ST, false>(v, xf, yf, tw.data() + TL::X0, tw.data() + TL::Y0);
    stage_store<ST, L>(v, tile.data(), xf, yf);
  }
  for (int i = 0; i < N * N; ++i)
    if (seen[i] != 1) {
      printf("L=%d stage0 coverage error at %d: %d\n", L, i, seen[i]);
      return 1;
    }
  for (int t = 0; t < NT; ++t) {
    using ST = typename P::S1;
    int xf, yf;
    fixed_coords<ST, P::WBITS>(t, xf, yf);
    float2 v[E];
    stage_load<ST, L>(v, tile.data(), xf, yf);
    stage_compute<ST, false>(v, xf, yf, tw.data() + TL::X1, tw.data() + TL::Y1);
    stage_store<ST, L>(v, tile.data(), xf, yf);
  }
  double num = 0, den = 0;
  std::fill(seen.begin(), seen.end(), 0);
  for (int t = 0; t < NT; ++t) {
    using ST = typename P::S2;
    int xf, yf;
    fixed_coords<ST, P::WBITS>(t, xf, yf);
    float2 v[E];
    stage_load<ST, L>(v, tile.data(), xf, yf);
    stage_compute<ST, false>(v, xf, yf, tw.data() + TL::X2, tw.data() + TL::Y2);
    for (int e = 0; e < E; ++e) {
      int dx, dy;
      elem_offset<ST>(e, dx, dy);
      const int kx = pos_to_freq_x<P>(xf | dx), ky = pos_to_freq_y<P>(yf | dy);
      seen[ky * N + kx]++;
      double er = v[e].x - rr[ky * N + kx], ei = v[e].y - ri[ky * N + kx];
      num += er * er + ei * ei;
      den += rr[ky * N + kx] * rr[ky * N + kx] + ri[ky * N + kx] * ri[ky * N + kx];
      regs[t][e] = v[e];
    }
  }
  for (int i = 0; i < N * N; ++i)
    if (seen[i] != 1) {
      printf("L=%d spectrum coverage error at %d: %d\n", L, i, seen[i]);
      return 1;
    }
  const double efwd = sqrt(num / den);
  // ---- inverse, starting from the registers of the last forward stage
  for (int t = 0; t < NT; ++t) {
    using ST = typename P::S2;
    int xf, yf;
    fixed_coords<ST, P::WBITS>(t, xf, yf);
    float2 v[E];
    for (int e = 0; e < E; ++e) v[e] = regs[t][e];
    stage_compute<ST, true>(v, xf, yf, tw.data() + TL::X2, tw.data() + TL::Y2);
    stage_store<ST, L>(v, tile.data(), xf, yf);
  }
  for (int t = 0; t < NT; ++t) {
    using ST = typename P::S1;
    int xf, yf;
    fixed_coords<ST, P::WBITS>(t, xf, yf);
    float2 v[E];
    stage_load<ST, L>(v, tile.data(), xf, yf);
    stage_compute<ST, true>(v, xf, yf, tw.data() + TL::X1, tw.data() + TL::Y1);
    stage_store<ST, L>(v, tile.data(), xf, yf);
  }
  num = den = 0;
  for (int t = 0; t < NT; ++t) {
    using ST = typename P::S0;
    int xf, yf;
    fixed_coords<ST, P::WBITS>(t, xf, yf);
    float2 v[E];
    stage_load<ST, L>(v, tile.data(), xf, yf);
    stage_compute<ST, true>(v, xf, yf, tw.data() + TL::X0, tw.data() + TL::Y0);
    for (int e = 0; e < E; ++e) {
      int dx, dy;
      elem_offset<ST>(e, dx, dy);
      const float2 want = in[(yf | dy) * N + (xf | dx)];
      double er = v[e].x / (double)(N * N) - want.x, ei = v[e].y / (double)(N * N) - want.y;
      num += er * er + ei * ei;
      den += want.x * want.x + want.y * want.y;
    }
  }
  const double einv = sqrt(num / den);
  // ---- bank-conflict audit: 64-bit accesses, 16 lanes per wavefront, bank pair = idx mod 16
  int worst = 1;
  auto audit = [&](auto st_tag) {
    using ST = decltype(st_tag);
    for (int w0 = 0; w0 < NT; w0 += 16)
      for (int e = 0; e < E; ++e) {
        int cnt[16] = {0};
        for (int l = 0; l < 16; ++l) {
          int xf, yf, dx, dy;
          fixed_coords<ST, P::WBITS>(w0 + l, xf, yf);
          elem_offset<ST>(e, dx, dy);
          cnt[G::idx(yf | dy, xf | dx) & 15]++;
        }
        for (int b = 0; b < 16; ++b) worst = cnt[b] > worst ? cnt[b] : worst;
      }
  };
  audit(typename P::S0{});
  audit(typename P::S1{});
  audit(typename P::S2{});
  printf("L=%d N=%d fwd_rel_l2=%.3e inv_rel_l2=%.3e worst_bank_conflict=%d\n", L, N, efwd, einv, worst);
  return (efwd < 2e-6 && einv < 2e-6) ? 0 : 1;
}

int main() {
  int rc = 0;
  rc |= check<6>();
  rc |= check<7>();
  printf(rc ? "EMU FAILED\n" : "EMU OK\n");
  return rc;
}
