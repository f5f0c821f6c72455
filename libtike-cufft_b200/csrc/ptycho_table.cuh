// Dispatch table of one detector-size kernel family, filled from three translation units per size
// (plan_lN.cu: operators + intensity, plan_lN_grad.cu: the fused gradient kernels,
// plan_lN_search.cu: line search + position correction) so that the family compiles in parallel.
#pragma once

#include "ptycho_register.cuh"

namespace ptx {

#define PTX_SET(id, ...)                                                                            \
  ops.kernels[id] = (const void*)(void (*)(const PassArgs, const CUtensorMap, const CUtensorMap))(__VA_ARGS__); \
  ops.names[id] = #__VA_ARGS__;

template <class P>
static void fill_tw_host(float2* tw) {
  fill_twiddles<P>(tw);
}

template <class P>
void fill_ops_base(PlanOps& ops) {
  ops.L = P::L;
  ops.N = P::N;
  ops.NT = P::NT;
  ops.RC = P::RC;
  ops.smem_bytes = Smem<P>::BYTES;
  ops.smem_bytes_nodata = Smem<P>::BYTES_NODATA;
  ops.smem_bytes_reg = RegGeom<P>::SMEM;
  ops.frame_per_cta = Scratch<P>::FRAME;
  ops.stash_per_cta = Scratch<P>::STASH;
  ops.accp_per_cta = Scratch<P>::ACCP;
  ops.slots_per_cta = Scratch<P>::SLOTS;
  ops.tw_total = TwLayout<P>::TOTAL;
  ops.fill_tw = fill_tw_host<P>;
  ops.NT_pipe = 0;
  ops.smem_bytes_pipe = 0;
  for (int k = K_PIPE_GAUSS; k <= K_PIPEMC_POIS; ++k) {
    ops.kernels[k] = nullptr;
    ops.names[k] = "(none)";
  }
  ops.patch_w = Patch<P>::TMA ? Patch<P>::W : 0;
  ops.patch_h = Patch<P>::TMA ? Patch<P>::H : 0;
  PTX_SET(K_FWD, k_fwd<P>)
  PTX_SET(K_NEAR, k_nearplane<P>)
  PTX_SET(K_ADJ_OBJ, k_adj<P, 0>)
  PTX_SET(K_ADJ_PRB, k_adj<P, 1>)
  PTX_SET(K_INT_GAUSS, k_intensity<P, 0>)
  PTX_SET(K_INT_POIS, k_intensity<P, 1>)
}

template <class P>
void fill_ops_grad(PlanOps& ops) {
  PTX_SET(K_GRAD_GAUSS_OBJ, k_grad<P, 0, 0, false>)
  PTX_SET(K_GRAD_GAUSS_PRB, k_grad<P, 0, 1, false>)
  PTX_SET(K_GRAD_POIS_OBJ, k_grad<P, 1, 0, false>)
  PTX_SET(K_GRAD_POIS_PRB, k_grad<P, 1, 1, false>)
  PTX_SET(K_GRADC_GAUSS_OBJ, k_grad<P, 0, 0, true>)
  PTX_SET(K_GRADC_GAUSS_PRB, k_grad<P, 0, 1, true>)
  PTX_SET(K_GRADC_POIS_OBJ, k_grad<P, 1, 0, true>)
  PTX_SET(K_GRADC_POIS_PRB, k_grad<P, 1, 1, true>)
}

template <class P>
void fill_ops_search(PlanOps& ops) {
  PTX_SET(K_LS_GAUSS, k_linesearch<P, 0, false, false>)
  PTX_SET(K_LS_POIS, k_linesearch<P, 1, false, false>)
  PTX_SET(K_LSAB_GAUSS, k_linesearch<P, 0, true, false>)
  PTX_SET(K_LSAB_POIS, k_linesearch<P, 1, true, false>)
  PTX_SET(K_LSC_GAUSS, k_linesearch<P, 0, false, true>)
  PTX_SET(K_LSC_POIS, k_linesearch<P, 1, false, true>)
  PTX_SET(K_LSCAB_GAUSS, k_linesearch<P, 0, true, true>)
  PTX_SET(K_LSCAB_POIS, k_linesearch<P, 1, true, true>)
  PTX_SET(K_REG_OBJ, k_register<P, 0>)
  PTX_SET(K_REG_FOURIER, k_register<P, 1>)
  PTX_SET(K_REG_REAL, k_register<P, 2>)
}

}  // namespace ptx
