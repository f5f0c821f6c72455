// Host-side glue between the C ABI (ptycho_api.cu) and the per-detector-size kernel families
// (plan_l6.cu ... plan_l9.cu, one translation unit each so that they compile in parallel).
#pragma once

#include <cuda.h>  // CUtensorMap (type only: the driver entry point is looked up at run time)
#include <cuda_runtime.h>
#include <stddef.h>

#include "ptycho_device.cuh"

namespace ptx {

// position correction (ptycho_register.cuh)
constexpr int REG_UMAX = 150;   // largest upsampled window: ceil(1.5 * 100)
constexpr int REG_EROWS = 256;
constexpr int REG_NQ = 24;      // terms of the Jacobi-Anger expansion of the window kernel (ptycho_register.cuh)
constexpr int REG_JROWS = 160;  // rows of the Chebyshev table T[j][n] (zero beyond U)  // rows of the E table (zero beyond U: padded chunks contribute |G| = 0)

// kernel parameter block, shared by every pass
struct PassArgs {
  Geo g;
  const float2* tw;        // twiddle tables (global, copied to smem by every CTA)
  // per-CTA global scratch, one CONTIGUOUS array per kind ([grid][...]; a CTA's slice is indexed by
  // blockIdx.x).  A plan may therefore only be used from one stream at a time.
  float2* frame;           // [grid][N*N]   staging frame of the N > 128 plans (kept L2-resident)
  float2* stash;           // [grid][N*N]   parked far field / probe accumulators (allocated on first use)
  float* accp;             // [grid][3*N*N] intensity / p1,p2,p3 accumulators, registration product
  double* slots;           // [grid][16*NT] thread-private running sums
  const float2* psi;       // [T,nz,n]
  const float2* psi_b;     // second object (line search)
  const float2* prb;       // probe base of the mode to use, angle stride prb_ts
  const float2* prb_b;
  size_t prb_ts, prb_b_ts;  // complex elements between angles
  size_t prb_ms, prb_b_ms;  // complex elements between modes (pair loop)
  const float2* scan;       // [T,S]
  const float* data;        // [T,S,N,N]
  const float* inten_in;    // [T,S,N,N] or null
  float* inten_out;         // [T,S,N,N] or null
  float2* far;              // [T,S,N,N] (k_fwd output; k_grad: optional far-field cache, pre-residual)
  const float2* far_in;     // k_adj input; k_linesearch: optional cached first far field of every pair
  size_t far_ms;            // complex elements between modes (pairs) of the cache
  float2* p23;              // k_linesearch: optional [T,S,N,N] output of (p2, p3) per pixel
  float2* grad;  // object gradient [T,nz,n] or probe gradient base (angle stride grad_ts)
  size_t grad_ts;
  const float* sc;  // device scalars
  double* red;
  int nmodes, npairs, c0, ncand;
  int use_tma;  // object patches arrive by TMA tensor copies (tensor maps are valid)
  int strip;    // single-tile plans: column-strip gather through the shared tile (ptycho_device.cuh)
  // position correction (ptycho_register.cuh)
  const double2* reg_E;  // [REG_EROWS][N] table W^(j k), W = exp(2 pi i / (uf N))
  double* reg_out;       // [npat][2] shifts (row, col)
  int reg_U, reg_uf;     // upsampled window size ceil(1.5 uf), upsampling factor
  const double* reg_A;   // [REG_NQ][N] eps_n J_n(theta_k)
  const double* reg_T;   // [REG_JROWS][REG_NQ] T_n(x_j)
  int reg_algo;          // 2: low-rank (Jacobi-Anger) on DMMA; 1: direct products on DMMA; 0: on DFMA
};

enum KernelId {
  K_FWD = 0, K_NEAR, K_ADJ_OBJ, K_ADJ_PRB, K_INT_GAUSS, K_INT_POIS,
  K_GRAD_GAUSS_OBJ, K_GRAD_GAUSS_PRB, K_GRAD_POIS_OBJ, K_GRAD_POIS_PRB, K_LS_GAUSS, K_LS_POIS,
  K_REG_OBJ, K_REG_FOURIER, K_REG_REAL,
  K_GRADC_GAUSS_OBJ, K_GRADC_GAUSS_PRB, K_GRADC_POIS_OBJ, K_GRADC_POIS_PRB,  // + far-field cache output
  K_LSAB_GAUSS, K_LSAB_POIS,  // + the next iteration's a, b sums of every candidate
  K_LSC_GAUSS, K_LSC_POIS, K_LSCAB_GAUSS, K_LSCAB_POIS,  // first far field read from the gradient pass's cache
  // warp-specialised, pipelined object-gradient kernels (ptycho_pipe.cuh; 128^2 plan only, else null)
  K_PIPE_GAUSS, K_PIPEC_GAUSS, K_PIPE_POIS, K_PIPEC_POIS,      // I = |F|^2 (one mode)
  K_PIPEM_GAUSS, K_PIPEMC_GAUSS, K_PIPEM_POIS, K_PIPEMC_POIS,  // I = the summed intensity map (several modes)
  K_COUNT
};

struct PlanOps {
  int L, N, NT, RC;
  size_t smem_bytes;       // dynamic shared memory of the kernels that read measured data
  size_t smem_bytes_nodata;  // ... of the others (no data tile: more of the SM's SRAM stays L1)
  size_t smem_bytes_reg;     // ... of the position-correction kernels
  int NT_pipe;               // threads per CTA of the pipelined kernels (0: none for this plan)
  size_t smem_bytes_pipe;    // ... and their dynamic shared memory
  size_t frame_per_cta, stash_per_cta;  // float2 (frame: 0 for single-tile plans)
  size_t accp_per_cta;     // floats
  size_t slots_per_cta;    // doubles
  int tw_total;            // float2
  void (*fill_tw)(float2*);
  int patch_w, patch_h;    // TMA box of the object patch in complex elements (0: no TMA gather)
  const void* kernels[K_COUNT];
  const char* names[K_COUNT];
};

const PlanOps* ops_l6();
const PlanOps* ops_l7();
const PlanOps* ops_l8();
const PlanOps* ops_l9();

}  // namespace ptx
