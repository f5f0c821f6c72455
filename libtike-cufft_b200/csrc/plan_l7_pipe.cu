// Kernel family for detector size 2^7, part 4: the warp-specialised, pipelined object-gradient kernels.
#include "ptycho_pipe.cuh"
#include "ptycho_table.cuh"

namespace ptx {
void fill_pipe_l7(PlanOps& ops) {
  ops.NT_pipe = Pipe::NTHREADS;
  ops.smem_bytes_pipe = Pipe::BYTES;
  PTX_SET(K_PIPE_GAUSS, k_grad_pipe<0, false, false>)
  PTX_SET(K_PIPEC_GAUSS, k_grad_pipe<0, true, false>)
  PTX_SET(K_PIPE_POIS, k_grad_pipe<1, false, false>)
  PTX_SET(K_PIPEC_POIS, k_grad_pipe<1, true, false>)
  PTX_SET(K_PIPEM_GAUSS, k_grad_pipe<0, false, true>)
  PTX_SET(K_PIPEMC_GAUSS, k_grad_pipe<0, true, true>)
  PTX_SET(K_PIPEM_POIS, k_grad_pipe<1, false, true>)
  PTX_SET(K_PIPEMC_POIS, k_grad_pipe<1, true, true>)
}
}  // namespace ptx
