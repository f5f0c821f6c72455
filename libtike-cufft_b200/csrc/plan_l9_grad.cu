// Kernel family for detector size 2^9, part 2 of 3: the fused gradient kernels (ptycho_passes.cuh).
#include "ptycho_table.cuh"

namespace ptx {
void fill_grad_l9(PlanOps& ops) { fill_ops_grad<Plan<9>>(ops); }
}  // namespace ptx
