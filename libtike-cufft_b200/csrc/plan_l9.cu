// Kernel family for detector size 2^9 (see ptycho_passes.cuh); one translation unit per size.
#include "ptycho_register.cuh"

namespace ptx {
const PlanOps* ops_l9() { return make_ops<Plan<9>>(); }
}  // namespace ptx
