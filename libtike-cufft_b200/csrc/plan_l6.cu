// Kernel family for detector size 2^6 (see ptycho_passes.cuh); one translation unit per size.
#include "ptycho_register.cuh"

namespace ptx {
const PlanOps* ops_l6() { return make_ops<Plan<6>>(); }
}  // namespace ptx
