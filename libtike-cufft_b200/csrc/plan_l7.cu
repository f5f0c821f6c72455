// Kernel family for detector size 2^7 (see ptycho_passes.cuh); one translation unit per size.
#include "ptycho_register.cuh"

namespace ptx {
const PlanOps* ops_l7() { return make_ops<Plan<7>>(); }
}  // namespace ptx
