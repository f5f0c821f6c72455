// Kernel family for detector size 2^8 (see ptycho_passes.cuh); one translation unit per size.
#include "ptycho_register.cuh"

namespace ptx {
const PlanOps* ops_l8() { return make_ops<Plan<8>>(); }
}  // namespace ptx
