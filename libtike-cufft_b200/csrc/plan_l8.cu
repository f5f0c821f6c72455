// Kernel family for detector size 2^8, part 1 of 3: operators, intensity pass, and the table itself.
#include "ptycho_table.cuh"

namespace ptx {
void fill_grad_l8(PlanOps& ops);    // plan_l8_grad.cu
void fill_search_l8(PlanOps& ops);  // plan_l8_search.cu
const PlanOps* ops_l8() {
  static PlanOps ops;
  static bool init = false;
  if (!init) {
    fill_ops_base<Plan<8>>(ops);
    fill_grad_l8(ops);
    fill_search_l8(ops);
    init = true;
  }
  return &ops;
}
}  // namespace ptx
