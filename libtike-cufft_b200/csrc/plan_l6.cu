// Kernel family for detector size 2^6, part 1 of 3: operators, intensity pass, and the table itself.
#include "ptycho_table.cuh"

namespace ptx {
void fill_grad_l6(PlanOps& ops);    // plan_l6_grad.cu
void fill_search_l6(PlanOps& ops);  // plan_l6_search.cu
const PlanOps* ops_l6() {
  static PlanOps ops;
  static bool init = false;
  if (!init) {
    fill_ops_base<Plan<6>>(ops);
    fill_grad_l6(ops);
    fill_search_l6(ops);
    init = true;
  }
  return &ops;
}
}  // namespace ptx
