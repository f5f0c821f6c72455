// Kernel family for detector size 2^9, part 1 of 3: operators, intensity pass, and the table itself.
#include "ptycho_table.cuh"

namespace ptx {
void fill_grad_l9(PlanOps& ops);    // plan_l9_grad.cu
void fill_search_l9(PlanOps& ops);  // plan_l9_search.cu
const PlanOps* ops_l9() {
  static PlanOps ops;
  static bool init = false;
  if (!init) {
    fill_ops_base<Plan<9>>(ops);
    fill_grad_l9(ops);
    fill_search_l9(ops);
    init = true;
  }
  return &ops;
}
}  // namespace ptx
