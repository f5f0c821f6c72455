// Kernel family for detector size 2^7, part 1 of 3: operators, intensity pass, and the table itself.
#include "ptycho_table.cuh"

namespace ptx {
void fill_grad_l7(PlanOps& ops);    // plan_l7_grad.cu
void fill_search_l7(PlanOps& ops);  // plan_l7_search.cu
void fill_pipe_l7(PlanOps& ops);    // plan_l7_pipe.cu
const PlanOps* ops_l7() {
  static PlanOps ops;
  static bool init = false;
  if (!init) {
    fill_ops_base<Plan<7>>(ops);
    fill_grad_l7(ops);
    fill_search_l7(ops);
    fill_pipe_l7(ops);
    init = true;
  }
  return &ops;
}
}  // namespace ptx
