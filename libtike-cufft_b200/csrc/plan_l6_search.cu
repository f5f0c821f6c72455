// Kernel family for detector size 2^6, part 3 of 3: line search and position correction.
#include "ptycho_table.cuh"

namespace ptx {
void fill_search_l6(PlanOps& ops) { fill_ops_search<Plan<6>>(ops); }
}  // namespace ptx
