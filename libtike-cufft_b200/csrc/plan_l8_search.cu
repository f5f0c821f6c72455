// Kernel family for detector size 2^8, part 3 of 3: line search and position correction.
#include "ptycho_table.cuh"

namespace ptx {
void fill_search_l8(PlanOps& ops) { fill_ops_search<Plan<8>>(ops); }
}  // namespace ptx
