// TEST SCAFFOLDING - stand-in for <cooperative_groups.h> (see cuda_runtime.h in this directory): a grid of ONE block,
// where grid.sync() is the block barrier.  The harness launches cooperative kernels that way itself.
#ifndef COCONS_TEST_COOP_GROUPS_EMUL_H
#define COCONS_TEST_COOP_GROUPS_EMUL_H
#include "cuda_runtime.h"
namespace cooperative_groups {
struct grid_group {
  void sync() const {
    if (gridDim.x * gridDim.y * gridDim.z != 1) {
      std::fprintf(stderr, "host emulation: grid.sync() needs a grid of one block\n");
      std::abort();
    }
    __syncthreads();
  }
};
inline grid_group this_grid() { return grid_group(); }
}  // namespace cooperative_groups
#endif
