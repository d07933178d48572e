// gpt_rooms_k3.cu — instantiations of rooms_step_kernel (see gpt_rooms_kernel.cuh)
#include "gpt_rooms_kernel.cuh"

namespace gpt {

void* rooms_pick_grid_small(int n, bool rgoal, bool replay, int variant) {
  return n == 3 ? pick_rr<GPT_OBS_GRID, 3>(rgoal, replay, variant) : pick_rr<GPT_OBS_GRID, 5>(rgoal, replay, variant);
}

}  // namespace gpt
