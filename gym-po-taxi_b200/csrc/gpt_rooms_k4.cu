// gpt_rooms_k4.cu — instantiations of rooms_step_kernel (see gpt_rooms_kernel.cuh)
#include "gpt_rooms_kernel.cuh"

namespace gpt {

void* rooms_pick_grid_large(int n, bool rgoal, bool replay, int variant) {
  return n == 7 ? pick_rr<GPT_OBS_GRID, 7>(rgoal, replay, variant) : pick_rr<GPT_OBS_GRID, 9>(rgoal, replay, variant);
}

}  // namespace gpt
