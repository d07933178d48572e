// gpt_rooms_k5.cu — instantiations of rooms_step_kernel (see gpt_rooms_kernel.cuh)
#include "gpt_rooms_kernel.cuh"

namespace gpt {

void* rooms_pick_grid_any(bool rgoal, bool replay, int variant) { return pick_rr<GPT_OBS_GRID, 0>(rgoal, replay, variant); }

}  // namespace gpt
