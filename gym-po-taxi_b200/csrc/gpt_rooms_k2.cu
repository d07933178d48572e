// gpt_rooms_k2.cu — instantiations of rooms_step_kernel (see gpt_rooms_kernel.cuh)
#include "gpt_rooms_kernel.cuh"

namespace gpt {

void* rooms_pick_vhansen(int obs, bool rgoal, bool replay, int variant) {
  return obs == GPT_OBS_VEC_HANSEN ? pick_rr<GPT_OBS_VEC_HANSEN, 0>(rgoal, replay, variant) : pick_rr<GPT_OBS_VEC_HANSEN_GOAL, 0>(rgoal, replay, variant);
}

}  // namespace gpt
