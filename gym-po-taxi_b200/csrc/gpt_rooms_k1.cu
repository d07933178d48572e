// gpt_rooms_k1.cu — instantiations of rooms_step_kernel (see gpt_rooms_kernel.cuh)
#include "gpt_rooms_kernel.cuh"

namespace gpt {

void* rooms_pick_vec(int obs, bool rgoal, bool replay, int variant) {
  switch (obs) {
    case GPT_OBS_VEC_MDP: return pick_rr<GPT_OBS_VEC_MDP, 0>(rgoal, replay, variant);
    case GPT_OBS_VEC_MDP_GOAL: return pick_rr<GPT_OBS_VEC_MDP_GOAL, 0>(rgoal, replay, variant);
    case GPT_OBS_HANSEN: return pick_rr<GPT_OBS_HANSEN, 0>(rgoal, replay, variant);
  }
  return nullptr;
}

}  // namespace gpt
