// gpt_rooms_k0.cu — instantiations of rooms_step_kernel (see gpt_rooms_kernel.cuh)
#include "gpt_rooms_kernel.cuh"

namespace gpt {

void* rooms_pick_table(int obs, bool rgoal, bool replay, int variant) {
  switch (obs) {
    case GPT_OBS_ROOM: return pick_rr<GPT_OBS_ROOM, 0>(rgoal, replay, variant);
    case GPT_OBS_ROOM_GOAL: return pick_rr<GPT_OBS_ROOM_GOAL, 0>(rgoal, replay, variant);
    case GPT_OBS_MDP: return pick_rr<GPT_OBS_MDP, 0>(rgoal, replay, variant);
    case GPT_OBS_MDP_GOAL: return pick_rr<GPT_OBS_MDP_GOAL, 0>(rgoal, replay, variant);
  }
  return nullptr;
}

}  // namespace gpt
