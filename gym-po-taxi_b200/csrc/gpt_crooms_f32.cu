// gpt_crooms_f32.cu — float32 fast-mode instantiations of the continuous ROOMS / Tag kernels (gpt_crooms_kernel.cuh).
// Compiled WITHOUT -fmad=false (unlike gpt_crooms.cu): FMA contraction is welcome here, the float32 mode is checked
// against the float64 oracle within a tolerance, not bit for bit.
#include "gpt_crooms_kernel.cuh"

namespace gpt {

void* crooms_pick_f32(int obs, bool replay, bool devctr, bool spec) { return crooms_pick_obs<float>(obs, replay, devctr, spec); }
void* tag_pick_f32(bool replay, bool devctr) { return tag_pick_rr<float>(replay, devctr); }

}  // namespace gpt
