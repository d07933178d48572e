#include "gpt_internal.h"
namespace gpt {
int crooms_create(gpt_env*, const gpt_config*) { return fail(GPT_E_ARG, "crooms: not built yet"); }
int tag_create(gpt_env*, const gpt_config*) { return fail(GPT_E_ARG, "tag: not built yet"); }
int crooms_launch(gpt_env*, const LaunchArgs&) { return fail(GPT_E_ARG, "crooms: not built yet"); }
int tag_launch(gpt_env*, const LaunchArgs&) { return fail(GPT_E_ARG, "tag: not built yet"); }
}
