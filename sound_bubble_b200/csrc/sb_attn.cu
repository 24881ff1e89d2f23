// placeholder until the attention path lands
#include "sb_common.cuh"
extern "C" size_t sb_attn_workspace_floats(int, int, int, int, int, int, int) { return 0; }
extern "C" int sb_attn_fwd(const sb_attn_args*, void*) {
    sb::set_error("sb_attn_fwd: not built yet");
    return SB_E_UNSUPP;
}
