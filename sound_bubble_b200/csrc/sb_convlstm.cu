// placeholder until the conv-LSTM intra path lands
#include "sb_common.cuh"
extern "C" int sb_intra_convlstm_fwd(const sb_intra_conv_args*, void*) {
    sb::set_error("sb_intra_convlstm_fwd: not built yet");
    return SB_E_UNSUPP;
}
