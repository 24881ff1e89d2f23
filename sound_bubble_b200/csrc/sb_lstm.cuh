// Argument block of the fused sequence (LSTM) kernels of sb_lstm.cu, shared with the conv-LSTM wrapper.
#pragma once
#include "sb_common.cuh"

namespace sb {

struct SeqArgs {
    const float* x0;
    const float* x1;            // optional second addend of the input (the two intra directions)
    const float* film_scale;    // [n_rows / film_row_div][n_steps][C] or NULL
    const float* film_shift;
    float* out[2];              // per direction.  PROJ: same addressing as x.  RAW_H: [row][pos][H]
    sb_lstm_dir w[2];
    const float* h0;            // [n_rows][H] or NULL (zero state)
    const float* c0;
    float* hN;                  // [n_rows][H] or NULL
    float* cN;
    int n_rows, n_steps, n_dirs;
    int rows_inner;             // row -> (row / rows_inner, row % rows_inner)
    long long stride_outer, stride_inner, stride_pos;      // in floats
    int film_row_div;
    int sum_dirs;               // both directions ADD their result into out[0] (zeroed by the caller): lstm_tcr_kernel only
};

__device__ __forceinline__ long long row_base(const SeqArgs& a, int row) {
    const int o = row / a.rows_inner;
    return (long long)o * a.stride_outer + (long long)(row - o * a.rows_inner) * a.stride_inner;
}

// raw_h = true: write h_t to out[dir] as [row][pos][H] instead of the projected, residual-added activation
int run_seq(const SeqArgs& a, int C, int H, bool raw_h, int algo, cudaStream_t st);
// tcgen05 variant (sb_lstm_tc.cu): C = 32, H = 64, projected mode only
int run_seq_tc(const SeqArgs& a, cudaStream_t st);
// warp-specialised tcgen05 + TMA variant (sb_lstm_tcp.cu): same conditions, plus a TMA-addressable activation layout
int run_seq_tcp(const SeqArgs& a, cudaStream_t st);
// the same with two 128-row tiles per CTA in ping-pong (cell warps alternate between the tiles)
int run_seq_tcq(const SeqArgs& a, cudaStream_t st);
bool seq_tcp_supported(const SeqArgs& a);
// would run_seq_tcp take lstm_tcr_kernel for this call (the only kernel that implements SeqArgs::sum_dirs)?
bool seq_tcr_selected(const SeqArgs& a);

}  // namespace sb
