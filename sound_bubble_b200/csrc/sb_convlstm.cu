// Conv-LSTM variant of the intra-frame path (conv_lstm = true; the Raspberry-Pi / "optim" models).
//
// Reference spans replaced (DE3 = src/models/tfgridnet_realtime_clean_dis_embd3/tfgridnet_causal.py :800-815,
// OPT = src/models/tfgridnet_realtime_clean_optim/tfgridnet_causal.py :684-697 with the deconv of :494-510):
//   FiLM -> Conv1d(C -> C, k = s = down) over frequency -> PReLU -> LayerNorm(C) -> BiLSTM over J steps
//   -> ConvTranspose1d(2H -> C, k = s = down) -> (pad 3 & crop | output_padding) -> + x'
// Three launches: convlstm_pre (conv + PReLU), the shared LSTM kernels in raw-h mode (LayerNorm fused into their
// load), convlstm_post (transposed conv of both directions + tail + residual).
#include "sb_common.cuh"
#include "sb_lstm.cuh"

namespace sb {

template <int C>
__global__ void __launch_bounds__(256) convlstm_pre_kernel(const sb_intra_conv_args a, const int J) {
    SB_DYN_SMEM(float, smem);
    const int F = a.F, k = a.down;
    float* w_s = smem;                              // [k][C][C]  (j, c, o)
    float* xs = w_s + k * C * C;                    // [J*k][C]   FiLM'd input rows the conv touches
    const int tid = threadIdx.x, bt = blockIdx.x, b = bt / a.T;
    for (int i = tid; i < k * C * C / 4; i += 256) st4(w_s + 4 * i, __ldg(reinterpret_cast<const float4*>(a.conv_w) + i));
    pdl_trigger();
    pdl_wait();
    const float* xr = a.x + (size_t)bt * F * C;
    for (int i = tid; i < J * k * C / 4; i += 256) {
        float4 v = ldg4_stream(xr + 4 * i);
        if (a.film_scale) {
            const float4 fs = __ldg(reinterpret_cast<const float4*>(a.film_scale + (size_t)b * F * C) + i);
            const float4 fb = __ldg(reinterpret_cast<const float4*>(a.film_shift + (size_t)b * F * C) + i);
            v.x = fmaf(v.x, fs.x, fb.x); v.y = fmaf(v.y, fs.y, fb.y); v.z = fmaf(v.z, fs.z, fb.z); v.w = fmaf(v.w, fs.w, fb.w);
        }
        st4(xs + 4 * i, v);
    }
    __syncthreads();
    const float slope = __ldg(a.prelu);
    float* z = a.ws + (size_t)bt * J * C;
    for (int i = tid; i < J * C; i += 256) {
        const int j = i / C, o = i - j * C;
        float acc = __ldg(a.conv_b + o);
        const float* xp = xs + j * k * C;
        for (int q = 0; q < k * C; ++q) acc = fmaf(xp[q], w_s[q * C + o], acc);     // q = (tap, c)
        z[i] = acc > 0.f ? acc : slope * acc;
    }
}

template <int C>
__global__ void __launch_bounds__(256) convlstm_post_kernel(const sb_intra_conv_args a, const int J, const float* hbuf) {
    constexpr int H = 64;
    SB_DYN_SMEM(float, hs);                          // [2][J][H]
    const int F = a.F, k = a.down;
    const int tid = threadIdx.x, bt = blockIdx.x, b = bt / a.T;
    const size_t BT = (size_t)a.B * a.T;
    pdl_trigger();
    pdl_wait();
    for (int i = tid; i < 2 * J * H / 4; i += 256) {
        const int d = i / (J * H / 4), r = i - d * (J * H / 4);
        st4(hs + 4 * i, ldg4_stream(hbuf + ((size_t)d * BT + bt) * J * H + 4 * r));
    }
    __syncthreads();
    const float* xr = a.x + (size_t)bt * F * C;
    float* yr = a.y + (size_t)bt * F * C;
    for (int i = tid; i < F * C; i += 256) {
        const int f = i / C, c = i - f * C;
        float v = xr[i];
        if (a.film_scale) v = fmaf(v, __ldg(a.film_scale + (size_t)b * F * C + i), __ldg(a.film_shift + (size_t)b * F * C + i));
        if (f < J * k) {
            const int j = f / k, tap = f - j * k;
            float acc = __ldg(a.deconv_b + c);
#pragma unroll
            for (int d = 0; d < 2; ++d) {
                const float* hp = hs + (d * J + j) * H;
                const float* wp = a.deconv_w + ((size_t)(d * k + tap) * H) * C + c;     // (d, tap, u, c)
#pragma unroll 8
                for (int u = 0; u < H; ++u) acc = fmaf(hp[u], __ldg(wp + u * C), acc);
            }
            v += acc;
        } else if (a.tail_mode == SB_CONVLSTM_OUTPAD) {
            v += __ldg(a.deconv_b + c);             // output_padding positions receive the bias only (OPT:506-510)
        }                                           // pad-and-crop positions receive zeros (DE3:810-813)
        yr[i] = v;
    }
}

template <int C>
static int run_convlstm(const sb_intra_conv_args& a, cudaStream_t st) {
    const int k = a.down, J = (a.F - k) / k + 1, BT = a.B * a.T, H = 64;
    float* z = a.ws;
    float* hbuf = a.ws + (size_t)BT * J * C;
    const size_t smem_pre = ((size_t)k * C * C + (size_t)J * k * C) * sizeof(float);
    SB_CHECK(launch("convlstm_pre", convlstm_pre_kernel<C>, dim3(BT), dim3(256), smem_pre, st, a, J));
    SeqArgs s{};
    s.x0 = z; s.x1 = nullptr; s.film_scale = nullptr; s.film_shift = nullptr;
    s.out[0] = hbuf; s.out[1] = hbuf + (size_t)BT * J * H;
    s.w[0] = a.dir[0]; s.w[1] = a.dir[1];
    s.n_rows = BT; s.n_steps = J; s.n_dirs = 2;
    s.rows_inner = BT; s.stride_outer = 0; s.stride_inner = (long long)J * C; s.stride_pos = C;
    s.film_row_div = 1;
    SB_CHECK(run_seq(s, C, a.H, true, a.algo, st));
    const size_t smem_post = (size_t)2 * J * H * sizeof(float);
    return launch("convlstm_post", convlstm_post_kernel<C>, dim3(BT), dim3(256), smem_post, st, a, J, (const float*)hbuf);
}

}  // namespace sb

extern "C" int sb_intra_convlstm_fwd(const sb_intra_conv_args* p, void* stream) {
    using namespace sb;
    SB_REQUIRE(p && p->x && p->y && p->conv_w && p->conv_b && p->prelu && p->deconv_w && p->deconv_b && p->ws, SB_E_BADARG,
               "sb_intra_convlstm_fwd: null pointer");
    SB_REQUIRE(p->B > 0 && p->T > 0 && p->F > 0, SB_E_BADARG, "sb_intra_convlstm_fwd: bad sizes");
    SB_REQUIRE(p->C == 16 || p->C == 32, SB_E_UNSUPP, "sb_intra_convlstm_fwd: C must be 16 or 32 (got %d)", p->C);
    SB_REQUIRE(p->H == 64, SB_E_UNSUPP, "sb_intra_convlstm_fwd: H must be 64 (got %d)", p->H);
    SB_REQUIRE(p->down >= 1 && p->down <= p->F, SB_E_BADARG, "sb_intra_convlstm_fwd: bad lstm_down %d", p->down);
    SB_REQUIRE((p->film_scale == nullptr) == (p->film_shift == nullptr), SB_E_BADARG, "film scale/shift must come together");
    const int J = (p->F - p->down) / p->down + 1;
    SB_REQUIRE(p->tail_mode == SB_CONVLSTM_OUTPAD || p->F - J * p->down <= 3, SB_E_UNSUPP,
               "sb_intra_convlstm_fwd: pad-and-crop tail covers at most 3 bins (F=%d, down=%d)", p->F, p->down);
    cudaStream_t st = (cudaStream_t)stream;
    return p->C == 32 ? run_convlstm<32>(*p, st) : run_convlstm<16>(*p, st);
}
