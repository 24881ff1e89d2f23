// The whole separator forward pass as one C call: TFGridNet.forward (DE3:433-552 / OPT:328-441).
//
// Launch sequence (all on the caller's stream, nothing allocated, no host synchronisation):
//   stft_features -> conv_in -> [film_params] -> n_blocks x { intra (BiLSTM | conv-LSTM) -> inter LSTM -> [attention] }
//   -> deconv_spec -> istft_ola
// Activations ping-pong between three [B][T][F][C] buffers of the workspace: a block reads X0, the two intra directions
// write X1 / X2, the inter path reads X1 + X2 and writes X0.
#include <mutex>
#include <vector>

#include "sb_common.cuh"

namespace sb {

// ---- optional per-stage timing (CUDA events on the launching stream) ------------------------------------------------
#ifndef SB_EMU
struct StageEvent { int kind; cudaEvent_t a, b; };
static std::mutex g_prof_mu;
static bool g_prof_on = false;
static std::vector<StageEvent> g_prof_events;

struct StageTimer {
    cudaStream_t st; int kind; bool on; cudaEvent_t a, b;
    StageTimer(void* stream, int k) : st((cudaStream_t)stream), kind(k), on(g_prof_on) {
        if (on) { cudaEventCreate(&a); cudaEventCreate(&b); cudaEventRecord(a, st); }
    }
    ~StageTimer() {
        if (on) {
            cudaEventRecord(b, st);
            std::lock_guard<std::mutex> lock(g_prof_mu);
            g_prof_events.push_back({kind, a, b});
        }
    }
};
#else
struct StageTimer { StageTimer(void*, int) {} };
#endif

struct Workspace {
    float *feats, *spec_in, *film, *x0, *x1, *x2, *spec_out, *extra;
    size_t total;
};

static size_t align_up(size_t n) { return (n + 63) & ~size_t(63); }      // 256-byte granules

static size_t extra_floats(const sb_net_desc* d, int B, int T) {
    size_t n = 0;
    if (d->conv_lstm) {
        const int J = (d->F - d->lstm_down) / d->lstm_down + 1;
        n = (size_t)B * T * J * (d->C + 2 * d->H);
    }
    if (d->use_attn) {
        const size_t m = sb_attn_workspace_floats(B, T, d->F, d->C, d->L, d->E, d->W);
        if (m > n) n = m;
    }
    return n;
}

static Workspace carve(const sb_net_desc* d, int B, int T, float* base) {
    Workspace w{};
    size_t off = 0;
    auto take = [&](size_t n) { float* p = base ? base + off : nullptr; off += align_up(n); return p; };
    const size_t act = (size_t)B * T * d->F * d->C;
    w.feats = take((size_t)B * T * d->F * d->Cin);
    w.spec_in = d->spectral_masking ? take((size_t)B * T * d->n_src * 2 * d->F) : nullptr;
    w.film = d->film_din > 0 && d->n_blocks > 1 ? take((size_t)(d->n_blocks - 1) * 2 * B * d->F * d->C) : nullptr;
    w.x0 = take(act);
    w.x1 = take(act);
    w.x2 = d->conv_lstm ? nullptr : take(act);
    w.spec_out = take((size_t)B * d->n_src * T * 2 * d->F);
    const size_t ex = extra_floats(d, B, T);
    w.extra = ex ? take(ex) : nullptr;
    w.total = off;
    return w;
}

}  // namespace sb

extern "C" size_t sb_workspace_floats(const sb_net_desc* d, int B, int T) {
    if (!d || B <= 0 || T <= 0) return 0;
    return sb::carve(d, B, T, nullptr).total;
}

// Units of the launch sequence: 0 = front-end (stft_features, conv_in, [film_params]); for GridNet block i: 1 + 2i =
// its intra-frame path, 2 + 2i = its inter-frame path [+ attention]; 2 n_blocks + 1 = back-end.  sb_net_forward runs all
// of them; a pipelined streaming session captures one CUDA graph per range so that consecutive chunks can overlap on
// several streams (sb_pipe.cu).  Carried state lives in units 0 (conv_buf), 2 + 2i (h, c, K/V) and the last (deconv /
// iSTFT history); the intra units carry none.
extern "C" int sb_net_forward_range(const sb_net_desc* d, const sb_net_io* io, int first_unit, int last_unit, void* stream) {
    using namespace sb;
    SB_REQUIRE(d && io, SB_E_BADARG, "sb_net_forward: null descriptor");
    SB_REQUIRE(first_unit >= 0 && first_unit <= last_unit && last_unit <= 2 * d->n_blocks + 1, SB_E_BADARG,
               "sb_net_forward_range: bad unit range [%d, %d]", first_unit, last_unit);
    SB_REQUIRE(io->wave && io->wave_out && io->workspace, SB_E_BADARG, "sb_net_forward: null wave / wave_out / workspace");
    SB_REQUIRE(io->B > 0 && io->T > 0, SB_E_BADARG, "sb_net_forward: bad B/T");
    SB_REQUIRE(d->n_blocks > 0 && d->n_blocks <= SB_MAX_BLOCKS, SB_E_UNSUPP, "sb_net_forward: n_blocks=%d out of range", d->n_blocks);
    SB_REQUIRE(d->film_din == 0 || io->dis_embed || io->film, SB_E_BADARG, "sb_net_forward: dis_embed is required by this model");
    SB_REQUIRE(((uintptr_t)io->workspace & 15) == 0, SB_E_BADARG, "sb_net_forward: workspace must be 16-byte aligned");
    const int B = io->B, T = io->T;
    const Workspace w = carve(d, B, T, io->workspace);

    const float* film = w.film;
    if (io->film && d->film_din > 0 && d->n_blocks > 1) film = io->film;
    SB_REQUIRE(first_unit == 0 || !w.film || film != w.film, SB_E_BADARG,
               "sb_net_forward_range: a range that skips the front-end needs io->film (precomputed FiLM table)");

    if (first_unit == 0) {
    sb_stft_args sa{};
    sa.wave = io->wave; sa.filt = d->enc_filt; sa.feats = w.feats; sa.spec = w.spec_in;
    sa.B = B; sa.M = d->M; sa.n_samples = d->stride * T + (d->n_fft - d->stride); sa.T = T;
    sa.n_fft = d->n_fft; sa.stride = d->stride; sa.F = d->F;
    sa.feat_mode = d->feat_mode; sa.Cin = d->Cin; sa.n_src = d->n_src;
    { StageTimer tm(stream, SB_STAGE_STFT); SB_CHECK(sb_stft_features_fwd(&sa, stream)); }

    sb_conv_in_args ca{};
    ca.feats = w.feats; ca.conv_buf_in = io->conv_buf_in; ca.conv_buf_out = io->conv_buf_out;
    ca.w_pack = d->conv_w_pack; ca.bias = d->conv_bias; ca.ln_g = d->conv_ln_g; ca.ln_b = d->conv_ln_b;
    ca.x = w.x0; ca.B = B; ca.T = T; ca.F = d->F; ca.Cin = d->Cin; ca.C = d->C;
    { StageTimer tm(stream, SB_STAGE_CONV_IN); SB_CHECK(sb_conv_in_fwd(&ca, stream)); }

    if (film == w.film && w.film) {
        sb_film_args fa{};
        fa.dis = io->dis_embed; fa.emb_w = d->emb_w; fa.emb_ln_g = d->emb_ln_g; fa.emb_ln_b = d->emb_ln_b;
        fa.w_w = d->film_w_w; fa.w_b = d->film_w_b; fa.b_w = d->film_b_w; fa.b_b = d->film_b_b;
        fa.film = w.film; fa.B = B; fa.F = d->F; fa.C = d->C; fa.Din = d->film_din; fa.n_layers = d->n_blocks - 1;
        fa.emb_mode = d->emb_mode;
        { StageTimer tm(stream, SB_STAGE_FILM); SB_CHECK(sb_film_params_fwd(&fa, stream)); }
    }
    }   // front-end

    const size_t film_stride = (size_t)B * d->F * d->C;
    for (int i = 0; i < d->n_blocks; ++i) {
        const bool do_intra = 1 + 2 * i >= first_unit && 1 + 2 * i <= last_unit;
        const bool do_inter = 2 + 2 * i >= first_unit && 2 + 2 * i <= last_unit;
        if (!do_intra && !do_inter) continue;
        const sb_block_desc& bd = d->blocks[i];
        const float* fscale = (film && i > 0) ? film + (size_t)(i - 1) * 2 * film_stride : nullptr;
        const float* fshift = fscale ? fscale + film_stride : nullptr;
        // Plain LSTM blocks: where the pipelined tensor-core kernel runs the intra unit, both directions are summed into ONE
        // buffer (TMA reduce stores) and the inter LSTM reads a single operand - which puts it on the same kernel.  The
        // decision depends only on sizes, pointers and options, so calls that run the two units separately (sb_pipe's
        // single-unit ranges) agree on it.
        sb_intra_args ia{};
        ia.x = w.x0; ia.film_scale = fscale; ia.film_shift = fshift; ia.y_fwd = w.x1; ia.y_bwd = w.x1;
        ia.dir[0] = bd.intra[0]; ia.dir[1] = bd.intra[1];
        ia.B = B; ia.T = T; ia.F = d->F; ia.C = d->C; ia.H = d->H; ia.algo = io->intra_algo;
        const bool summed = !d->conv_lstm && sb_intra_sum_supported(&ia);
        if (!summed) ia.y_bwd = w.x2;
        const float* inter_x1 = (d->conv_lstm || summed) ? nullptr : w.x2;   // else: the second intra direction, added on load
        if (do_intra && d->conv_lstm) {
            sb_intra_conv_args ca{};
            ca.x = w.x0; ca.film_scale = fscale; ca.film_shift = fshift; ca.y = w.x1;
            ca.conv_w = bd.cl_conv_w; ca.conv_b = bd.cl_conv_b; ca.prelu = bd.cl_prelu;
            ca.deconv_w = bd.cl_deconv_w; ca.deconv_b = bd.cl_deconv_b;
            ca.dir[0] = bd.intra[0]; ca.dir[1] = bd.intra[1];
            ca.ws = w.extra; ca.B = B; ca.T = T; ca.F = d->F; ca.C = d->C; ca.H = d->H;
            ca.down = d->lstm_down; ca.tail_mode = d->tail_mode; ca.algo = io->intra_algo;
            { StageTimer tm(stream, SB_STAGE_INTRA); SB_CHECK(sb_intra_convlstm_fwd(&ca, stream)); }
        } else if (do_intra) {
            { StageTimer tm(stream, SB_STAGE_INTRA); SB_CHECK(sb_intra_lstm_fwd(&ia, stream)); }
        }
        if (!do_inter) continue;
        sb_inter_args na{};
        na.x0 = w.x1; na.x1 = inter_x1; na.y = w.x0;
        na.h0 = io->h_in[i]; na.c0 = io->c_in[i]; na.hN = io->h_out[i]; na.cN = io->c_out[i];
        na.dir = bd.inter; na.B = B; na.T = T; na.F = d->F; na.C = d->C; na.H = d->H; na.algo = io->inter_algo;
        { StageTimer tm(stream, SB_STAGE_INTER); SB_CHECK(sb_inter_lstm_fwd(&na, stream)); }
        if (d->use_attn) {
            sb_attn_args aa{};
            aa.x = w.x0; aa.y = w.x0;
            aa.q = bd.attn_q; aa.k = bd.attn_k; aa.v = bd.attn_v; aa.o = bd.attn_o;
            aa.K_buf_in = io->K_in[i]; aa.K_buf_out = io->K_out[i];
            aa.V_buf_in = io->V_in[i]; aa.V_buf_out = io->V_out[i];
            aa.ws = w.extra; aa.B = B; aa.T = T; aa.F = d->F; aa.C = d->C; aa.L = d->L; aa.E = d->E; aa.W = d->W;
            { StageTimer tm(stream, SB_STAGE_ATTN); SB_CHECK(sb_attn_fwd(&aa, stream)); }
        }
    }

    if (last_unit <= 2 * d->n_blocks) return 0;
    sb_backend_args ba{};
    ba.x = w.x0; ba.deconv_buf_in = io->deconv_buf_in; ba.deconv_buf_out = io->deconv_buf_out;
    ba.istft_buf_in = io->istft_buf_in; ba.istft_buf_out = io->istft_buf_out;
    ba.w = d->deconv_w; ba.bias = d->deconv_bias; ba.filt = d->dec_filt;
    ba.mask_spec = w.spec_in; ba.wave_out = io->wave_out; ba.ws = w.spec_out;
    ba.B = B; ba.T = T; ba.F = d->F; ba.C = d->C; ba.n_src = d->n_src; ba.n_fft = d->n_fft; ba.stride = d->stride;
    StageTimer tm(stream, SB_STAGE_BACKEND);
    return sb_backend_fwd(&ba, stream);
}

extern "C" int sb_net_forward(const sb_net_desc* d, const sb_net_io* io, void* stream) {
    SB_REQUIRE(d, SB_E_BADARG, "sb_net_forward: null descriptor");
    return sb_net_forward_range(d, io, 0, 2 * d->n_blocks + 1, stream);
}

extern "C" int sb_profile_begin(void) {
#ifndef SB_EMU
    std::lock_guard<std::mutex> lock(sb::g_prof_mu);
    sb::g_prof_on = true;
#endif
    return 0;
}

extern "C" int sb_profile_end(double* ms, int64_t* calls) {
    using namespace sb;
    SB_REQUIRE(ms && calls, SB_E_BADARG, "sb_profile_end: null output");
#ifndef SB_EMU
    std::lock_guard<std::mutex> lock(g_prof_mu);
    g_prof_on = false;
    int rc = 0;
    for (auto& e : g_prof_events) {
        float t = 0.f;
        cudaError_t err = cudaEventSynchronize(e.b);
        if (err == cudaSuccess) err = cudaEventElapsedTime(&t, e.a, e.b);
        if (err != cudaSuccess) { set_error("sb_profile_end: %s", cudaGetErrorString(err)); rc = (int)err; }
        else if (e.kind >= 0 && e.kind < SB_STAGE_COUNT) { ms[e.kind] += t; calls[e.kind] += 1; }
        cudaEventDestroy(e.a);
        cudaEventDestroy(e.b);
    }
    g_prof_events.clear();
    return rc;
#else
    return 0;
#endif
}
