/*
 * soundbubble.h — C ABI of libsoundbubble_sm100a.so
 *
 * B200-native (sm_100a) kernels for the Sound Bubble separator forward pass: STFT front-end + inter-microphone
 * features + causal conv-in, FiLM distance conditioning, intra-frame BiLSTM across frequency (plain and conv-LSTM
 * variants), inter-frame LSTM across time with carried state, sliding-window full-band attention, causal deconv +
 * iSTFT overlap-add.
 *
 * The reference (chentuochao/Sound_Bubble) is pure PyTorch and has no FFI; the interface each entry point
 * replaces is therefore a span of the reference's Python hot path, cited per function as
 *   DE3 = src/models/tfgridnet_realtime_clean_dis_embd3/tfgridnet_causal.py
 *   OPT = src/models/tfgridnet_realtime_clean_optim/tfgridnet_causal.py
 * The binding a maintainer would add on the reference side (ctypes) is shown in INTEGRATION.md.
 *
 * Conventions
 *   - every pointer is a DEVICE pointer to float32 unless it says otherwise; nothing is allocated or freed here;
 *   - `stream` is a cudaStream_t passed as void*; all work is enqueued on it and the call returns immediately;
 *   - return value 0 = ok; > 0 = cudaError_t of the failed launch; < 0 = argument error (SB_E_*);
 *     sb_last_error_string() describes the last non-zero return of the calling thread;
 *   - activations between stages live in one layout: X[B][T][F][C], channel innermost ("BTFC").
 *   - streaming state is read and written in the reference's own layouts (DE3:403-421, 696-720) so a state dict
 *     can be handed back and forth between the reference module and this library.
 */
#ifndef SOUNDBUBBLE_H_
#define SOUNDBUBBLE_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SB_VERSION 100          /* 0.1.0 */

#define SB_E_BADARG   (-1)      /* null pointer / non-positive size */
#define SB_E_UNSUPP   (-2)      /* a (C, H, mics, ...) combination no kernel is instantiated for */
#define SB_E_SMEM     (-3)      /* device cannot provide the shared memory a kernel needs */

#define SB_MAX_BLOCKS 16
#define SB_MAX_MICS   8

/* LSTM scheduling: which kernel family runs the recurrence */
#define SB_ALGO_AUTO  0
#define SB_ALGO_TILE  1         /* 8 sequences per warp, weights in shared memory, register-tiled gate GEMM      */
#define SB_ALGO_LANE1 2         /* 1 sequence  per CTA, one gate column per thread, weights in registers        */
#define SB_ALGO_LANE2 3         /* 2 sequences per CTA                                                           */
#define SB_ALGO_LANE4 4         /* 4 sequences per CTA                                                           */
#define SB_ALGO_TILE4 6         /* tile family with 4 sequences per warp (C = 32 only; chosen by TILE for 1-2 step calls) */
#define SB_ALGO_TC    7         /* tcgen05: 128 sequences per CTA, gate GEMM on the tensor cores as a bf16 hi/lo split   */
#define SB_ALGO_TCP   9         /* the same arithmetic as a warp-specialised pipeline: TMA tensor loads of the TF grid   */
                                /* (cp.async.bulk.tensor), LayerNorm / output warps, one MMA-issuing thread, 8 cell-update */
                                /* warps, mbarriers instead of block barriers (sb_lstm_tcp.cu; C = 32, projected mode)     */
                                /* (3 products, fp32 accumulation in TMEM), cell update from TMEM (C = 32, projected mode) */
#define SB_ALGO_TCQ   10        /* SB_ALGO_TCP with two 128-row tiles per CTA in ping-pong: the cell warps alternate between the tiles  */
                                /* while the tensor pipe and the stream group prepare the other one (sb_lstm_tcp.cu::lstm_tcq_kernel)  */
#define SB_ALGO_WS2   8         /* the WS kernel with 2 sequences per CTA sharing the weights in registers, their steps     */
                                /* interleaved phase by phase: 1.6x the latency, 0.81x the SM-time per sequence             */
#define SB_ALGO_WS    5         /* 1 sequence per CTA, warp-specialised: 4 recurrence warps (2 units x 4 gates x K/4 per   */
                                /* thread, weights in registers, packed FFMA2) + 4 helper warps (loads, LayerNorm, input   */
                                /* gates, stores)                                                                           */

/* feature modes of the front-end (DE3:486-507) */
#define SB_FEAT_NONE        0   /* merge_method "None": conv-in sees [Re, Im] only                               */
#define SB_FEAT_OMNI        1   /* MC_features_OMNX  (DE3:72-93): (M-1) ILD + (M-1) (sin, cos) IPD vs mic 0      */
#define SB_FEAT_DIRECTIONAL 2   /* MC_features_direct (DE3:176-207): 1 + 3 ILD, 5 (sin, cos) IPD (6 mics)        */

/* how the distance embedding is normalised / laid out (DE3:114-173) */
#define SB_EMB_CONV    0        /* Dis_Embed_Conv:   view [B,F,Din] -> LayerNorm(Din) -> transpose                */
#define SB_EMB_LINEAR  1        /* Dis_Embed_Linear: LayerNorm over the whole F*Din vector -> view [B,Din,F]      */

/* variant of the intra-frame conv-LSTM tail (a9') */
#define SB_CONVLSTM_PADCROP 0   /* DE3:810-813  deconv, pad 3 zeros, crop to F                                    */
#define SB_CONVLSTM_OUTPAD  1   /* OPT:506-510  deconv with output_padding = F - (F/k)*k                          */

/* process-wide options */
#define SB_OPT_PDL 1            /* 1 = launch kernels with programmatic dependent launch (prologue overlap)       */
#define SB_OPT_ATTN_TC 2        /* 0 = keep the attention core of whole-utterance calls on the SIMT kernel (default 1:  */
                                /* tcgen05 core for T >= 64, W <= 129, F*E <= 320)                                    */
#define SB_OPT_TRAIN_ONE_ROW 3  /* 1 = training LSTM kernels with one gate row (forward) / one W_hh column (BPTT) per thread  */
                                /* and 256 threads (the first version, kept for comparison; default 0: two per thread, 128) */
#define SB_OPT_TRAIN_FFMA2 4    /* packed fp32x2 FMAs in the training GEMM kernels (rowgemm / outer); default 1 (measured 2.7 % per step) */
#define SB_OPT_TC_V1 5          /* 1 = SB_ALGO_TC runs the first tcgen05 kernel (lstm_tc_kernel, LDG operand loads); default 0: the   */
                                /* warp-specialised TMA pipeline lstm_tcp_kernel whenever a tensor map can address the grid (A/B knob) */
#define SB_OPT_TC_CELL7 6       /* lstm_tcp_kernel's cell update with shared reciprocals (5 ex2 + 2 rcp per cell instead of 5 + 5;  */
                                /* the kernel is bound by the XU pipe).  Default set from the measured A/B, see DESIGN.md section 4.  */
#define SB_OPT_TRAIN_TC 7       /* 1 (default) = LSTM weight gradients (dW_ih, dW_hh, db) of the C = 32 paths as a tcgen05 reduction GEMM */
                                /* with bf16 hi/lo three-term operands (sb_train_tc.cu); 0 = the fp32 SIMT outer_kernel                 */
#define SB_OPT_TC_PIPE 8        /* single-addend calls of SB_ALGO_TC run lstm_tcr_kernel: the h part of step s + 1 is issued chunk by   */
                                /* chunk while the cell update of step s is still running (default 1)                                   */
#define SB_OPT_TC_CW16 9        /* lstm_tcr_kernel with 16 cell-update warps (768 threads, setmaxnreg) instead of 8                    */
#define SB_OPT_FRONT_TC 10      /* conv-in of calls with T >= 4 frames on the tensor cores (conv_in_tc_kernel), default 1              */
int sb_set_option(int option, int value);

/* ---------------------------------------------------------------------------------------------------------- */
/* One LSTM direction, packed by sound_bubble_b200/packing.py from the checkpoint tensors                       */
/*   intra_rnn.{weight_ih,weight_hh,bias_ih,bias_hh}_l0[_reverse], intra_norm.norm, intra_linear  (DE3:623-629)  */
/*   inter_rnn.*, inter_norm.norm, inter_linear                                                 (DE3:631-637)  */
/* ---------------------------------------------------------------------------------------------------------- */
typedef struct sb_lstm_dir {
    const float* w_tile;    /* [C+H][4H]  gate-GEMM operand, column p(g,u) = (g/2)*2H + 4*(u/2) + 2*(g%2) + u%2  */
    const float* b_tile;    /* [4H]       b_ih + b_hh in the same column order                                  */
    const float* w_lane;    /* [(C+H)/4][4H][4]  slot s = 4u+g holds row g*H+u of [W_ih | W_hh], 4 k per float4  */
    const float* b_lane;    /* [4H]       b_ih + b_hh, slot order 4u+g                                           */
    const float* w_rec;     /* [16][2][2H] float4: thread t=(ur=t/4,kq=t%4), k, A|B -> W_hh[g*H+ur+32(A|B)][16kq+k], g=0..3 */
    const float* w_xp;      /* [C/4][2][2H] float4: same thread map, k -> W_ih[g*H+unit][(C/4)kq+k] for g = 0..3             */
    const float* w_prj;     /* [4][2H] float4: thread t -> lin[ur%C][16kq+j], j<16, zero outside its plane ur/C             */
    const float* tc_w;      /* raw bf16 words: gate matrix (rows n = 4u+g, K = C+H) hi then lo, projection [C][H] hi then lo, */
                            /* each as the no-swizzle K-major UMMA image [K/8][rows/8][8][8]  (lstm_tc_kernel, C = 32)     */
    const float* tc_b;      /* [4H]       b_ih + b_hh in n = 4u+g order                                          */
    const float* lin_t;     /* [H][C]     output projection, transposed (this direction's half for the BiLSTM)  */
    const float* lin_n;     /* [C][H]     output projection, natural                                             */
    const float* lin_b;     /* [C]        projection bias (added by direction 0 only)                            */
    const float* ln_g;      /* [C]        LayerNorm gain applied to the LSTM input                               */
    const float* ln_b;      /* [C]                                                                               */
} sb_lstm_dir;

/* ---------------------------------------------------------------------------------------------------------- */
/* a3-a5: STFT + re/im regroup + inter-microphone features.                                                     */
/* Replaces self.enc(input) (DE3:475, asteroid Encoder∘STFTFB), the split/cat (:482-484) and                    */
/* MC_features_OMNX / MC_features_direct (:72-93, :176-207).                                                    */
/*   wave  [B][M][n_samples], frame t covers samples [t*stride, t*stride + n_fft); T = (n_samples-n_fft)/stride+1 */
/*   filt  [2F][n_fft]  the checkpoint buffer enc.filterbank._filters (rows 0..F-1 real, F..2F-1 imaginary)      */
/*   feats [B][T][F][Cin], Cin = 2M (+ features): channels [Re m0..mM-1, Im m0..mM-1, ILD..., sin1,cos1,...]     */
/*   spec  optional [B][T][S][2F]: STFT of the first `n_src` microphones (only for spectral masking, :529-530)   */
/* ---------------------------------------------------------------------------------------------------------- */
typedef struct sb_stft_args {
    const float* wave;
    const float* filt;
    float*       feats;
    float*       spec;          /* may be NULL */
    int B, M, n_samples, T;
    int n_fft, stride, F;
    int feat_mode, Cin, n_src;
} sb_stft_args;
int sb_stft_features_fwd(const sb_stft_args* a, void* stream);

/* ---------------------------------------------------------------------------------------------------------- */
/* a6: causal Conv2d(Cin -> C, k=(3 time, 3 freq), pad (0,1)) over [2 history frames ; T frames] + LayerNorm(C). */
/* Replaces torch.cat((conv_buf, batch)) / self.conv (DE3:504-507, :332-354, LayerNormPermuted :219-231).        */
/*   w_pack [3][Cin][3][C] = conv.0.weight[o][c][kt][kf] re-ordered (kt, c, kf, o); bias [C]; ln_g/ln_b NULL if  */
/*   use_first_ln is false.  conv_buf_in/out [B][Cin][2][F] (reference layout); in and out must NOT alias.       */
/* ---------------------------------------------------------------------------------------------------------- */
typedef struct sb_conv_in_args {
    const float* feats;         /* [B][T][F][Cin] */
    const float* conv_buf_in;
    float*       conv_buf_out;
    const float* w_pack;
    const float* bias;
    const float* ln_g;
    const float* ln_b;
    float*       x;             /* [B][T][F][C] */
    int B, T, F, Cin, C;
} sb_conv_in_args;
int sb_conv_in_fwd(const sb_conv_in_args* a, void* stream);

/* ---------------------------------------------------------------------------------------------------------- */
/* a7 + a8 (parameter part): Dis_Embed_Conv / Dis_Embed_Linear (DE3:114-173) and the two 1x1 convs of every      */
/* FilmLayer (:51-68).                                                                                           */
/*   dis [B][3]; emb_w [F*Din][3]; emb_ln_g/b [Din] (conv) or [F*Din] (linear)                                    */
/*   out film [n_layers][2][B][F][C]  (index 0 = scale, 1 = shift)                                               */
/* ---------------------------------------------------------------------------------------------------------- */
typedef struct sb_film_args {
    const float* dis;
    const float* emb_w;
    const float* emb_ln_g;
    const float* emb_ln_b;
    const float* w_w;           /* [n_layers][C][Din]  embeds.j.weight.weight */
    const float* w_b;           /* [n_layers][C]       embeds.j.weight.bias   */
    const float* b_w;           /* [n_layers][C][Din]  embeds.j.bias.weight   */
    const float* b_b;           /* [n_layers][C]       embeds.j.bias.bias     */
    float*       film;
    int B, F, C, Din, n_layers;
    int emb_mode;               /* SB_EMB_CONV / SB_EMB_LINEAR */
} sb_film_args;
int sb_film_params_fwd(const sb_film_args* a, void* stream);

/* ---------------------------------------------------------------------------------------------------------- */
/* a8 (apply) + a9: FiLM, LayerNorm(C), BiLSTM across frequency (zero initial state every frame),                */
/* Linear(2H -> C), residual.  Replaces FilmLayer.forward (DE3:51-68, called :509-513) and the intra branch of   */
/* GridNetBlock.forward (DE3:794-827, conv_lstm=false).                                                          */
/*   x [B][T][F][C] -> y_fwd, y_bwd [B][T][F][C] with  y_fwd + y_bwd = intra_linear(BiLSTM(LN(x'))) + x',        */
/*   x' = x*film_scale + film_shift (film_* NULL for block 0 / the OPT variant).  The two directions are written */
/*   by different CTAs to different buffers; the consumer (sb_inter_lstm_fwd) adds them while loading.           */
/*   SUM MODE, y_bwd == y_fwd: the call zeroes the buffer and both directions add their rows into it (TMA reduce  */
/*   stores, the sum of two addends onto zero is order-independent), so the consumer reads ONE operand.  Only the */
/*   pipelined tensor-core kernel implements it: sb_intra_sum_supported() says whether this call would run; a    */
/*   call it answers 0 for fails with SB_E_UNSUPP.                                                                */
/* ---------------------------------------------------------------------------------------------------------- */
typedef struct sb_intra_args {
    const float* x;
    const float* film_scale;    /* [B][F][C] or NULL */
    const float* film_shift;
    float*       y_fwd;
    float*       y_bwd;
    sb_lstm_dir  dir[2];
    int B, T, F, C, H;
    int algo;
} sb_intra_args;
int sb_intra_lstm_fwd(const sb_intra_args* a, void* stream);
int sb_intra_sum_supported(const sb_intra_args* a);     /* 1 / 0, no launch */

/* ---------------------------------------------------------------------------------------------------------- */
/* a8 (apply) + a9': the conv-LSTM intra branch (DE3:800-815, OPT:684-697, 494-510):                             */
/*   FiLM -> Conv1d(C -> C, k = s = down) over frequency -> PReLU -> LayerNorm(C) -> BiLSTM over J = (F-k)/k + 1  */
/*   steps -> ConvTranspose1d(2H -> C, k = s = down) -> (pad&crop | output_padding) -> + x'.                      */
/*   conv_w [down][C][C] = blocks.i.conv.weight[o][c][j] re-ordered (j, c, o); deconv_w [2][down][H][C] =        */
/*   blocks.i.deconv.weight[d*H+u][c][j] re-ordered (d, j, u, c).  The LSTM entries lin_* of dir[] are unused.   */
/*   ws: workspace of B*T*J*(C + 2H) floats.  y [B][T][F][C].                                                    */
/* ---------------------------------------------------------------------------------------------------------- */
typedef struct sb_intra_conv_args {
    const float* x;
    const float* film_scale;    /* [B][F][C] or NULL */
    const float* film_shift;
    float*       y;
    const float* conv_w;
    const float* conv_b;        /* [C] */
    const float* prelu;         /* [1] */
    const float* deconv_w;
    const float* deconv_b;      /* [C] */
    sb_lstm_dir  dir[2];        /* ln_g/ln_b = blocks.i.norm.norm */
    float*       ws;
    int B, T, F, C, H;
    int down, tail_mode;
    int algo;
} sb_intra_conv_args;
int sb_intra_convlstm_fwd(const sb_intra_conv_args* a, void* stream);

/* ---------------------------------------------------------------------------------------------------------- */
/* a10: LayerNorm(C), LSTM across time from (h0, c0), Linear(H -> C), residual.  Replaces the inter branch of    */
/* GridNetBlock.forward (DE3:829-849).  Input is x0 (+ x1 if not NULL, the two intra directions).                */
/*   h0, c0, hN, cN [B*F][H] with row b*F + f (DE3:833,840); hN/cN may alias h0/c0.                              */
/* ---------------------------------------------------------------------------------------------------------- */
typedef struct sb_inter_args {
    const float* x0;
    const float* x1;            /* may be NULL */
    float*       y;             /* [B][T][F][C] */
    const float* h0;
    const float* c0;
    float*       hN;
    float*       cN;
    sb_lstm_dir  dir;
    int B, T, F, C, H;
    int algo;
} sb_inter_args;
int sb_inter_lstm_fwd(const sb_inter_args* a, void* stream);

/* ---------------------------------------------------------------------------------------------------------- */
/* a11: sliding-window full-band self-attention (DE3:639-684, 856-898, 722-744), heads L, window W:              */
/*   Q,K = LN_{F*E}(PReLU(Linear(C -> L*E))) per head, V = LN_{F*Vd}(PReLU(Linear(C -> C))) per head (Vd = C/L); */
/*   K,V <- [history (W-1 frames, zero-initialised, NOT masked) ; current]; per frame softmax(q.K^T/sqrt(F*E)) V; */
/*   heads regrouped -> PReLU(Linear(C -> C)) -> LN_{F*C} -> + x.                                                 */
/*   K_buf_in/out [B*L][W-1][F*E], V_buf_in/out [B*L][W-1][F*Vd] (reference layout); in/out must NOT alias.      */
/*   ws: workspace of sb_attn_workspace_floats(B, T, F, C, L, E, W) floats.                                      */
/* ---------------------------------------------------------------------------------------------------------- */
typedef struct sb_attn_proj {
    const float* w;             /* [out][C]  attn_conv_*.0.weight      */
    const float* b;             /* [out]                                */
    const float* prelu;         /* [1]       attn_conv_*.1.weight       */
    const float* ln_g;          /* [F*out/L] attn_conv_*.3.norm.weight  */
    const float* ln_b;
} sb_attn_proj;
typedef struct sb_attn_args {
    const float* x;             /* [B][T][F][C] */
    float*       y;             /* [B][T][F][C]; may alias x */
    sb_attn_proj q, k, v, o;
    const float* K_buf_in;  float* K_buf_out;
    const float* V_buf_in;  float* V_buf_out;
    float*       ws;
    int B, T, F, C, L, E, W;
} sb_attn_args;
size_t sb_attn_workspace_floats(int B, int T, int F, int C, int L, int E, int W);
int    sb_attn_fwd(const sb_attn_args* a, void* stream);

/* ---------------------------------------------------------------------------------------------------------- */
/* a13-a15: causal ConvTranspose2d(C -> 2S, k=(3,3), pad (2,1)) over [2 history frames ; T frames], optional     */
/* spectral masking, iSTFT with the previous frame carried, overlap-add, crops.  Replaces DE3:517-542.           */
/*   w [C][2S][3][3] = deconv.weight as stored; bias [2S]; filt [2F][n_fft] = dec.filterbank._filters            */
/*   deconv_buf_in/out [B][C][2][F]; istft_buf_in/out [B][S][2F] (reference shape [B,S,2F,1]); no in/out alias   */
/*   wave_out [B][S][stride*T]                                                                                   */
/* ---------------------------------------------------------------------------------------------------------- */
typedef struct sb_backend_args {
    const float* x;             /* [B][T][F][C] */
    const float* deconv_buf_in;
    float*       deconv_buf_out;
    const float* istft_buf_in;
    float*       istft_buf_out;
    const float* w;
    const float* bias;
    const float* filt;
    const float* mask_spec;     /* optional [B][T][S][2F] from sb_stft_features_fwd, NULL = no spectral masking */
    float*       wave_out;
    float*       ws;            /* workspace of B*(T+1)*S*2F floats (the output spectrum incl. the carried frame) */
    int B, T, F, C, n_src;
    int n_fft, stride;
} sb_backend_args;
int sb_backend_fwd(const sb_backend_args* a, void* stream);

/* ---------------------------------------------------------------------------------------------------------- */
/* The whole path in one call: TFGridNet.forward (DE3:433-552 / OPT:328-441)                                     */
/* ---------------------------------------------------------------------------------------------------------- */
typedef struct sb_block_desc {
    sb_lstm_dir  intra[2];
    sb_lstm_dir  inter;
    /* conv-LSTM intra (conv_lstm = true), else NULL */
    const float* cl_conv_w; const float* cl_conv_b; const float* cl_prelu;
    const float* cl_deconv_w; const float* cl_deconv_b;
    /* attention (use_attn = true), else w == NULL */
    sb_attn_proj attn_q, attn_k, attn_v, attn_o;
} sb_block_desc;

typedef struct sb_net_desc {
    int M, n_fft, stride, F;            /* microphones, window, hop, n_fft/2+1              */
    int C, H, n_blocks, n_src;          /* emb_dim D, lstm hidden, B, num_src               */
    int feat_mode, Cin;
    int film_din;                       /* 0 = no distance embedding (OPT variant)          */
    int emb_mode;
    int spectral_masking;
    int conv_lstm, lstm_down, tail_mode;
    int use_attn, L, E, W;
    const float* enc_filt;
    const float* dec_filt;
    const float* conv_w_pack;
    const float* conv_bias;
    const float* conv_ln_g;             /* NULL if use_first_ln false */
    const float* conv_ln_b;
    const float* emb_w;
    const float* emb_ln_g;
    const float* emb_ln_b;
    const float* film_w_w;
    const float* film_w_b;
    const float* film_b_w;
    const float* film_b_b;
    const float* deconv_w;
    const float* deconv_bias;
    sb_block_desc blocks[SB_MAX_BLOCKS];
} sb_net_desc;

typedef struct sb_net_io {
    const float* wave;                  /* [B][M][n_samples], n_samples = stride*T + (n_fft - stride)            */
    const float* dis_embed;             /* [B][3] or NULL                                                         */
    const float* film;                  /* optional FiLM table [n_blocks-1][2][B][F][C] precomputed by              */
                                        /* sb_film_params_fwd for these dis_embed rows (a streaming session computes */
                                        /* it once); NULL = computed from dis_embed inside every call                */
    float*       wave_out;              /* [B][S][stride*T]                                                       */
    const float* conv_buf_in;   float* conv_buf_out;
    const float* deconv_buf_in; float* deconv_buf_out;
    const float* istft_buf_in;  float* istft_buf_out;
    const float* h_in[SB_MAX_BLOCKS];   float* h_out[SB_MAX_BLOCKS];      /* may alias */
    const float* c_in[SB_MAX_BLOCKS];   float* c_out[SB_MAX_BLOCKS];
    const float* K_in[SB_MAX_BLOCKS];   float* K_out[SB_MAX_BLOCKS];      /* must not alias */
    const float* V_in[SB_MAX_BLOCKS];   float* V_out[SB_MAX_BLOCKS];
    float*       workspace;             /* sb_workspace_floats(desc, B, T) floats                                 */
    int B, T;
    int intra_algo, inter_algo;
} sb_net_io;

size_t sb_workspace_floats(const sb_net_desc* d, int B, int T);
int    sb_net_forward(const sb_net_desc* d, const sb_net_io* io, void* stream);
/* A contiguous part of the same launch sequence.  Units: 0 = front-end (stft_features, conv_in, [film_params]); for */
/* GridNet block i: 1 + 2i = intra-frame path, 2 + 2i = inter-frame path [+ attention]; 2 n_blocks + 1 = back-end;      */
/* sb_net_forward == range [0, 2 n_blocks + 1].  Activations travel between ranges through io->workspace, so the ranges */
/* of one call must use the same sb_net_io.  A range that skips unit 0 of a FiLM model needs io->film.  Used by the     */
/* pipelined streaming session: consecutive chunks depend on each other per unit only (DE3:403-421, 696-720: every      */
/* state tensor belongs to one unit), so chunk t+1 unit u can run next to chunk t unit u+1 on another stream.           */
int    sb_net_forward_range(const sb_net_desc* d, const sb_net_io* io, int first_unit, int last_unit, void* stream);

/* ========================================================================================================== */
/* TRAINING PATH (the *_bwd twins of the stages above; configs 4, 5: src/train_pt.py -> PLModule._step ->           */
/* self.model(inputs) -> loss.backward(), src/hl_modules/distance_based_hl_module.py:303-330, 437-441).              */
/* Training calls run from zero state (the reference trains with input_state=None, DE3/net.py:84-93), read the       */
/* parameters in their CHECKPOINT layouts (they change every step, so nothing is re-packed), keep what the backward  */
/* pass needs in a caller-owned `saved` buffer and ACCUMULATE (+=) into the caller's gradient buffers with fp32      */
/* atomics (zero them first; summation order, hence the last bits, vary from run to run).  Plain BiLSTM / LSTM       */
/* blocks, the conv-LSTM intra path and (first version) the attention.                                                */
/* ========================================================================================================== */

/* One recurrent path of a GridNet block:  y = x + Linear(LSTM(LayerNorm_C(x))).                                     */
/*   inter = 0: intra-frame path (DE3:794-827): rows (b,t), F steps, two directions, lin_w [C][2H]                   */
/*   inter = 1: inter-frame path (DE3:829-849): rows (b,f), T steps, one direction from zero (h0, c0), lin_w [C][H]  */
/*   w_ih [4H][C], w_hh [4H][H], b_ih/b_hh [4H] exactly as nn.LSTM stores them (gate order i, f, g, o).               */
/*   saved: sb_path_train_saved_floats() floats = normalised input, LayerNorm output, 1/std, and per direction the    */
/*   activated gates, c_t and h_t of every step (sequence-major [row][step][.]).  The backward call CONSUMES it (the   */
/*   gate slots are overwritten with the pre-activation gradients).                                                   */
typedef struct sb_path_train_args {
    const float* x;             /* [B][T][F][C] */
    float*       y;             /* [B][T][F][C]; must not alias x */
    const float* ln_g;
    const float* ln_b;
    const float* w_ih[2];
    const float* w_hh[2];
    const float* b_ih[2];
    const float* b_hh[2];
    const float* lin_w;
    const float* lin_b;
    float*       saved;
    int B, T, F, C, H;
    int inter;
} sb_path_train_args;
size_t sb_path_train_saved_floats(int B, int T, int F, int C, int H, int inter);
int    sb_intra_lstm_train_fwd(const sb_path_train_args* a, void* stream);     /* requires inter == 0 */
int    sb_inter_lstm_train_fwd(const sb_path_train_args* a, void* stream);     /* requires inter == 1 */

typedef struct sb_path_bwd_args {
    sb_path_train_args f;       /* the forward call's arguments (y unused) */
    const float* gy;            /* [B][T][F][C] dL/dy */
    float*       gx;            /* [B][T][F][C] dL/dx (written; may alias gy) */
    float* g_ln_g;  float* g_ln_b;
    float* g_w_ih[2]; float* g_w_hh[2]; float* g_b_ih[2]; float* g_b_hh[2];
    float* g_lin_w; float* g_lin_b;
    float* ws;                  /* sb_path_bwd_workspace_floats() floats */
} sb_path_bwd_args;
size_t sb_path_bwd_workspace_floats(int B, int T, int F, int C, int H, int inter);
int    sb_intra_lstm_bwd(const sb_path_bwd_args* a, void* stream);
int    sb_inter_lstm_bwd(const sb_path_bwd_args* a, void* stream);

/* The conv-LSTM intra-frame path (a9', DE3:800-815 / OPT:684-697, 494-510) for training:                              */
/*   y = x + tail(ConvTranspose1d(BiLSTM(LayerNorm(PReLU(Conv1d(x))))))   with k = s = down over frequency.               */
/*   conv_w [C][C][down], deconv_w [2H][C][down], prelu [1], ln_* = blocks.i.norm.norm, all as stored in the checkpoint.  */
/*   saved: sb_convpath_train_saved_floats(); the backward call consumes it.  FiLM is applied before (sb_film_apply_*).  */
typedef struct sb_convpath_train_args {
    const float* x;             /* [B][T][F][C] */
    float*       y;             /* [B][T][F][C]; must not alias x */
    const float* conv_w;
    const float* conv_b;
    const float* prelu;
    const float* ln_g;
    const float* ln_b;
    const float* w_ih[2];
    const float* w_hh[2];
    const float* b_ih[2];
    const float* b_hh[2];
    const float* deconv_w;
    const float* deconv_b;
    float*       saved;
    int B, T, F, C, H;
    int down, tail_mode;        /* SB_CONVLSTM_PADCROP / SB_CONVLSTM_OUTPAD */
} sb_convpath_train_args;
size_t sb_convpath_train_saved_floats(int B, int T, int F, int C, int H, int down);
int    sb_intra_convlstm_train_fwd(const sb_convpath_train_args* a, void* stream);

typedef struct sb_convpath_bwd_args {
    sb_convpath_train_args f;   /* the forward call's arguments (x = the forward input, y unused) */
    const float* gy;            /* [B][T][F][C] */
    float*       gx;            /* [B][T][F][C]; may alias gy */
    float* g_conv_w; float* g_conv_b; float* g_prelu; float* g_ln_g; float* g_ln_b;
    float* g_w_ih[2]; float* g_w_hh[2]; float* g_b_ih[2]; float* g_b_hh[2];
    float* g_deconv_w; float* g_deconv_b;
    float* ws;                  /* sb_convpath_bwd_workspace_floats() floats */
} sb_convpath_bwd_args;
size_t sb_convpath_bwd_workspace_floats(int B, int T, int F, int C, int H, int down);
int    sb_intra_convlstm_bwd(const sb_convpath_bwd_args* a, void* stream);

/* a11 for training, from zero K / V history (DE3:639-684, 856-898): y = x + LN(PReLU(Linear(attention(x)))).           */
/*   Projection parameters as in sb_attn_fwd.  saved: sb_attn_train_saved_floats(); the backward call consumes it.       */
/*   First version (plain loops, one CTA per head-row): the attention is dormant in every shipped configuration.         */
typedef struct sb_attn_train_args {
    const float* x;             /* [B][T][F][C] */
    float*       y;             /* [B][T][F][C]; must not alias x */
    sb_attn_proj q, k, v, o;
    float*       saved;
    int B, T, F, C, L, E, W;
} sb_attn_train_args;
typedef struct sb_attn_proj_grad {
    float* w; float* b; float* prelu; float* ln_g; float* ln_b;
} sb_attn_proj_grad;
typedef struct sb_attn_bwd_args {
    sb_attn_train_args f;       /* the forward call's arguments (y unused) */
    const float* gy;
    float*       gx;            /* may alias gy */
    sb_attn_proj_grad gq, gk, gv, go;
    float*       ws;            /* sb_attn_bwd_workspace_floats() floats */
} sb_attn_bwd_args;
size_t sb_attn_train_saved_floats(const sb_attn_train_args* a);
size_t sb_attn_bwd_workspace_floats(const sb_attn_train_args* a);
int    sb_attn_train_fwd(const sb_attn_train_args* a, void* stream);
int    sb_attn_bwd(const sb_attn_bwd_args* a, void* stream);

/* FilmLayer.forward (DE3:51-68, :509-513) as its own stage: y = x * scale[b,f,c] + shift[b,f,c]; the backward also    */
/* accumulates dL/dscale, dL/dshift [B][F][C] (sums over frames), which sb_film_params_bwd turns into parameter grads.*/
typedef struct sb_film_apply_args {
    const float* x;             /* [B][T][F][C] */
    const float* film_scale;    /* [B][F][C] */
    const float* film_shift;
    float*       y;             /* fwd: output.  bwd: unused */
    const float* gy;            /* bwd only */
    float*       gx;            /* bwd only; may alias gy */
    float*       g_scale;       /* bwd only, [B][F][C], accumulated */
    float*       g_shift;
    int B, T, F, C;
} sb_film_apply_args;
int sb_film_apply_fwd(const sb_film_apply_args* a, void* stream);
int sb_film_apply_bwd(const sb_film_apply_args* a, void* stream);

/* Backward of sb_film_params_fwd (both emb_modes): g_film [n_layers][2][B][F][C] -> gradients of                       */
/* embeds.j.{weight,bias}.{weight,bias}, the embedding LayerNorm (dis_norm / dis_embedding.1) and dis_embedding.0.weight */
typedef struct sb_film_bwd_args {
    sb_film_args f;             /* forward arguments (film unused) */
    const float* g_film;
    float* g_emb_w; float* g_emb_ln_g; float* g_emb_ln_b;
    float* g_w_w; float* g_w_b; float* g_b_w; float* g_b_b;
} sb_film_bwd_args;
int sb_film_params_bwd(const sb_film_bwd_args* a, void* stream);

/* a6 for training: Conv2d(Cin -> C, (3,3), pad (0,1)) from zero history + LayerNorm(C); w [C][Cin][3][3] as stored.  */
/*   saved: B*T*F*(2C+1) floats (conv output, normalised conv output, 1/std) when ln_g != NULL, else unused.         */
typedef struct sb_conv_in_train_args {
    const float* feats;         /* [B][T][F][Cin] */
    const float* w;
    const float* bias;
    const float* ln_g;          /* NULL if use_first_ln is false */
    const float* ln_b;
    float*       x;             /* [B][T][F][C] */
    float*       saved;
    const float* gx;            /* bwd only: dL/dx */
    float* g_w; float* g_bias; float* g_ln_g; float* g_ln_b;    /* bwd only, accumulated */
    float*       ws;            /* bwd only: B*T*F*C floats */
    int B, T, F, Cin, C;
} sb_conv_in_train_args;
int sb_conv_in_train_fwd(const sb_conv_in_train_args* a, void* stream);
int sb_conv_in_bwd(const sb_conv_in_train_args* a, void* stream);

/* Backward of sb_backend_fwd from zero history (DE3:517-542): g_wave [B][S][stride*T] -> gx [B][T][F][C], g_w, g_bias. */
/*   ws: B*T*S*2F floats (the spectrum gradient).  mask_spec as in the forward call or NULL.                           */
typedef struct sb_backend_bwd_args {
    const float* x;             /* [B][T][F][C] the forward input */
    const float* g_wave;
    const float* w;             /* [C][2S][3][3] */
    const float* filt;          /* [2F][n_fft] */
    const float* mask_spec;
    float*       gx;
    float*       g_w;
    float*       g_bias;
    float*       ws;
    int B, T, F, C, n_src;
    int n_fft, stride;
} sb_backend_bwd_args;
int sb_backend_bwd(const sb_backend_bwd_args* a, void* stream);

/* ---------------------------------------------------------------------------------------------------------- */
/* Batch assembly on the device: src/datasets/general_multisrc_dataset_dis_embed.py:112-218 for PCM already in HBM  */
/* (int16 -> float32, target = sum of the in-bubble voices at the reference microphone, one-hot radius) and the      */
/* per-channel perturbations of src/datasets/perturbations (SampleShift = torch.roll, ChannelGain, ChannelDrop,      */
/* PeakNorm); the random draws stay on the host.  Order: shift, gain, drop (they commute), then the peak scale.      */
/* ---------------------------------------------------------------------------------------------------------- */
typedef struct sb_prepare_args {
    const int16_t* mix;         /* [B][M][N] PCM16 mixture                                                          */
    const int16_t* voices;      /* [B][V][N] PCM16 solo voices at the reference microphone (mic 0), NULL if V == 0  */
    const uint8_t* inside;      /* [B][V]    1 = voice within the bubble radius -> part of the target               */
    const float*   gain;        /* [B][M]    linear channel gains or NULL                                           */
    const int*     shift;       /* [B][M]    circular shifts in samples (torch.roll) or NULL                        */
    const uint8_t* drop;        /* [B][M]    1 = channel zeroed (never channel 0) or NULL                           */
    const float*   peak_scale;  /* [B]       PeakNorm's drawn scale, 0 = not applied to this row; or NULL           */
    const int*     radius_idx;  /* [B]       0: 1 m, 1: 1.5 m, 2: 2 m; NULL = 1 m                                   */
    float*         mixture;     /* [B][M][N] out                                                                    */
    float*         target;      /* [B][1][N] out                                                                    */
    float*         dis_embed;   /* [B][3]    out                                                                    */
    float*         peak_ws;     /* sb_prepare_workspace_floats(B, M, N) floats (only read / written with peak_scale) */
    int B, M, V, N;
} sb_prepare_args;
size_t sb_prepare_workspace_floats(int B, int M, int N);
int    sb_prepare_batch_fwd(const sb_prepare_args* a, void* stream);

/* ---------------------------------------------------------------------------------------------------------- */
/* Pipelined streaming (throughput mode of the edge/causal_infer.py:28-47 protocol): one sb_pipe_feed per 8 ms     */
/* chunk, state carried, consecutive chunks overlapping on `depth` streams owned by the pipe.  Chunk t runs on      */
/* stream t % depth as one CUDA graph per unit range (captured at creation from sb_net_forward_range); range j of   */
/* chunk t waits for range j of chunk t-1.  The caller owns all device memory: ios[k] (k = t % n_ios, n_ios = depth */
/* if even else 2*depth) holds the slot's wave / wave_out / workspace (shared by the entries of slot k % depth; the  */
/* same T >= 1 frames per chunk everywhere: T = 1 is the 8 ms protocol, larger T pipelines an offline utterance in    */
/* time slices) and reads the state arena t % 2 / writes the other one; io.film must be set for FiLM models.  Run one */
/* eager sb_net_forward per kernel configuration before sb_pipe_create (shared-memory opt-ins cannot happen in a capture). */
/* A range made of one intra-frame unit carries no state (DE3:819-827) and does not wait for its predecessor.           */
/* Grouped throughput mode: with T = G > 1 frames per sb_net_io the caller may still feed ONE 8 ms window per call        */
/* (sb_pipe_feed_chunk); the pipe gathers G consecutive windows of the rolling protocol (edge/causal_infer.py:39-40)     */
/* into the slot's wave buffer and launches them as one T = G call; sb_pipe_flush / sb_pipe_end run a partial group.      */
/* The only entry points of the library that create CUDA objects (streams, events, graphs); not thread-safe per pipe.*/
/* ---------------------------------------------------------------------------------------------------------- */
typedef struct sb_pipe sb_pipe;
int       sb_pipe_create(const sb_net_desc* d, const sb_net_io* ios, int n_ios, int depth, const int* range_first,
                         const int* range_last, int n_ranges, sb_pipe** out);
int       sb_pipe_destroy(sb_pipe* p);
int       sb_pipe_begin(sb_pipe* p, void* caller_stream);   /* pipe streams wait for the caller's stream            */
int       sb_pipe_feed(sb_pipe* p, const float* window, float* out);  /* host (pinned) or device pointers, or NULL  */
int       sb_pipe_feed_chunk(sb_pipe* p, const float* window, float* out);  /* one [B][M][n_fft] window of a group of T */
int       sb_pipe_flush(sb_pipe* p);                        /* launch the chunks of a partial group now             */
int       sb_pipe_end(sb_pipe* p, void* caller_stream);     /* flush + the caller's stream waits for every chunk fed */
int       sb_pipe_reset(sb_pipe* p);                        /* chunk counter back to 0 (caller resets the state)    */
long long sb_pipe_calls(const sb_pipe* p);

/* ---------------------------------------------------------------------------------------------------------- */
/* Per-stage device timing of sb_net_forward with CUDA events on the launching stream (bench.py's roofline leg).  */
/* sb_profile_begin() arms it; every sb_net_forward until sb_profile_end() brackets each stage with events (do    */
/* not use while the stream is being captured into a graph).  sb_profile_end() synchronises the events and adds,  */
/* per stage kind, the total milliseconds into ms[SB_STAGE_COUNT] and the launch count into calls[SB_STAGE_COUNT].*/
/* ---------------------------------------------------------------------------------------------------------- */
#define SB_STAGE_STFT    0
#define SB_STAGE_CONV_IN 1
#define SB_STAGE_FILM    2
#define SB_STAGE_INTRA   3
#define SB_STAGE_INTER   4
#define SB_STAGE_ATTN    5
#define SB_STAGE_BACKEND 6
#define SB_STAGE_COUNT   7
int sb_profile_begin(void);
int sb_profile_end(double* ms, int64_t* calls);

/* ---------------------------------------------------------------------------------------------------------- */
int         sb_version(void);
const char* sb_last_error_string(void);
/* number of kernel launches issued through this library by the calling process (bench.py's gpu_launches)      */
uint64_t    sb_launch_count(void);
/* sizeof() of the structs above as compiled, so the ctypes mirror can be checked without a GPU                 */
/*   0 lstm_dir 1 stft 2 conv_in 3 film 4 intra 5 inter 6 backend 7 net_desc 8 net_io 9 intra_conv 10 attn_proj */
/*   11 attn 12 block_desc 13 prepare 14 path_train 15 path_bwd 16 film_apply 17 film_bwd 18 conv_in_train     */
/*   19 backend_bwd 20 convpath_train 21 convpath_bwd 22 attn_train 23 attn_proj_grad 24 attn_bwd              */
int         sb_abi_sizeof(int which);

#ifdef __cplusplus
}
#endif
#endif /* SOUNDBUBBLE_H_ */
