// Front-end of the separator: windowed STFT + inter-microphone features, causal conv-in + LayerNorm, FiLM parameters.
//
// Reference spans replaced (DE3 = src/models/tfgridnet_realtime_clean_dis_embd3/tfgridnet_causal.py):
//   stft_features_kernel : self.enc(input) :475 (asteroid Encoder∘STFTFB = conv1d with the [2F,1,n_fft] basis buffer),
//                          re/im regroup :482-484, MC_features_OMNX :72-93 (+IPD_OMNX :32-48), MC_features_direct :176-207
//   conv_in_kernel       : history cat :504-505, Conv2d k=(3,3) pad (0,1) :332-347, LayerNormPermuted :219-231
//   film_params_kernel   : Dis_Embed_Conv :150-173 / Dis_Embed_Linear :114-147, FilmLayer 1x1 convs :51-68
#include "sb_common.cuh"

namespace sb {

// =============================================================================================================
// STFT + features
// =============================================================================================================
// CTA = (32-bin chunk, 8-frame tile, utterance).  The basis rows of the chunk (32 real + 32 imaginary, straight from
// the checkpoint buffer) and the wave samples of the tile sit in shared memory; each thread owns one bin of two
// consecutive frames for all M microphones (2*2*M accumulators), so a basis float4 feeds 2*M*4 FMAs and the wave reads
// are warp-wide broadcasts.  Features are computed in registers and staged through shared memory so the [t][f][Cin]
// rows leave as contiguous runs.
constexpr int kStftTT = 8;      // frames per CTA
constexpr int kStftFC = 32;     // bins per CTA

template <int M>
__global__ void __launch_bounds__(128, 2) stft_features_kernel(const sb_stft_args a) {
    SB_DYN_SMEM(float, smem);
    const int n_fft = a.n_fft, stride = a.stride, F = a.F, Cin = a.Cin;
    const int BS = n_fft + 4;                                   // basis row stride (== 4 mod 32 for n_fft = 288)
    const int WT = (kStftTT - 1) * stride + n_fft;              // wave samples per microphone in the tile
    float* bs = smem;                                           // [64][BS]
    float* ws = bs + 2 * kStftFC * BS;                          // [M][WT]
    float* fs = bs;                                             // staging [TT][32][Cin], reuses the basis tile

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int f0 = blockIdx.x * kStftFC, t0 = blockIdx.y * kStftTT, b = blockIdx.z;
    const int nvalid = min(kStftTT, a.T - t0);
    const int nf = min(kStftFC, F - f0);

    // basis chunk (constant data): warp w stages rows w, w+4, ... with asynchronous 16-byte copies (all in flight)
    {
        const int q = n_fft / 4;
        for (int row = warp; row < 2 * kStftFC; row += 4) {
            const int ff = row & (kStftFC - 1);
            float* dst = bs + row * BS;
            if (ff < nf) {
                const float* src = a.filt + (size_t)((row < kStftFC) ? f0 + ff : F + f0 + ff) * n_fft;
                for (int k4 = lane; k4 < q; k4 += 32) cp_async_16(dst + 4 * k4, src + 4 * k4);
            } else {
                for (int k4 = lane; k4 < q; k4 += 32) st4(dst + 4 * k4, make_float4(0.f, 0.f, 0.f, 0.f));
            }
        }
    }
    pdl_trigger();
    pdl_wait();
    {   // wave tile: only the samples the valid frames touch are read; what lies beyond feeds frames that are never
        // stored, so it is left as is.  Warp w stages microphones w, w+4, ...
        const int wvalid = (nvalid - 1) * stride + n_fft;
        const float* wsrc = a.wave + (size_t)b * M * a.n_samples + (size_t)t0 * stride;
        const bool vec = ((a.n_samples & 3) == 0) && ((reinterpret_cast<uintptr_t>(a.wave) & 15) == 0);
        for (int m = warp; m < M; m += 4) {
            const float* src = wsrc + (size_t)m * a.n_samples;
            float* dst = ws + m * WT;
            if (vec) {
                for (int n4 = lane; n4 < wvalid / 4; n4 += 32) cp_async_16(dst + 4 * n4, src + 4 * n4);
            } else {
                for (int n = lane; n < wvalid; n += 32) cp_async_4(dst + n, src + n);
            }
        }
        cp_async_wait_all();
    }
    __syncthreads();

    // streaming-sized tiles (one frame pair): the four warps split the window instead of idling, partial sums meet in
    // shared memory.  Otherwise warp w owns frame pair w.
    const bool ksplit = nvalid <= 2 && (n_fft % 16) == 0;
    const int fpair = ksplit ? 0 : warp;
    const int kbeg = ksplit ? warp * (n_fft / 4) : 0;
    const int kend = ksplit ? kbeg + n_fft / 4 : n_fft;
    float re[2][M], im[2][M];
#pragma unroll
    for (int j = 0; j < 2; ++j)
#pragma unroll
        for (int m = 0; m < M; ++m) { re[j][m] = 0.f; im[j][m] = 0.f; }
    const float* br = bs + lane * BS;
    const float* bi = bs + (kStftFC + lane) * BS;
    const float* wa = ws + (2 * fpair) * stride;
    const float* wb = wa + stride;
    for (int k = kbeg; k < kend; k += 4) {
        const float4 r4 = ld4(br + k), i4 = ld4(bi + k);
#pragma unroll
        for (int m = 0; m < M; ++m) {
            const float4 xa = ld4(wa + m * WT + k);
            const float4 xb = ld4(wb + m * WT + k);
            re[0][m] = fmaf(xa.x, r4.x, re[0][m]); re[0][m] = fmaf(xa.y, r4.y, re[0][m]);
            re[0][m] = fmaf(xa.z, r4.z, re[0][m]); re[0][m] = fmaf(xa.w, r4.w, re[0][m]);
            im[0][m] = fmaf(xa.x, i4.x, im[0][m]); im[0][m] = fmaf(xa.y, i4.y, im[0][m]);
            im[0][m] = fmaf(xa.z, i4.z, im[0][m]); im[0][m] = fmaf(xa.w, i4.w, im[0][m]);
            re[1][m] = fmaf(xb.x, r4.x, re[1][m]); re[1][m] = fmaf(xb.y, r4.y, re[1][m]);
            re[1][m] = fmaf(xb.z, r4.z, re[1][m]); re[1][m] = fmaf(xb.w, r4.w, re[1][m]);
            im[1][m] = fmaf(xb.x, i4.x, im[1][m]); im[1][m] = fmaf(xb.y, i4.y, im[1][m]);
            im[1][m] = fmaf(xb.z, i4.z, im[1][m]); im[1][m] = fmaf(xb.w, i4.w, im[1][m]);
        }
    }
    __syncthreads();                    // everyone is done with the basis tile; reuse it as the staging buffer
    if (ksplit) {
        float* red = bs + kStftTT * kStftFC * Cin;          // [4 warps][4M][32], behind the feature staging area
#pragma unroll
        for (int j = 0; j < 2; ++j)
#pragma unroll
            for (int m = 0; m < M; ++m) {
                red[((warp * 4 * M) + (2 * j) * M + m) * 32 + lane] = re[j][m];
                red[((warp * 4 * M) + (2 * j + 1) * M + m) * 32 + lane] = im[j][m];
            }
        __syncthreads();
        if (warp == 0) {
#pragma unroll
            for (int j = 0; j < 2; ++j)
#pragma unroll
                for (int m = 0; m < M; ++m) {
                    float sr = 0.f, si = 0.f;
#pragma unroll
                    for (int w4 = 0; w4 < 4; ++w4) {
                        sr += red[((w4 * 4 * M) + (2 * j) * M + m) * 32 + lane];
                        si += red[((w4 * 4 * M) + (2 * j + 1) * M + m) * 32 + lane];
                    }
                    re[j][m] = sr; im[j][m] = si;
                }
        }
    }
    const bool writer = !ksplit || warp == 0;

    const float eps = 1e-6f;
#pragma unroll
    for (int j = 0; j < 2; ++j) {
        const int tt = 2 * fpair + j;
        if (!writer || tt >= nvalid) continue;            // frames past the end of the call are never stored
        float* o = fs + (tt * kStftFC + lane) * Cin;
#pragma unroll
        for (int m = 0; m < M; ++m) { o[m] = re[j][m]; o[M + m] = im[j][m]; }
        if (a.feat_mode != SB_FEAT_NONE) {
            float mag[M];
#pragma unroll
            for (int m = 0; m < M; ++m) mag[m] = sqrtf(re[j][m] * re[j][m] + im[j][m] * im[j][m]);
            int c = 2 * M;
            if (a.feat_mode == SB_FEAT_OMNI) {
#pragma unroll
                for (int m = 1; m < M; ++m) o[c++] = log10f((mag[m] + eps) / (mag[0] + eps));
            } else if (M >= 6) {        // SB_FEAT_DIRECTIONAL: mic2/mic3, then mics 1, 4, 5 against mic 0
                o[c++] = log10f((mag[2] + eps) / (mag[3] + eps));
                o[c++] = log10f((mag[1] + eps) / (mag[0] + eps));
                o[c++] = log10f((mag[4] + eps) / (mag[0] + eps));
                o[c++] = log10f((mag[5] + eps) / (mag[0] + eps));
            }
#pragma unroll
            for (int m = 1; m < M; ++m) {
                const float den = mag[m] * mag[0] + eps;
                o[c++] = (re[j][0] * im[j][m] - im[j][0] * re[j][m]) / den;             // sin
                o[c++] = (re[j][m] * re[j][0] + im[j][m] * im[j][0]) / den;             // cos
            }
        }
        if (a.spec && tt < nvalid && lane < nf) {
            float* sp = a.spec + ((size_t)(b * a.T + t0 + tt) * a.n_src) * 2 * F + f0 + lane;
#pragma unroll
            for (int m = 0; m < M; ++m)
                if (m < a.n_src) { sp[(size_t)m * 2 * F] = re[j][m]; sp[(size_t)m * 2 * F + F] = im[j][m]; }
        }
    }
    __syncthreads();
    for (int tt = 0; tt < nvalid; ++tt) {
        float* dst = a.feats + ((size_t)(b * a.T + t0 + tt) * F + f0) * Cin;
        const float* src = fs + tt * kStftFC * Cin;
        for (int i = tid; i < nf * Cin; i += 128) dst[i] = src[i];
    }
}

template <int M>
static int launch_stft(const sb_stft_args& a, cudaStream_t st) {
    const size_t smem = (size_t)(2 * kStftFC * (a.n_fft + 4) + M * ((kStftTT - 1) * a.stride + a.n_fft)) * sizeof(float);
    SB_REQUIRE(smem <= 227 * 1024, SB_E_SMEM, "sb_stft_features_fwd: n_fft=%d stride=%d needs %zu bytes of shared memory", a.n_fft, a.stride, smem);
    dim3 grid(ceil_div(a.F, kStftFC), ceil_div(a.T, kStftTT), a.B);
    return launch("stft_features", stft_features_kernel<M>, grid, dim3(128), smem, st, a);
}

// =============================================================================================================
// conv-in + LayerNorm
// =============================================================================================================
// CTA = (tile of TT frames, utterance).  The TT+2 input frames (history from conv_buf for t < 0) are transposed into
// shared memory as [frame][c][F+2] with zero columns for the frequency padding; the packed weights [kt][c][kf][C]
// follow.  One thread computes 8 output channels of one (t, f): per input value 1 LDS + 2 broadcast LDS.128 + 8 FMA.
// The C/8 threads of a (t, f) are adjacent lanes, so LayerNorm(C) is two quad shuffles and the store is coalesced.
//
// PP = 4 (whole-utterance and grouped calls): a thread owns 8 output channels of FOUR adjacent bins, so the two weight
// LDS.128 of a tap feed 4 x 8 FMAs and the six input values of a (frame, channel) row arrive as two LDS.128: 8 shared-memory
// loads per 96 FMAs instead of 9 per 24 (the PP = 1 form was bound by the shared-memory pipe at 18 % of the FMA rate and had
// grown to 18 % of the grouped step's SM-time, profiles/r02_launches_bench.txt).  Same summation order: bit-identical.
template <int C, int PP>
__global__ void __launch_bounds__(640, 1) conv_in_kernel(const sb_conv_in_args a, const int TT, const int FC, const int FP) {
    constexpr int NOG = C / 8;
    SB_DYN_SMEM(float, smem);
    const int F = a.F, Cin = a.Cin;
    float* w_s = smem;                                          // [3][Cin][3][C]
    float* in_s = w_s + 9 * Cin * C;                            // [TT+2][Cin][FP]; column j <-> bin f0 - 1 + j
    const int tid = threadIdx.x;
    const int t0 = blockIdx.x * TT, b = blockIdx.y, f0 = blockIdx.z * FC;
    const int nvalid = min(TT, a.T - t0);
    const int nf = min(FC, F - f0);

    stage_f4(w_s, a.w_pack, 9 * Cin * C / 4, tid, blockDim.x);
    pdl_trigger();
    pdl_wait();

    const int nfr = nvalid + 2;
    const int ncol = nf + 2;
    const int nwarps = blockDim.x >> 5, warp = tid >> 5, lane = tid & 31;
    for (int fr = 0; fr < nfr; ++fr) {                          // asynchronous 4-byte copies: a transposing gather
        const int ft = t0 - 2 + fr;
        float* dst = in_s + fr * Cin * FP;
        if (ft >= 0) {
            const float* src = a.feats + (size_t)(b * a.T + ft) * F * Cin;
            for (int j = warp; j < ncol; j += nwarps) {         // one bin (Cin contiguous floats in global) per warp pass
                const int f = f0 - 1 + j;
                if (f >= 0 && f < F) {
                    for (int c = lane; c < Cin; c += 32) cp_async_4(dst + c * FP + j, src + (size_t)f * Cin + c);
                } else {
                    for (int c = lane; c < Cin; c += 32) dst[c * FP + j] = 0.0f;
                }
            }
        } else {
            const float* src = a.conv_buf_in + (size_t)b * Cin * 2 * F + (size_t)(2 + ft) * F;
            for (int c = warp; c < Cin; c += nwarps) {          // one channel row (bins contiguous in global) per warp pass
                for (int j = lane; j < ncol; j += 32) {
                    const int f = f0 - 1 + j;
                    if (f >= 0 && f < F) cp_async_4(dst + c * FP + j, src + (size_t)c * 2 * F + f);
                    else dst[c * FP + j] = 0.0f;
                }
            }
        }
    }
    if (PP > 1) {                                               // columns past the tile: read by the last bin group, never stored
        const int npad = FP - ncol;
        for (int i = tid; i < nfr * Cin * npad; i += blockDim.x) {
            const int row = i / npad;
            in_s[row * FP + ncol + (i - row * npad)] = 0.0f;
        }
    }
    cp_async_wait_all();
    __syncthreads();

    if constexpr (PP == 4) {
        const int NFG = (nf + 3) >> 2;
        const int n_items4 = nvalid * NFG * NOG;
        for (int base = 0; base < n_items4; base += blockDim.x) {
            const int item = base + tid;
            const bool valid = item < n_items4;
            const int it = valid ? item : 0;
            const int og = it % NOG, pf = it / NOG;
            const int tt = pf / NFG, fl0 = (pf - tt * NFG) << 2;
            float2 acc[4][4];
            {
                const float4 b0 = __ldg(reinterpret_cast<const float4*>(a.bias) + 2 * og);
                const float4 b1 = __ldg(reinterpret_cast<const float4*>(a.bias) + 2 * og + 1);
#pragma unroll
                for (int p = 0; p < 4; ++p) {
                    acc[p][0] = make_float2(b0.x, b0.y); acc[p][1] = make_float2(b0.z, b0.w);
                    acc[p][2] = make_float2(b1.x, b1.y); acc[p][3] = make_float2(b1.z, b1.w);
                }
            }
            for (int kt = 0; kt < 3; ++kt) {
                const float* ip = in_s + (tt + kt) * Cin * FP + fl0;
                const float* wp = w_s + (kt * Cin * 3) * C + og * 8;
#pragma unroll 3
                for (int c = 0; c < Cin; ++c) {
                    const float4 a0 = ld4(ip + c * FP), a1 = ld4(ip + c * FP + 4);
                    const float v[6] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y};
#pragma unroll
                    for (int kf = 0; kf < 3; ++kf) {
                        const float4 w0 = ld4(wp + (c * 3 + kf) * C);
                        const float4 w1 = ld4(wp + (c * 3 + kf) * C + 4);
#pragma unroll
                        for (int p = 0; p < 4; ++p) {
                            ffma2(acc[p][0], make_float2(w0.x, w0.y), v[p + kf]);
                            ffma2(acc[p][1], make_float2(w0.z, w0.w), v[p + kf]);
                            ffma2(acc[p][2], make_float2(w1.x, w1.y), v[p + kf]);
                            ffma2(acc[p][3], make_float2(w1.z, w1.w), v[p + kf]);
                        }
                    }
                }
            }
#pragma unroll
            for (int p = 0; p < 4; ++p) {
                float o[8] = {acc[p][0].x, acc[p][0].y, acc[p][1].x, acc[p][1].y, acc[p][2].x, acc[p][2].y, acc[p][3].x, acc[p][3].y};
                if (a.ln_g) {
                    float s1 = 0.f;
#pragma unroll
                    for (int j = 0; j < 8; ++j) s1 += o[j];
                    const float mean = group_sum<NOG>(s1) * (1.0f / C);
                    float s2 = 0.f;
#pragma unroll
                    for (int j = 0; j < 8; ++j) { o[j] -= mean; s2 = fmaf(o[j], o[j], s2); }
                    const float rstd = rsqrtf(group_sum<NOG>(s2) * (1.0f / C) + kLnEps);
#pragma unroll
                    for (int j = 0; j < 8; ++j) o[j] = fmaf(o[j] * rstd, __ldg(a.ln_g + og * 8 + j), __ldg(a.ln_b + og * 8 + j));
                }
                if (valid && fl0 + p < nf) {
                    float* dst = a.x + ((size_t)(b * a.T + t0 + tt) * F + f0 + fl0 + p) * C + og * 8;
                    st4(dst, make_float4(o[0], o[1], o[2], o[3]));
                    st4(dst + 4, make_float4(o[4], o[5], o[6], o[7]));
                }
            }
        }
    } else {
    const int n_items = nvalid * nf * NOG;
    for (int base = 0; base < n_items; base += blockDim.x) {
        const int item = base + tid;
        const bool valid = item < n_items;
        const int it = valid ? item : 0;
        const int og = it % NOG, pf = it / NOG;
        const int tt = pf / nf, fl = pf - tt * nf;
        float acc[8];
        {
            const float4 b0 = __ldg(reinterpret_cast<const float4*>(a.bias) + 2 * og);
            const float4 b1 = __ldg(reinterpret_cast<const float4*>(a.bias) + 2 * og + 1);
            acc[0] = b0.x; acc[1] = b0.y; acc[2] = b0.z; acc[3] = b0.w;
            acc[4] = b1.x; acc[5] = b1.y; acc[6] = b1.z; acc[7] = b1.w;
        }
        for (int kt = 0; kt < 3; ++kt) {
            const float* ip = in_s + (tt + kt) * Cin * FP + fl;
            const float* wp = w_s + (kt * Cin * 3) * C + og * 8;
#pragma unroll 3
            for (int c = 0; c < Cin; ++c) {
#pragma unroll
                for (int kf = 0; kf < 3; ++kf) {
                    const float v = ip[c * FP + kf];
                    const float4 w0 = ld4(wp + (c * 3 + kf) * C);
                    const float4 w1 = ld4(wp + (c * 3 + kf) * C + 4);
                    acc[0] = fmaf(v, w0.x, acc[0]); acc[1] = fmaf(v, w0.y, acc[1]);
                    acc[2] = fmaf(v, w0.z, acc[2]); acc[3] = fmaf(v, w0.w, acc[3]);
                    acc[4] = fmaf(v, w1.x, acc[4]); acc[5] = fmaf(v, w1.y, acc[5]);
                    acc[6] = fmaf(v, w1.z, acc[6]); acc[7] = fmaf(v, w1.w, acc[7]);
                }
            }
        }
        if (a.ln_g) {
            float s1 = 0.f;
#pragma unroll
            for (int j = 0; j < 8; ++j) s1 += acc[j];
            const float mean = group_sum<NOG>(s1) * (1.0f / C);
            float s2 = 0.f;
#pragma unroll
            for (int j = 0; j < 8; ++j) { acc[j] -= mean; s2 = fmaf(acc[j], acc[j], s2); }
            const float rstd = rsqrtf(group_sum<NOG>(s2) * (1.0f / C) + kLnEps);
#pragma unroll
            for (int j = 0; j < 8; ++j)
                acc[j] = fmaf(acc[j] * rstd, __ldg(a.ln_g + og * 8 + j), __ldg(a.ln_b + og * 8 + j));
        }
        if (valid) {
            float* dst = a.x + ((size_t)(b * a.T + t0 + tt) * F + f0 + fl) * C + og * 8;
            st4(dst, make_float4(acc[0], acc[1], acc[2], acc[3]));
            st4(dst + 4, make_float4(acc[4], acc[5], acc[6], acc[7]));
        }
    }

    }   // PP == 1

    if (t0 + nvalid == a.T) {           // this CTA holds the last two frames of [history ; feats]: new conv_buf
        float* dst = a.conv_buf_out + (size_t)b * Cin * 2 * F;
        for (int i = tid; i < Cin * 2 * nf; i += blockDim.x) {
            const int c = i / (2 * nf), r = i - c * 2 * nf;
            const int j = r / nf, fl = r - j * nf;
            dst[(size_t)c * 2 * F + (size_t)j * F + f0 + fl] = in_s[((nvalid + j) * Cin + c) * FP + 1 + fl];
        }
    }
}

// =============================================================================================================
// distance embedding + FiLM parameters
// =============================================================================================================
// One CTA per (utterance, FiLM layer): e = dis @ W^T (F*Din values, recomputed per layer: it is tiny), LayerNorm (per
// Din group, or over the whole vector for the Linear variants), then scale/shift = the layer's two 1x1 convs.
// film[(j*2 + k)*B*F*C + (b*F + f)*C + c].
__global__ void __launch_bounds__(256) film_params_kernel(const sb_film_args a) {
    SB_DYN_SMEM(float, e);              // [F*Din] (+ 64 floats of reduction scratch)
    __shared__ float red[64];
    const int tid = threadIdx.x, b = blockIdx.x;
    const int F = a.F, Din = a.Din, C = a.C, n = F * Din;
    pdl_trigger();
    pdl_wait();
    const float d0 = __ldg(a.dis + b * 3), d1 = __ldg(a.dis + b * 3 + 1), d2 = __ldg(a.dis + b * 3 + 2);
    for (int i = tid; i < n; i += 256)
        e[i] = d0 * __ldg(a.emb_w + 3 * i) + d1 * __ldg(a.emb_w + 3 * i + 1) + d2 * __ldg(a.emb_w + 3 * i + 2);
    __syncthreads();
    if (a.emb_mode == SB_EMB_CONV) {
        for (int f = tid; f < F; f += 256) {
            float s1 = 0.f;
            for (int d = 0; d < Din; ++d) s1 += e[f * Din + d];
            const float mean = s1 / Din;
            float s2 = 0.f;
            for (int d = 0; d < Din; ++d) { const float t = e[f * Din + d] - mean; s2 += t * t; }
            const float rstd = rsqrtf(s2 / Din + kLnEps);
            for (int d = 0; d < Din; ++d)
                e[f * Din + d] = (e[f * Din + d] - mean) * rstd * __ldg(a.emb_ln_g + d) + __ldg(a.emb_ln_b + d);
        }
    } else {
        float s1 = 0.f;
        for (int i = tid; i < n; i += 256) s1 += e[i];
        s1 = group_sum<32>(s1);
        if ((tid & 31) == 0) red[tid >> 5] = s1;
        __syncthreads();
        float tot = 0.f;
        for (int w = 0; w < 8; ++w) tot += red[w];
        const float mean = tot / n;
        float s2 = 0.f;
        for (int i = tid; i < n; i += 256) { const float t = e[i] - mean; s2 += t * t; }
        s2 = group_sum<32>(s2);
        if ((tid & 31) == 0) red[32 + (tid >> 5)] = s2;
        __syncthreads();
        float tot2 = 0.f;
        for (int w = 0; w < 8; ++w) tot2 += red[32 + w];
        const float rstd = rsqrtf(tot2 / n + kLnEps);
        for (int i = tid; i < n; i += 256)
            e[i] = (e[i] - mean) * rstd * __ldg(a.emb_ln_g + i) + __ldg(a.emb_ln_b + i);
    }
    __syncthreads();
    const int per = F * C;
    const int j = blockIdx.y;
    for (int i = tid; i < 2 * per; i += 256) {
        const int k = i / per, r = i - k * per;
        const int jk = 2 * j + k;
        const int f = r / C, c = r - f * C;
        const float* ww = (k ? a.b_w : a.w_w) + ((size_t)j * C + c) * Din;
        float v = __ldg((k ? a.b_b : a.w_b) + j * C + c);
        for (int d = 0; d < Din; ++d) {
            const float ev = (a.emb_mode == SB_EMB_CONV) ? e[f * Din + d] : e[d * F + f];
            v = fmaf(__ldg(ww + d), ev, v);
        }
        a.film[((size_t)jk * a.B + b) * per + r] = v;
    }
}

}  // namespace sb

extern "C" int sb_stft_features_fwd(const sb_stft_args* p, void* stream) {
    using namespace sb;
    SB_REQUIRE(p && p->wave && p->filt && p->feats, SB_E_BADARG, "sb_stft_features_fwd: null pointer");
    SB_REQUIRE(p->B > 0 && p->T > 0 && p->M > 0 && p->F > 0, SB_E_BADARG, "sb_stft_features_fwd: bad sizes");
    SB_REQUIRE(p->M <= SB_MAX_MICS, SB_E_UNSUPP, "sb_stft_features_fwd: at most %d microphones (got %d)", SB_MAX_MICS, p->M);
    SB_REQUIRE(p->n_fft % 4 == 0 && p->stride % 4 == 0, SB_E_UNSUPP,
               "sb_stft_features_fwd: n_fft and stride must be multiples of 4 (got %d, %d)", p->n_fft, p->stride);
    SB_REQUIRE(kStftTT * kStftFC * p->Cin + 16 * p->M * 32 <= 2 * kStftFC * (p->n_fft + 4), SB_E_UNSUPP,
               "sb_stft_features_fwd: n_fft=%d too small for Cin=%d", p->n_fft, p->Cin);
    SB_REQUIRE(p->F == p->n_fft / 2 + 1, SB_E_BADARG, "sb_stft_features_fwd: F must be n_fft/2+1");
    SB_REQUIRE((long long)(p->T - 1) * p->stride + p->n_fft <= p->n_samples, SB_E_BADARG, "sb_stft_features_fwd: T frames do not fit n_samples");
    int feat = 0;
    if (p->feat_mode == SB_FEAT_OMNI) feat = 3 * (p->M - 1);
    else if (p->feat_mode == SB_FEAT_DIRECTIONAL) {
        SB_REQUIRE(p->M == 6, SB_E_UNSUPP, "directional features need 6 microphones (got %d)", p->M);
        feat = 3 * (p->M - 1) - 1;
    } else SB_REQUIRE(p->feat_mode == SB_FEAT_NONE, SB_E_BADARG, "unknown feat_mode %d", p->feat_mode);
    SB_REQUIRE(p->Cin == 2 * p->M + feat, SB_E_BADARG, "sb_stft_features_fwd: Cin=%d does not match 2M+features=%d", p->Cin, 2 * p->M + feat);
    SB_REQUIRE(!p->spec || (p->n_src > 0 && p->n_src <= p->M), SB_E_BADARG, "sb_stft_features_fwd: bad n_src");
    cudaStream_t st = (cudaStream_t)stream;
    switch (p->M) {
        case 1: return launch_stft<1>(*p, st);
        case 2: return launch_stft<2>(*p, st);
        case 3: return launch_stft<3>(*p, st);
        case 4: return launch_stft<4>(*p, st);
        case 5: return launch_stft<5>(*p, st);
        case 6: return launch_stft<6>(*p, st);
        case 7: return launch_stft<7>(*p, st);
        default: return launch_stft<8>(*p, st);
    }
}

extern "C" int sb_conv_in_fwd(const sb_conv_in_args* p, void* stream) {
    using namespace sb;
    SB_REQUIRE(p && p->feats && p->conv_buf_in && p->conv_buf_out && p->w_pack && p->bias && p->x, SB_E_BADARG, "sb_conv_in_fwd: null pointer");
    SB_REQUIRE(p->B > 0 && p->T > 0 && p->F > 0 && p->Cin > 0, SB_E_BADARG, "sb_conv_in_fwd: bad sizes");
    SB_REQUIRE(p->C == 16 || p->C == 32, SB_E_UNSUPP, "sb_conv_in_fwd: C must be 16 or 32 (got %d)", p->C);
    SB_REQUIRE((p->ln_g == nullptr) == (p->ln_b == nullptr), SB_E_BADARG, "sb_conv_in_fwd: ln_g/ln_b must come together");
    SB_REQUIRE(p->conv_buf_in != p->conv_buf_out, SB_E_BADARG, "sb_conv_in_fwd: conv_buf_in and conv_buf_out must not alias");
    // offline: 4 frames x all bins per CTA; streaming-sized calls: 1 frame x 16 bins per CTA so the few frames still
    // spread over the whole chip (B * T * ceil(F/16) CTAs)
    if (front_tc_enabled() && conv_in_tc_supported(*p)) return run_conv_in_tc(*p, (cudaStream_t)stream);
    const bool big = p->T >= 4;
    const int TT = big ? 4 : 1;
    const int FC = big ? p->F : 16;
    const int FP = big ? (FC + 3) / 4 * 4 + 8 : FC + 2;        // PP = 4 reads columns fl0 .. fl0 + 7 of a 16-byte aligned row
    const int threads = big ? 640 : 128;
    const size_t smem = ((size_t)9 * p->Cin * p->C + (size_t)(TT + 2) * p->Cin * FP) * sizeof(float);
    SB_REQUIRE(smem <= 227 * 1024, SB_E_SMEM, "sb_conv_in_fwd: Cin=%d F=%d needs %zu bytes of shared memory", p->Cin, p->F, smem);
    dim3 grid(ceil_div(p->T, TT), p->B, ceil_div(p->F, FC));
    cudaStream_t st = (cudaStream_t)stream;
    if (big) {
        if (p->C == 32) return launch("conv_in", conv_in_kernel<32, 4>, grid, dim3(threads), smem, st, *p, TT, FC, FP);
        return launch("conv_in", conv_in_kernel<16, 4>, grid, dim3(threads), smem, st, *p, TT, FC, FP);
    }
    if (p->C == 32) return launch("conv_in", conv_in_kernel<32, 1>, grid, dim3(threads), smem, st, *p, TT, FC, FP);
    return launch("conv_in", conv_in_kernel<16, 1>, grid, dim3(threads), smem, st, *p, TT, FC, FP);
}

extern "C" int sb_film_params_fwd(const sb_film_args* p, void* stream) {
    using namespace sb;
    SB_REQUIRE(p && p->dis && p->emb_w && p->emb_ln_g && p->emb_ln_b && p->w_w && p->w_b && p->b_w && p->b_b && p->film,
               SB_E_BADARG, "sb_film_params_fwd: null pointer");
    SB_REQUIRE(p->B > 0 && p->F > 0 && p->C > 0 && p->Din > 0 && p->n_layers > 0, SB_E_BADARG, "sb_film_params_fwd: bad sizes");
    SB_REQUIRE(p->emb_mode == SB_EMB_CONV || p->emb_mode == SB_EMB_LINEAR, SB_E_BADARG, "sb_film_params_fwd: bad emb_mode");
    const size_t smem = (size_t)p->F * p->Din * sizeof(float);
    return launch("film_params", film_params_kernel, dim3(p->B, p->n_layers), dim3(256), smem, (cudaStream_t)stream, *p);
}
