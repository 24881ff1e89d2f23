// Back-end of the separator: causal deconv to the output spectrum (+ optional spectral masking), iSTFT, overlap-add.
//
// Reference spans replaced (DE3 = src/models/tfgridnet_realtime_clean_dis_embd3/tfgridnet_causal.py):
//   deconv_spec_kernel : history cat :517-518, ConvTranspose2d(C -> 2S, (3,3), pad (2,1)) :401,520, re/im view :521-526,
//                        spectral masking :529-530, new deconv_buf / istft_buf :518,534
//   istft_ola_kernel   : previous-frame cat :533, self.dec (asteroid Decoder = conv_transpose1d with the [2F,1,n_fft]
//                        basis buffer, stride = hop) :537, crops :538,542
#include "sb_common.cuh"

namespace sb {

// out[o,t,f] = bias[o] + sum_{c,kt,kf} W[c,o,kt,kf] * in[c, t+2-kt, f+1-kf]   (in = [2 history frames ; x])
// Eight lanes share one (t, f): lane cq reads channels 4cq..4cq+3 of the nine taps as float4 (each tap is one full
// 128-byte line per position) with its 4 x (2S) x 9 weights in registers; an 8-lane shuffle tree finishes the sum.
template <int C, int NO>
__global__ void __launch_bounds__(256) deconv_spec_kernel(const sb_backend_args a) {
    constexpr int LPP = C / 4;                  // lanes per position
    constexpr int PPB = 256 / LPP;              // positions per CTA pass
    const int tid = threadIdx.x;
    const int cq = tid % LPP, pslot = tid / LPP;
    const int b = blockIdx.y, F = a.F, T = a.T;

    float w[NO][9][4];
#pragma unroll
    for (int o = 0; o < NO; ++o)
#pragma unroll
        for (int k = 0; k < 9; ++k)
#pragma unroll
            for (int j = 0; j < 4; ++j) w[o][k][j] = __ldg(a.w + ((size_t)(4 * cq + j) * NO + o) * 9 + k);
    float bias[NO];
#pragma unroll
    for (int o = 0; o < NO; ++o) bias[o] = __ldg(a.bias + o);
    pdl_trigger();
    pdl_wait();

    const int n_pos = T * F;
    const float* xb = a.x + (size_t)b * T * F * C;
    const float* hist = a.deconv_buf_in + (size_t)b * C * 2 * F;
    for (int base = blockIdx.x * PPB; base < n_pos; base += gridDim.x * PPB) {
        const int pos = base + pslot;
        const bool valid = pos < n_pos;
        const int pp = valid ? pos : 0;
        const int t = pp / F, f = pp - t * F;
        float acc[NO];
#pragma unroll
        for (int o = 0; o < NO; ++o) acc[o] = 0.f;
#pragma unroll
        for (int kt = 0; kt < 3; ++kt) {
            const int ft = t - kt;              // actual frame of tap kt (in-index t+2-kt, minus the 2 history frames)
#pragma unroll
            for (int kf = 0; kf < 3; ++kf) {
                const int ff = f + 1 - kf;
                if (ff < 0 || ff >= F) continue;
                float4 v;
                if (ft >= 0) {
                    v = __ldg(reinterpret_cast<const float4*>(xb + ((size_t)ft * F + ff) * C) + cq);
                } else {                        // history frame 2+ft of deconv_buf [C][2][F]
                    const float* hp = hist + (size_t)(2 + ft) * F + ff;
                    v.x = __ldg(hp + (size_t)(4 * cq + 0) * 2 * F); v.y = __ldg(hp + (size_t)(4 * cq + 1) * 2 * F);
                    v.z = __ldg(hp + (size_t)(4 * cq + 2) * 2 * F); v.w = __ldg(hp + (size_t)(4 * cq + 3) * 2 * F);
                }
#pragma unroll
                for (int o = 0; o < NO; ++o) {
                    acc[o] = fmaf(v.x, w[o][kt * 3 + kf][0], acc[o]); acc[o] = fmaf(v.y, w[o][kt * 3 + kf][1], acc[o]);
                    acc[o] = fmaf(v.z, w[o][kt * 3 + kf][2], acc[o]); acc[o] = fmaf(v.w, w[o][kt * 3 + kf][3], acc[o]);
                }
            }
        }
#pragma unroll
        for (int o = 0; o < NO; ++o) acc[o] = group_sum<LPP>(acc[o]) + bias[o];
        if (valid && cq == 0) {
#pragma unroll
            for (int o = 0; o < NO; ++o) {      // channel o -> (source s = o/2, re/im = o%2)   (view :521)
                const size_t in_frame = (size_t)(o & 1) * F + f;
                float v = acc[o];
                if (a.mask_spec) v *= __ldg(a.mask_spec + ((size_t)(b * T + t) * (NO / 2) + (o >> 1)) * 2 * F + in_frame);
                a.ws[(((size_t)b * (NO / 2) + (o >> 1)) * T + t) * 2 * F + in_frame] = v;      // ws [B][S][T][2F]
                if (t == T - 1) a.istft_buf_out[((size_t)b * (NO / 2) + (o >> 1)) * 2 * F + (size_t)(o & 1) * F + f] = v;
            }
        }
    }

    if (blockIdx.x == 0) {                      // new deconv_buf = last two frames of [history ; x], layout [C][2][F]
        float* dst = a.deconv_buf_out + (size_t)b * C * 2 * F;
        for (int i = tid; i < C * 2 * F; i += 256) {
            const int c = i / (2 * F), r = i - c * 2 * F;
            const int j = r / F, f = r - j * F;
            const int ft = T - 2 + j;
            dst[i] = (ft >= 0) ? __ldg(xb + ((size_t)ft * F + f) * C + c) : __ldg(hist + (size_t)c * 2 * F + (size_t)(2 + ft) * F + f);
        }
    }
}

// The same arithmetic with the input tile in shared memory: CTA = (4 frames + the 2 frames of history, utterance), the
// TT + 2 rows of x are one contiguous block of HBM (18 560 B per frame) copied with 16-byte loads, after which the nine taps of
// a position are LDS.128 instead of L2 round trips (the kernel above waits on the long scoreboard for 59 % of its samples,
// profiles/r02_prof_frontback.txt).  111 KB of shared memory: two CTAs per SM.  Used for calls of more than kIstftTT frames.
constexpr int kDeconvTT = 4;
template <int C, int NO>
__global__ void __launch_bounds__(256, 2) deconv_spec_tile_kernel(const sb_backend_args a) {
    constexpr int LPP = C / 4, PPB = 256 / LPP;
    SB_DYN_SMEM(float, in_s);                   // [kDeconvTT + 2][F][C]; frame slot j <-> frame t0 - 2 + j
    const int tid = threadIdx.x;
    const int cq = tid % LPP, pslot = tid / LPP;
    const int b = blockIdx.y, F = a.F, T = a.T;
    const int t0 = blockIdx.x * kDeconvTT;
    const int nvalid = min(kDeconvTT, T - t0);

    float w[NO][9][4];
#pragma unroll
    for (int o = 0; o < NO; ++o)
#pragma unroll
        for (int k = 0; k < 9; ++k)
#pragma unroll
            for (int j = 0; j < 4; ++j) w[o][k][j] = __ldg(a.w + ((size_t)(4 * cq + j) * NO + o) * 9 + k);
    float bias[NO];
#pragma unroll
    for (int o = 0; o < NO; ++o) bias[o] = __ldg(a.bias + o);
    pdl_trigger();
    pdl_wait();

    const float* xb = a.x + (size_t)b * T * F * C;
    const float* hist = a.deconv_buf_in + (size_t)b * C * 2 * F;
    const int fc4 = F * C / 4;                  // float4 per frame
    for (int j = 0; j < nvalid + 2; ++j) {
        const int ft = t0 - 2 + j;
        float* dst = in_s + (size_t)j * F * C;
        if (ft >= 0) {
            const float4* src = reinterpret_cast<const float4*>(xb + (size_t)ft * F * C);
            for (int i = tid; i < fc4; i += 256) st4(dst + 4 * i, ldg4_stream(reinterpret_cast<const float*>(src + i)));
        } else {                                // history frame 2 + ft of deconv_buf [C][2][F]: a transposing gather
            const float* hp = hist + (size_t)(2 + ft) * F;
            for (int i = tid; i < F * C; i += 256) {
                const int c = i / F, f = i - c * F;
                dst[f * C + c] = __ldg(hp + (size_t)c * 2 * F + f);
            }
        }
    }
    __syncthreads();

    const int n_pos = nvalid * F;
    for (int base = 0; base < n_pos; base += PPB) {
        const int pos = base + pslot;
        const bool valid = pos < n_pos;
        const int pp = valid ? pos : 0;
        const int tt = pp / F, f = pp - tt * F;
        const int t = t0 + tt;
        float acc[NO];
#pragma unroll
        for (int o = 0; o < NO; ++o) acc[o] = 0.f;
#pragma unroll
        for (int kt = 0; kt < 3; ++kt) {
            const float* fr = in_s + (size_t)(tt + 2 - kt) * F * C;      // frame t - kt sits in slot tt + 2 - kt
#pragma unroll
            for (int kf = 0; kf < 3; ++kf) {
                const int ff = f + 1 - kf;
                if (ff < 0 || ff >= F) continue;
                const float4 v = ld4(fr + (size_t)ff * C + 4 * cq);
#pragma unroll
                for (int o = 0; o < NO; ++o) {
                    acc[o] = fmaf(v.x, w[o][kt * 3 + kf][0], acc[o]); acc[o] = fmaf(v.y, w[o][kt * 3 + kf][1], acc[o]);
                    acc[o] = fmaf(v.z, w[o][kt * 3 + kf][2], acc[o]); acc[o] = fmaf(v.w, w[o][kt * 3 + kf][3], acc[o]);
                }
            }
        }
#pragma unroll
        for (int o = 0; o < NO; ++o) acc[o] = group_sum<LPP>(acc[o]) + bias[o];
        if (valid && cq == 0) {
#pragma unroll
            for (int o = 0; o < NO; ++o) {      // channel o -> (source s = o/2, re/im = o%2)   (view :521)
                const size_t in_frame = (size_t)(o & 1) * F + f;
                float v = acc[o];
                if (a.mask_spec) v *= __ldg(a.mask_spec + ((size_t)(b * T + t) * (NO / 2) + (o >> 1)) * 2 * F + in_frame);
                a.ws[(((size_t)b * (NO / 2) + (o >> 1)) * T + t) * 2 * F + in_frame] = v;      // ws [B][S][T][2F]
                if (t == T - 1) a.istft_buf_out[((size_t)b * (NO / 2) + (o >> 1)) * 2 * F + (size_t)(o & 1) * F + f] = v;
            }
        }
    }

    if (t0 + nvalid == T) {                     // this CTA holds the last two frames of [history ; x]: new deconv_buf [C][2][F]
        float* dst = a.deconv_buf_out + (size_t)b * C * 2 * F;
        for (int i = tid; i < C * 2 * F; i += 256) {
            const int c = i / (2 * F), r = i - c * 2 * F;
            const int j = r / F, f = r - j * F;
            dst[i] = in_s[((size_t)(nvalid + j) * F + f) * C + c];      // frame T - 2 + j sits in slot nvalid + j
        }
    }
}

// wave[t*hop + r] = sum_k spec_t[k] basis[k][r]  +  (r < n_fft-hop) sum_k spec_{t-1}[k] basis[k][hop + r]
// CTA = (tile of TT frames, (utterance, source)); thread r owns sample r of every frame in the tile (+ the carried
// one): the basis value basis[k][r] is read once per k (coalesced over r) and reused for TT+1 frames whose spectra are
// broadcast from shared memory.  The two overlapping halves meet through shared memory.
constexpr int kIstftTT = 8;

__global__ void __launch_bounds__(1024) istft_ola_kernel(const sb_backend_args a) {
    constexpr int NF = kIstftTT + 1, NFP = 12;
    SB_DYN_SMEM(float, smem);
    const int F2 = 2 * a.F, n_fft = a.n_fft, hop = a.stride;
    float* sp = smem;                           // [F2][NFP]  spectra of frames t0-1 .. t0+nv-1
    float* ola = sp + F2 * NFP;                 // [NF][n_fft]
    const int tid = threadIdx.x;
    const int t0 = blockIdx.x * kIstftTT, bs = blockIdx.y;
    const int nv = min(kIstftTT, a.T - t0);
    pdl_trigger();
    pdl_wait();
    for (int i = tid; i < F2 * NF; i += blockDim.x) {
        const int fr = i / F2, k = i - fr * F2;
        float v = 0.f;
        if (fr <= nv) {
            const int ft = t0 - 1 + fr;
            v = (ft >= 0) ? ldg1_stream(a.ws + ((size_t)bs * a.T + ft) * F2 + k)
                          : ldg1_stream(a.istft_buf_in + (size_t)bs * F2 + k);
        }
        sp[k * NFP + fr] = v;
    }
    __syncthreads();
    // the k loop is a chain of dependent basis loads: KS thread groups (blockDim = KS x n_fft rounded to warps) take a
    // slice of k each and meet in shared memory (ola [KS][NF][n_fft])
    const int nth = (n_fft + 31) & ~31, KS = blockDim.x / nth;
    const int grp = tid / nth, r0 = tid - grp * nth;
    if (grp < KS && r0 < n_fft) {
        float acc[NF];
#pragma unroll
        for (int i = 0; i < NF; ++i) acc[i] = 0.f;
        const float* bp = a.filt + r0;
        const int kper = (F2 + KS - 1) / KS, k0 = grp * kper, k1 = min(F2, k0 + kper);
        // every CTA reads the same basis rows: start each CTA at a different k so that they do not all queue on the same
        // L2 lines at the same moment (the kernel waits on these loads for 79 % of its samples, profiles/r02_prof_frontback.txt)
        const int nk = k1 - k0, rot = nk > 0 ? (int)((blockIdx.x * 13u + blockIdx.y * 29u) % (unsigned)nk) : 0;
#pragma unroll 16
        for (int i = 0; i < nk; ++i) {
            const int k = k0 + (i + rot >= nk ? i + rot - nk : i + rot);
            const float bv = __ldg(bp + (size_t)k * n_fft);
            const float4 s0 = ld4(sp + k * NFP), s1 = ld4(sp + k * NFP + 4);
            const float s8 = sp[k * NFP + 8];
            acc[0] = fmaf(s0.x, bv, acc[0]); acc[1] = fmaf(s0.y, bv, acc[1]);
            acc[2] = fmaf(s0.z, bv, acc[2]); acc[3] = fmaf(s0.w, bv, acc[3]);
            acc[4] = fmaf(s1.x, bv, acc[4]); acc[5] = fmaf(s1.y, bv, acc[5]);
            acc[6] = fmaf(s1.z, bv, acc[6]); acc[7] = fmaf(s1.w, bv, acc[7]);
            acc[8] = fmaf(s8, bv, acc[8]);
        }
#pragma unroll
        for (int i = 0; i < NF; ++i) ola[(grp * NF + i) * n_fft + r0] = acc[i];
    }
    __syncthreads();
    const int look = n_fft - hop;
    float* dst = a.wave_out + (size_t)bs * a.T * hop + (size_t)t0 * hop;
    for (int i = tid; i < nv * hop; i += blockDim.x) {
        const int fr = i / hop, r = i - fr * hop;
        float v = 0.f;
        for (int g2 = 0; g2 < KS; ++g2) {
            v += ola[(g2 * NF + fr + 1) * n_fft + r];
            if (r < look) v += ola[(g2 * NF + fr) * n_fft + hop + r];
        }
        dst[i] = v;
    }
}

// Streaming-sized calls (T <= kIstftTT frames): deconv, masking, iSTFT and overlap-add in ONE launch per utterance; the
// output spectrum never leaves shared memory.  Same arithmetic and mappings as the two kernels above.
template <int C, int NO>
__global__ void __launch_bounds__(512) backend_small_kernel(const sb_backend_args a) {
    constexpr int LPP = C / 4, NG = 512 / LPP, NS = NO / 2, NF = kIstftTT + 1, NFP = 12;
    SB_DYN_SMEM(float, smem);
    const int F = a.F, F2 = 2 * F, T = a.T, n_fft = a.n_fft, hop = a.stride;
    float* w_s = smem;                          // [C][NO][9]
    float* sp = w_s + C * NO * 9;               // [NS][F2][NFP]: slot 0 = carried frame, slot 1 + t = frame t
    float* ola = sp + NS * F2 * NFP;            // [NF][n_fft]
    const int tid = threadIdx.x, b = blockIdx.x;
    const int cq = tid % LPP, grp = tid / LPP;
    for (int i = tid; i < C * NO * 9; i += 512) w_s[i] = __ldg(a.w + i);
    float bias[NO];
#pragma unroll
    for (int o = 0; o < NO; ++o) bias[o] = __ldg(a.bias + o);
    pdl_trigger();
    pdl_wait();
    for (int i = tid; i < NS * F2; i += 512) {
        float* row = sp + i * NFP;
        row[0] = ldg1_stream(a.istft_buf_in + (size_t)b * NS * F2 + i);
        for (int fr = T + 1; fr < NFP; ++fr) row[fr] = 0.f;
    }
    __syncthreads();
    float w[NO][9][4];
#pragma unroll
    for (int o = 0; o < NO; ++o)
#pragma unroll
        for (int k = 0; k < 9; ++k)
#pragma unroll
            for (int j = 0; j < 4; ++j) w[o][k][j] = w_s[((4 * cq + j) * NO + o) * 9 + k];

    const int n_pos = T * F;
    const float* xb = a.x + (size_t)b * T * F * C;
    const float* hist = a.deconv_buf_in + (size_t)b * C * 2 * F;
    for (int base = 0; base < n_pos; base += NG) {
        const int pos = base + grp;
        const bool valid = pos < n_pos;
        const int pp = valid ? pos : 0;
        const int t = pp / F, f = pp - t * F;
        float acc[NO];
#pragma unroll
        for (int o = 0; o < NO; ++o) acc[o] = 0.f;
#pragma unroll
        for (int kt = 0; kt < 3; ++kt) {
            const int ft = t - kt;
#pragma unroll
            for (int kf = 0; kf < 3; ++kf) {
                const int ff = f + 1 - kf;
                if (ff < 0 || ff >= F) continue;
                float4 v;
                if (ft >= 0) {
                    v = __ldg(reinterpret_cast<const float4*>(xb + ((size_t)ft * F + ff) * C) + cq);
                } else {
                    const float* hp = hist + (size_t)(2 + ft) * F + ff;
                    v.x = __ldg(hp + (size_t)(4 * cq + 0) * 2 * F); v.y = __ldg(hp + (size_t)(4 * cq + 1) * 2 * F);
                    v.z = __ldg(hp + (size_t)(4 * cq + 2) * 2 * F); v.w = __ldg(hp + (size_t)(4 * cq + 3) * 2 * F);
                }
#pragma unroll
                for (int o = 0; o < NO; ++o) {
                    acc[o] = fmaf(v.x, w[o][kt * 3 + kf][0], acc[o]); acc[o] = fmaf(v.y, w[o][kt * 3 + kf][1], acc[o]);
                    acc[o] = fmaf(v.z, w[o][kt * 3 + kf][2], acc[o]); acc[o] = fmaf(v.w, w[o][kt * 3 + kf][3], acc[o]);
                }
            }
        }
#pragma unroll
        for (int o = 0; o < NO; ++o) acc[o] = group_sum<LPP>(acc[o]) + bias[o];
        if (valid && cq == 0) {
#pragma unroll
            for (int o = 0; o < NO; ++o) {
                const int in_frame = (o & 1) * F + f;
                float v = acc[o];
                if (a.mask_spec) v *= __ldg(a.mask_spec + ((size_t)(b * T + t) * NS + (o >> 1)) * F2 + in_frame);
                sp[((o >> 1) * F2 + in_frame) * NFP + 1 + t] = v;
                if (t == T - 1) a.istft_buf_out[((size_t)b * NS + (o >> 1)) * F2 + in_frame] = v;
            }
        }
    }
    {                                           // new deconv_buf = last two frames of [history ; x], layout [C][2][F]
        float* dst = a.deconv_buf_out + (size_t)b * C * 2 * F;
        for (int i = tid; i < C * 2 * F; i += 512) {
            const int c = i / (2 * F), r = i - c * 2 * F;
            const int j = r / F, f = r - j * F;
            const int ft = T - 2 + j;
            dst[i] = (ft >= 0) ? __ldg(xb + ((size_t)ft * F + f) * C + c) : __ldg(hist + (size_t)c * 2 * F + (size_t)(2 + ft) * F + f);
        }
    }
    __syncthreads();
    const int look = n_fft - hop;
    for (int s = 0; s < NS; ++s) {
        if (tid < n_fft) {
            float acc[NF];
#pragma unroll
            for (int i = 0; i < NF; ++i) acc[i] = 0.f;
            const float* bp = a.filt + tid;
            const float* sps = sp + s * F2 * NFP;
#pragma unroll 16
            for (int k = 0; k < F2; ++k) {
                const float bv = __ldg(bp + (size_t)k * n_fft);
                const float4 s0 = ld4(sps + k * NFP), s1 = ld4(sps + k * NFP + 4);
                const float s8 = sps[k * NFP + 8];
                acc[0] = fmaf(s0.x, bv, acc[0]); acc[1] = fmaf(s0.y, bv, acc[1]);
                acc[2] = fmaf(s0.z, bv, acc[2]); acc[3] = fmaf(s0.w, bv, acc[3]);
                acc[4] = fmaf(s1.x, bv, acc[4]); acc[5] = fmaf(s1.y, bv, acc[5]);
                acc[6] = fmaf(s1.z, bv, acc[6]); acc[7] = fmaf(s1.w, bv, acc[7]);
                acc[8] = fmaf(s8, bv, acc[8]);
            }
#pragma unroll
            for (int i = 0; i < NF; ++i) ola[i * n_fft + tid] = acc[i];
        }
        __syncthreads();
        float* dst = a.wave_out + ((size_t)b * NS + s) * T * hop;
        for (int i = tid; i < T * hop; i += 512) {
            const int fr = i / hop, r = i - fr * hop;
            float v = ola[(fr + 1) * n_fft + r];
            if (r < look) v += ola[fr * n_fft + hop + r];
            dst[i] = v;
        }
        __syncthreads();
    }
}

template <int C>
static int launch_backend(const sb_backend_args& a, cudaStream_t st) {
    if (a.T <= kIstftTT && a.n_fft <= 512) {           // streaming-sized call: one fused launch
        const size_t smem = ((size_t)C * 2 * a.n_src * 9 + (size_t)a.n_src * 2 * a.F * 12 + (size_t)(kIstftTT + 1) * a.n_fft) * sizeof(float);
        if (a.n_src == 1) return launch("backend_small", backend_small_kernel<C, 2>, dim3(a.B), dim3(512), smem, st, a);
        if (a.n_src == 2) return launch("backend_small", backend_small_kernel<C, 4>, dim3(a.B), dim3(512), smem, st, a);
        set_error("sb_backend_fwd: n_src must be 1 or 2 (got %d)", a.n_src);
        return SB_E_UNSUPP;
    }
    const int LPP = C / 4, PPB = 256 / LPP;
    const int n_pos = a.T * a.F;
    // every CTA first pulls its 72 weights into registers: give it several 32-position passes to amortise that over (one pass
    // per CTA made the launch 25x slower than its L1 traffic allows, profiles/r02_launches_bench.txt) - about four CTAs per SM
    int gx = ceil_div(n_pos, PPB);
    const int cap = ceil_div(4 * sm_count(), a.B);
    if (gx > cap) gx = cap;
    dim3 grid(gx, a.B);
    const size_t smem_t = (size_t)(kDeconvTT + 2) * a.F * C * sizeof(float);
    const bool tiled = smem_t <= 113 * 1024;    // shared-memory tiles (two CTAs per SM); the direct kernel for wider grids
    dim3 grid_t(ceil_div(a.T, kDeconvTT), a.B);
    switch (a.n_src) {
        case 1:
            if (tiled) SB_CHECK(launch("deconv_spec", deconv_spec_tile_kernel<C, 2>, grid_t, dim3(256), smem_t, st, a));
            else SB_CHECK(launch("deconv_spec", deconv_spec_kernel<C, 2>, grid, dim3(256), 0, st, a));
            break;
        case 2:
            if (tiled) SB_CHECK(launch("deconv_spec", deconv_spec_tile_kernel<C, 4>, grid_t, dim3(256), smem_t, st, a));
            else SB_CHECK(launch("deconv_spec", deconv_spec_kernel<C, 4>, grid, dim3(256), 0, st, a));
            break;
        default:
            set_error("sb_backend_fwd: n_src must be 1 or 2 (got %d)", a.n_src);
            return SB_E_UNSUPP;
    }
    const int nth = ceil_div(a.n_fft, 32) * 32;
    const int KS = nth * 3 <= 1024 ? 3 : (nth * 2 <= 1024 ? 2 : 1);
    const int threads = nth * KS;
    const size_t smem = ((size_t)2 * a.F * 12 + (size_t)KS * (kIstftTT + 1) * a.n_fft) * sizeof(float);
    return launch("istft_ola", istft_ola_kernel, dim3(ceil_div(a.T, kIstftTT), a.B * a.n_src), dim3(threads), smem, st, a);
}

}  // namespace sb

extern "C" int sb_backend_fwd(const sb_backend_args* p, void* stream) {
    using namespace sb;
    SB_REQUIRE(p && p->x && p->deconv_buf_in && p->deconv_buf_out && p->istft_buf_in && p->istft_buf_out && p->w &&
               p->bias && p->filt && p->wave_out && p->ws, SB_E_BADARG, "sb_backend_fwd: null pointer");
    SB_REQUIRE(p->B > 0 && p->T > 0 && p->F > 0, SB_E_BADARG, "sb_backend_fwd: bad sizes");
    SB_REQUIRE(p->C == 16 || p->C == 32, SB_E_UNSUPP, "sb_backend_fwd: C must be 16 or 32 (got %d)", p->C);
    SB_REQUIRE(p->F == p->n_fft / 2 + 1, SB_E_BADARG, "sb_backend_fwd: F must be n_fft/2+1");
    SB_REQUIRE(p->n_fft <= 2 * p->stride && p->n_fft > p->stride && p->n_fft <= 1024, SB_E_UNSUPP,
               "sb_backend_fwd: overlap-add is written for hop < n_fft <= 2*hop, n_fft <= 1024 (n_fft=%d hop=%d)", p->n_fft, p->stride);
    SB_REQUIRE(p->deconv_buf_in != p->deconv_buf_out && p->istft_buf_in != p->istft_buf_out, SB_E_BADARG,
               "sb_backend_fwd: state in/out buffers must not alias");
    cudaStream_t st = (cudaStream_t)stream;
    return p->C == 32 ? launch_backend<32>(*p, st) : launch_backend<16>(*p, st);
}
