// Sliding-window full-band self-attention of GridNetBlock (use_attn = true).
//
// Reference spans replaced (DE3 = src/models/tfgridnet_realtime_clean_dis_embd3/tfgridnet_causal.py):
//   attn_qkv_kernel  : attn_conv_{Q,K,V} = Linear -> PReLU -> head split -> LayerNorm(F*E | F*Vd)   :639-675, 856-863
//   attn_copy_kernel : K/V history cat and the new K_buf / V_buf                                    :864-873
//   attn_core_kernel : get_lookahead_mask-free local attention: for every frame softmax(q K_win^T / sqrt(F*E)) V_win
//                      over the last W frames (zero-initialised history is NOT masked)              :722-744, 875-888
//   attn_out_kernel  : head regroup, attn_concat_proj = Linear -> PReLU -> LayerNorm(F*C), residual  :676-684, 889-898
//
// fp32 SIMT throughout.  The core kernel tiles TQ consecutive queries per CTA so a K/V row is read once per tile
// instead of once per query (W + TQ - 1 rows instead of TQ * W).
#include <math.h>

#include "sb_common.cuh"

namespace sb {

// sb_attn_tc.cu: the same core on the tensor cores (tcgen05) for T >= 64
bool attn_core_tc_supported(const sb_attn_args& a);
int attn_core_tc(const sb_attn_args& a, const float* Q, const float* Kc, const float* Vc, float* AO, cudaStream_t st);

struct AttnWs {
    float *Q, *Kc, *Vc, *AO, *P;
    size_t total;
};

static size_t a64(size_t n) { return (n + 63) & ~size_t(63); }

static AttnWs attn_carve(float* base, int B, int T, int F, int C, int L, int E, int W) {
    AttnWs w{};
    const size_t BL = (size_t)B * L, DK = (size_t)F * E, DV = (size_t)F * (C / L);
    size_t off = 0;
    auto take = [&](size_t n) { float* p = base ? base + off : nullptr; off += a64(n); return p; };
    w.Q = take(BL * T * DK);
    w.Kc = take(BL * (T + W - 1) * DK);
    w.Vc = take(BL * (T + W - 1) * DV);
    w.AO = take(BL * T * DV);
    w.P = take(BL * W);                     // streaming path: softmax probabilities of the one query per head-row
    w.total = off;
    return w;
}

__device__ __forceinline__ float block_sum_256(float v, float* red) {      // red: 8 floats; all 256 threads call
    v = group_sum<32>(v);
    __syncthreads();
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
    __syncthreads();
    float t = 0.f;
#pragma unroll
    for (int w = 0; w < 8; ++w) t += red[w];
    return t;
}

// ---- Q, K, V projections + per-head LayerNorm; one CTA per (t, b, head) ---------------------------------------------
// Head l needs output rows l*E.. of Q and K and l*Vd.. of V: NOH = 2E + Vd rows of C weights.  The weights sit in shared
// memory TRANSPOSED ([c][NOH]: the threads of a warp differ in the output row, so they read consecutive words), x[f][:]
// is a broadcast.  The three LayerNorms (F*E, F*E, F*Vd elements) are block-wide, every thread a few elements.
__global__ void __launch_bounds__(256) attn_qkv_kernel(const sb_attn_args a, float* Q, float* Kc, float* Vc) {
    SB_DYN_SMEM(float, smem);
    __shared__ float red[8];
    const int F = a.F, C = a.C, L = a.L, E = a.E, Vd = C / L, W = a.W, T = a.T;
    const int NOH = 2 * E + Vd, DK = F * E, DV = F * Vd;
    float* xs = smem;                       // [F][C]
    float* wt = xs + F * C;                 // [C][NOH]  columns: Q (E), K (E), V (Vd) of this head
    float* bs = wt + C * NOH;               // [NOH]
    float* st = bs + ((NOH + 3) & ~3);      // staging: q [DK], k [DK], v [DV]
    const int tid = threadIdx.x;
    const int t = blockIdx.x, b = blockIdx.y, l = blockIdx.z;
    for (int i = tid; i < NOH * C; i += 256) {
        const int o = i / C, c = i - o * C;
        const float wv = o < E ? __ldg(a.q.w + (l * E + o) * C + c)
                               : (o < 2 * E ? __ldg(a.k.w + (l * E + o - E) * C + c) : __ldg(a.v.w + (l * Vd + o - 2 * E) * C + c));
        wt[c * NOH + o] = wv;
    }
    for (int o = tid; o < NOH; o += 256)
        bs[o] = o < E ? __ldg(a.q.b + l * E + o) : (o < 2 * E ? __ldg(a.k.b + l * E + o - E) : __ldg(a.v.b + l * Vd + o - 2 * E));
    const float sq = __ldg(a.q.prelu), sk = __ldg(a.k.prelu), sv = __ldg(a.v.prelu);
    pdl_trigger();
    pdl_wait();
    const float* xr = a.x + ((size_t)b * T + t) * F * C;
    for (int i = tid; i < F * C / 4; i += 256) st4(xs + 4 * i, ldg4_stream(xr + 4 * i));
    __syncthreads();
    for (int i = tid; i < F * NOH; i += 256) {
        const int f = i / NOH, o = i - f * NOH;
        float acc = bs[o];
        const float* xp = xs + f * C;
#pragma unroll 8
        for (int c = 0; c < C; ++c) acc = fmaf(xp[c], wt[c * NOH + o], acc);
        if (o < E) {
            st[f * E + o] = acc > 0.f ? acc : sq * acc;
        } else if (o < 2 * E) {
            st[DK + f * E + (o - E)] = acc > 0.f ? acc : sk * acc;
        } else {
            st[2 * DK + f * Vd + (o - 2 * E)] = acc > 0.f ? acc : sv * acc;
        }
    }
    __syncthreads();
    for (int which = 0; which < 3; ++which) {               // q, k, v rows of this head
        const int n = which == 2 ? DV : DK;
        const float* row = st + (which == 2 ? 2 * DK : which * DK);
        const sb_attn_proj& pj = which == 0 ? a.q : (which == 1 ? a.k : a.v);
        float s1 = 0.f;
        for (int i = tid; i < n; i += 256) s1 += row[i];
        const float mean = block_sum_256(s1, red) / n;
        float s2 = 0.f;
        for (int i = tid; i < n; i += 256) { const float d = row[i] - mean; s2 = fmaf(d, d, s2); }
        const float rstd = rsqrtf(block_sum_256(s2, red) / n + kLnEps);
        float* dst;
        if (which == 0) dst = Q + (((size_t)b * L + l) * T + t) * DK;
        else if (which == 1) dst = Kc + (((size_t)b * L + l) * (T + W - 1) + W - 1 + t) * DK;
        else dst = Vc + (((size_t)b * L + l) * (T + W - 1) + W - 1 + t) * DV;
#pragma unroll 4
        for (int i = tid; i < n; i += 256) dst[i] = fmaf((row[i] - mean) * rstd, __ldg(pj.ln_g + i), __ldg(pj.ln_b + i));
    }
}

// ---- row-block copies: history -> head of the concatenated K/V, tail of the concatenated K/V -> new history ---------
// dst[r][0..rows_copy*D) = src[r][src_off .. ), for r < n_outer
__global__ void __launch_bounds__(256) attn_copy_kernel(const float* src, float* dst, int n_outer, long long src_stride,
                                                        long long dst_stride, long long src_off, long long dst_off, long long n) {
    pdl_trigger();
    pdl_wait();
    const long long total = (long long)n_outer * n;
    for (long long i = (long long)blockIdx.x * 256 + threadIdx.x; i < total; i += (long long)gridDim.x * 256) {
        const long long r = i / n, k = i - r * n;
        dst[r * dst_stride + dst_off + k] = src[r * src_stride + src_off + k];
    }
}

// ---- windowed attention; CTA = (tile of TQ queries, head-row bl) -----------------------------------------------------
constexpr int kAttnTQ = 8;

__global__ void __launch_bounds__(256) attn_core_kernel(const sb_attn_args a, const float* Q, const float* Kc, const float* Vc, float* AO) {
    SB_DYN_SMEM(float, smem);
    constexpr int TQ = kAttnTQ;
    const int F = a.F, L = a.L, E = a.E, Vd = a.C / L, W = a.W, T = a.T;
    const int DK = F * E, DV = F * Vd, NJ = W + TQ - 1, PS = W + 1;
    float* qs = smem;                       // [TQ][DK]
    float* ps = qs + TQ * DK;               // [TQ][PS] logits -> probabilities
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int t0 = blockIdx.x * TQ, bl = blockIdx.y;
    const int nq = min(TQ, T - t0);
    pdl_trigger();
    pdl_wait();
    const float* qsrc = Q + ((size_t)bl * T + t0) * DK;
    for (int i = tid; i < nq * DK; i += 256) qs[i] = ldg1_stream(qsrc + i);
    __syncthreads();
    const float scale = 1.0f / sqrtf((float)DK);
    const float* kbase = Kc + ((size_t)bl * (T + W - 1) + t0) * DK;
    for (int jl = warp; jl < NJ; jl += 8) {                 // key row t0 + jl of the concatenated sequence
        const int i_lo = max(0, jl - W + 1), i_hi = min(nq - 1, jl);
        if (i_lo > i_hi) continue;                          // warp-uniform
        const float* kr = kbase + (size_t)jl * DK;
        for (int i = i_lo; i <= i_hi; ++i) {
            float acc = 0.f;
            for (int k = lane; k < DK; k += 32) acc = fmaf(qs[i * DK + k], __ldg(kr + k), acc);
            acc = group_sum<32>(acc);
            if (lane == 0) ps[i * PS + (jl - i)] = acc * scale;
        }
    }
    __syncthreads();
    for (int i = warp; i < nq; i += 8) {                    // softmax over the W keys of query i
        float m = -INFINITY;
        for (int w = lane; w < W; w += 32) m = fmaxf(m, ps[i * PS + w]);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
        float s = 0.f;
        for (int w = lane; w < W; w += 32) { const float e = expf(ps[i * PS + w] - m); ps[i * PS + w] = e; s += e; }
        const float inv = 1.0f / group_sum<32>(s);
        for (int w = lane; w < W; w += 32) ps[i * PS + w] *= inv;
    }
    __syncthreads();
    const float* vbase = Vc + ((size_t)bl * (T + W - 1) + t0) * DV;
    for (int c4 = tid; c4 < DV / 4; c4 += 256) {
        float4 acc[TQ];
#pragma unroll
        for (int i = 0; i < TQ; ++i) acc[i] = make_float4(0.f, 0.f, 0.f, 0.f);
        for (int jl = 0; jl < NJ; ++jl) {
            if (jl - (nq - 1) >= W) break;
            const float4 v = __ldg(reinterpret_cast<const float4*>(vbase + (size_t)jl * DV) + c4);
#pragma unroll
            for (int i = 0; i < TQ; ++i) {
                const int w = jl - i;
                if (i < nq && w >= 0 && w < W) {
                    const float p = ps[i * PS + w];
                    acc[i].x = fmaf(p, v.x, acc[i].x); acc[i].y = fmaf(p, v.y, acc[i].y);
                    acc[i].z = fmaf(p, v.z, acc[i].z); acc[i].w = fmaf(p, v.w, acc[i].w);
                }
            }
        }
#pragma unroll
        for (int i = 0; i < TQ; ++i)
            if (i < nq) st4(AO + ((size_t)bl * T + t0 + i) * DV + 4 * c4, acc[i]);
    }
}

// =============================================================================================================
// streaming path (T == 1): one query per head-row against its W-frame window
// =============================================================================================================
// A chunk's attention is a stream over the K / V history: (W-1) * (DK + DV) floats per head-row are read once, used for
// one dot product / one weighted sum, and written back one row earlier (the reference's cat + slice, DE3:864-873).  The two
// kernels below do exactly that and nothing else: HBM-bound, 2 * (W-1) * (DK + DV) * 4 bytes per head-row per chunk.

// scores + softmax; the K history moves up one row on the way.  CTA = head-row bl, warp = key rows w, w + 8, ...
__global__ void __launch_bounds__(256) attn_stream_scores_kernel(const sb_attn_args a, const float* Q, const float* Kc, float* P) {
    SB_DYN_SMEM(float, smem);
    const int F = a.F, E = a.E, W = a.W;
    const int DK = F * E, DK2 = DK / 2;
    float* qs = smem;                       // [DK]
    float* ps = qs + ((DK + 3) & ~3);       // [W]
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int bl = blockIdx.x;
    pdl_trigger();
    pdl_wait();
    for (int i = tid; i < DK; i += 256) qs[i] = ldg1_stream(Q + (size_t)bl * DK + i);
    __syncthreads();
    const float scale = 1.0f / sqrtf((float)DK);
    const float* hist = a.K_buf_in + (size_t)bl * (W - 1) * DK;
    const float* knew = Kc + ((size_t)bl * W + (W - 1)) * DK;
    float* hout = a.K_buf_out + (size_t)bl * (W - 1) * DK;
    constexpr int KI = 5;                                   // float2 per lane and row held in registers (DK <= 320)
    for (int r0 = warp; r0 < W; r0 += 16) {                 // two key rows per warp at a time, every load issued before use
        const int r1 = r0 + 8;
        const bool two = r1 < W;
        const float* src0 = r0 < W - 1 ? hist + (size_t)r0 * DK : knew;
        const float* src1 = !two ? knew : (r1 < W - 1 ? hist + (size_t)r1 * DK : knew);
        float* dst0 = r0 >= 1 ? hout + (size_t)(r0 - 1) * DK : nullptr;
        float* dst1 = two ? hout + (size_t)(r1 - 1) * DK : nullptr;
        float acc0 = 0.f, acc1 = 0.f;
        for (int kb = 0; kb < DK2; kb += 32 * KI) {
            float2 v0[KI], v1[KI];
#pragma unroll
            for (int i = 0; i < KI; ++i) {
                const int k = kb + 32 * i + lane;
                v0[i] = k < DK2 ? ldg2_stream(src0 + 2 * k) : make_float2(0.f, 0.f);
                v1[i] = k < DK2 ? ldg2_stream(src1 + 2 * k) : make_float2(0.f, 0.f);
            }
#pragma unroll
            for (int i = 0; i < KI; ++i) {
                const int k = kb + 32 * i + lane;
                if (k < DK2) {
                    const float2 q2 = ld2(qs + 2 * k);
                    acc0 = fmaf(q2.x, v0[i].x, acc0); acc0 = fmaf(q2.y, v0[i].y, acc0);
                    acc1 = fmaf(q2.x, v1[i].x, acc1); acc1 = fmaf(q2.y, v1[i].y, acc1);
                    if (dst0) st2(dst0 + 2 * k, v0[i]);
                    if (dst1) st2(dst1 + 2 * k, v1[i]);
                }
            }
        }
        acc0 = group_sum<32>(acc0);
        acc1 = group_sum<32>(acc1);
        if (lane == 0) {
            ps[r0] = acc0 * scale;
            if (two) ps[r1] = acc1 * scale;
        }
    }
    __syncthreads();
    if (warp == 0) {
        float m = -INFINITY;
        for (int w = lane; w < W; w += 32) m = fmaxf(m, ps[w]);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
        float s = 0.f;
        for (int w = lane; w < W; w += 32) { const float e = expf(ps[w] - m); ps[w] = e; s += e; }
        const float inv = 1.0f / group_sum<32>(s);
        for (int w = lane; w < W; w += 32) P[(size_t)bl * W + w] = ps[w] * inv;
    }
}

// out = P V; the V history moves up one row on the way.  grid (column chunks, head-rows), thread = one float4 column
__global__ void __launch_bounds__(128) attn_stream_pv_kernel(const sb_attn_args a, const float* Vc, const float* P, float* AO) {
    SB_DYN_SMEM(float, smem);               // [W] probabilities
    const int W = a.W, DV = a.F * (a.C / a.L), DV4 = DV / 4;
    const int bl = blockIdx.y, c4 = blockIdx.x * 128 + threadIdx.x;
    pdl_trigger();
    pdl_wait();
    for (int w = threadIdx.x; w < W; w += 128) smem[w] = P[(size_t)bl * W + w];
    __syncthreads();
    if (c4 >= DV4) return;
    const float4* hist = reinterpret_cast<const float4*>(a.V_buf_in + (size_t)bl * (W - 1) * DV) + c4;
    const float4* vnew = reinterpret_cast<const float4*>(Vc + ((size_t)bl * W + (W - 1)) * DV) + c4;
    float4* hout = reinterpret_cast<float4*>(a.V_buf_out + (size_t)bl * (W - 1) * DV) + c4;
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    constexpr int U = 8;                    // rows in flight per thread
    int r = 0;
    for (; r + U <= W - 1; r += U) {
        float4 v[U];
#pragma unroll
        for (int i = 0; i < U; ++i) v[i] = ldg4_stream(reinterpret_cast<const float*>(hist + (size_t)(r + i) * DV4));
#pragma unroll
        for (int i = 0; i < U; ++i) {
            const float p = smem[r + i];
            acc.x = fmaf(p, v[i].x, acc.x); acc.y = fmaf(p, v[i].y, acc.y);
            acc.z = fmaf(p, v[i].z, acc.z); acc.w = fmaf(p, v[i].w, acc.w);
            if (r + i >= 1) hout[(size_t)(r + i - 1) * DV4] = v[i];
        }
    }
    for (; r < W; ++r) {
        const float4 v = r < W - 1 ? ldg4_stream(reinterpret_cast<const float*>(hist + (size_t)r * DV4)) : *vnew;
        const float p = smem[r];
        acc.x = fmaf(p, v.x, acc.x); acc.y = fmaf(p, v.y, acc.y);
        acc.z = fmaf(p, v.z, acc.z); acc.w = fmaf(p, v.w, acc.w);
        if (r >= 1) hout[(size_t)(r - 1) * DV4] = v;
    }
    st4(AO + (size_t)bl * DV + 4 * c4, acc);
}

// ---- head regroup + output projection + LayerNorm(F*C) + residual; one CTA per (t, b) --------------------------------
__global__ void __launch_bounds__(256) attn_out_kernel(const sb_attn_args a, const float* AO) {
    SB_DYN_SMEM(float, smem);
    __shared__ float red[8];
    const int F = a.F, C = a.C, L = a.L, Vd = C / L, T = a.T, DV = F * Vd, n = F * C;
    float* os = smem;                       // [F][C]  regrouped attention output
    float* ys = os + n;                     // [F][C]  projected
    float* ws = ys + n;                     // [C][C]
    const int tid = threadIdx.x;
    const int t = blockIdx.x, b = blockIdx.y;
    for (int i = tid; i < C * C; i += 256) ws[(i % C) * C + i / C] = __ldg(a.o.w + i);     // transposed: [c][o]
    const float slope = __ldg(a.o.prelu);
    pdl_trigger();
    pdl_wait();
    for (int i = tid; i < L * DV; i += 256) {               // AO[b*L + l][t][f*Vd + vd] -> os[f][l*Vd + vd]
        const int l = i / DV, r = i - l * DV;
        const int f = r / Vd, vd = r - f * Vd;
        os[f * C + l * Vd + vd] = ldg1_stream(AO + (((size_t)b * L + l) * T + t) * DV + r);
    }
    __syncthreads();
    // thread = output channel o (tid % C) for the frequencies f = tid / C, tid / C + 256 / C, ...: its weight column stays in
    // registers and every os[f][:] row arrives as broadcast LDS.128, so the projection costs 1/4 shared-memory load per FMA
    float s1 = 0.f;
    if (C == 32) {
        const int o = tid & 31;
        float wcol[32];
#pragma unroll
        for (int c = 0; c < 32; ++c) wcol[c] = ws[c * 32 + o];
        const float bo = __ldg(a.o.b + o);
        for (int f = tid >> 5; f < F; f += 8) {
            float acc = bo;
#pragma unroll
            for (int c4 = 0; c4 < 8; ++c4) {
                const float4 v = ld4(os + f * 32 + 4 * c4);
                acc = fmaf(v.x, wcol[4 * c4], acc); acc = fmaf(v.y, wcol[4 * c4 + 1], acc);
                acc = fmaf(v.z, wcol[4 * c4 + 2], acc); acc = fmaf(v.w, wcol[4 * c4 + 3], acc);
            }
            acc = acc > 0.f ? acc : slope * acc;
            ys[f * 32 + o] = acc;
            s1 += acc;
        }
    } else {
        for (int i = tid; i < n; i += 256) {
            const int f = i / C, o = i - f * C;             // a warp: one or two f (broadcast), consecutive o (conflict-free)
            float acc = __ldg(a.o.b + o);
#pragma unroll 8
            for (int c = 0; c < C; ++c) acc = fmaf(os[f * C + c], ws[c * C + o], acc);
            acc = acc > 0.f ? acc : slope * acc;
            ys[i] = acc;
            s1 += acc;
        }
    }
    const float mean = block_sum_256(s1, red) / n;
    float s2 = 0.f;
    for (int i = tid; i < n; i += 256) { const float d = ys[i] - mean; s2 = fmaf(d, d, s2); }
    const float rstd = rsqrtf(block_sum_256(s2, red) / n + kLnEps);
    const float* xr = a.x + ((size_t)b * T + t) * n;
    float* yr = a.y + ((size_t)b * T + t) * n;
#pragma unroll 4
    for (int i = tid; i < n; i += 256)
        yr[i] = xr[i] + fmaf((ys[i] - mean) * rstd, __ldg(a.o.ln_g + i), __ldg(a.o.ln_b + i));
}

}  // namespace sb

extern "C" size_t sb_attn_workspace_floats(int B, int T, int F, int C, int L, int E, int W) {
    if (B <= 0 || T <= 0 || F <= 0 || C <= 0 || L <= 0 || E <= 0 || W <= 0) return 0;
    return sb::attn_carve(nullptr, B, T, F, C, L, E, W).total;
}

extern "C" int sb_attn_fwd(const sb_attn_args* p, void* stream) {
    using namespace sb;
    SB_REQUIRE(p && p->x && p->y && p->ws && p->K_buf_in && p->K_buf_out && p->V_buf_in && p->V_buf_out, SB_E_BADARG, "sb_attn_fwd: null pointer");
    SB_REQUIRE(p->q.w && p->k.w && p->v.w && p->o.w, SB_E_BADARG, "sb_attn_fwd: null projection weights");
    SB_REQUIRE(p->B > 0 && p->T > 0 && p->F > 0 && p->L > 0 && p->E > 0 && p->W > 1, SB_E_BADARG, "sb_attn_fwd: bad sizes");
    SB_REQUIRE(p->C % p->L == 0 && p->C % 4 == 0, SB_E_UNSUPP, "sb_attn_fwd: C=%d must be a multiple of L=%d and of 4", p->C, p->L);
    SB_REQUIRE((p->F * (p->C / p->L)) % 4 == 0, SB_E_UNSUPP, "sb_attn_fwd: F*C/L must be a multiple of 4");
    SB_REQUIRE(p->K_buf_in != p->K_buf_out && p->V_buf_in != p->V_buf_out, SB_E_BADARG, "sb_attn_fwd: K/V in/out must not alias");
    cudaStream_t st = (cudaStream_t)stream;
    const int B = p->B, T = p->T, F = p->F, C = p->C, L = p->L, E = p->E, W = p->W, Vd = C / L;
    const int BL = B * L, DK = F * E, DV = F * Vd, TT = T + W - 1;
    const AttnWs w = attn_carve(p->ws, B, T, F, C, L, E, W);
    if (T == 1 && DK % 2 == 0) {            // streaming chunk: stream the history once, shifting it on the way
        const int NOH1 = 2 * E + Vd;
        const size_t smem_qkv1 = ((size_t)F * C + (size_t)NOH1 * C + ((NOH1 + 3) & ~3) + (size_t)2 * DK + DV) * sizeof(float);
        SB_CHECK(launch("attn_qkv", attn_qkv_kernel, dim3(T, B, L), dim3(256), smem_qkv1, st, *p, w.Q, w.Kc, w.Vc));
        SB_CHECK(launch("attn_stream_scores", attn_stream_scores_kernel, dim3(BL), dim3(256),
                        (size_t)(((DK + 3) & ~3) + W) * sizeof(float), st, *p, (const float*)w.Q, (const float*)w.Kc, w.P));
        SB_CHECK(launch("attn_stream_pv", attn_stream_pv_kernel, dim3(ceil_div(DV / 4, 128), BL), dim3(128),
                        (size_t)W * sizeof(float), st, *p, (const float*)w.Vc, (const float*)w.P, w.AO));
        const size_t smem_out1 = ((size_t)2 * F * C + (size_t)C * C) * sizeof(float);
        return launch("attn_out", attn_out_kernel, dim3(T, B), dim3(256), smem_out1, st, *p, (const float*)w.AO);
    }
    const int cgrid = 2 * sm_count();
    // history -> first W-1 rows of the concatenated K / V
    SB_CHECK(launch("attn_copy", attn_copy_kernel, dim3(cgrid), dim3(256), 0, st, p->K_buf_in, w.Kc, BL,
                    (long long)(W - 1) * DK, (long long)TT * DK, 0LL, 0LL, (long long)(W - 1) * DK));
    SB_CHECK(launch("attn_copy", attn_copy_kernel, dim3(cgrid), dim3(256), 0, st, p->V_buf_in, w.Vc, BL,
                    (long long)(W - 1) * DV, (long long)TT * DV, 0LL, 0LL, (long long)(W - 1) * DV));
    const int NOH = 2 * E + Vd;
    const size_t smem_qkv = ((size_t)F * C + (size_t)NOH * C + ((NOH + 3) & ~3) + (size_t)2 * DK + DV) * sizeof(float);
    SB_CHECK(launch("attn_qkv", attn_qkv_kernel, dim3(T, B, L), dim3(256), smem_qkv, st, *p, w.Q, w.Kc, w.Vc));
    if (attn_tc_enabled() && attn_core_tc_supported(*p)) {
        SB_CHECK(attn_core_tc(*p, w.Q, w.Kc, w.Vc, w.AO, st));
    } else {
        const size_t smem_core = ((size_t)kAttnTQ * DK + (size_t)kAttnTQ * (W + 1)) * sizeof(float);
        SB_CHECK(launch("attn_core", attn_core_kernel, dim3(ceil_div(T, kAttnTQ), BL), dim3(256), smem_core, st, *p,
                        (const float*)w.Q, (const float*)w.Kc, (const float*)w.Vc, w.AO));
    }
    // last W-1 rows of the concatenated K / V -> new history
    SB_CHECK(launch("attn_copy", attn_copy_kernel, dim3(cgrid), dim3(256), 0, st, (const float*)w.Kc, p->K_buf_out, BL,
                    (long long)TT * DK, (long long)(W - 1) * DK, (long long)T * DK, 0LL, (long long)(W - 1) * DK));
    SB_CHECK(launch("attn_copy", attn_copy_kernel, dim3(cgrid), dim3(256), 0, st, (const float*)w.Vc, p->V_buf_out, BL,
                    (long long)TT * DV, (long long)(W - 1) * DV, (long long)T * DV, 0LL, (long long)(W - 1) * DV));
    const size_t smem_out = ((size_t)2 * F * C + (size_t)C * C) * sizeof(float);
    return launch("attn_out", attn_out_kernel, dim3(T, B), dim3(256), smem_out, st, *p, (const float*)w.AO);
}
