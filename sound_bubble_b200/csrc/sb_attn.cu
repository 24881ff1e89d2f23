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

struct AttnWs {
    float *Q, *Kc, *Vc, *AO;
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

// ---- Q, K, V projections + per-head LayerNorm; one CTA per (t, b) ---------------------------------------------------
__global__ void __launch_bounds__(256) attn_qkv_kernel(const sb_attn_args a, float* Q, float* Kc, float* Vc) {
    SB_DYN_SMEM(float, smem);
    const int F = a.F, C = a.C, L = a.L, E = a.E, Vd = C / L, W = a.W, T = a.T;
    const int LE = L * E, NO = 2 * LE + C, DK = F * E, DV = F * Vd;
    float* xs = smem;                       // [F][C]
    float* ws = xs + F * C;                 // [NO][C]  rows: Q (LE), K (LE), V (C)
    float* bs = ws + NO * C;                // [NO]
    float* st = bs + ((NO + 3) & ~3);       // staging: Q [L][DK], K [L][DK], V [L][DV]
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int t = blockIdx.x, b = blockIdx.y;
    for (int i = tid; i < NO * C; i += 256) {
        const int o = i / C, c = i - o * C;
        ws[i] = o < LE ? __ldg(a.q.w + o * C + c) : (o < 2 * LE ? __ldg(a.k.w + (o - LE) * C + c) : __ldg(a.v.w + (o - 2 * LE) * C + c));
    }
    for (int o = tid; o < NO; o += 256)
        bs[o] = o < LE ? __ldg(a.q.b + o) : (o < 2 * LE ? __ldg(a.k.b + o - LE) : __ldg(a.v.b + o - 2 * LE));
    const float sq = __ldg(a.q.prelu), sk = __ldg(a.k.prelu), sv = __ldg(a.v.prelu);
    pdl_trigger();
    pdl_wait();
    const float* xr = a.x + ((size_t)b * T + t) * F * C;
    for (int i = tid; i < F * C / 4; i += 256) st4(xs + 4 * i, ldg4_stream(xr + 4 * i));
    __syncthreads();
    for (int i = tid; i < F * NO; i += 256) {
        const int f = i / NO, o = i - f * NO;
        float acc = bs[o];
        const float* xp = xs + f * C;
        const float* wp = ws + o * C;
        for (int c = 0; c < C; ++c) acc = fmaf(xp[c], wp[c], acc);
        if (o < LE) {
            acc = acc > 0.f ? acc : sq * acc;
            st[(o / E) * DK + f * E + (o % E)] = acc;
        } else if (o < 2 * LE) {
            const int oo = o - LE;
            acc = acc > 0.f ? acc : sk * acc;
            st[L * DK + (oo / E) * DK + f * E + (oo % E)] = acc;
        } else {
            const int oo = o - 2 * LE;
            acc = acc > 0.f ? acc : sv * acc;
            st[2 * L * DK + (oo / Vd) * DV + f * Vd + (oo % Vd)] = acc;
        }
    }
    __syncthreads();
    // 3L rows to normalise: row r -> (tensor = r / L, head = r % L); one warp per row
    for (int r = warp; r < 3 * L; r += 8) {
        const int which = r / L, l = r - which * L;
        const int n = which == 2 ? DV : DK;
        float* row = st + (which == 2 ? 2 * L * DK + l * DV : (which * L + l) * DK);
        const sb_attn_proj& pj = which == 0 ? a.q : (which == 1 ? a.k : a.v);
        float s1 = 0.f;
        for (int i = lane; i < n; i += 32) s1 += row[i];
        const float mean = group_sum<32>(s1) / n;
        float s2 = 0.f;
        for (int i = lane; i < n; i += 32) { const float d = row[i] - mean; s2 = fmaf(d, d, s2); }
        const float rstd = rsqrtf(group_sum<32>(s2) / n + kLnEps);
        float* dst;
        if (which == 0) dst = Q + (((size_t)b * L + l) * T + t) * DK;
        else if (which == 1) dst = Kc + (((size_t)b * L + l) * (T + W - 1) + W - 1 + t) * DK;
        else dst = Vc + (((size_t)b * L + l) * (T + W - 1) + W - 1 + t) * DV;
        for (int i = lane; i < n; i += 32) dst[i] = fmaf((row[i] - mean) * rstd, __ldg(pj.ln_g + i), __ldg(pj.ln_b + i));
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
    for (int i = tid; i < C * C; i += 256) ws[i] = __ldg(a.o.w + i);
    const float slope = __ldg(a.o.prelu);
    pdl_trigger();
    pdl_wait();
    for (int i = tid; i < L * DV; i += 256) {               // AO[b*L + l][t][f*Vd + vd] -> os[f][l*Vd + vd]
        const int l = i / DV, r = i - l * DV;
        const int f = r / Vd, vd = r - f * Vd;
        os[f * C + l * Vd + vd] = ldg1_stream(AO + (((size_t)b * L + l) * T + t) * DV + r);
    }
    __syncthreads();
    float s1 = 0.f;
    for (int i = tid; i < n; i += 256) {
        const int f = i / C, o = i - f * C;
        float acc = __ldg(a.o.b + o);
        for (int c = 0; c < C; ++c) acc = fmaf(os[f * C + c], ws[o * C + c], acc);
        acc = acc > 0.f ? acc : slope * acc;
        ys[i] = acc;
        s1 += acc;
    }
    const float mean = block_sum_256(s1, red) / n;
    float s2 = 0.f;
    for (int i = tid; i < n; i += 256) { const float d = ys[i] - mean; s2 = fmaf(d, d, s2); }
    const float rstd = rsqrtf(block_sum_256(s2, red) / n + kLnEps);
    const float* xr = a.x + ((size_t)b * T + t) * n;
    float* yr = a.y + ((size_t)b * T + t) * n;
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
    const int cgrid = 2 * sm_count();
    // history -> first W-1 rows of the concatenated K / V
    SB_CHECK(launch("attn_copy", attn_copy_kernel, dim3(cgrid), dim3(256), 0, st, p->K_buf_in, w.Kc, BL,
                    (long long)(W - 1) * DK, (long long)TT * DK, 0LL, 0LL, (long long)(W - 1) * DK));
    SB_CHECK(launch("attn_copy", attn_copy_kernel, dim3(cgrid), dim3(256), 0, st, p->V_buf_in, w.Vc, BL,
                    (long long)(W - 1) * DV, (long long)TT * DV, 0LL, 0LL, (long long)(W - 1) * DV));
    const int NO = 2 * L * E + C;
    const size_t smem_qkv = ((size_t)F * C + (size_t)NO * C + ((NO + 3) & ~3) + (size_t)2 * L * DK + (size_t)L * DV) * sizeof(float);
    SB_CHECK(launch("attn_qkv", attn_qkv_kernel, dim3(T, B), dim3(256), smem_qkv, st, *p, w.Q, w.Kc, w.Vc));
    const size_t smem_core = ((size_t)kAttnTQ * DK + (size_t)kAttnTQ * (W + 1)) * sizeof(float);
    SB_CHECK(launch("attn_core", attn_core_kernel, dim3(ceil_div(T, kAttnTQ), BL), dim3(256), smem_core, st, *p,
                    (const float*)w.Q, (const float*)w.Kc, (const float*)w.Vc, w.AO));
    // last W-1 rows of the concatenated K / V -> new history
    SB_CHECK(launch("attn_copy", attn_copy_kernel, dim3(cgrid), dim3(256), 0, st, (const float*)w.Kc, p->K_buf_out, BL,
                    (long long)TT * DK, (long long)(W - 1) * DK, (long long)T * DK, 0LL, (long long)(W - 1) * DK));
    SB_CHECK(launch("attn_copy", attn_copy_kernel, dim3(cgrid), dim3(256), 0, st, (const float*)w.Vc, p->V_buf_out, BL,
                    (long long)TT * DV, (long long)(W - 1) * DV, (long long)T * DV, 0LL, (long long)(W - 1) * DV));
    const size_t smem_out = ((size_t)2 * F * C + (size_t)C * C) * sizeof(float);
    return launch("attn_out", attn_out_kernel, dim3(T, B), dim3(256), smem_out, st, *p, (const float*)w.AO);
}
