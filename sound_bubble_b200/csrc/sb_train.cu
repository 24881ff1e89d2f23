// Training path: forward kernels that keep what back-propagation needs, and the backward kernels (the *_bwd twins).
//
// Reference spans (DE3 = src/models/tfgridnet_realtime_clean_dis_embd3/tfgridnet_causal.py); what is differentiated is
// exactly what PLModule._step runs under autograd (src/hl_modules/distance_based_hl_module.py:303-330, 437-441):
//   conv-in + LayerNorm      :332-354, 499-507     conv_in_train_kernel, ln_fwd_kernel / ln_bwd_kernel, conv_in_wgrad_kernel
//   FilmLayer                :51-68, 509-513       film_apply_*_kernel, film_params_bwd_kernel (Dis_Embed_Conv :150-173)
//   intra / inter LSTM paths :794-827, 829-849     ln_fwd -> lstm_train_fwd -> rowgemm (Linear + residual);
//                                                  rowgemm (dL/dh) -> lstm_train_bwd (BPTT, serial part only: the cell
//                                                  derivatives and W_hh^T dz) -> rowgemm (dL/dLN(x)) -> ln_bwd, and the
//                                                  weight gradients as one reduction over all (row, step) pairs (outer_kernel)
//   conv-LSTM intra path     :800-815 (OPT :684-697, 494-510)   convpre_* / prelu_* / convpost_* around the same LSTM kernels
//   deconv + iSTFT/OLA       :401, 517-542         istft_bwd_kernel, deconv_bwd_x_kernel, deconv_wgrad_kernel
// The STFT basis buffers and the input features carry no gradient (no parameter upstream of conv-in).
//
// Layout: the LSTM-side buffers are sequence-major, n = row * S + step ("[row][step][.]"); the activations stay in
// X[B][T][F][C].  For the intra-frame path the two orders coincide (row = (b,t), step = f); for the inter-frame path
// (row = (b,f), step = t) RowMap converts.  Gradients are accumulated with fp32 atomics.
#include "sb_common.cuh"

namespace sb {

constexpr int kH = 64;

struct RowMap {
    int S, F, inter;
};
// sequence-major index n -> index of the same (b, t, f) position in X[B][T][F][.]
__device__ __forceinline__ long long map_pos(const RowMap& m, long long n) {
    if (!m.inter) return n;
    const long long row = n / m.S;
    const int s = (int)(n - row * m.S);
    const long long b = row / m.F;
    const int f = (int)(row - b * m.F);
    return (b * m.S + s) * m.F + f;
}

#ifdef SB_EMU
__device__ __forceinline__ void atomic_add(float* p, float v) { emu_atomic_add(p, v); }
#else
__device__ __forceinline__ void atomic_add(float* p, float v) { atomicAdd(p, v); }
#endif

// ------------------------------------------------------------------------------------------------------------
// LayerNorm over C: C / 4 adjacent lanes per position, one float4 each (a warp reads 512 contiguous bytes; with one position
// per thread every load instruction touched 32 different 128-byte lines and the kernels ran at a third of the HBM rate).
// in: rows of C floats at map_pos(n) (or n); out: xhat[n], xn[n], rstd[n].  Grid: ln_rows_per_block<C>() positions per block.
// ------------------------------------------------------------------------------------------------------------
template <int C>
constexpr int ln_rows_per_block() { return 256 / (C / 4); }

template <int C>
__global__ void __launch_bounds__(256) ln_fwd_kernel(const float* x, RowMap map, const float* g, const float* bta,
                                                     float* xhat, float* xn, float* rstd, long long N, float eps) {
    constexpr int LPR = C / 4, RPB = 256 / LPR;
    pdl_wait();
    const int li = threadIdx.x % LPR;
    const long long n = (long long)blockIdx.x * RPB + threadIdx.x / LPR;
    const bool live = n < N;                                // uniform over the lanes of a position
    const long long nc = live ? n : N - 1;
    const float4 t = ldg4_stream(x + map_pos(map, nc) * C + 4 * li);
    const float mean = group_sum<LPR>((t.x + t.y) + (t.z + t.w)) * (1.0f / C);
    const float dx = t.x - mean, dy = t.y - mean, dz = t.z - mean, dw = t.w - mean;
    const float var = group_sum<LPR>(fmaf(dx, dx, fmaf(dy, dy, fmaf(dz, dz, dw * dw))));
    const float r = rsqrtf(var * (1.0f / C) + eps);
    if (!live) return;
    if (li == 0) rstd[n] = r;
    const float4 h = make_float4(dx * r, dy * r, dz * r, dw * r);
    const float4 gg = __ldg(reinterpret_cast<const float4*>(g) + li), bb = __ldg(reinterpret_cast<const float4*>(bta) + li);
    st4(xhat + n * C + 4 * li, h);
    st4(xn + n * C + 4 * li, make_float4(fmaf(h.x, gg.x, bb.x), fmaf(h.y, gg.y, bb.y), fmaf(h.z, gg.z, bb.z), fmaf(h.w, gg.w, bb.w)));
}

// dL/dLN-output (dxn[n]) -> dL/dx at map_pos(n) (+ the residual branch's gradient gy), and the gain / bias gradients.
template <int C>
__global__ void __launch_bounds__(256) ln_bwd_kernel(const float* dxn, const float* xhat, const float* rstd, const float* g,
                                                     const float* gy, float* gx, RowMap map, float* g_g, float* g_b,
                                                     long long N) {
    constexpr int LPR = C / 4, RPB = 256 / LPR;
    __shared__ float red[2 * C];
    pdl_wait();
    if (threadIdx.x < 2 * C) red[threadIdx.x] = 0.f;
    __syncthreads();
    const int li = threadIdx.x % LPR;
    const float4 gw = __ldg(reinterpret_cast<const float4*>(g) + li);
    float4 ag = make_float4(0.f, 0.f, 0.f, 0.f), ab = ag;   // this lane's four channels, over the positions it visits
    const long long n_iter = (N + (long long)gridDim.x * RPB - 1) / ((long long)gridDim.x * RPB);
    for (long long it = 0; it < n_iter; ++it) {             // same trip count for every thread: the shuffles stay converged
        const long long n = (it * gridDim.x + blockIdx.x) * RPB + threadIdx.x / LPR;
        const bool live = n < N;
        const long long nc = live ? n : N - 1;
        float4 d = ldg4_stream(dxn + nc * C + 4 * li);
        const float4 h = ldg4_stream(xhat + nc * C + 4 * li);
        if (live) {
            ag.x = fmaf(d.x, h.x, ag.x); ag.y = fmaf(d.y, h.y, ag.y); ag.z = fmaf(d.z, h.z, ag.z); ag.w = fmaf(d.w, h.w, ag.w);
            ab.x += d.x; ab.y += d.y; ab.z += d.z; ab.w += d.w;
        }
        d.x *= gw.x; d.y *= gw.y; d.z *= gw.z; d.w *= gw.w;
        const float m1 = group_sum<LPR>((d.x + d.y) + (d.z + d.w)) * (1.0f / C);
        const float m2 = group_sum<LPR>(fmaf(d.x, h.x, fmaf(d.y, h.y, fmaf(d.z, h.z, d.w * h.w)))) * (1.0f / C);
        if (live) {
            const float r = rstd[n];
            const long long p = map_pos(map, n) * C + 4 * li;
            float4 o = make_float4(r * (d.x - m1 - h.x * m2), r * (d.y - m1 - h.y * m2), r * (d.z - m1 - h.z * m2), r * (d.w - m1 - h.w * m2));
            if (gy) {
                const float4 t = ld_plain4(gy + p);
                o.x += t.x; o.y += t.y; o.z += t.z; o.w += t.w;
            }
            st4(gx + p, o);
        }
    }
    // lanes li, li + LPR, ... of a warp hold the same four channels
    float a4[4] = {ag.x, ag.y, ag.z, ag.w}, b4[4] = {ab.x, ab.y, ab.z, ab.w};
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        float a = a4[i], b = b4[i];
#pragma unroll
        for (int o = 16; o >= LPR; o >>= 1) { a += __shfl_xor_sync(0xffffffffu, a, o); b += __shfl_xor_sync(0xffffffffu, b, o); }
        if ((threadIdx.x & 31) < LPR) { atomic_add(&red[4 * li + i], a); atomic_add(&red[C + 4 * li + i], b); }
    }
    __syncthreads();
    if (threadIdx.x < C) atomic_add(g_g + threadIdx.x, red[threadIdx.x]);
    else if (threadIdx.x < 2 * C) atomic_add(g_b + threadIdx.x - C, red[threadIdx.x]);
}

// ------------------------------------------------------------------------------------------------------------
// LSTM forward that stores the activated gates, c_t and h_t of every step.  CTA = 4 sequences of one direction; thread j
// owns gate column j (row j of [W_ih | W_hh], in registers) for the four sequences; [x_t | h_{t-1}] of the four sequences
// sits in shared memory as one float4 per k (a broadcast LDS.128 feeds four FMAs); threads (q, u) do the cell update.
// ------------------------------------------------------------------------------------------------------------
struct LstmTrain {
    const float* xn;                    // [N][C]
    const float* w_ih[2];
    const float* w_hh[2];
    const float* b_ih[2];
    const float* b_hh[2];
    float* gates[2];                    // [N][4H]  fwd: i, f, g, o (activated)   bwd: overwritten with dL/dz
    float* c[2];                        // [N][H]
    float* h[2];                        // [N][H]
    const float* dh[2];                 // bwd: dL/dh_t from the projection, [N][H]
    int R, S;
};

template <int C>
__global__ void __launch_bounds__(256) lstm_train_fwd_kernel(const LstmTrain a) {
    constexpr int H = kH, NS = 4, K = C + H;
    __shared__ float4 a_s[K];
    __shared__ float z_s[NS][4 * H];
    const int tid = threadIdx.x, d = blockIdx.y, r0 = blockIdx.x * NS, S = a.S;
    float w[K];
#pragma unroll
    for (int k = 0; k < C; k += 4) {
        const float4 t = __ldg(reinterpret_cast<const float4*>(a.w_ih[d] + (size_t)tid * C + k));
        w[k] = t.x; w[k + 1] = t.y; w[k + 2] = t.z; w[k + 3] = t.w;
    }
#pragma unroll
    for (int k = 0; k < H; k += 4) {
        const float4 t = __ldg(reinterpret_cast<const float4*>(a.w_hh[d] + (size_t)tid * H + k));
        w[C + k] = t.x; w[C + k + 1] = t.y; w[C + k + 2] = t.z; w[C + k + 3] = t.w;
    }
    const float bias = __ldg(a.b_ih[d] + tid) + __ldg(a.b_hh[d] + tid);
    const int gate = tid / H;
    // x loader role (tid < NS*C): sequence xq, channel xc.  cell role: sequence q, unit u.
    const int xq = tid / C, xc = tid - xq * C;
    const bool loader = tid < NS * C;
    const long long xrow = min(r0 + (loader ? xq : 0), a.R - 1);
    const int q = tid / H, u = tid - q * H;
    const bool valid = r0 + q < a.R;
    const long long crow = min(r0 + q, a.R - 1);
    reinterpret_cast<float*>(&a_s[C + u])[q] = 0.f;
    float c_reg = 0.f;
    pdl_wait();
    float xv = 0.f;
    if (loader) xv = ldg1_stream(a.xn + (xrow * S + (d ? S - 1 : 0)) * C + xc);
    for (int s = 0; s < S; ++s) {
        const int se = d ? S - 1 - s : s;
        if (loader) reinterpret_cast<float*>(&a_s[xc])[xq] = xv;
        __syncthreads();
        if (loader && s + 1 < S) xv = ldg1_stream(a.xn + (xrow * S + (d ? se - 1 : se + 1)) * C + xc);
        float acc0 = bias, acc1 = bias, acc2 = bias, acc3 = bias;
#pragma unroll
        for (int k = 0; k < K; ++k) {
            const float4 v = a_s[k];
            acc0 = fmaf(w[k], v.x, acc0); acc1 = fmaf(w[k], v.y, acc1); acc2 = fmaf(w[k], v.z, acc2); acc3 = fmaf(w[k], v.w, acc3);
        }
        if (gate == 2) { acc0 = tanh_f(acc0); acc1 = tanh_f(acc1); acc2 = tanh_f(acc2); acc3 = tanh_f(acc3); }
        else { acc0 = sigmoid_f(acc0); acc1 = sigmoid_f(acc1); acc2 = sigmoid_f(acc2); acc3 = sigmoid_f(acc3); }
        z_s[0][tid] = acc0; z_s[1][tid] = acc1; z_s[2][tid] = acc2; z_s[3][tid] = acc3;
        {
            float* gp = a.gates[d] + ((long long)r0 * S + se) * (4 * H) + tid;
            const long long seq = (long long)S * 4 * H;
            if (r0 + 0 < a.R) gp[0] = acc0;
            if (r0 + 1 < a.R) gp[seq] = acc1;
            if (r0 + 2 < a.R) gp[2 * seq] = acc2;
            if (r0 + 3 < a.R) gp[3 * seq] = acc3;
        }
        __syncthreads();
        const float gi = z_s[q][u], gf = z_s[q][H + u], gg = z_s[q][2 * H + u], go = z_s[q][3 * H + u];
        c_reg = fmaf(gf, c_reg, gi * gg);
        const float hh = go * tanh_f(c_reg);
        reinterpret_cast<float*>(&a_s[C + u])[q] = hh;
        if (valid) {
            const long long n = crow * S + se;
            a.c[d][n * H + u] = c_reg;
            a.h[d][n * H + u] = hh;
        }
    }
}

// Same computation with TWO gate rows per thread (rows t and t + 2H: an (i|f) row and a (g|o) row), 128 threads: one
// broadcast LDS.128 now feeds 8 FMAs, which halves the shared-memory traffic that bounds the one-row version (C = 32).
template <int C>
__global__ void __launch_bounds__(128) lstm_train_fwd2_kernel(const LstmTrain a) {
    constexpr int H = kH, NS = 4, K = C + H;
    static_assert(NS * C <= 128, "x loader");
    __shared__ float4 a_s[K];
    __shared__ float z_s[NS][4 * H];
    const int tid = threadIdx.x, d = blockIdx.y, r0 = blockIdx.x * NS, S = a.S;
    float w0[K], w1[K];
#pragma unroll
    for (int k = 0; k < C; k += 4) {
        const float4 t = __ldg(reinterpret_cast<const float4*>(a.w_ih[d] + (size_t)tid * C + k));
        const float4 u = __ldg(reinterpret_cast<const float4*>(a.w_ih[d] + (size_t)(tid + 2 * H) * C + k));
        w0[k] = t.x; w0[k + 1] = t.y; w0[k + 2] = t.z; w0[k + 3] = t.w;
        w1[k] = u.x; w1[k + 1] = u.y; w1[k + 2] = u.z; w1[k + 3] = u.w;
    }
#pragma unroll
    for (int k = 0; k < H; k += 4) {
        const float4 t = __ldg(reinterpret_cast<const float4*>(a.w_hh[d] + (size_t)tid * H + k));
        const float4 u = __ldg(reinterpret_cast<const float4*>(a.w_hh[d] + (size_t)(tid + 2 * H) * H + k));
        w0[C + k] = t.x; w0[C + k + 1] = t.y; w0[C + k + 2] = t.z; w0[C + k + 3] = t.w;
        w1[C + k] = u.x; w1[C + k + 1] = u.y; w1[C + k + 2] = u.z; w1[C + k + 3] = u.w;
    }
    const float bias0 = __ldg(a.b_ih[d] + tid) + __ldg(a.b_hh[d] + tid);
    const float bias1 = __ldg(a.b_ih[d] + tid + 2 * H) + __ldg(a.b_hh[d] + tid + 2 * H);
    const bool row1_tanh = tid < H;                 // row tid + 2H is a g row for tid < H, an o row otherwise
    const int xq = tid / C, xc = tid - xq * C;
    const bool loader = tid < NS * C;
    const long long xrow = min(r0 + (loader ? xq : 0), a.R - 1);
    // cell role: unit u of sequences q0 and q0 + 2
    const int q0 = tid / H, u = tid - q0 * H;
    reinterpret_cast<float*>(&a_s[C + u])[q0] = 0.f;
    reinterpret_cast<float*>(&a_s[C + u])[q0 + 2] = 0.f;
    float c_reg[2] = {0.f, 0.f};
    pdl_wait();
    float xv = 0.f;
    if (loader) xv = ldg1_stream(a.xn + (xrow * S + (d ? S - 1 : 0)) * C + xc);
    for (int s = 0; s < S; ++s) {
        const int se = d ? S - 1 - s : s;
        if (loader) reinterpret_cast<float*>(&a_s[xc])[xq] = xv;
        __syncthreads();
        if (loader && s + 1 < S) xv = ldg1_stream(a.xn + (xrow * S + (d ? se - 1 : se + 1)) * C + xc);
        float2 acc2[2][2] = {{make_float2(bias0, bias0), make_float2(bias0, bias0)}, {make_float2(bias1, bias1), make_float2(bias1, bias1)}};
#pragma unroll
        for (int k = 0; k < K; ++k) {               // one broadcast LDS.128 -> 4 packed FFMA2 = 8 FMA
            const float4 v = a_s[k];
            ffma2(acc2[0][0], make_float2(v.x, v.y), w0[k]);
            ffma2(acc2[0][1], make_float2(v.z, v.w), w0[k]);
            ffma2(acc2[1][0], make_float2(v.x, v.y), w1[k]);
            ffma2(acc2[1][1], make_float2(v.z, v.w), w1[k]);
        }
        float acc[2][4] = {{acc2[0][0].x, acc2[0][0].y, acc2[0][1].x, acc2[0][1].y}, {acc2[1][0].x, acc2[1][0].y, acc2[1][1].x, acc2[1][1].y}};
#pragma unroll
        for (int q = 0; q < NS; ++q) {
            acc[0][q] = sigmoid_f(acc[0][q]);
            acc[1][q] = row1_tanh ? tanh_f(acc[1][q]) : sigmoid_f(acc[1][q]);
            z_s[q][tid] = acc[0][q];
            z_s[q][tid + 2 * H] = acc[1][q];
            if (r0 + q < a.R) {
                float* gp = a.gates[d] + ((long long)(r0 + q) * S + se) * (4 * H) + tid;
                gp[0] = acc[0][q];
                gp[2 * H] = acc[1][q];
            }
        }
        __syncthreads();
#pragma unroll
        for (int i = 0; i < 2; ++i) {
            const int q = q0 + 2 * i;
            const float gi = z_s[q][u], gf = z_s[q][H + u], gg = z_s[q][2 * H + u], go = z_s[q][3 * H + u];
            c_reg[i] = fmaf(gf, c_reg[i], gi * gg);
            const float hh = go * tanh_f(c_reg[i]);
            reinterpret_cast<float*>(&a_s[C + u])[q] = hh;
            if (r0 + q < a.R) {
                const long long n = (long long)(r0 + q) * S + se;
                a.c[d][n * H + u] = c_reg[i];
                a.h[d][n * H + u] = hh;
            }
        }
    }
}

// BPTT, serial part: per step the cell derivatives (dz, written over the stored gates) and dh_{t-1} += W_hh^T dz.
// Thread (q, u) differentiates the cell of sequence q, unit u; thread (quarter, k) then sums its quarter of the 4H gate
// rows of column k of W_hh (64 weights in registers) for the four sequences.
__global__ void __launch_bounds__(256) lstm_train_bwd_kernel(const LstmTrain a) {
    constexpr int H = kH, NS = 4;
    __shared__ float4 dz_s[4 * H];
    __shared__ float part_s[4][NS][H];
    const int tid = threadIdx.x, d = blockIdx.y, r0 = blockIdx.x * NS, S = a.S;
    const int q = tid / H, u = tid - q * H;
    float wq[H];
#pragma unroll
    for (int j = 0; j < H; ++j) wq[j] = __ldg(a.w_hh[d] + (size_t)(q * H + j) * H + u);
#pragma unroll
    for (int i = 0; i < 4; ++i) part_s[i][q][u] = 0.f;
    const bool valid = r0 + q < a.R;
    const long long row = min(r0 + q, a.R - 1);
    float* gates = a.gates[d];
    const float* cc = a.c[d];
    const float* dh = a.dh[d];
    pdl_wait();
    __syncthreads();
    float dc = 0.f;
    // software pipeline: the loads of step s+1 are issued before the W_hh^T product of step s
    int se = d ? 0 : S - 1;
    long long n = row * S + se;
    float gi = gates[n * 4 * H + u], gf = gates[n * 4 * H + H + u], gg = gates[n * 4 * H + 2 * H + u], go = gates[n * 4 * H + 3 * H + u];
    float ct = cc[n * H + u], dhv = dh[n * H + u];
    float cp = S > 1 ? cc[(d ? n + 1 : n - 1) * H + u] : 0.f;
    for (int s = 0; s < S; ++s) {
        const float dht = dhv + ((part_s[0][q][u] + part_s[1][q][u]) + (part_s[2][q][u] + part_s[3][q][u]));
        const float tc = tanh_f(ct);
        const float dcv = fmaf(dht * go, 1.f - tc * tc, dc);
        const float dzi = dcv * gg * gi * (1.f - gi);
        const float dzf = dcv * cp * gf * (1.f - gf);
        const float dzg = dcv * gi * (1.f - gg * gg);
        const float dzo = dht * tc * go * (1.f - go);
        dc = dcv * gf;
        reinterpret_cast<float*>(&dz_s[u])[q] = dzi;
        reinterpret_cast<float*>(&dz_s[H + u])[q] = dzf;
        reinterpret_cast<float*>(&dz_s[2 * H + u])[q] = dzg;
        reinterpret_cast<float*>(&dz_s[3 * H + u])[q] = dzo;
        if (valid) {
            gates[n * 4 * H + u] = dzi; gates[n * 4 * H + H + u] = dzf; gates[n * 4 * H + 2 * H + u] = dzg; gates[n * 4 * H + 3 * H + u] = dzo;
        }
        if (s + 1 < S) {
            se = d ? se + 1 : se - 1;
            n = row * S + se;
            gi = gates[n * 4 * H + u]; gf = gates[n * 4 * H + H + u]; gg = gates[n * 4 * H + 2 * H + u]; go = gates[n * 4 * H + 3 * H + u];
            ct = cp;
            dhv = dh[n * H + u];
            cp = s + 2 < S ? cc[(d ? n + 1 : n - 1) * H + u] : 0.f;
        }
        __syncthreads();
        float p0 = 0.f, p1 = 0.f, p2 = 0.f, p3 = 0.f;
#pragma unroll
        for (int j = 0; j < H; ++j) {
            const float4 v = dz_s[q * H + j];
            p0 = fmaf(wq[j], v.x, p0); p1 = fmaf(wq[j], v.y, p1); p2 = fmaf(wq[j], v.z, p2); p3 = fmaf(wq[j], v.w, p3);
        }
        part_s[q][0][u] = p0; part_s[q][1][u] = p1; part_s[q][2][u] = p2; part_s[q][3][u] = p3;
        __syncthreads();
    }
}

// BPTT with TWO columns of W_hh per thread (k and k + 32) and 128 threads: one broadcast LDS.128 of dz feeds 8 FMAs.
// Thread (quarter, kp) sums its quarter of the 4H gate rows for columns kp and kp + 32; thread (q0, u) differentiates the
// cells of unit u for sequences q0 and q0 + 2.
__global__ void __launch_bounds__(128) lstm_train_bwd2_kernel(const LstmTrain a) {
    constexpr int H = kH, NS = 4;
    __shared__ float4 dz_s[4 * H];
    __shared__ float part_s[4][NS][H];
    const int tid = threadIdx.x, d = blockIdx.y, r0 = blockIdx.x * NS, S = a.S;
    const int qtr = tid / 32, kp = tid % 32;
    float wq[2][H];
#pragma unroll
    for (int j = 0; j < H; ++j) {
        wq[0][j] = __ldg(a.w_hh[d] + (size_t)(qtr * H + j) * H + kp);
        wq[1][j] = __ldg(a.w_hh[d] + (size_t)(qtr * H + j) * H + kp + 32);
    }
    const int q0 = tid / H, u = tid - q0 * H;
#pragma unroll
    for (int i = 0; i < 4; ++i) { part_s[i][q0][u] = 0.f; part_s[i][q0 + 2][u] = 0.f; }
    float* gates = a.gates[d];
    const float* cc = a.c[d];
    const float* dh = a.dh[d];
    bool valid[2];
    long long row[2];
#pragma unroll
    for (int i = 0; i < 2; ++i) { valid[i] = r0 + q0 + 2 * i < a.R; row[i] = min(r0 + q0 + 2 * i, a.R - 1); }
    pdl_wait();
    __syncthreads();
    // processed step p (p = 0 is the last step of the forward recurrence) sits at position se(p) of its sequence.
    // Register pipeline two steps deep: the loads of step p + 2 are issued while step p is worked on.
    auto nof = [&](int i, int p) { return row[i] * S + (d ? p : S - 1 - p); };
    struct Stage { float gi, gf, gg, go, dh; };
    Stage cur[2], nxt[2];
    float c0[2], c1[2], c2[2], dc[2] = {0.f, 0.f};
    auto load_stage = [&](Stage& st, int i, int p) {
        const long long n = nof(i, p);
        st.gi = gates[n * 4 * H + u]; st.gf = gates[n * 4 * H + H + u]; st.gg = gates[n * 4 * H + 2 * H + u]; st.go = gates[n * 4 * H + 3 * H + u];
        st.dh = dh[n * H + u];
    };
#pragma unroll
    for (int i = 0; i < 2; ++i) {
        load_stage(cur[i], i, 0);
        nxt[i] = cur[i];
        if (S > 1) load_stage(nxt[i], i, 1);
        c0[i] = cc[nof(i, 0) * H + u];
        c1[i] = S > 1 ? cc[nof(i, 1) * H + u] : 0.f;
        c2[i] = S > 2 ? cc[nof(i, 2) * H + u] : 0.f;
    }
    for (int s = 0; s < S; ++s) {
#pragma unroll
        for (int i = 0; i < 2; ++i) {
            const int q = q0 + 2 * i;
            const long long n = nof(i, s);
            const Stage g = cur[i];
            const float dht = g.dh + ((part_s[0][q][u] + part_s[1][q][u]) + (part_s[2][q][u] + part_s[3][q][u]));
            const float tc = tanh_f(c0[i]);
            const float dcv = fmaf(dht * g.go, 1.f - tc * tc, dc[i]);
            const float dzi = dcv * g.gg * g.gi * (1.f - g.gi);
            const float dzf = dcv * c1[i] * g.gf * (1.f - g.gf);
            const float dzg = dcv * g.gi * (1.f - g.gg * g.gg);
            const float dzo = dht * tc * g.go * (1.f - g.go);
            dc[i] = dcv * g.gf;
            reinterpret_cast<float*>(&dz_s[u])[q] = dzi;
            reinterpret_cast<float*>(&dz_s[H + u])[q] = dzf;
            reinterpret_cast<float*>(&dz_s[2 * H + u])[q] = dzg;
            reinterpret_cast<float*>(&dz_s[3 * H + u])[q] = dzo;
            if (valid[i]) {
                gates[n * 4 * H + u] = dzi; gates[n * 4 * H + H + u] = dzf; gates[n * 4 * H + 2 * H + u] = dzg; gates[n * 4 * H + 3 * H + u] = dzo;
            }
            cur[i] = nxt[i];
            c0[i] = c1[i];
            c1[i] = c2[i];
            if (s + 2 < S) load_stage(nxt[i], i, s + 2);
            c2[i] = s + 3 < S ? cc[nof(i, s + 3) * H + u] : 0.f;
        }
        __syncthreads();
        float2 p[2][2] = {{make_float2(0.f, 0.f), make_float2(0.f, 0.f)}, {make_float2(0.f, 0.f), make_float2(0.f, 0.f)}};
#pragma unroll
        for (int j = 0; j < H; ++j) {               // one broadcast LDS.128 -> 4 packed FFMA2 = 8 FMA
            const float4 v = dz_s[qtr * H + j];
            ffma2(p[0][0], make_float2(v.x, v.y), wq[0][j]);
            ffma2(p[0][1], make_float2(v.z, v.w), wq[0][j]);
            ffma2(p[1][0], make_float2(v.x, v.y), wq[1][j]);
            ffma2(p[1][1], make_float2(v.z, v.w), wq[1][j]);
        }
        part_s[qtr][0][kp] = p[0][0].x; part_s[qtr][1][kp] = p[0][0].y; part_s[qtr][2][kp] = p[0][1].x; part_s[qtr][3][kp] = p[0][1].y;
        part_s[qtr][0][kp + 32] = p[1][0].x; part_s[qtr][1][kp + 32] = p[1][0].y; part_s[qtr][2][kp + 32] = p[1][1].x; part_s[qtr][3][kp + 32] = p[1][1].y;
        __syncthreads();
    }
}

// ------------------------------------------------------------------------------------------------------------
// out[o(n)][0..P) = bias + res[o(n)] + sum_a sum_k A_a[i(n)][k] * W_a(k, p)       (tall-skinny GEMM, W in shared memory)
// CTA tile = ROWS x P with an 8 x 8 register tile per thread; A is staged k-major through shared memory in chunks of KB
// columns (coalesced float4 loads, the next chunk prefetched into registers while the current one is multiplied), so one
// k costs 2 + 2 LDS.128 for 64 FMA.
// ------------------------------------------------------------------------------------------------------------
struct RowGemm {
    const float* A[2];
    const float* W[2];
    int nA, K, lda, ldw, w_trans;       // W(k, p) = w_trans ? W[p*ldw + k] : W[k*ldw + p]
    const float* bias;
    const float* res;
    float* out;
    RowMap map;
    int a_mapped, o_mapped;
    long long N;
};

template <int P>
struct RowGemmCfg {
    static constexpr int CG = P / 8, RG = 256 / CG, ROWS = RG * 8, KB = 8;
    static constexpr int AS = ROWS + 4, LPT = ROWS * KB / 4 / 256;
    static size_t smem_bytes(int nA, int K) { return ((size_t)nA * K * P + (size_t)KB * AS) * sizeof(float); }
};

template <int P, bool PACK>
__global__ void __launch_bounds__(256, 2) rowgemm_kernel(const RowGemm g) {
    using Cfg = RowGemmCfg<P>;
    constexpr int CG = Cfg::CG, ROWS = Cfg::ROWS, KB = Cfg::KB, AS = Cfg::AS, LPT = Cfg::LPT;
    SB_DYN_SMEM(float, ws);                         // [nA][K][P]
    const int tid = threadIdx.x, K = g.K;
    float* As = ws + (size_t)g.nA * K * P;          // [KB][AS]
    for (int a = 0; a < g.nA; ++a) {
        const float* wsrc = a ? g.W[1] : g.W[0];
        for (int i = tid; i < K * P; i += 256) {
            const int k = i / P, p = i - k * P;
            ws[(a * K + k) * P + p] = __ldg(wsrc + (g.w_trans ? (size_t)p * g.ldw + k : (size_t)k * g.ldw + p));
        }
    }
    const long long row_base = (long long)blockIdx.x * ROWS;
    // loader role: float4 number tid + 256 i of the chunk = (row lrow, columns 4*lk4 .. +3)
    long long aoff[LPT];
    auto lrow = [&](int i) { return (tid + 256 * i) / (KB / 4); };
    auto lk = [&](int i) { return 4 * ((tid + 256 * i) % (KB / 4)); };
#pragma unroll
    for (int i = 0; i < LPT; ++i) {
        const long long n = row_base + lrow(i);
        aoff[i] = n < g.N ? (g.a_mapped ? map_pos(g.map, n) : n) * g.lda + lk(i) : -1;
    }
    const int p0 = (tid % CG) * 8, r0 = (tid / CG) * 8;
    float2 acc2[8][4];                              // [row][column pair]
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc2[i][j] = make_float2(0.f, 0.f);
    pdl_wait();
    const int chunks = K / KB, total = g.nA * chunks;
    float4 v[LPT];
    auto fetch = [&](int c) {
        const float* base = (c >= chunks ? g.A[1] : g.A[0]) + (c % chunks) * KB;
#pragma unroll
        for (int i = 0; i < LPT; ++i) v[i] = aoff[i] >= 0 ? ld4(base + aoff[i]) : make_float4(0.f, 0.f, 0.f, 0.f);
    };
    fetch(0);
    for (int c = 0; c < total; ++c) {
        __syncthreads();                            // the previous chunk has been consumed (and W is staged)
#pragma unroll
        for (int i = 0; i < LPT; ++i) {
            float* d = As + lk(i) * AS + lrow(i);
            d[0] = v[i].x; d[AS] = v[i].y; d[2 * AS] = v[i].z; d[3 * AS] = v[i].w;
        }
        __syncthreads();
        if (c + 1 < total) fetch(c + 1);
        const float* wa = ws + ((size_t)(c / chunks) * K + (c % chunks) * KB) * P + p0;
#pragma unroll 2
        for (int k = 0; k < KB; ++k) {
            const float4 a0 = ld4(As + k * AS + r0), a1 = ld4(As + k * AS + r0 + 4);
            const float4 w0 = ld4(wa + k * P), w1 = ld4(wa + k * P + 4);
            const float av[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
            const float wv[8] = {w0.x, w0.y, w0.z, w0.w, w1.x, w1.y, w1.z, w1.w};
#pragma unroll
            for (int i = 0; i < 8; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    if (PACK) ffma2(acc2[i][j], make_float2(wv[2 * j], wv[2 * j + 1]), av[i]);
                    else { acc2[i][j].x = fmaf(av[i], wv[2 * j], acc2[i][j].x); acc2[i][j].y = fmaf(av[i], wv[2 * j + 1], acc2[i][j].y); }
                }
        }
    }
    float acc[8][8];
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) { acc[i][2 * j] = acc2[i][j].x; acc[i][2 * j + 1] = acc2[i][j].y; }
    float bv[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) bv[j] = g.bias ? __ldg(g.bias + p0 + j) : 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const long long n = row_base + r0 + i;
        if (n >= g.N) break;
        const long long no = (g.o_mapped ? map_pos(g.map, n) : n) * P + p0;
        float4 o0 = make_float4(acc[i][0] + bv[0], acc[i][1] + bv[1], acc[i][2] + bv[2], acc[i][3] + bv[3]);
        float4 o1 = make_float4(acc[i][4] + bv[4], acc[i][5] + bv[5], acc[i][6] + bv[6], acc[i][7] + bv[7]);
        if (g.res) {
            const float4 q0 = ldg4_stream(g.res + no), q1 = ldg4_stream(g.res + no + 4);
            o0.x += q0.x; o0.y += q0.y; o0.z += q0.z; o0.w += q0.w; o1.x += q1.x; o1.y += q1.y; o1.z += q1.z; o1.w += q1.w;
        }
        st4(g.out + no, o0);
        st4(g.out + no + 4, o1);
    }
}

// ------------------------------------------------------------------------------------------------------------
// dW[j][k] += sum_n A[n][j] * B[n'][k],  db[j] += sum_n A[n][j]          (reduction over all (row, step) pairs)
//   columns k < kc0 of B come from Bm:  b_mode 0: n' = n;  1: n' = map_pos(n) (A mapped the same way if a_mapped);
//   columns k >= kc0 come from Bm2 at the previous step of the same sequence in processing order (n - 1, or n + 1 for
//   the reverse direction; a zero row at the sequence start) and go to dW2: one pass over dz yields dW_ih and dW_hh.
// Each CTA walks a contiguous range of n in slabs of 16 rows staged in shared memory (the next slab is prefetched into
// registers while the current one is multiplied); thread tile TJ x TK in registers.
// ------------------------------------------------------------------------------------------------------------
struct Outer {
    const float* A;
    const float* Bm;
    const float* Bm2;
    float* dW;
    float* dW2;
    float* db;
    float* db2;
    RowMap map;
    int a_mapped, b_mode, reverse, ldw, ldw2, kc0;
    long long N, rows_per_cta;
};

template <int J, int KC, int TJ, int TK, bool PACK>
__global__ void __launch_bounds__(256) outer_kernel(const Outer o) {
    constexpr int RB = 16, NTJ = J / TJ, NTK = KC / TK;
    constexpr int NA = (RB * J / 4 + 255) / 256, NB = (RB * KC / 4 + 255) / 256;
    static_assert(NTJ * NTK <= 256 && TJ % 2 == 0 && TK % 2 == 0, "tile");
    __shared__ __align__(16) float As[RB][J];
    __shared__ __align__(16) float Bs[RB][KC];
    const int tid = threadIdx.x;
    const int tj = tid % NTJ, tk = tid / NTJ;
    const bool worker = tk < NTK;
    float2 acc2[TJ][TK / 2];
    float accb[TJ];
#pragma unroll
    for (int i = 0; i < TJ; ++i) {
        accb[i] = 0.f;
#pragma unroll
        for (int k = 0; k < TK / 2; ++k) acc2[i][k] = make_float2(0.f, 0.f);
    }
    const long long n_begin = (long long)blockIdx.x * o.rows_per_cta;
    const long long n_end = n_begin + o.rows_per_cta < o.N ? n_begin + o.rows_per_cta : o.N;
    const int kc0 = o.kc0;
    pdl_wait();
    float4 va[NA], vb[NB];
    auto fetch = [&](long long n0) {
#pragma unroll
        for (int q = 0; q < NA; ++q) {
            const int i = tid + 256 * q, r = i / (J / 4), c4 = i - r * (J / 4);
            const long long n = n0 + r;
            va[q] = make_float4(0.f, 0.f, 0.f, 0.f);
            if (i < RB * (J / 4) && n < n_end) va[q] = ldg4_stream(o.A + (o.a_mapped ? map_pos(o.map, n) : n) * J + 4 * c4);
        }
#pragma unroll
        for (int q = 0; q < NB; ++q) {
            const int i = tid + 256 * q, r = i / (KC / 4), col = 4 * (i - r * (KC / 4));
            const long long n = n0 + r;
            vb[q] = make_float4(0.f, 0.f, 0.f, 0.f);
            if (i < RB * (KC / 4) && n < n_end) {
                if (col < kc0) {
                    vb[q] = ldg4_stream(o.Bm + (o.b_mode == 1 ? map_pos(o.map, n) : n) * kc0 + col);
                } else {
                    const int s = (int)(n % o.map.S);
                    const bool zero = o.reverse ? s == o.map.S - 1 : s == 0;
                    if (!zero) vb[q] = ldg4_stream(o.Bm2 + (o.reverse ? n + 1 : n - 1) * (KC - kc0) + (col - kc0));
                }
            }
        }
    };
    if (n_begin < n_end) fetch(n_begin);
    for (long long n0 = n_begin; n0 < n_end; n0 += RB) {
        __syncthreads();
#pragma unroll
        for (int q = 0; q < NA; ++q) {
            const int i = tid + 256 * q;
            if (i < RB * (J / 4)) st4(&As[0][0] + 4 * i, va[q]);
        }
#pragma unroll
        for (int q = 0; q < NB; ++q) {
            const int i = tid + 256 * q;
            if (i < RB * (KC / 4)) st4(&Bs[0][0] + 4 * i, vb[q]);
        }
        __syncthreads();
        if (n0 + RB < n_end) fetch(n0 + RB);
        if (worker) {
#pragma unroll 4
            for (int r = 0; r < RB; ++r) {
                float av[TJ], bv[TK];
#pragma unroll
                for (int i = 0; i < TJ; i += 2) { const float2 t = ld2(&As[r][tj * TJ + i]); av[i] = t.x; av[i + 1] = t.y; }
#pragma unroll
                for (int k = 0; k < TK; k += 2) { const float2 t = ld2(&Bs[r][tk * TK + k]); bv[k] = t.x; bv[k + 1] = t.y; }
#pragma unroll
                for (int i = 0; i < TJ; ++i) {
                    if (tk == 0) accb[i] += av[i];
#pragma unroll
                    for (int k = 0; k < TK / 2; ++k) {
                        if (PACK) ffma2(acc2[i][k], make_float2(bv[2 * k], bv[2 * k + 1]), av[i]);
                        else { acc2[i][k].x = fmaf(av[i], bv[2 * k], acc2[i][k].x); acc2[i][k].y = fmaf(av[i], bv[2 * k + 1], acc2[i][k].y); }
                    }
                }
            }
        }
    }
    float acc[TJ][TK];
#pragma unroll
    for (int i = 0; i < TJ; ++i)
#pragma unroll
        for (int k = 0; k < TK / 2; ++k) { acc[i][2 * k] = acc2[i][k].x; acc[i][2 * k + 1] = acc2[i][k].y; }
    if (worker) {
#pragma unroll
        for (int i = 0; i < TJ; ++i) {
            const int j = tj * TJ + i;
#pragma unroll
            for (int k = 0; k < TK; ++k) {
                const int col = tk * TK + k;
                if (col < kc0) atomic_add(o.dW + (size_t)j * o.ldw + col, acc[i][k]);
                else atomic_add(o.dW2 + (size_t)j * o.ldw2 + (col - kc0), acc[i][k]);
            }
            if (tk == 0) {
                if (o.db) atomic_add(o.db + j, accb[i]);
                if (o.db2) atomic_add(o.db2 + j, accb[i]);
            }
        }
    }
}

// ------------------------------------------------------------------------------------------------------------
// FiLM apply
// ------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) film_apply_fwd_kernel(const sb_film_apply_args a) {
    pdl_wait();
    const long long n4 = (long long)a.B * a.T * a.F * a.C / 4, fc4 = (long long)a.F * a.C / 4, tfc4 = fc4 * a.T;
    for (long long i = (long long)blockIdx.x * 256 + threadIdx.x; i < n4; i += (long long)gridDim.x * 256) {
        const long long b = i / tfc4, r = i % fc4;
        const float4 v = ldg4_stream(a.x + 4 * i);
        const float4 s = __ldg(reinterpret_cast<const float4*>(a.film_scale) + b * fc4 + r);
        const float4 h = __ldg(reinterpret_cast<const float4*>(a.film_shift) + b * fc4 + r);
        st4(a.y + 4 * i, make_float4(fmaf(v.x, s.x, h.x), fmaf(v.y, s.y, h.y), fmaf(v.z, s.z, h.z), fmaf(v.w, s.w, h.w)));
    }
}

// thread (b, f, c) walks a slice of the frames: dL/dx = g * scale; dL/dscale += sum_t g * x; dL/dshift += sum_t g.
// blockIdx.y = one of gridDim.y frame slices (one thread per (b, f, c) over all 625 frames of a 5 s clip was a serial chain of
// 625 dependent round trips on 145 CTAs: 282 us for 280 MB); the slices meet in the accumulating gradient buffers by atomics.
__global__ void __launch_bounds__(256) film_apply_bwd_kernel(const sb_film_apply_args a) {
    pdl_wait();
    const long long fc = (long long)a.F * a.C, total = fc * a.B;
    const long long i = (long long)blockIdx.x * 256 + threadIdx.x;
    if (i >= total) return;
    const long long b = i / fc, r = i - b * fc;
    const float sc = __ldg(a.film_scale + i);
    const int per = (a.T + (int)gridDim.y - 1) / (int)gridDim.y;
    const int t_begin = (int)blockIdx.y * per, t_end = min(a.T, t_begin + per);
    float gs = 0.f, gh = 0.f;
    long long p = (b * a.T + t_begin) * fc + r;
#pragma unroll 4
    for (int t = t_begin; t < t_end; ++t, p += fc) {
        const float g = ld_plain(a.gy + p), x = ldg1_stream(a.x + p);
        gs = fmaf(g, x, gs);
        gh += g;
        a.gx[p] = g * sc;
    }
    if (gridDim.y == 1) {
        a.g_scale[i] += gs;
        a.g_shift[i] += gh;
    } else if (t_begin < t_end) {
        atomic_add(a.g_scale + i, gs);
        atomic_add(a.g_shift + i, gh);
    }
}

// Dis_Embed_Conv / Dis_Embed_Linear + the 1x1 convs of every FilmLayer, backward.  CTA = one utterance.
// Embedding entry (f, i) lives at flat index pos = f*Din + i (conv: view [B,F,Din], LayerNorm over Din) or i*F + f
// (linear: LayerNorm over the whole F*Din vector, view [B,Din,F]); pos is also its row of the embedding weight.
__device__ __forceinline__ float block_sum_256(float v, float* red) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    __syncthreads();
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
    __syncthreads();
    float t = 0.f;
    for (int w = 0; w < 8; ++w) t += red[w];
    return t;
}

__global__ void __launch_bounds__(256) film_params_bwd_kernel(const sb_film_bwd_args a) {
    SB_DYN_SMEM(float, sm);
    __shared__ float red[8];
    const sb_film_args& f = a.f;
    const int F = f.F, C = f.C, Din = f.Din, L = f.n_layers, b = blockIdx.x, tid = threadIdx.x, n = F * Din;
    const bool conv = f.emb_mode == SB_EMB_CONV;
    float* e_s = sm;                                // [n]  embedding (LayerNorm output)
    float* eh_s = e_s + n;                          // [n]  normalised
    float* rs_s = eh_s + n;                         // [F]  1/std per bin (conv mode)
    float* de_s = rs_s + F;                         // [n]
    auto pos = [&](int fq, int i) { return conv ? fq * Din + i : i * F + fq; };
    pdl_wait();
    const float d0 = __ldg(f.dis + b * 3), d1 = __ldg(f.dis + b * 3 + 1), d2 = __ldg(f.dis + b * 3 + 2);
    for (int i = tid; i < n; i += 256) {
        const float* w = f.emb_w + (size_t)i * 3;
        eh_s[i] = fmaf(__ldg(w + 2), d2, fmaf(__ldg(w + 1), d1, __ldg(w) * d0));
        de_s[i] = 0.f;
    }
    __syncthreads();
    float rstd_all = 0.f;
    if (conv) {
        for (int fq = tid; fq < F; fq += 256) {
            float mean = 0.f;
            for (int i = 0; i < Din; ++i) mean += eh_s[fq * Din + i];
            mean /= Din;
            float var = 0.f;
            for (int i = 0; i < Din; ++i) { const float v = eh_s[fq * Din + i] - mean; eh_s[fq * Din + i] = v; var = fmaf(v, v, var); }
            const float r = rsqrtf(var / Din + 1e-5f);
            rs_s[fq] = r;
            for (int i = 0; i < Din; ++i) {
                const float h = eh_s[fq * Din + i] * r;
                eh_s[fq * Din + i] = h;
                e_s[fq * Din + i] = fmaf(h, __ldg(f.emb_ln_g + i), __ldg(f.emb_ln_b + i));
            }
        }
    } else {
        float s1 = 0.f;
        for (int i = tid; i < n; i += 256) s1 += eh_s[i];
        const float mean = block_sum_256(s1, red) / n;
        float s2 = 0.f;
        for (int i = tid; i < n; i += 256) { const float v = eh_s[i] - mean; s2 = fmaf(v, v, s2); }
        rstd_all = rsqrtf(block_sum_256(s2, red) / n + 1e-5f);
        for (int i = tid; i < n; i += 256) {
            const float h = (eh_s[i] - mean) * rstd_all;
            eh_s[i] = h;
            e_s[i] = fmaf(h, __ldg(f.emb_ln_g + i), __ldg(f.emb_ln_b + i));
        }
    }
    __syncthreads();
    const size_t plane = (size_t)f.B * F * C;
    // (1) thread (layer, c): gradients of the two 1x1 convs, summed over f
    for (int lc = tid; lc < L * C; lc += 256) {
        const int l = lc / C, c = lc - l * C;
        const float* gs = a.g_film + ((size_t)(2 * l) * f.B + b) * F * C + c;
        const float* gh = gs + plane;
        float sw = 0.f, sb_ = 0.f;
        for (int i = 0; i < Din; ++i) {
            float aw = 0.f, ab = 0.f;
            for (int fq = 0; fq < F; ++fq) {
                const float e = e_s[pos(fq, i)];
                aw = fmaf(__ldg(gs + (size_t)fq * C), e, aw);
                ab = fmaf(__ldg(gh + (size_t)fq * C), e, ab);
            }
            atomic_add(a.g_w_w + (size_t)lc * Din + i, aw);
            atomic_add(a.g_b_w + (size_t)lc * Din + i, ab);
        }
        for (int fq = 0; fq < F; ++fq) { sw += __ldg(gs + (size_t)fq * C); sb_ += __ldg(gh + (size_t)fq * C); }
        atomic_add(a.g_w_b + lc, sw);
        atomic_add(a.g_b_b + lc, sb_);
    }
    // (2) thread f: dL/de
    for (int fq = tid; fq < F; fq += 256) {
        for (int l = 0; l < L; ++l) {
            const float* gs = a.g_film + (((size_t)(2 * l) * f.B + b) * F + fq) * C;
            const float* gh = gs + plane;
            for (int c = 0; c < C; ++c) {
                const float s = __ldg(gs + c), h = __ldg(gh + c);
                for (int i = 0; i < Din; ++i)
                    de_s[pos(fq, i)] += s * __ldg(f.w_w + (size_t)(l * C + c) * Din + i) + h * __ldg(f.b_w + (size_t)(l * C + c) * Din + i);
            }
        }
    }
    __syncthreads();
    // (3) LayerNorm backward and dL/d(embedding weight)
    if (conv) {
        for (int fq = tid; fq < F; fq += 256) {
            float m1 = 0.f, m2 = 0.f;
            for (int i = 0; i < Din; ++i) {
                const float de = de_s[fq * Din + i], h = eh_s[fq * Din + i];
                atomic_add(a.g_emb_ln_g + i, de * h);
                atomic_add(a.g_emb_ln_b + i, de);
                const float dg = de * __ldg(f.emb_ln_g + i);
                de_s[fq * Din + i] = dg;
                m1 += dg;
                m2 = fmaf(dg, h, m2);
            }
            m1 /= Din;
            m2 /= Din;
            for (int i = 0; i < Din; ++i) {
                const float dl = rs_s[fq] * (de_s[fq * Din + i] - m1 - eh_s[fq * Din + i] * m2);
                float* gw = a.g_emb_w + (size_t)(fq * Din + i) * 3;
                atomic_add(gw, dl * d0);
                atomic_add(gw + 1, dl * d1);
                atomic_add(gw + 2, dl * d2);
            }
        }
    } else {
        float m1 = 0.f, m2 = 0.f;
        for (int i = tid; i < n; i += 256) {
            const float de = de_s[i], h = eh_s[i];
            atomic_add(a.g_emb_ln_g + i, de * h);
            atomic_add(a.g_emb_ln_b + i, de);
            const float dg = de * __ldg(f.emb_ln_g + i);
            de_s[i] = dg;
            m1 += dg;
            m2 = fmaf(dg, h, m2);
        }
        m1 = block_sum_256(m1, red) / n;
        m2 = block_sum_256(m2, red) / n;
        for (int i = tid; i < n; i += 256) {
            const float dl = rstd_all * (de_s[i] - m1 - eh_s[i] * m2);
            float* gw = a.g_emb_w + (size_t)i * 3;
            atomic_add(gw, dl * d0);
            atomic_add(gw + 1, dl * d1);
            atomic_add(gw + 2, dl * d2);
        }
    }
}

// ------------------------------------------------------------------------------------------------------------
// conv-in for training (zero history): y[b,t,f,o] = bias[o] + sum_{c,kt,kf} w[o][c][kt][kf] * feats[b, t+kt-2, f+kf-1, c]
// ------------------------------------------------------------------------------------------------------------
template <int C>
__global__ void __launch_bounds__(256) conv_in_train_kernel(const sb_conv_in_train_args a, float* y) {
    SB_DYN_SMEM(float, w_s);                        // [kt][kf][cin][o]
    const int Cin = a.Cin, F = a.F, T = a.T, tid = threadIdx.x;
    for (int i = tid; i < 9 * Cin * C; i += 256) {
        const int o = i % C, c = (i / C) % Cin, tap = i / (C * Cin);
        w_s[i] = __ldg(a.w + ((size_t)o * Cin + c) * 9 + tap);
    }
    pdl_wait();
    __syncthreads();
    constexpr int TPR = C / 8, ROWS = 256 / TPR;
    const long long N = (long long)a.B * T * F;
    const long long n = (long long)blockIdx.x * ROWS + tid / TPR;
    if (n >= N) return;
    const int o0 = (tid % TPR) * 8;
    const int fq = (int)(n % F), t = (int)((n / F) % T);
    float acc[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) acc[i] = __ldg(a.bias + o0 + i);
    for (int kt = 0; kt < 3; ++kt) {
        if (t + kt - 2 < 0) continue;
        for (int kf = 0; kf < 3; ++kf) {
            const int ff = fq + kf - 1;
            if (ff < 0 || ff >= F) continue;
            const float* xr = a.feats + (n + (long long)(kt - 2) * F + (kf - 1)) * Cin;
            const float* wr = w_s + (size_t)(kt * 3 + kf) * Cin * C + o0;
            for (int c = 0; c < Cin; ++c) {
                const float v = __ldg(xr + c);
                const float4 w0 = ld4(wr + c * C), w1 = ld4(wr + c * C + 4);
                acc[0] = fmaf(v, w0.x, acc[0]); acc[1] = fmaf(v, w0.y, acc[1]); acc[2] = fmaf(v, w0.z, acc[2]); acc[3] = fmaf(v, w0.w, acc[3]);
                acc[4] = fmaf(v, w1.x, acc[4]); acc[5] = fmaf(v, w1.y, acc[5]); acc[6] = fmaf(v, w1.z, acc[6]); acc[7] = fmaf(v, w1.w, acc[7]);
            }
        }
    }
    st4(y + n * C + o0, make_float4(acc[0], acc[1], acc[2], acc[3]));
    st4(y + n * C + o0 + 4, make_float4(acc[4], acc[5], acc[6], acc[7]));
}

// dL/dw[o][c][kt][kf] += sum_n g[n][o] * feats[n shifted by (kt-2, kf-1)][c];  dL/dbias[o] += sum_n g[n][o].
// Thread = one (tap, c) pair with C accumulators; the g rows of 32 positions are staged in shared memory.
template <int C>
__global__ void __launch_bounds__(256) conv_in_wgrad_kernel(const sb_conv_in_train_args a, const float* g, long long rows_per_cta) {
    constexpr int RB = 32;
    __shared__ __align__(16) float g_s[RB][C];
    const int Cin = a.Cin, F = a.F, T = a.T, tid = threadIdx.x;
    const long long N = (long long)a.B * T * F;
    const long long n_begin = (long long)blockIdx.x * rows_per_cta;
    const long long n_end = n_begin + rows_per_cta < N ? n_begin + rows_per_cta : N;
    pdl_wait();
    for (int pair0 = 0; pair0 < 9 * Cin; pair0 += 256) {
        const int pair = pair0 + tid;
        const bool worker = pair < 9 * Cin;
        const int tap = worker ? pair / Cin : 0, c = worker ? pair - tap * Cin : 0, kt = tap / 3, kf = tap - kt * 3;
        float acc[C], accb = 0.f;
#pragma unroll
        for (int o = 0; o < C; ++o) acc[o] = 0.f;
        for (long long n0 = n_begin; n0 < n_end; n0 += RB) {
            __syncthreads();
            for (int i = tid; i < RB * C / 4; i += 256) {
                const long long n = n0 + i / (C / 4);
                st4(&g_s[0][0] + 4 * i, n < n_end ? ldg4_stream(g + n * C + 4 * (i % (C / 4))) : make_float4(0.f, 0.f, 0.f, 0.f));
            }
            __syncthreads();
            if (worker) {
                // (t, f) of the block's first position once, then carried along (two 64-bit divisions per position were most
                // of this kernel's 1.6 ms); the g row of a position arrives as eight LDS.128
                int fq = (int)(n0 % F), t = (int)((n0 / F) % T);
                float vr[RB];                               // the block's 32 input values first: all loads in flight together
                const float* fp = a.feats + (n0 + (long long)(kt - 2) * F + (kf - 1)) * Cin + c;
#pragma unroll
                for (int r = 0; r < RB; ++r) {
                    const long long n = n0 + r;
                    const int ff = fq + kf - 1;
                    const bool in = n < n_end && t + kt - 2 >= 0 && ff >= 0 && ff < F;
                    vr[r] = in ? __ldg(fp + r * Cin) : 0.f;
                    if (++fq == F) {
                        fq = 0;
                        if (++t == T) t = 0;
                    }
                }
#pragma unroll
                for (int r = 0; r < RB; ++r) {              // (a position outside the input contributes v = 0: acc unchanged)
                    const float v = vr[r];
#pragma unroll
                    for (int o = 0; o < C; o += 4) {
                        const float4 gv = ld4(&g_s[r][o]);
                        acc[o] = fmaf(v, gv.x, acc[o]); acc[o + 1] = fmaf(v, gv.y, acc[o + 1]);
                        acc[o + 2] = fmaf(v, gv.z, acc[o + 2]); acc[o + 3] = fmaf(v, gv.w, acc[o + 3]);
                    }
                }
            }
            if (pair0 == 0 && tid < C)
                for (int r = 0; r < RB; ++r) accb += g_s[r][tid];
        }
        if (worker)
#pragma unroll
            for (int o = 0; o < C; ++o) atomic_add(a.g_w + ((size_t)o * Cin + c) * 9 + tap, acc[o]);
        if (pair0 == 0 && tid < C) atomic_add(a.g_bias + tid, accb);
    }
}

// ------------------------------------------------------------------------------------------------------------
// back-end, backward.  Forward (zero history): spec[o=(s,ri)][t][f] = bias[o] + sum_{c,kt,kf} x[c][t-kt][f+1-kf] w[c][o][kt][kf],
// wave[192 t' + r] += sum_k spec_k[t'] filt[k][r] over frames t' in {t, t-1} (iSTFT/OLA, first `stride` samples dropped).
// ------------------------------------------------------------------------------------------------------------
// g_spec[b][t][s][k] = sum_r filt[k][r] * g_wave[b][s][stride*t + r]   (samples past stride*T are the cropped look-ahead)
// Thread k walks the n_fft taps of basis row k; the basis is staged through shared memory 32 taps at a time with coalesced
// reads (thread k reading filt[k][r] straight from global touched 32 different lines per load instruction, and ten warps of
// four resident CTAs thrashed the L1: 280 us for a 0.4 GMAC job).
__global__ void __launch_bounds__(320) istft_bwd_kernel(const sb_backend_bwd_args a) {
    constexpr int TT = 8, RC = 32;
    SB_DYN_SMEM(float, g_s);                        // [(TT-1)*stride + n_fft] then the basis chunk [2F][RC + 1]
    const int F2 = 2 * a.F, S = a.n_src, t0 = blockIdx.x * TT, bs = blockIdx.y, tid = threadIdx.x;
    const int span = (TT - 1) * a.stride + a.n_fft, len = a.stride * a.T;
    float* f_s = g_s + span;
    pdl_wait();
    const float* gw = a.g_wave + (size_t)bs * len;
    for (int i = tid; i < span; i += 320) {
        const long long p = (long long)t0 * a.stride + i;
        g_s[i] = p < len ? ldg1_stream(gw + p) : 0.f;
    }
    const int b = bs / S, s = bs - b * S;
    float acc[2][TT];                               // basis rows tid and tid + 320 (2F <= 640)
#pragma unroll
    for (int h = 0; h < 2; ++h)
#pragma unroll
        for (int j = 0; j < TT; ++j) acc[h][j] = 0.f;
    for (int r0 = 0; r0 < a.n_fft; r0 += RC) {
        const int nr = min(RC, a.n_fft - r0);
        __syncthreads();                            // g_s staged / the previous chunk consumed
        for (int i = tid; i < F2 * RC; i += 320) {
            const int k = i / RC, rr = i - k * RC;
            f_s[k * (RC + 1) + rr] = rr < nr ? __ldg(a.filt + (size_t)k * a.n_fft + r0 + rr) : 0.f;
        }
        __syncthreads();
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            const int k = tid + 320 * h;
            if (k < F2) {
                for (int rr = 0; rr < nr; ++rr) {
                    const float w = f_s[k * (RC + 1) + rr];
#pragma unroll
                    for (int j = 0; j < TT; ++j) acc[h][j] = fmaf(w, g_s[j * a.stride + r0 + rr], acc[h][j]);
                }
            }
        }
    }
#pragma unroll
    for (int h = 0; h < 2; ++h) {
        const int k = tid + 320 * h;
        if (k >= F2) continue;
#pragma unroll
        for (int j = 0; j < TT; ++j) {
            const int t = t0 + j;
            if (t >= a.T) break;
            const size_t idx = (((size_t)b * a.T + t) * S + s) * F2 + k;
            a.ws[idx] = a.mask_spec ? acc[h][j] * __ldg(a.mask_spec + idx) : acc[h][j];
        }
    }
}

// dL/dx[b][t][f][c] = sum_{o,kt,kf} g_spec_o[t+kt][f-1+kf] * w[c][o][kt][kf]; one thread per (position, channel)
template <int C>
__global__ void __launch_bounds__(256) deconv_bwd_x_kernel(const sb_backend_bwd_args a) {
    __shared__ float w_s[4 * 9 * C];                // [o][tap][c]
    const int O = 2 * a.n_src, F = a.F, T = a.T, tid = threadIdx.x;
    for (int i = tid; i < O * 9 * C; i += 256) {
        const int c = i % C, tap = (i / C) % 9, o = i / (9 * C);
        w_s[i] = __ldg(a.w + ((size_t)c * O + o) * 9 + tap);
    }
    pdl_wait();
    __syncthreads();
    const long long N = (long long)a.B * T * F;
    const long long n = (long long)blockIdx.x * (256 / C) + tid / C;
    if (n >= N) return;
    const int c = tid % C, fq = (int)(n % F), t = (int)((n / F) % T);
    const long long b = n / ((long long)F * T);
    float acc = 0.f;
    for (int o = 0; o < O; ++o) {
        const int s = o >> 1, ri = o & 1;
        for (int kt = 0; kt < 3; ++kt) {
            if (t + kt >= T) break;
            const float* gr = a.ws + (((size_t)b * T + t + kt) * a.n_src + s) * 2 * F + ri * F;
            for (int kf = 0; kf < 3; ++kf) {
                const int ff = fq - 1 + kf;
                if (ff < 0 || ff >= F) continue;
                acc = fmaf(__ldg(gr + ff), w_s[(o * 9 + kt * 3 + kf) * C + c], acc);
            }
        }
    }
    a.gx[n * C + c] = acc;
}

// dL/dw[c][o][kt][kf] += sum x[b][t-kt][f+1-kf][c] * g_spec_o[b][t][f];  dL/dbias[o] += sum g_spec_o.  Thread = (tap, c).
template <int C>
__global__ void __launch_bounds__(320) deconv_wgrad_kernel(const sb_backend_bwd_args a, long long rows_per_cta) {
    const int O = 2 * a.n_src, F = a.F, T = a.T, tid = threadIdx.x;
    if (tid >= 9 * C) return;
    const int tap = tid / C, c = tid - tap * C, kt = tap / 3, kf = tap - kt * 3;
    const long long N = (long long)a.B * T * F;
    const long long n_begin = (long long)blockIdx.x * rows_per_cta;
    const long long n_end = n_begin + rows_per_cta < N ? n_begin + rows_per_cta : N;
    float acc[4] = {0.f, 0.f, 0.f, 0.f}, accb[4] = {0.f, 0.f, 0.f, 0.f};
    pdl_wait();
    // (b, t, f) of the position are carried along instead of being re-derived with two 64-bit divisions per position, which
    // were most of this kernel's 1.9 ms (profiles/r02_launches_train.txt)
    int fq = (int)(n_begin % F), t = (int)((n_begin / F) % T);
    long long b = n_begin / ((long long)F * T);
    // four positions per pass (20 loads in flight together), and the loads of pass p + 1 are issued before the FMAs of pass p:
    // a warp issues in order, so without the second register set every pass exposed one full memory latency
    // (pointers are advanced, not recomputed: the 64-bit multiplies of the index arithmetic were the bulk of the instructions)
    const float* xp = a.x + (n_begin + (long long)(-kt) * F + (1 - kf)) * C + c;
    const float* gp = a.ws + ((size_t)b * T + t) * a.n_src * 2 * F + fq;
    const long long g_wrap = (long long)a.n_src * 2 * F - F;
    const long long g_src = 2 * F;
    auto load_pass = [&](long long n, float (&v)[4], float (&gq)[4][4]) {
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const bool live = n + i < n_end;
            const int ff = fq + 1 - kf;
            const bool in = live && t - kt >= 0 && ff >= 0 && ff < F;
            v[i] = in ? __ldg(xp) : 0.f;
#pragma unroll
            for (int o = 0; o < 4; ++o) gq[i][o] = (live && o < O) ? __ldg(gp + (o >> 1) * g_src + (o & 1) * F) : 0.f;
            xp += C;
            ++gp;
            if (++fq == F) {
                fq = 0;
                gp += g_wrap;
                if (++t == T) { t = 0; ++b; }
            }
        }
    };
    auto fma_pass = [&](const float (&v)[4], const float (&gq)[4][4]) {
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int o = 0; o < 4; ++o) {
                acc[o] = fmaf(v[i], gq[i][o], acc[o]);
                accb[o] += gq[i][o];
            }
    };
    float v0[4], g0[4][4], v1[4], g1[4][4];
    load_pass(n_begin, v0, g0);
    for (long long n = n_begin; n < n_end; n += 8) {
        load_pass(n + 4, v1, g1);                           // (positions past n_end load nothing and contribute zeros)
        fma_pass(v0, g0);
        load_pass(n + 8, v0, g0);
        fma_pass(v1, g1);
    }
#pragma unroll
    for (int o = 0; o < 4; ++o) {
        if (o < O) {
            atomic_add(a.g_w + ((size_t)c * O + o) * 9 + tap, acc[o]);
            if (tid == 0) atomic_add(a.g_bias + o, accb[o]);
        }
    }
}

// ------------------------------------------------------------------------------------------------------------
// conv-LSTM intra-frame path for training (a9': DE3 :800-815, OPT :684-697 with the deconv of :494-510):
//   x -> Conv1d(C -> C, k = s = down) over frequency -> PReLU -> LayerNorm(C) -> BiLSTM over J steps
//     -> ConvTranspose1d(2H -> C, k = s = down) -> (pad & crop | output_padding) -> + x
// k = s: the convolution is a per-group linear map, group n' = (b, t, j) covering bins j*k .. j*k + k - 1.
// ------------------------------------------------------------------------------------------------------------
struct ConvPath {
    const float* x;             // [B*T][F][C]
    const float* conv_w;        // [C][C][k]   (o, c, tap) as stored
    const float* conv_b;
    const float* deconv_w;      // [2H][C][k]  (d*H + u, c, tap) as stored
    const float* deconv_b;
    int BT, F, C, J, k, outpad;
};

// zraw[n'][o] = conv_b[o] + sum_{tap, c} x[bt][j*k + tap][c] * conv_w[o][c][tap]; thread = (group, 8 outputs)
template <int C>
__global__ void __launch_bounds__(256) convpre_train_kernel(const ConvPath a, float* zraw) {
    SB_DYN_SMEM(float, w_s);                        // [tap][c][o]
    const int k = a.k, tid = threadIdx.x;
    for (int i = tid; i < k * C * C; i += 256) {
        const int o = i % C, c = (i / C) % C, tap = i / (C * C);
        w_s[i] = __ldg(a.conv_w + ((size_t)o * C + c) * k + tap);
    }
    pdl_wait();
    __syncthreads();
    constexpr int TPR = C / 8;
    const long long N = (long long)a.BT * a.J;
    const long long n = (long long)blockIdx.x * (256 / TPR) + tid / TPR;
    if (n >= N) return;
    const int o0 = (tid % TPR) * 8;
    const long long bt = n / a.J;
    const int j = (int)(n - bt * a.J);
    const float* xr = a.x + ((size_t)bt * a.F + (size_t)j * k) * C;       // k*C contiguous floats
    float acc[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) acc[i] = __ldg(a.conv_b + o0 + i);
    for (int q = 0; q < k * C; ++q) {
        const float v = __ldg(xr + q);
        const float4 w0 = ld4(w_s + q * C + o0), w1 = ld4(w_s + q * C + o0 + 4);
        acc[0] = fmaf(v, w0.x, acc[0]); acc[1] = fmaf(v, w0.y, acc[1]); acc[2] = fmaf(v, w0.z, acc[2]); acc[3] = fmaf(v, w0.w, acc[3]);
        acc[4] = fmaf(v, w1.x, acc[4]); acc[5] = fmaf(v, w1.y, acc[5]); acc[6] = fmaf(v, w1.z, acc[6]); acc[7] = fmaf(v, w1.w, acc[7]);
    }
    st4(zraw + n * C + o0, make_float4(acc[0], acc[1], acc[2], acc[3]));
    st4(zraw + n * C + o0 + 4, make_float4(acc[4], acc[5], acc[6], acc[7]));
}

__global__ void __launch_bounds__(256) prelu_fwd_kernel(const float* z, const float* slope, float* out, long long n4) {
    pdl_wait();
    const float a = __ldg(slope);
    for (long long i = (long long)blockIdx.x * 256 + threadIdx.x; i < n4; i += (long long)gridDim.x * 256) {
        float4 v = ldg4_stream(z + 4 * i);
        v.x = v.x > 0.f ? v.x : a * v.x; v.y = v.y > 0.f ? v.y : a * v.y; v.z = v.z > 0.f ? v.z : a * v.z; v.w = v.w > 0.f ? v.w : a * v.w;
        st4(out + 4 * i, v);
    }
}

// in place: g <- g * (z > 0 ? 1 : slope);  dL/dslope += sum g * z over z <= 0
__global__ void __launch_bounds__(256) prelu_bwd_kernel(float* g, const float* z, const float* slope, float* g_slope, long long n) {
    __shared__ float red[8];
    pdl_wait();
    const float a = __ldg(slope);
    float acc = 0.f;
    for (long long i = (long long)blockIdx.x * 256 + threadIdx.x; i < n; i += (long long)gridDim.x * 256) {
        const float zv = z[i], gv = g[i];
        if (zv > 0.f) continue;
        acc = fmaf(gv, zv, acc);
        g[i] = gv * a;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
        float t = 0.f;
        for (int i = 0; i < 8; ++i) t += red[i];
        atomic_add(g_slope, t);
    }
}

// y[bt][f][c] = x + (f < J*k ? deconv_b[c] + sum_{d,u} h_d[n'][u] * deconv_w[d*H+u][c][tap] : tail); thread = (position, 8 channels)
template <int C>
__global__ void __launch_bounds__(256) convpost_train_kernel(const ConvPath a, const float* h0, const float* h1, float* y) {
    SB_DYN_SMEM(float, w_s);                        // [tap][2H][c]
    constexpr int H = kH;
    const int k = a.k, tid = threadIdx.x;
    for (int i = tid; i < k * 2 * H * C; i += 256) {
        const int c = i % C, du = (i / C) % (2 * H), tap = i / (C * 2 * H);
        w_s[i] = __ldg(a.deconv_w + ((size_t)du * C + c) * k + tap);
    }
    pdl_wait();
    __syncthreads();
    constexpr int TPR = C / 8;
    const long long N = (long long)a.BT * a.F;
    const long long n = (long long)blockIdx.x * (256 / TPR) + tid / TPR;
    if (n >= N) return;
    const int c0 = (tid % TPR) * 8;
    const long long bt = n / a.F;
    const int f = (int)(n - bt * a.F);
    float acc[8];
    const float4 x0 = ldg4_stream(a.x + n * C + c0), x1 = ldg4_stream(a.x + n * C + c0 + 4);
    acc[0] = x0.x; acc[1] = x0.y; acc[2] = x0.z; acc[3] = x0.w; acc[4] = x1.x; acc[5] = x1.y; acc[6] = x1.z; acc[7] = x1.w;
    if (f < a.J * k || a.outpad) {
#pragma unroll
        for (int i = 0; i < 8; ++i) acc[i] += __ldg(a.deconv_b + c0 + i);
    }
    if (f < a.J * k) {
        const int j = f / k, tap = f - j * k;
        const long long np = bt * a.J + j;
#pragma unroll
        for (int d = 0; d < 2; ++d) {
            const float* hr = (d ? h1 : h0) + np * H;
            const float* wr = w_s + ((size_t)tap * 2 * H + d * H) * C + c0;
            for (int u = 0; u < H; ++u) {
                const float v = hr[u];
                const float4 w0 = ld4(wr + u * C), w1 = ld4(wr + u * C + 4);
                acc[0] = fmaf(v, w0.x, acc[0]); acc[1] = fmaf(v, w0.y, acc[1]); acc[2] = fmaf(v, w0.z, acc[2]); acc[3] = fmaf(v, w0.w, acc[3]);
                acc[4] = fmaf(v, w1.x, acc[4]); acc[5] = fmaf(v, w1.y, acc[5]); acc[6] = fmaf(v, w1.z, acc[6]); acc[7] = fmaf(v, w1.w, acc[7]);
            }
        }
    }
    st4(y + n * C + c0, make_float4(acc[0], acc[1], acc[2], acc[3]));
    st4(y + n * C + c0 + 4, make_float4(acc[4], acc[5], acc[6], acc[7]));
}

// dL/dh_d[n'][u] = sum_{tap, c} gy[bt][j*k + tap][c] * deconv_w[d*H+u][c][tap]; thread = (group, 8 of the 2H units)
template <int C>
__global__ void __launch_bounds__(256) convpost_dh_kernel(const ConvPath a, const float* gy, float* dh0, float* dh1) {
    SB_DYN_SMEM(float, w_s);                        // [tap][c][2H]
    constexpr int H = kH;
    const int k = a.k, tid = threadIdx.x;
    for (int i = tid; i < k * C * 2 * H; i += 256) {
        const int du = i % (2 * H), c = (i / (2 * H)) % C, tap = i / (2 * H * C);
        w_s[i] = __ldg(a.deconv_w + ((size_t)du * C + c) * k + tap);
    }
    pdl_wait();
    __syncthreads();
    const long long N = (long long)a.BT * a.J;
    const long long n = (long long)blockIdx.x * 16 + tid / 16;
    if (n >= N) return;
    const int u0 = (tid % 16) * 8;                  // 0 .. 127 over both directions
    const long long bt = n / a.J;
    const int j = (int)(n - bt * a.J);
    const float* gr = gy + ((size_t)bt * a.F + (size_t)j * k) * C;
    float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    for (int q = 0; q < k * C; ++q) {
        const float v = gr[q];
        const float4 w0 = ld4(w_s + q * 2 * H + u0), w1 = ld4(w_s + q * 2 * H + u0 + 4);
        acc[0] = fmaf(v, w0.x, acc[0]); acc[1] = fmaf(v, w0.y, acc[1]); acc[2] = fmaf(v, w0.z, acc[2]); acc[3] = fmaf(v, w0.w, acc[3]);
        acc[4] = fmaf(v, w1.x, acc[4]); acc[5] = fmaf(v, w1.y, acc[5]); acc[6] = fmaf(v, w1.z, acc[6]); acc[7] = fmaf(v, w1.w, acc[7]);
    }
    float* out = (u0 < H ? dh0 + n * H + u0 : dh1 + n * H + (u0 - H));
    st4(out, make_float4(acc[0], acc[1], acc[2], acc[3]));
    st4(out + 4, make_float4(acc[4], acc[5], acc[6], acc[7]));
}

// dL/ddeconv_w[du][c][tap] += sum_{n'} h[n'][du] * gy[bt][j*k+tap][c];  dL/ddeconv_b[c] += sum over the bins that receive
// the bias.  Thread = (tap, c) with 2H accumulators; the h rows of 8 groups are staged in shared memory.
template <int C>
__global__ void __launch_bounds__(256) convpost_wgrad_kernel(const ConvPath a, const float* gy, const float* h0, const float* h1,
                                                            float* g_w, float* g_b, long long groups_per_cta) {
    constexpr int H = kH, RB = 8;
    __shared__ __align__(16) float h_s[RB][2 * H];
    const int k = a.k, tid = threadIdx.x;
    const bool worker = tid < k * C;
    const int tap = worker ? tid / C : 0, c = worker ? tid - tap * C : 0;
    const long long N = (long long)a.BT * a.J;
    const long long n_begin = (long long)blockIdx.x * groups_per_cta;
    const long long n_end = n_begin + groups_per_cta < N ? n_begin + groups_per_cta : N;
    float acc[2 * H], accb = 0.f;
#pragma unroll
    for (int i = 0; i < 2 * H; ++i) acc[i] = 0.f;
    pdl_wait();
    for (long long n0 = n_begin; n0 < n_end; n0 += RB) {
        __syncthreads();
        for (int i = tid; i < RB * 2 * H / 4; i += 256) {
            const int r = i / (2 * H / 4), q = i - r * (2 * H / 4);
            const long long n = n0 + r;
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            if (n < n_end) v = q < H / 4 ? ldg4_stream(h0 + n * H + 4 * q) : ldg4_stream(h1 + n * H + 4 * (q - H / 4));
            st4(&h_s[r][4 * q], v);
        }
        __syncthreads();
        if (worker) {
            for (int r = 0; r < RB; ++r) {
                const long long n = n0 + r;
                if (n >= n_end) break;
                const long long bt = n / a.J;
                const int j = (int)(n - bt * a.J);
                const float g = __ldg(gy + ((size_t)bt * a.F + (size_t)j * k + tap) * C + c);
                accb += g;
                if (a.outpad && tap == 0 && j == a.J - 1)          // output_padding bins receive the bias only
                    for (int f = a.J * k; f < a.F; ++f) accb += __ldg(gy + ((size_t)bt * a.F + f) * C + c);
#pragma unroll
                for (int i = 0; i < 2 * H; i += 4) {
                    const float4 hv = ld4(&h_s[r][i]);
                    acc[i] = fmaf(g, hv.x, acc[i]); acc[i + 1] = fmaf(g, hv.y, acc[i + 1]);
                    acc[i + 2] = fmaf(g, hv.z, acc[i + 2]); acc[i + 3] = fmaf(g, hv.w, acc[i + 3]);
                }
            }
        }
    }
    if (worker) {
#pragma unroll
        for (int i = 0; i < 2 * H; ++i) atomic_add(g_w + ((size_t)i * C + c) * k + tap, acc[i]);
        atomic_add(g_b + c, accb);
    }
}

// dL/dx[bt][f][c] = gy + (f < J*k ? sum_o dz[n'][o] * conv_w[o][c][tap] : 0); thread = (position, 8 channels)
template <int C>
__global__ void __launch_bounds__(256) convpre_bwd_x_kernel(const ConvPath a, const float* gy, const float* dz, float* gx) {
    SB_DYN_SMEM(float, w_s);                        // [tap][o][c]
    const int k = a.k, tid = threadIdx.x;
    for (int i = tid; i < k * C * C; i += 256) {
        const int c = i % C, o = (i / C) % C, tap = i / (C * C);
        w_s[i] = __ldg(a.conv_w + ((size_t)o * C + c) * k + tap);
    }
    pdl_wait();
    __syncthreads();
    constexpr int TPR = C / 8;
    const long long N = (long long)a.BT * a.F;
    const long long n = (long long)blockIdx.x * (256 / TPR) + tid / TPR;
    if (n >= N) return;
    const int c0 = (tid % TPR) * 8;
    const long long bt = n / a.F;
    const int f = (int)(n - bt * a.F);
    const float4 g0 = ld_plain4(gy + n * C + c0), g1 = ld_plain4(gy + n * C + c0 + 4);
    float acc[8] = {g0.x, g0.y, g0.z, g0.w, g1.x, g1.y, g1.z, g1.w};
    if (f < a.J * k) {
        const int j = f / k, tap = f - j * k;
        const float* zr = dz + (bt * a.J + j) * C;
        const float* wr = w_s + (size_t)tap * C * C + c0;
        for (int o = 0; o < C; ++o) {
            const float v = zr[o];
            const float4 w0 = ld4(wr + o * C), w1 = ld4(wr + o * C + 4);
            acc[0] = fmaf(v, w0.x, acc[0]); acc[1] = fmaf(v, w0.y, acc[1]); acc[2] = fmaf(v, w0.z, acc[2]); acc[3] = fmaf(v, w0.w, acc[3]);
            acc[4] = fmaf(v, w1.x, acc[4]); acc[5] = fmaf(v, w1.y, acc[5]); acc[6] = fmaf(v, w1.z, acc[6]); acc[7] = fmaf(v, w1.w, acc[7]);
        }
    }
    st4(gx + n * C + c0, make_float4(acc[0], acc[1], acc[2], acc[3]));
    st4(gx + n * C + c0 + 4, make_float4(acc[4], acc[5], acc[6], acc[7]));
}

// dL/dconv_w[o][c][tap] += sum_{n'} dz[n'][o] * x[bt][j*k+tap][c];  dL/dconv_b[o] += sum dz[n'][o].  Thread = (tap, c).
template <int C>
__global__ void __launch_bounds__(256) convpre_wgrad_kernel(const ConvPath a, const float* dz, float* g_w, float* g_b, long long groups_per_cta) {
    constexpr int RB = 32;
    __shared__ __align__(16) float g_s[RB][C];
    const int k = a.k, tid = threadIdx.x;
    const bool worker = tid < k * C;
    const int tap = worker ? tid / C : 0, c = worker ? tid - tap * C : 0;
    const long long N = (long long)a.BT * a.J;
    const long long n_begin = (long long)blockIdx.x * groups_per_cta;
    const long long n_end = n_begin + groups_per_cta < N ? n_begin + groups_per_cta : N;
    float acc[C], accb = 0.f;
#pragma unroll
    for (int o = 0; o < C; ++o) acc[o] = 0.f;
    pdl_wait();
    for (long long n0 = n_begin; n0 < n_end; n0 += RB) {
        __syncthreads();
        for (int i = tid; i < RB * C / 4; i += 256) {
            const long long n = n0 + i / (C / 4);
            st4(&g_s[0][0] + 4 * i, n < n_end ? ldg4_stream(dz + n * C + 4 * (i % (C / 4))) : make_float4(0.f, 0.f, 0.f, 0.f));
        }
        __syncthreads();
        if (worker) {
            for (int r = 0; r < RB; ++r) {
                const long long n = n0 + r;
                if (n >= n_end) break;
                const long long bt = n / a.J;
                const int j = (int)(n - bt * a.J);
                const float v = __ldg(a.x + ((size_t)bt * a.F + (size_t)j * k + tap) * C + c);
#pragma unroll
                for (int o = 0; o < C; ++o) acc[o] = fmaf(v, g_s[r][o], acc[o]);
            }
        }
        if (tid < C)
            for (int r = 0; r < RB; ++r) accb += g_s[r][tid];
    }
    if (worker)
#pragma unroll
        for (int o = 0; o < C; ++o) atomic_add(g_w + ((size_t)o * C + c) * k + tap, acc[o]);
    if (tid < C) atomic_add(g_b + tid, accb);
}

// ------------------------------------------------------------------------------------------------------------
// Sliding-window attention for training (a11: DE3 :639-684, 856-898, 722-744) from zero K / V history.  First version:
// one CTA per head-row, every contraction a plain loop (the attention is dormant in every shipped configuration).
//   head-row r = (b*L + l)*T + t;  element i of a head-row <-> bin f = i / ed, channel l*ed + e of position (b, t, f)
// ------------------------------------------------------------------------------------------------------------
struct AttnDims {
    int B, T, F, C, L, E, W, Vd;
};

// zq/zk [n][L*E], zv [n][C] = Linear(x[n]) before the PReLU; thread = position
template <int C>
__global__ void __launch_bounds__(128) attn_proj_train_kernel(const float* x, const float* wq, const float* bq, const float* wk, const float* bk,
                                                              const float* wv, const float* bv, float* zq, float* zk, float* zv, int LE, long long N) {
    __shared__ float w_s[(2 * 32 + C) * C];
    __shared__ float b_s[2 * 32 + C];
    const int tid = threadIdx.x, O = 2 * LE + C;
    for (int i = tid; i < O * C; i += 128) {
        const int o = i / C, c = i - o * C;
        w_s[i] = o < LE ? __ldg(wq + o * C + c) : o < 2 * LE ? __ldg(wk + (o - LE) * C + c) : __ldg(wv + (o - 2 * LE) * C + c);
    }
    for (int o = tid; o < O; o += 128) b_s[o] = o < LE ? __ldg(bq + o) : o < 2 * LE ? __ldg(bk + o - LE) : __ldg(bv + o - 2 * LE);
    pdl_wait();
    __syncthreads();
    const long long n = (long long)blockIdx.x * 128 + tid;
    if (n >= N) return;
    float xv[C];
#pragma unroll
    for (int c = 0; c < C; c += 4) { const float4 t = ldg4_stream(x + n * C + c); xv[c] = t.x; xv[c + 1] = t.y; xv[c + 2] = t.z; xv[c + 3] = t.w; }
    for (int o = 0; o < O; ++o) {
        float acc = b_s[o];
#pragma unroll
        for (int c = 0; c < C; ++c) acc = fmaf(xv[c], w_s[o * C + c], acc);
        if (o < LE) zq[n * LE + o] = acc;
        else if (o < 2 * LE) zk[n * LE + o - LE] = acc;
        else zv[n * C + o - 2 * LE] = acc;
    }
}

// rows of D elements: PReLU -> LayerNorm(D).  head = 1: row r is a head-row, element i comes from z[pos(b,t,f)][l*ed + e];
// head = 0: row = (b, t), element i = z[row*D + i] and the result is added to res (the block's residual).
struct RowLn {
    const float* z;             // pre-PReLU values
    const float* slope;
    const float* g;
    const float* b;
    const float* res;           // head = 0 only
    float* out;                 // [row][D]
    float* stats;               // [row][2] mean, 1/std
    const float* gout;          // bwd: dL/d(out)
    float* dz;                  // bwd: dL/dz, same addressing as z
    float* g_g; float* g_b; float* g_slope;
    int D, head, ed, zstride, L, T, F;
};
__device__ __forceinline__ long long rowln_src(const RowLn& a, long long row, int i) {
    if (!a.head) return row * a.D + i;
    const long long b = row / ((long long)a.L * a.T);
    const int l = (int)((row / a.T) % a.L), t = (int)(row % a.T);
    const int f = i / a.ed, e = i - f * a.ed;
    return ((b * a.T + t) * a.F + f) * a.zstride + l * a.ed + e;
}
__global__ void __launch_bounds__(256) prelu_ln_rows_fwd_kernel(const RowLn a) {
    __shared__ float red[8];
    pdl_wait();
    const long long row = blockIdx.x;
    const float sl = __ldg(a.slope);
    float s1 = 0.f;
    for (int i = threadIdx.x; i < a.D; i += 256) { const float v = a.z[rowln_src(a, row, i)]; s1 += v > 0.f ? v : sl * v; }
    const float mean = block_sum_256(s1, red) / a.D;
    float s2 = 0.f;
    for (int i = threadIdx.x; i < a.D; i += 256) {
        float v = a.z[rowln_src(a, row, i)];
        v = (v > 0.f ? v : sl * v) - mean;
        s2 = fmaf(v, v, s2);
    }
    const float rstd = rsqrtf(block_sum_256(s2, red) / a.D + 1e-5f);
    if (threadIdx.x == 0) { a.stats[2 * row] = mean; a.stats[2 * row + 1] = rstd; }
    for (int i = threadIdx.x; i < a.D; i += 256) {
        float v = a.z[rowln_src(a, row, i)];
        v = ((v > 0.f ? v : sl * v) - mean) * rstd;
        v = fmaf(v, __ldg(a.g + i), __ldg(a.b + i));
        if (a.res) v += a.res[row * a.D + i];
        a.out[row * a.D + i] = v;
    }
}
__global__ void __launch_bounds__(256) prelu_ln_rows_bwd_kernel(const RowLn a) {
    __shared__ float red[8];
    pdl_wait();
    const long long row = blockIdx.x;
    const float sl = __ldg(a.slope), mean = a.stats[2 * row], rstd = a.stats[2 * row + 1];
    float m1 = 0.f, m2 = 0.f;
    for (int i = threadIdx.x; i < a.D; i += 256) {
        const float z = a.z[rowln_src(a, row, i)];
        const float h = ((z > 0.f ? z : sl * z) - mean) * rstd;
        const float g = a.gout[row * a.D + i];
        atomic_add(a.g_g + i, g * h);
        atomic_add(a.g_b + i, g);
        const float dh = g * __ldg(a.g + i);
        m1 += dh;
        m2 = fmaf(dh, h, m2);
    }
    m1 = block_sum_256(m1, red) / a.D;
    m2 = block_sum_256(m2, red) / a.D;
    float ds = 0.f;
    for (int i = threadIdx.x; i < a.D; i += 256) {
        const long long src = rowln_src(a, row, i);
        const float z = a.z[src];
        const float h = ((z > 0.f ? z : sl * z) - mean) * rstd;
        const float dp = rstd * (a.gout[row * a.D + i] * __ldg(a.g + i) - m1 - h * m2);
        if (z > 0.f) a.dz[src] = dp;
        else { a.dz[src] = dp * sl; ds = fmaf(dp, z, ds); }
    }
    ds = block_sum_256(ds, red);
    if (threadIdx.x == 0) atomic_add(a.g_slope, ds);
}

// forward core: att[r][w] = softmax_w(q_r . k_{s(w)} / sqrt(DQ)), s(w) = t - W + 1 + w (frames < 0: zero key and value,
// NOT masked: they take part in the softmax with logit 0); o = sum_w att[w] v_{s(w)} written in [B][T][F][C] order
__global__ void __launch_bounds__(128) attn_core_train_fwd_kernel(const AttnDims d, const float* qn, const float* kn, const float* vn,
                                                                  float* att, float* ob) {
    SB_DYN_SMEM(float, sm);
    const int DQ = d.F * d.E, DV = d.F * d.Vd, W = d.W, tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    float* q_s = sm;                        // [DQ]
    float* l_s = q_s + DQ;                  // [W]
    pdl_wait();
    const long long r = blockIdx.x;
    const long long bl = r / d.T;
    const int t = (int)(r - bl * d.T);
    for (int i = tid; i < DQ; i += 128) q_s[i] = qn[r * DQ + i];
    __syncthreads();
    const float scale = rsqrtf((float)DQ);
    for (int w = warp; w < W; w += 4) {
        const int s = t - W + 1 + w;
        float acc = 0.f;
        if (s >= 0) {
            const float* kr = kn + (bl * d.T + s) * DQ;
            for (int i = lane; i < DQ; i += 32) acc = fmaf(q_s[i], kr[i], acc);
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
        if (lane == 0) l_s[w] = acc * scale;
    }
    __syncthreads();
    if (warp == 0) {
        float mx = -3.4e38f;
        for (int w = lane; w < W; w += 32) mx = fmaxf(mx, l_s[w]);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
        float sum = 0.f;
        for (int w = lane; w < W; w += 32) { const float e = __expf(l_s[w] - mx); l_s[w] = e; sum += e; }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
        const float inv = 1.0f / sum;
        for (int w = lane; w < W; w += 32) { const float p = l_s[w] * inv; l_s[w] = p; att[r * W + w] = p; }
    }
    __syncthreads();
    const long long b = bl / d.L;
    const int l = (int)(bl - b * d.L);
    const int w0 = t - W + 1 < 0 ? W - 1 - t : 0;      // first window slot that holds a real frame
    for (int i = tid; i < DV; i += 128) {
        float acc = 0.f;
        for (int w = w0; w < W; ++w) acc = fmaf(l_s[w], vn[(bl * d.T + (t - W + 1 + w)) * DV + i], acc);
        const int f = i / d.Vd, v = i - f * d.Vd;
        ob[((b * d.T + t) * d.F + f) * d.C + l * d.Vd + v] = acc;
    }
}

// backward, query side: dl[r][w] = att (datt - sum att datt), datt[w] = dO_r . v_s;  dq_r = sum_w dl[w] k_s / sqrt(DQ)
__global__ void __launch_bounds__(128) attn_core_bwd_q_kernel(const AttnDims d, const float* kn, const float* vn, const float* att,
                                                              const float* dob, float* dl, float* dqn) {
    SB_DYN_SMEM(float, sm);
    __shared__ float red[4];
    const int DQ = d.F * d.E, DV = d.F * d.Vd, W = d.W, tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    float* do_s = sm;                       // [DV]
    float* l_s = do_s + DV;                 // [W]
    pdl_wait();
    const long long r = blockIdx.x;
    const long long bl = r / d.T;
    const int t = (int)(r - bl * d.T);
    const long long b = bl / d.L;
    const int l = (int)(bl - b * d.L);
    for (int i = tid; i < DV; i += 128) {
        const int f = i / d.Vd, v = i - f * d.Vd;
        do_s[i] = dob[((b * d.T + t) * d.F + f) * d.C + l * d.Vd + v];
    }
    __syncthreads();
    for (int w = warp; w < W; w += 4) {
        const int s = t - W + 1 + w;
        float acc = 0.f;
        if (s >= 0) {
            const float* vr = vn + (bl * d.T + s) * DV;
            for (int i = lane; i < DV; i += 32) acc = fmaf(do_s[i], vr[i], acc);
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
        if (lane == 0) l_s[w] = acc;
    }
    __syncthreads();
    float part = 0.f;
    for (int w = tid; w < W; w += 128) part = fmaf(att[r * W + w], l_s[w], part);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) part += __shfl_xor_sync(0xffffffffu, part, o);
    if (lane == 0) red[warp] = part;
    __syncthreads();
    const float dot = (red[0] + red[1]) + (red[2] + red[3]);
    __syncthreads();
    for (int w = tid; w < W; w += 128) {
        const float v = att[r * W + w] * (l_s[w] - dot);
        l_s[w] = v;
        dl[r * W + w] = v;
    }
    __syncthreads();
    const float scale = rsqrtf((float)DQ);
    const int w0 = t - W + 1 < 0 ? W - 1 - t : 0;
    for (int i = tid; i < DQ; i += 128) {
        float acc = 0.f;
        for (int w = w0; w < W; ++w) acc = fmaf(l_s[w], kn[(bl * d.T + (t - W + 1 + w)) * DQ + i], acc);
        dqn[r * DQ + i] = acc * scale;
    }
}

// backward, key / value side: frame s is slot w = s - t + W - 1 of the queries t = s .. min(s + W - 1, T - 1)
__global__ void __launch_bounds__(128) attn_core_bwd_kv_kernel(const AttnDims d, const float* qn, const float* att, const float* dl,
                                                               const float* dob, float* dkn, float* dvn) {
    const int DQ = d.F * d.E, DV = d.F * d.Vd, W = d.W, tid = threadIdx.x;
    pdl_wait();
    const long long r = blockIdx.x;
    const long long bl = r / d.T;
    const int s = (int)(r - bl * d.T);
    const long long b = bl / d.L;
    const int l = (int)(bl - b * d.L);
    const int t_end = s + W - 1 < d.T - 1 ? s + W - 1 : d.T - 1;
    const float scale = rsqrtf((float)DQ);
    for (int i = tid; i < DQ; i += 128) {
        float acc = 0.f;
        for (int t = s; t <= t_end; ++t) acc = fmaf(dl[(bl * d.T + t) * W + (s - t + W - 1)], qn[(bl * d.T + t) * DQ + i], acc);
        dkn[r * DQ + i] = acc * scale;
    }
    for (int i = tid; i < DV; i += 128) {
        const int f = i / d.Vd, v = i - f * d.Vd;
        float acc = 0.f;
        for (int t = s; t <= t_end; ++t)
            acc = fmaf(att[(bl * d.T + t) * W + (s - t + W - 1)], dob[((b * d.T + t) * d.F + f) * d.C + l * d.Vd + v], acc);
        dvn[r * DV + i] = acc;
    }
}

// dL/dx[n] = gy[n] + dzq[n] Wq + dzk[n] Wk + dzv[n] Wv; thread = position
template <int C>
__global__ void __launch_bounds__(128) attn_proj_bwd_x_kernel(const float* gy, const float* dzq, const float* dzk, const float* dzv,
                                                              const float* wq, const float* wk, const float* wv, float* gx, int LE, long long N) {
    __shared__ float w_s[(2 * 32 + C) * C];
    const int tid = threadIdx.x, O = 2 * LE + C;
    for (int i = tid; i < O * C; i += 128) {
        const int o = i / C, c = i - o * C;
        w_s[i] = o < LE ? __ldg(wq + o * C + c) : o < 2 * LE ? __ldg(wk + (o - LE) * C + c) : __ldg(wv + (o - 2 * LE) * C + c);
    }
    pdl_wait();
    __syncthreads();
    const long long n = (long long)blockIdx.x * 128 + tid;
    if (n >= N) return;
    float acc[C];
#pragma unroll
    for (int c = 0; c < C; c += 4) { const float4 t = ld_plain4(gy + n * C + c); acc[c] = t.x; acc[c + 1] = t.y; acc[c + 2] = t.z; acc[c + 3] = t.w; }
    for (int o = 0; o < O; ++o) {
        const float g = o < LE ? dzq[n * LE + o] : o < 2 * LE ? dzk[n * LE + o - LE] : dzv[n * C + o - 2 * LE];
#pragma unroll
        for (int c = 0; c < C; ++c) acc[c] = fmaf(g, w_s[o * C + c], acc[c]);
    }
#pragma unroll
    for (int c = 0; c < C; c += 4) st4(gx + n * C + c, make_float4(acc[c], acc[c + 1], acc[c + 2], acc[c + 3]));
}

// ------------------------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------------------------
struct PathDims {
    long long N, R;
    int S, nd;
    RowMap map;
};
static PathDims path_dims(const sb_path_train_args& a) {
    PathDims d;
    d.nd = a.inter ? 1 : 2;
    d.S = a.inter ? a.T : a.F;
    d.R = a.inter ? (long long)a.B * a.F : (long long)a.B * a.T;
    d.N = d.R * d.S;
    d.map = RowMap{d.S, a.F, a.inter};
    return d;
}
struct SavedView {
    float *xhat, *xn, *rstd, *gates[2], *c[2], *h[2];
};
static SavedView saved_view(const sb_path_train_args& a, const PathDims& d) {
    SavedView v{};
    float* p = a.saved;
    v.xhat = p; p += d.N * a.C;
    v.xn = p; p += d.N * a.C;
    v.rstd = p; p += (d.N + 3) / 4 * 4;
    for (int k = 0; k < d.nd; ++k) {
        v.gates[k] = p; p += d.N * 4 * a.H;
        v.c[k] = p; p += d.N * a.H;
        v.h[k] = p; p += d.N * a.H;
    }
    return v;
}

static int check_path(const sb_path_train_args* p, int inter, const char* who) {
    SB_REQUIRE(p && p->x && p->ln_g && p->ln_b && p->lin_w && p->lin_b && p->saved, SB_E_BADARG, "%s: null pointer", who);
    SB_REQUIRE(p->inter == inter, SB_E_BADARG, "%s: args.inter must be %d", who, inter);
    SB_REQUIRE(p->B > 0 && p->T > 0 && p->F > 0, SB_E_BADARG, "%s: bad sizes", who);
    SB_REQUIRE(p->C == 16 || p->C == 32, SB_E_UNSUPP, "%s: C must be 16 or 32 (got %d)", who, p->C);
    SB_REQUIRE(p->H == kH, SB_E_UNSUPP, "%s: H must be 64 (got %d)", who, p->H);
    for (int d = 0; d < (inter ? 1 : 2); ++d)
        SB_REQUIRE(p->w_ih[d] && p->w_hh[d] && p->b_ih[d] && p->b_hh[d], SB_E_BADARG, "%s: null LSTM parameter", who);
    return 0;
}

template <int P>
static int run_rowgemm(const RowGemm& g, cudaStream_t st, const char* name) {
    using Cfg = RowGemmCfg<P>;
    if (g.K % Cfg::KB != 0) {
        set_error("%s: K = %d is not a multiple of %d", name, g.K, Cfg::KB);
        return SB_E_UNSUPP;
    }
    if (train_ffma2_enabled())
        return launch(name, rowgemm_kernel<P, true>, dim3((unsigned)ceil_div_ll(g.N, Cfg::ROWS)), dim3(256), Cfg::smem_bytes(g.nA, g.K), st, g);
    return launch(name, rowgemm_kernel<P, false>, dim3((unsigned)ceil_div_ll(g.N, Cfg::ROWS)), dim3(256), Cfg::smem_bytes(g.nA, g.K), st, g);
}
static int rowgemm(const RowGemm& g, int P, cudaStream_t st, const char* name) {
    switch (P) {
        case 16: return run_rowgemm<16>(g, st, name);
        case 32: return run_rowgemm<32>(g, st, name);
        case 64: return run_rowgemm<64>(g, st, name);
    }
    set_error("%s: unsupported width %d", name, P);
    return SB_E_UNSUPP;
}

static long long reduction_rows(long long N, int rb) {
    const long long ctas = 2LL * sm_count();
    long long rows = ceil_div_ll(N, ctas);
    rows = ceil_div_ll(rows, rb) * rb;
    return rows < rb ? rb : rows;
}
template <int J, int KC, int TJ, int TK>
static int run_outer(Outer o, cudaStream_t st, const char* name) {
    o.rows_per_cta = reduction_rows(o.N, 16);
    if (train_ffma2_enabled())
        return launch(name, outer_kernel<J, KC, TJ, TK, true>, dim3((unsigned)ceil_div_ll(o.N, o.rows_per_cta)), dim3(256), 0, st, o);
    return launch(name, outer_kernel<J, KC, TJ, TK, false>, dim3((unsigned)ceil_div_ll(o.N, o.rows_per_cta)), dim3(256), 0, st, o);
}

static int path_train_fwd(const sb_path_train_args& a, cudaStream_t st) {
    const PathDims d = path_dims(a);
    const SavedView v = saved_view(a, d);
    const unsigned gN = (unsigned)ceil_div_ll(d.N, a.C == 32 ? ln_rows_per_block<32>() : ln_rows_per_block<16>());
    if (a.C == 32) SB_CHECK(launch("ln_fwd", ln_fwd_kernel<32>, dim3(gN), dim3(256), 0, st, a.x, d.map, a.ln_g, a.ln_b, v.xhat, v.xn, v.rstd, d.N, 1e-5f));
    else SB_CHECK(launch("ln_fwd", ln_fwd_kernel<16>, dim3(gN), dim3(256), 0, st, a.x, d.map, a.ln_g, a.ln_b, v.xhat, v.xn, v.rstd, d.N, 1e-5f));
    LstmTrain l{};
    l.xn = v.xn;
    for (int k = 0; k < d.nd; ++k) {
        l.w_ih[k] = a.w_ih[k]; l.w_hh[k] = a.w_hh[k]; l.b_ih[k] = a.b_ih[k]; l.b_hh[k] = a.b_hh[k];
        l.gates[k] = v.gates[k]; l.c[k] = v.c[k]; l.h[k] = v.h[k];
    }
    l.R = (int)d.R; l.S = d.S;
    const dim3 grid((unsigned)ceil_div_ll(d.R, 4), d.nd);
    if (a.C == 32 && !train_one_row_enabled()) SB_CHECK(launch("lstm_train_fwd", lstm_train_fwd2_kernel<32>, grid, dim3(128), 0, st, l));
    else if (a.C == 32) SB_CHECK(launch("lstm_train_fwd", lstm_train_fwd_kernel<32>, grid, dim3(256), 0, st, l));
    else SB_CHECK(launch("lstm_train_fwd", lstm_train_fwd_kernel<16>, grid, dim3(256), 0, st, l));
    RowGemm g{};
    g.nA = d.nd; g.K = a.H; g.lda = a.H; g.ldw = d.nd * a.H; g.w_trans = 1;
    for (int k = 0; k < d.nd; ++k) { g.A[k] = v.h[k]; g.W[k] = a.lin_w + k * a.H; }
    g.bias = a.lin_b; g.res = a.x; g.out = a.y; g.map = d.map; g.a_mapped = 0; g.o_mapped = 1; g.N = d.N;
    return rowgemm(g, a.C, st, "path_linear");
}

static int path_bwd(const sb_path_bwd_args& b, cudaStream_t st) {
    const sb_path_train_args& a = b.f;
    const PathDims d = path_dims(a);
    const SavedView v = saved_view(a, d);
    float* dh[2] = {b.ws, b.ws + d.N * a.H};
    float* dxn = b.ws + (long long)d.nd * d.N * a.H;
    // (1) projection: dL/dh per direction, dL/dlin_w, dL/dlin_b
    for (int k = 0; k < d.nd; ++k) {
        RowGemm g{};
        g.nA = 1; g.K = a.C; g.lda = a.C; g.ldw = d.nd * a.H; g.w_trans = 0;
        g.A[0] = b.gy; g.W[0] = a.lin_w + k * a.H;
        g.out = dh[k]; g.map = d.map; g.a_mapped = 1; g.o_mapped = 0; g.N = d.N;
        SB_CHECK(rowgemm(g, a.H, st, "path_linear_bwd"));
        Outer o{};
        o.A = b.gy; o.Bm = v.h[k]; o.kc0 = a.H; o.dW = b.g_lin_w + k * a.H; o.ldw = d.nd * a.H;
        o.db = k == 0 ? b.g_lin_b : nullptr; o.db2 = nullptr;
        o.map = d.map; o.a_mapped = 1; o.b_mode = 0; o.reverse = 0; o.N = d.N;
        if (a.C == 32) SB_CHECK((run_outer<32, 64, 4, 2>(o, st, "path_linear_wgrad")));
        else SB_CHECK((run_outer<16, 64, 2, 2>(o, st, "path_linear_wgrad")));
    }
    // (2) BPTT: gates -> dz in place
    LstmTrain l{};
    for (int k = 0; k < d.nd; ++k) { l.w_hh[k] = a.w_hh[k]; l.gates[k] = v.gates[k]; l.c[k] = v.c[k]; l.dh[k] = dh[k]; }
    l.R = (int)d.R; l.S = d.S;
    if (train_one_row_enabled()) SB_CHECK(launch("lstm_train_bwd", lstm_train_bwd_kernel, dim3((unsigned)ceil_div_ll(d.R, 4), d.nd), dim3(256), 0, st, l));
    else SB_CHECK(launch("lstm_train_bwd", lstm_train_bwd2_kernel, dim3((unsigned)ceil_div_ll(d.R, 4), d.nd), dim3(128), 0, st, l));
    // (3) weight gradients in one pass over dz: dW_ih = dz^T LN(x), dW_hh = dz^T h_prev, db = sum dz
    for (int k = 0; k < d.nd; ++k) {
#ifndef SB_EMU
        if (a.C == 32 && train_tc_enabled()) {              // tcgen05 reduction GEMM (sb_train_tc.cu)
            WgradTc w{};
            w.dz = v.gates[k]; w.xn = v.xn; w.h = v.h[k];
            w.dW_ih = b.g_w_ih[k]; w.dW_hh = b.g_w_hh[k]; w.db = b.g_b_ih[k]; w.db2 = b.g_b_hh[k];
            w.S = d.S; w.reverse = k; w.N = d.N;
            w.scratch = b.ws;                               // the workspace is idle between BPTT (dh consumed) and lstm_dx (dxn written)
            w.scratch_floats = (long long)d.nd * d.N * a.H + d.N * a.C;
            SB_CHECK(run_wgrad_tc(w, st));
            continue;
        }
#endif
        Outer o{};
        o.A = v.gates[k]; o.map = d.map; o.a_mapped = 0; o.reverse = k; o.N = d.N;
        o.Bm = v.xn; o.b_mode = 0; o.kc0 = a.C; o.dW = b.g_w_ih[k]; o.ldw = a.C; o.db = b.g_b_ih[k]; o.db2 = b.g_b_hh[k];
        o.Bm2 = v.h[k]; o.dW2 = b.g_w_hh[k]; o.ldw2 = a.H;
        if (a.C == 32) SB_CHECK((run_outer<256, 96, 8, 12>(o, st, "lstm_wgrad")));
        else SB_CHECK((run_outer<256, 80, 8, 10>(o, st, "lstm_wgrad")));
    }
    // (4) dL/dLN(x) = sum_dir dz W_ih, then LayerNorm backward + the residual branch
    RowGemm g{};
    g.nA = d.nd; g.K = 4 * a.H; g.lda = 4 * a.H; g.ldw = a.C; g.w_trans = 0;
    for (int k = 0; k < d.nd; ++k) { g.A[k] = v.gates[k]; g.W[k] = a.w_ih[k]; }
    g.out = dxn; g.map = d.map; g.a_mapped = 0; g.o_mapped = 0; g.N = d.N;
    SB_CHECK(rowgemm(g, a.C, st, "lstm_dx"));
    const unsigned grid = (unsigned)(ceil_div_ll(d.N, 32) < 8LL * sm_count() ? ceil_div_ll(d.N, 32) : 8LL * sm_count());   // ln_bwd: eight resident CTAs per SM
    if (a.C == 32) return launch("ln_bwd", ln_bwd_kernel<32>, dim3(grid), dim3(256), 0, st, (const float*)dxn, (const float*)v.xhat, (const float*)v.rstd, a.ln_g, b.gy, b.gx, d.map, b.g_ln_g, b.g_ln_b, d.N);
    return launch("ln_bwd", ln_bwd_kernel<16>, dim3(grid), dim3(256), 0, st, (const float*)dxn, (const float*)v.xhat, (const float*)v.rstd, a.ln_g, b.gy, b.gx, d.map, b.g_ln_g, b.g_ln_b, d.N);
}

static int check_path_bwd(const sb_path_bwd_args* p, int inter, const char* who) {
    SB_REQUIRE(p, SB_E_BADARG, "%s: null pointer", who);
    SB_CHECK(check_path(&p->f, inter, who));
    SB_REQUIRE(p->gy && p->gx && p->g_ln_g && p->g_ln_b && p->g_lin_w && p->g_lin_b && p->ws, SB_E_BADARG, "%s: null pointer", who);
    for (int d = 0; d < (inter ? 1 : 2); ++d)
        SB_REQUIRE(p->g_w_ih[d] && p->g_w_hh[d] && p->g_b_ih[d] && p->g_b_hh[d], SB_E_BADARG, "%s: null gradient buffer", who);
    return 0;
}

// ---- conv-LSTM intra path ---------------------------------------------------------------------------------
struct ConvSaved {
    float *zraw, *pz, *xhat, *xn, *rstd, *gates[2], *c[2], *h[2];
};
static ConvSaved conv_saved_view(const sb_convpath_train_args& a, long long NP) {
    ConvSaved v{};
    float* p = a.saved;
    v.zraw = p; p += NP * a.C;
    v.pz = p; p += NP * a.C;
    v.xhat = p; p += NP * a.C;
    v.xn = p; p += NP * a.C;
    v.rstd = p; p += (NP + 3) / 4 * 4;
    for (int k = 0; k < 2; ++k) {
        v.gates[k] = p; p += NP * 4 * a.H;
        v.c[k] = p; p += NP * a.H;
        v.h[k] = p; p += NP * a.H;
    }
    return v;
}
static ConvPath conv_path_of(const sb_convpath_train_args& a) {
    ConvPath c{};
    c.x = a.x; c.conv_w = a.conv_w; c.conv_b = a.conv_b; c.deconv_w = a.deconv_w; c.deconv_b = a.deconv_b;
    c.BT = a.B * a.T; c.F = a.F; c.C = a.C; c.k = a.down; c.J = (a.F - a.down) / a.down + 1;
    c.outpad = a.tail_mode == SB_CONVLSTM_OUTPAD;
    return c;
}
static int check_convpath(const sb_convpath_train_args* p, const char* who) {
    SB_REQUIRE(p && p->x && p->conv_w && p->conv_b && p->prelu && p->ln_g && p->ln_b && p->deconv_w && p->deconv_b && p->saved,
               SB_E_BADARG, "%s: null pointer", who);
    SB_REQUIRE(p->B > 0 && p->T > 0 && p->F > 0, SB_E_BADARG, "%s: bad sizes", who);
    SB_REQUIRE(p->C == 16 || p->C == 32, SB_E_UNSUPP, "%s: C must be 16 or 32 (got %d)", who, p->C);
    SB_REQUIRE(p->H == kH, SB_E_UNSUPP, "%s: H must be 64 (got %d)", who, p->H);
    SB_REQUIRE(p->down >= 1 && p->down <= p->F && p->down * p->C <= 256, SB_E_UNSUPP, "%s: unsupported lstm_down %d", who, p->down);
    const int J = (p->F - p->down) / p->down + 1;
    SB_REQUIRE(p->tail_mode == SB_CONVLSTM_OUTPAD || p->F - J * p->down <= 3, SB_E_UNSUPP,
               "%s: pad-and-crop tail covers at most 3 bins (F=%d, down=%d)", who, p->F, p->down);
    for (int d = 0; d < 2; ++d)
        SB_REQUIRE(p->w_ih[d] && p->w_hh[d] && p->b_ih[d] && p->b_hh[d], SB_E_BADARG, "%s: null LSTM parameter", who);
    return 0;
}

template <int C>
static int convpath_fwd(const sb_convpath_train_args& a, cudaStream_t st) {
    const ConvPath cp = conv_path_of(a);
    const long long NP = (long long)cp.BT * cp.J, N = (long long)cp.BT * cp.F;
    const ConvSaved v = conv_saved_view(a, NP);
    constexpr int RPC = 256 / (C / 8);
    SB_CHECK(launch("convpre_train", convpre_train_kernel<C>, dim3((unsigned)ceil_div_ll(NP, RPC)), dim3(256),
                    (size_t)cp.k * C * C * sizeof(float), st, cp, v.zraw));
    const long long n4 = NP * C / 4;
    const unsigned ge = (unsigned)(ceil_div_ll(n4, 256) < 8LL * sm_count() ? ceil_div_ll(n4, 256) : 8LL * sm_count());
    SB_CHECK(launch("prelu_fwd", prelu_fwd_kernel, dim3(ge), dim3(256), 0, st, (const float*)v.zraw, a.prelu, v.pz, n4));
    const RowMap ident{cp.J, cp.F, 0};
    SB_CHECK(launch("ln_fwd", ln_fwd_kernel<C>, dim3((unsigned)ceil_div_ll(NP, ln_rows_per_block<C>())), dim3(256), 0, st, (const float*)v.pz, ident, a.ln_g,
                    a.ln_b, v.xhat, v.xn, v.rstd, NP, 1e-5f));
    LstmTrain l{};
    l.xn = v.xn;
    for (int k = 0; k < 2; ++k) {
        l.w_ih[k] = a.w_ih[k]; l.w_hh[k] = a.w_hh[k]; l.b_ih[k] = a.b_ih[k]; l.b_hh[k] = a.b_hh[k];
        l.gates[k] = v.gates[k]; l.c[k] = v.c[k]; l.h[k] = v.h[k];
    }
    l.R = cp.BT; l.S = cp.J;
    const dim3 grid((unsigned)ceil_div(cp.BT, 4), 2);
    if (C == 32 && !train_one_row_enabled()) SB_CHECK(launch("lstm_train_fwd", lstm_train_fwd2_kernel<32>, grid, dim3(128), 0, st, l));
    else SB_CHECK(launch("lstm_train_fwd", lstm_train_fwd_kernel<C>, grid, dim3(256), 0, st, l));
    return launch("convpost_train", convpost_train_kernel<C>, dim3((unsigned)ceil_div_ll(N, RPC)), dim3(256),
                  (size_t)cp.k * 2 * kH * C * sizeof(float), st, cp, (const float*)v.h[0], (const float*)v.h[1], a.y);
}

template <int C>
static int convpath_bwd(const sb_convpath_bwd_args& b, cudaStream_t st) {
    const sb_convpath_train_args& a = b.f;
    const ConvPath cp = conv_path_of(a);
    const long long NP = (long long)cp.BT * cp.J, N = (long long)cp.BT * cp.F;
    const ConvSaved v = conv_saved_view(a, NP);
    float* dh[2] = {b.ws, b.ws + NP * a.H};
    float* dxn = b.ws + 2 * NP * a.H;
    const size_t smem_d = (size_t)cp.k * 2 * kH * C * sizeof(float);
    SB_CHECK(launch("convpost_dh", convpost_dh_kernel<C>, dim3((unsigned)ceil_div_ll(NP, 16)), dim3(256), smem_d, st, cp, b.gy, dh[0], dh[1]));
    const long long gpc = reduction_rows(NP, 8);
    SB_CHECK(launch("convpost_wgrad", convpost_wgrad_kernel<C>, dim3((unsigned)ceil_div_ll(NP, gpc)), dim3(256), 0, st, cp, b.gy,
                    (const float*)v.h[0], (const float*)v.h[1], b.g_deconv_w, b.g_deconv_b, gpc));
    LstmTrain l{};
    for (int k = 0; k < 2; ++k) { l.w_hh[k] = a.w_hh[k]; l.gates[k] = v.gates[k]; l.c[k] = v.c[k]; l.dh[k] = dh[k]; }
    l.R = cp.BT; l.S = cp.J;
    const dim3 grid((unsigned)ceil_div(cp.BT, 4), 2);
    if (train_one_row_enabled()) SB_CHECK(launch("lstm_train_bwd", lstm_train_bwd_kernel, grid, dim3(256), 0, st, l));
    else SB_CHECK(launch("lstm_train_bwd", lstm_train_bwd2_kernel, grid, dim3(128), 0, st, l));
    const RowMap ident{cp.J, cp.F, 0};
    for (int k = 0; k < 2; ++k) {
        Outer o{};
        o.A = v.gates[k]; o.map = ident; o.a_mapped = 0; o.reverse = k; o.N = NP;
        o.Bm = v.xn; o.b_mode = 0; o.kc0 = C; o.dW = b.g_w_ih[k]; o.ldw = C; o.db = b.g_b_ih[k]; o.db2 = b.g_b_hh[k];
        o.Bm2 = v.h[k]; o.dW2 = b.g_w_hh[k]; o.ldw2 = a.H;
        if (C == 32) SB_CHECK((run_outer<256, 96, 8, 12>(o, st, "lstm_wgrad")));
        else SB_CHECK((run_outer<256, 80, 8, 10>(o, st, "lstm_wgrad")));
    }
    RowGemm g{};
    g.nA = 2; g.K = 4 * a.H; g.lda = 4 * a.H; g.ldw = C; g.w_trans = 0;
    for (int k = 0; k < 2; ++k) { g.A[k] = v.gates[k]; g.W[k] = a.w_ih[k]; }
    g.out = dxn; g.map = ident; g.a_mapped = 0; g.o_mapped = 0; g.N = NP;
    SB_CHECK(rowgemm(g, C, st, "lstm_dx"));
    // LayerNorm backward -> gradient of the PReLU output (written over the PReLU output slot), PReLU backward in place
    const unsigned gl = (unsigned)(ceil_div_ll(NP, 32) < 8LL * sm_count() ? ceil_div_ll(NP, 32) : 8LL * sm_count());
    SB_CHECK(launch("ln_bwd", ln_bwd_kernel<C>, dim3(gl), dim3(256), 0, st, (const float*)dxn, (const float*)v.xhat, (const float*)v.rstd, a.ln_g,
                    (const float*)nullptr, v.pz, ident, b.g_ln_g, b.g_ln_b, NP));
    const long long ne = NP * C;
    const unsigned ge = (unsigned)(ceil_div_ll(ne, 256) < 4LL * sm_count() ? ceil_div_ll(ne, 256) : 4LL * sm_count());
    SB_CHECK(launch("prelu_bwd", prelu_bwd_kernel, dim3(ge), dim3(256), 0, st, v.pz, (const float*)v.zraw, a.prelu, b.g_prelu, ne));
    const long long gpc2 = reduction_rows(NP, 32);
    SB_CHECK(launch("convpre_wgrad", convpre_wgrad_kernel<C>, dim3((unsigned)ceil_div_ll(NP, gpc2)), dim3(256), 0, st, cp, (const float*)v.pz,
                    b.g_conv_w, b.g_conv_b, gpc2));
    constexpr int RPC = 256 / (C / 8);
    return launch("convpre_bwd_x", convpre_bwd_x_kernel<C>, dim3((unsigned)ceil_div_ll(N, RPC)), dim3(256), (size_t)cp.k * C * C * sizeof(float),
                  st, cp, b.gy, (const float*)v.pz, b.gx);
}

// ---- attention (training) -----------------------------------------------------------------------------------
static long long al4(long long n) { return (n + 3) / 4 * 4; }
struct AttnSaved {
    float *zq, *zk, *zv, *qn, *kn, *vn, *att, *ob, *zo, *st_q, *st_k, *st_v, *st_o;
};
struct AttnSizes {
    long long N, R, BT;
    int LE, DQ, DV;
};
static AttnSizes attn_sizes(const sb_attn_train_args& a) {
    AttnSizes z;
    z.N = (long long)a.B * a.T * a.F; z.R = (long long)a.B * a.L * a.T; z.BT = (long long)a.B * a.T;
    z.LE = a.L * a.E; z.DQ = a.F * a.E; z.DV = a.F * (a.C / a.L);
    return z;
}
static AttnSaved attn_saved_view(const sb_attn_train_args& a, const AttnSizes& z) {
    AttnSaved v{};
    float* p = a.saved;
    v.zq = p; p += al4(z.N * z.LE);
    v.zk = p; p += al4(z.N * z.LE);
    v.zv = p; p += al4(z.N * a.C);
    v.qn = p; p += al4(z.N * z.LE);
    v.kn = p; p += al4(z.N * z.LE);
    v.vn = p; p += al4(z.N * a.C);
    v.att = p; p += al4(z.R * a.W);
    v.ob = p; p += al4(z.N * a.C);
    v.zo = p; p += al4(z.N * a.C);
    v.st_q = p; p += al4(2 * z.R);
    v.st_k = p; p += al4(2 * z.R);
    v.st_v = p; p += al4(2 * z.R);
    v.st_o = p; p += al4(2 * z.BT);
    return v;
}
static size_t attn_saved_floats(const sb_attn_train_args& a) {
    const AttnSizes z = attn_sizes(a);
    return (size_t)(4 * al4(z.N * z.LE) + 4 * al4(z.N * a.C) + al4(z.R * a.W) + 3 * al4(2 * z.R) + al4(2 * z.BT));
}
static int check_attn_train(const sb_attn_train_args* p, const char* who) {
    SB_REQUIRE(p && p->x && p->saved, SB_E_BADARG, "%s: null pointer", who);
    const sb_attn_proj* pr[4] = {&p->q, &p->k, &p->v, &p->o};
    for (int i = 0; i < 4; ++i)
        SB_REQUIRE(pr[i]->w && pr[i]->b && pr[i]->prelu && pr[i]->ln_g && pr[i]->ln_b, SB_E_BADARG, "%s: null projection parameter", who);
    SB_REQUIRE(p->B > 0 && p->T > 0 && p->F > 0 && p->L > 0 && p->E > 0 && p->W > 0, SB_E_BADARG, "%s: bad sizes", who);
    SB_REQUIRE(p->C == 16 || p->C == 32, SB_E_UNSUPP, "%s: C must be 16 or 32 (got %d)", who, p->C);
    SB_REQUIRE(p->C % p->L == 0, SB_E_BADARG, "%s: C must be a multiple of L", who);
    SB_REQUIRE(p->L * p->E == 8 || p->L * p->E == 16, SB_E_UNSUPP, "%s: L*E must be 8 or 16 (got %d)", who, p->L * p->E);
    SB_REQUIRE((size_t)(p->F * (p->C / p->L) + p->W) * sizeof(float) <= 200 * 1024, SB_E_UNSUPP, "%s: head-row too long", who);
    return 0;
}
static AttnDims attn_dims(const sb_attn_train_args& a) { return AttnDims{a.B, a.T, a.F, a.C, a.L, a.E, a.W, a.C / a.L}; }

static RowLn rowln_head(const sb_attn_train_args& a, const sb_attn_proj& pr, float* z, int ed, int zstride, float* out, float* stats) {
    RowLn r{};
    r.z = z; r.slope = pr.prelu; r.g = pr.ln_g; r.b = pr.ln_b; r.out = out; r.stats = stats;
    r.D = a.F * ed; r.head = 1; r.ed = ed; r.zstride = zstride; r.L = a.L; r.T = a.T; r.F = a.F;
    return r;
}

template <int C>
static int attn_train_fwd(const sb_attn_train_args& a, cudaStream_t st) {
    const AttnSizes z = attn_sizes(a);
    const AttnSaved v = attn_saved_view(a, z);
    const AttnDims d = attn_dims(a);
    SB_CHECK(launch("attn_proj_train", attn_proj_train_kernel<C>, dim3((unsigned)ceil_div_ll(z.N, 128)), dim3(128), 0, st, a.x, a.q.w, a.q.b,
                    a.k.w, a.k.b, a.v.w, a.v.b, v.zq, v.zk, v.zv, z.LE, z.N));
    SB_CHECK(launch("attn_head_ln", prelu_ln_rows_fwd_kernel, dim3((unsigned)z.R), dim3(256), 0, st, rowln_head(a, a.q, v.zq, a.E, z.LE, v.qn, v.st_q)));
    SB_CHECK(launch("attn_head_ln", prelu_ln_rows_fwd_kernel, dim3((unsigned)z.R), dim3(256), 0, st, rowln_head(a, a.k, v.zk, a.E, z.LE, v.kn, v.st_k)));
    SB_CHECK(launch("attn_head_ln", prelu_ln_rows_fwd_kernel, dim3((unsigned)z.R), dim3(256), 0, st, rowln_head(a, a.v, v.zv, d.Vd, a.C, v.vn, v.st_v)));
    SB_CHECK(launch("attn_core_train_fwd", attn_core_train_fwd_kernel, dim3((unsigned)z.R), dim3(128), (size_t)(z.DQ + a.W) * sizeof(float), st, d,
                    (const float*)v.qn, (const float*)v.kn, (const float*)v.vn, v.att, v.ob));
    RowGemm g{};
    g.nA = 1; g.K = C; g.lda = C; g.ldw = C; g.w_trans = 1;
    g.A[0] = v.ob; g.W[0] = a.o.w; g.bias = a.o.b; g.out = v.zo; g.map = RowMap{1, a.F, 0}; g.N = z.N;
    SB_CHECK(rowgemm(g, C, st, "attn_out_proj"));
    RowLn r{};
    r.z = v.zo; r.slope = a.o.prelu; r.g = a.o.ln_g; r.b = a.o.ln_b; r.res = a.x; r.out = a.y; r.stats = v.st_o; r.D = a.F * C; r.head = 0;
    return launch("attn_out_ln", prelu_ln_rows_fwd_kernel, dim3((unsigned)z.BT), dim3(256), 0, st, r);
}

template <int J, int KC>
static int attn_wgrad(const float* A, const float* Bm, float* dW, float* db, long long N, cudaStream_t st) {
    Outer o{};
    o.A = A; o.Bm = Bm; o.kc0 = KC; o.dW = dW; o.ldw = KC; o.db = db; o.map = RowMap{1, 1, 0}; o.N = N;
    return run_outer<J, KC, (J >= 32 ? 4 : 2), 2>(o, st, "attn_wgrad");
}

template <int C>
static int attn_train_bwd(const sb_attn_bwd_args& b, cudaStream_t st) {
    const sb_attn_train_args& a = b.f;
    const AttnSizes z = attn_sizes(a);
    const AttnSaved v = attn_saved_view(a, z);
    const AttnDims d = attn_dims(a);
    float* p = b.ws;
    float* dob = p; p += al4(z.N * C);
    float* dqn = p; p += al4(z.N * z.LE);
    float* dkn = p; p += al4(z.N * z.LE);
    float* dvn = p; p += al4(z.N * C);
    float* dl = p;
    // output LayerNorm + PReLU (dz over zo in place), output projection
    RowLn r{};
    r.z = v.zo; r.slope = a.o.prelu; r.g = a.o.ln_g; r.stats = v.st_o; r.D = a.F * C; r.head = 0;
    r.gout = b.gy; r.dz = v.zo; r.g_g = b.go.ln_g; r.g_b = b.go.ln_b; r.g_slope = b.go.prelu;
    SB_CHECK(launch("attn_out_ln_bwd", prelu_ln_rows_bwd_kernel, dim3((unsigned)z.BT), dim3(256), 0, st, r));
    RowGemm g{};
    g.nA = 1; g.K = C; g.lda = C; g.ldw = C; g.w_trans = 0;
    g.A[0] = v.zo; g.W[0] = a.o.w; g.out = dob; g.map = RowMap{1, a.F, 0}; g.N = z.N;
    SB_CHECK(rowgemm(g, C, st, "attn_out_proj_bwd"));
    SB_CHECK((attn_wgrad<C, C>(v.zo, v.ob, b.go.w, b.go.b, z.N, st)));
    // attention core
    SB_CHECK(launch("attn_core_bwd_q", attn_core_bwd_q_kernel, dim3((unsigned)z.R), dim3(128), (size_t)(z.DV + a.W) * sizeof(float), st, d,
                    (const float*)v.kn, (const float*)v.vn, (const float*)v.att, (const float*)dob, dl, dqn));
    SB_CHECK(launch("attn_core_bwd_kv", attn_core_bwd_kv_kernel, dim3((unsigned)z.R), dim3(128), 0, st, d, (const float*)v.qn, (const float*)v.att,
                    (const float*)dl, (const float*)dob, dkn, dvn));
    // head LayerNorms + PReLUs (dz over zq / zk / zv in place)
    const sb_attn_proj* pr[3] = {&a.q, &a.k, &a.v};
    const sb_attn_proj_grad* gr[3] = {&b.gq, &b.gk, &b.gv};
    float* zs[3] = {v.zq, v.zk, v.zv};
    float* gs[3] = {dqn, dkn, dvn};
    float* sts[3] = {v.st_q, v.st_k, v.st_v};
    for (int i = 0; i < 3; ++i) {
        RowLn h = rowln_head(a, *pr[i], zs[i], i < 2 ? a.E : d.Vd, i < 2 ? z.LE : C, nullptr, sts[i]);
        h.gout = gs[i]; h.dz = zs[i]; h.g_g = gr[i]->ln_g; h.g_b = gr[i]->ln_b; h.g_slope = gr[i]->prelu;
        SB_CHECK(launch("attn_head_ln_bwd", prelu_ln_rows_bwd_kernel, dim3((unsigned)z.R), dim3(256), 0, st, h));
    }
    // projections
    if (z.LE == 8) {
        SB_CHECK((attn_wgrad<8, C>(v.zq, a.x, b.gq.w, b.gq.b, z.N, st)));
        SB_CHECK((attn_wgrad<8, C>(v.zk, a.x, b.gk.w, b.gk.b, z.N, st)));
    } else {
        SB_CHECK((attn_wgrad<16, C>(v.zq, a.x, b.gq.w, b.gq.b, z.N, st)));
        SB_CHECK((attn_wgrad<16, C>(v.zk, a.x, b.gk.w, b.gk.b, z.N, st)));
    }
    SB_CHECK((attn_wgrad<C, C>(v.zv, a.x, b.gv.w, b.gv.b, z.N, st)));
    return launch("attn_proj_bwd_x", attn_proj_bwd_x_kernel<C>, dim3((unsigned)ceil_div_ll(z.N, 128)), dim3(128), 0, st, b.gy, (const float*)v.zq,
                  (const float*)v.zk, (const float*)v.zv, a.q.w, a.k.w, a.v.w, b.gx, z.LE, z.N);
}

}  // namespace sb

using namespace sb;

extern "C" size_t sb_path_train_saved_floats(int B, int T, int F, int C, int H, int inter) {
    const long long N = (long long)B * T * F;
    const int nd = inter ? 1 : 2;
    return (size_t)(2 * N * C + (N + 3) / 4 * 4 + nd * N * 6 * H);
}
extern "C" size_t sb_path_bwd_workspace_floats(int B, int T, int F, int C, int H, int inter) {
    const long long N = (long long)B * T * F;
    return (size_t)((inter ? 1 : 2) * N * H + N * C);
}
extern "C" int sb_intra_lstm_train_fwd(const sb_path_train_args* p, void* stream) {
    SB_CHECK(check_path(p, 0, "sb_intra_lstm_train_fwd"));
    SB_REQUIRE(p->y && p->y != p->x, SB_E_BADARG, "sb_intra_lstm_train_fwd: y must be a separate buffer");
    return path_train_fwd(*p, (cudaStream_t)stream);
}
extern "C" int sb_inter_lstm_train_fwd(const sb_path_train_args* p, void* stream) {
    SB_CHECK(check_path(p, 1, "sb_inter_lstm_train_fwd"));
    SB_REQUIRE(p->y && p->y != p->x, SB_E_BADARG, "sb_inter_lstm_train_fwd: y must be a separate buffer");
    return path_train_fwd(*p, (cudaStream_t)stream);
}
extern "C" int sb_intra_lstm_bwd(const sb_path_bwd_args* p, void* stream) {
    SB_CHECK(check_path_bwd(p, 0, "sb_intra_lstm_bwd"));
    return path_bwd(*p, (cudaStream_t)stream);
}
extern "C" int sb_inter_lstm_bwd(const sb_path_bwd_args* p, void* stream) {
    SB_CHECK(check_path_bwd(p, 1, "sb_inter_lstm_bwd"));
    return path_bwd(*p, (cudaStream_t)stream);
}

static int check_film_apply(const sb_film_apply_args* p, const char* who) {
    SB_REQUIRE(p && p->x && p->film_scale && p->film_shift, SB_E_BADARG, "%s: null pointer", who);
    SB_REQUIRE(p->B > 0 && p->T > 0 && p->F > 0 && p->C > 0 && p->C % 4 == 0, SB_E_BADARG, "%s: bad sizes", who);
    return 0;
}
extern "C" int sb_film_apply_fwd(const sb_film_apply_args* p, void* stream) {
    SB_CHECK(check_film_apply(p, "sb_film_apply_fwd"));
    SB_REQUIRE(p->y, SB_E_BADARG, "sb_film_apply_fwd: null output");
    const long long n4 = (long long)p->B * p->T * p->F * p->C / 4;
    const long long blocks = ceil_div_ll(n4, 256);
    const unsigned grid = (unsigned)(blocks < 8LL * sm_count() ? blocks : 8LL * sm_count());
    return launch("film_apply_fwd", film_apply_fwd_kernel, dim3(grid), dim3(256), 0, (cudaStream_t)stream, *p);
}
extern "C" int sb_film_apply_bwd(const sb_film_apply_args* p, void* stream) {
    SB_CHECK(check_film_apply(p, "sb_film_apply_bwd"));
    SB_REQUIRE(p->gy && p->gx && p->g_scale && p->g_shift, SB_E_BADARG, "sb_film_apply_bwd: null pointer");
    const long long total = (long long)p->B * p->F * p->C;
    const int slices = p->T >= 64 ? 8 : 1;                      // whole clips: eight frame slices per (b, f, c)
    return launch("film_apply_bwd", film_apply_bwd_kernel, dim3((unsigned)ceil_div_ll(total, 256), slices), dim3(256), 0, (cudaStream_t)stream, *p);
}

extern "C" int sb_film_params_bwd(const sb_film_bwd_args* p, void* stream) {
    SB_REQUIRE(p && p->g_film && p->f.dis && p->f.emb_w && p->f.emb_ln_g && p->f.emb_ln_b && p->f.w_w && p->f.b_w, SB_E_BADARG,
               "sb_film_params_bwd: null pointer");
    SB_REQUIRE(p->g_emb_w && p->g_emb_ln_g && p->g_emb_ln_b && p->g_w_w && p->g_w_b && p->g_b_w && p->g_b_b, SB_E_BADARG,
               "sb_film_params_bwd: null gradient buffer");
    SB_REQUIRE(p->f.emb_mode == SB_EMB_CONV || p->f.emb_mode == SB_EMB_LINEAR, SB_E_BADARG, "sb_film_params_bwd: unknown emb_mode %d", p->f.emb_mode);
    SB_REQUIRE(p->f.B > 0 && p->f.F > 0 && p->f.C > 0 && p->f.Din > 0 && p->f.n_layers > 0, SB_E_BADARG, "sb_film_params_bwd: bad sizes");
    const size_t smem = ((size_t)3 * p->f.F * p->f.Din + p->f.F) * sizeof(float);
    return launch("film_params_bwd", film_params_bwd_kernel, dim3(p->f.B), dim3(256), smem, (cudaStream_t)stream, *p);
}

static int check_conv_in_train(const sb_conv_in_train_args* p, const char* who) {
    SB_REQUIRE(p && p->feats && p->w && p->bias, SB_E_BADARG, "%s: null pointer", who);
    SB_REQUIRE((p->ln_g == nullptr) == (p->ln_b == nullptr), SB_E_BADARG, "%s: LayerNorm gain and bias must come together", who);
    SB_REQUIRE(!p->ln_g || p->saved, SB_E_BADARG, "%s: saved buffer missing", who);
    SB_REQUIRE(p->B > 0 && p->T > 0 && p->F > 0 && p->Cin > 0, SB_E_BADARG, "%s: bad sizes", who);
    SB_REQUIRE(p->C == 16 || p->C == 32, SB_E_UNSUPP, "%s: C must be 16 or 32 (got %d)", who, p->C);
    return 0;
}
extern "C" int sb_conv_in_train_fwd(const sb_conv_in_train_args* p, void* stream) {
    SB_CHECK(check_conv_in_train(p, "sb_conv_in_train_fwd"));
    SB_REQUIRE(p->x, SB_E_BADARG, "sb_conv_in_train_fwd: null output");
    cudaStream_t st = (cudaStream_t)stream;
    const long long N = (long long)p->B * p->T * p->F;
    const int C = p->C;
    // with the LayerNorm the raw conv output goes to the head of `saved` and ln_fwd writes x from it
    float* raw = p->ln_g ? p->saved : p->x;
    const size_t smem = (size_t)9 * p->Cin * C * sizeof(float);
    const unsigned grid = (unsigned)ceil_div_ll(N, 256 / (C / 8));
    if (front_tc_enabled() && conv_in_tc_train_supported(p->B, p->T, p->F, p->Cin, C))      // tcgen05 implicit GEMM (sb_frontend_tc.cu)
        SB_CHECK(run_conv_in_tc_train(p->feats, p->w, p->bias, raw, p->B, p->T, p->F, p->Cin, st));
    else if (C == 32) SB_CHECK(launch("conv_in_train", conv_in_train_kernel<32>, dim3(grid), dim3(256), smem, st, *p, raw));
    else SB_CHECK(launch("conv_in_train", conv_in_train_kernel<16>, dim3(grid), dim3(256), smem, st, *p, raw));
    if (!p->ln_g) return 0;
    float* xhat = p->saved + N * C;
    float* rstd = p->saved + 2 * N * C;
    const RowMap map{1, p->F, 0};
    const unsigned gN = (unsigned)ceil_div_ll(N, C == 32 ? ln_rows_per_block<32>() : ln_rows_per_block<16>());
    if (C == 32) return launch("ln_fwd", ln_fwd_kernel<32>, dim3(gN), dim3(256), 0, st, (const float*)raw, map, p->ln_g, p->ln_b, xhat, p->x, rstd, N, 1e-5f);
    return launch("ln_fwd", ln_fwd_kernel<16>, dim3(gN), dim3(256), 0, st, (const float*)raw, map, p->ln_g, p->ln_b, xhat, p->x, rstd, N, 1e-5f);
}
extern "C" int sb_conv_in_bwd(const sb_conv_in_train_args* p, void* stream) {
    SB_CHECK(check_conv_in_train(p, "sb_conv_in_bwd"));
    SB_REQUIRE(p->gx && p->g_w && p->g_bias && p->ws, SB_E_BADARG, "sb_conv_in_bwd: null pointer");
    SB_REQUIRE(!p->ln_g || (p->g_ln_g && p->g_ln_b), SB_E_BADARG, "sb_conv_in_bwd: null LayerNorm gradient buffer");
    SB_REQUIRE(9 * p->Cin <= 512, SB_E_UNSUPP, "sb_conv_in_bwd: too many input channels (%d)", p->Cin);
    cudaStream_t st = (cudaStream_t)stream;
    const long long N = (long long)p->B * p->T * p->F;
    const int C = p->C;
    const float* g = p->gx;
    if (p->ln_g) {
        const RowMap map{1, p->F, 0};
        const unsigned grid = (unsigned)(ceil_div_ll(N, 32) < 8LL * sm_count() ? ceil_div_ll(N, 32) : 8LL * sm_count());
        const float* xhat = p->saved + N * C;
        const float* rstd = p->saved + 2 * N * C;
        if (C == 32) SB_CHECK(launch("ln_bwd", ln_bwd_kernel<32>, dim3(grid), dim3(256), 0, st, p->gx, xhat, rstd, p->ln_g, (const float*)nullptr, p->ws, map, p->g_ln_g, p->g_ln_b, N));
        else SB_CHECK(launch("ln_bwd", ln_bwd_kernel<16>, dim3(grid), dim3(256), 0, st, p->gx, xhat, rstd, p->ln_g, (const float*)nullptr, p->ws, map, p->g_ln_g, p->g_ln_b, N));
        g = p->ws;
    }
    const long long rows = reduction_rows(N, 32);
    const unsigned grid = (unsigned)ceil_div_ll(N, rows);
    if (C == 32) return launch("conv_in_wgrad", conv_in_wgrad_kernel<32>, dim3(grid), dim3(256), 0, st, *p, g, rows);
    return launch("conv_in_wgrad", conv_in_wgrad_kernel<16>, dim3(grid), dim3(256), 0, st, *p, g, rows);
}

extern "C" int sb_backend_bwd(const sb_backend_bwd_args* p, void* stream) {
    SB_REQUIRE(p && p->x && p->g_wave && p->w && p->filt && p->gx && p->g_w && p->g_bias && p->ws, SB_E_BADARG, "sb_backend_bwd: null pointer");
    SB_REQUIRE(p->B > 0 && p->T > 0 && p->F > 0 && p->n_fft > 0 && p->stride > 0 && p->n_fft >= p->stride, SB_E_BADARG, "sb_backend_bwd: bad sizes");
    SB_REQUIRE(p->C == 16 || p->C == 32, SB_E_UNSUPP, "sb_backend_bwd: C must be 16 or 32 (got %d)", p->C);
    SB_REQUIRE(p->n_src == 1 || p->n_src == 2, SB_E_UNSUPP, "sb_backend_bwd: 1 or 2 sources (got %d)", p->n_src);
    cudaStream_t st = (cudaStream_t)stream;
    SB_REQUIRE(2 * p->F <= 640, SB_E_UNSUPP, "sb_backend_bwd: more than 640 basis rows (F = %d)", p->F);
    const size_t smem = ((size_t)7 * p->stride + p->n_fft + (size_t)2 * p->F * 33) * sizeof(float);
    SB_CHECK(launch("istft_bwd", istft_bwd_kernel, dim3(ceil_div(p->T, 8), p->B * p->n_src), dim3(320), smem, st, *p));
    const long long N = (long long)p->B * p->T * p->F;
    const long long rows = ceil_div_ll(N, 6LL * sm_count()) < 1 ? 1 : ceil_div_ll(N, 6LL * sm_count());      // deconv_wgrad: three resident CTAs per SM, two rounds
    if (p->C == 32) {
        SB_CHECK(launch("deconv_bwd_x", deconv_bwd_x_kernel<32>, dim3((unsigned)ceil_div_ll(N, 8)), dim3(256), 0, st, *p));
        return launch("deconv_wgrad", deconv_wgrad_kernel<32>, dim3((unsigned)ceil_div_ll(N, rows)), dim3(288), 0, st, *p, rows);
    }
    SB_CHECK(launch("deconv_bwd_x", deconv_bwd_x_kernel<16>, dim3((unsigned)ceil_div_ll(N, 16)), dim3(256), 0, st, *p));
    return launch("deconv_wgrad", deconv_wgrad_kernel<16>, dim3((unsigned)ceil_div_ll(N, rows)), dim3(160), 0, st, *p, rows);
}

extern "C" size_t sb_convpath_train_saved_floats(int B, int T, int F, int C, int H, int down) {
    const long long NP = (long long)B * T * ((F - down) / down + 1);
    return (size_t)(4 * NP * C + (NP + 3) / 4 * 4 + 2 * NP * 6 * H);
}
extern "C" size_t sb_convpath_bwd_workspace_floats(int B, int T, int F, int C, int H, int down) {
    const long long NP = (long long)B * T * ((F - down) / down + 1);
    return (size_t)(2 * NP * H + NP * C);
}
extern "C" int sb_intra_convlstm_train_fwd(const sb_convpath_train_args* p, void* stream) {
    SB_CHECK(check_convpath(p, "sb_intra_convlstm_train_fwd"));
    SB_REQUIRE(p->y && p->y != p->x, SB_E_BADARG, "sb_intra_convlstm_train_fwd: y must be a separate buffer");
    return p->C == 32 ? convpath_fwd<32>(*p, (cudaStream_t)stream) : convpath_fwd<16>(*p, (cudaStream_t)stream);
}
extern "C" int sb_intra_convlstm_bwd(const sb_convpath_bwd_args* p, void* stream) {
    SB_REQUIRE(p, SB_E_BADARG, "sb_intra_convlstm_bwd: null pointer");
    SB_CHECK(check_convpath(&p->f, "sb_intra_convlstm_bwd"));
    SB_REQUIRE(p->gy && p->gx && p->ws && p->g_conv_w && p->g_conv_b && p->g_prelu && p->g_ln_g && p->g_ln_b && p->g_deconv_w && p->g_deconv_b,
               SB_E_BADARG, "sb_intra_convlstm_bwd: null pointer");
    for (int d = 0; d < 2; ++d)
        SB_REQUIRE(p->g_w_ih[d] && p->g_w_hh[d] && p->g_b_ih[d] && p->g_b_hh[d], SB_E_BADARG, "sb_intra_convlstm_bwd: null gradient buffer");
    return p->f.C == 32 ? convpath_bwd<32>(*p, (cudaStream_t)stream) : convpath_bwd<16>(*p, (cudaStream_t)stream);
}

extern "C" size_t sb_attn_train_saved_floats(const sb_attn_train_args* p) { return p ? attn_saved_floats(*p) : 0; }
extern "C" size_t sb_attn_bwd_workspace_floats(const sb_attn_train_args* p) {
    if (!p) return 0;
    const AttnSizes z = attn_sizes(*p);
    return (size_t)(2 * al4(z.N * p->C) + 2 * al4(z.N * z.LE) + al4(z.R * p->W));
}
extern "C" int sb_attn_train_fwd(const sb_attn_train_args* p, void* stream) {
    SB_CHECK(check_attn_train(p, "sb_attn_train_fwd"));
    SB_REQUIRE(p->y && p->y != p->x, SB_E_BADARG, "sb_attn_train_fwd: y must be a separate buffer");
    return p->C == 32 ? attn_train_fwd<32>(*p, (cudaStream_t)stream) : attn_train_fwd<16>(*p, (cudaStream_t)stream);
}
extern "C" int sb_attn_bwd(const sb_attn_bwd_args* p, void* stream) {
    SB_REQUIRE(p, SB_E_BADARG, "sb_attn_bwd: null pointer");
    SB_CHECK(check_attn_train(&p->f, "sb_attn_bwd"));
    SB_REQUIRE(p->gy && p->gx && p->ws, SB_E_BADARG, "sb_attn_bwd: null pointer");
    const sb_attn_proj_grad* gr[4] = {&p->gq, &p->gk, &p->gv, &p->go};
    for (int i = 0; i < 4; ++i)
        SB_REQUIRE(gr[i]->w && gr[i]->b && gr[i]->prelu && gr[i]->ln_g && gr[i]->ln_b, SB_E_BADARG, "sb_attn_bwd: null gradient buffer");
    return p->f.C == 32 ? attn_train_bwd<32>(*p, (cudaStream_t)stream) : attn_train_bwd<16>(*p, (cudaStream_t)stream);
}
