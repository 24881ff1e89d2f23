// Intra-frame (across frequency, bidirectional) and inter-frame (across time, carried state) LSTM paths.
//
// Replaces, per GridNetBlock (reference DE3 = src/models/tfgridnet_realtime_clean_dis_embd3/tfgridnet_causal.py):
//   intra:  FilmLayer (:51-68, applied :509-513) -> intra_norm -> intra_rnn (BiLSTM) -> intra_linear -> +x  (:794-827)
//   inter:  inter_norm -> inter_rnn (h0, c0 carried) -> inter_linear -> +x                                  (:829-849)
// Both are the same fused "sequence" kernel: [FiLM] -> LayerNorm(C) -> x W_ih^T + h W_hh^T + b -> gates -> Linear ->
// residual, one launch per (block, path); every activation is read once and written once.
//
// The kernel families share one argument block; pick_algo() chooses per call:
//   lstm_tile_kernel : throughput path.  [W_ih|W_hh]^T (K x 256 fp32) lives in shared memory; every WARP owns 8 (or 4)
//                      sequences end to end (A operand, gates, cell state, projection), so the step loop has no block
//                      barrier and the warps of an SM drift apart and overlap FMA / MUFU / LDS phases.  Each lane
//                      holds an 8-row x 8-column accumulator tile = all four gates of two hidden units (no shuffles
//                      for the cell update); the output projection of step s-1 rides in the h-part of step s's GEMM.
//   lstm_ws_kernel   : latency path for streaming chunks (B sequences x 145 serial steps).  One sequence per CTA,
//                      warp-specialised: recurrence warps with the recurrent weights in registers, helper warps for
//                      loads / LayerNorm / input gates / stores.
//   lstm_lane_kernel : the earlier latency path (1/2/4 sequences per CTA, one gate column per thread); selectable.
//   lstm_tc_kernel   : sb_lstm_tc.cu, tcgen05 gate GEMM for large batches of sequences.
#include <cmath>
#include <type_traits>

#include "sb_common.cuh"
#include "sb_lstm.cuh"

namespace sb {


// =============================================================================================================
// tile kernel
// =============================================================================================================
template <int C, int RW_>
struct TileCfg {
    static constexpr int H = 64, K = C + H, RW = RW_;
    static constexpr int AS = K + ((16 - K % 32 + 32) % 32);    // A row stride: == 16 (mod 32) -> conflict-free stores
    static constexpr int RS = (C == 32) ? 48 : 16;              // residual row stride
    static constexpr int warp_floats = RW * AS + 2 * RW * RS;
    static constexpr size_t smem_floats(int nwarps) { return (size_t)K * 256 + H * C + (size_t)nwarps * warp_floats; }
};

template <int C, bool RAW_H, int RW_>
__global__ void __launch_bounds__(256, 1) lstm_tile_kernel(const SeqArgs a) {
    using Cfg = TileCfg<C, RW_>;
    constexpr int H = Cfg::H, K = Cfg::K, RW = Cfg::RW, AS = Cfg::AS, RS = Cfg::RS;
    constexpr int LPR = 32 / RW;            // lanes per row in the load / LayerNorm mapping
    constexpr int NV = C / (4 * LPR);       // float4 per lane there: channels (4*LPR)*v + 4*q .. +3
    static_assert(NV >= 1, "RW = 4 needs C = 32");
    constexpr int ORW = RW * C / 32;        // rows per lane in the projection / output mapping (lane -> channel)
    static_assert(AS % 32 == 16 && AS % 4 == 0, "A stride");
    SB_DYN_SMEM(float, smem);
    float* Wt = smem;                       // [K][256]   column p(g,u) = (g/2)*128 + 4*(u/2) + 2*(g%2) + u%2
    float* WlT = Wt + K * 256;              // [H][C]     projection, transposed

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int dir = blockIdx.y;
    const sb_lstm_dir& w = a.w[dir];
    const int S = a.n_steps;

    // stage the weights once per CTA with TMA bulk copies (constant data: allowed before pdl_wait)
    __shared__ BulkBarrier wbar;
    if (tid == 0) bulk_barrier_init(&wbar);
    __syncthreads();
    if (tid == 0) {
        bulk_expect(&wbar, (unsigned)((K * 256 + (RAW_H ? 0 : H * C)) * sizeof(float)));
        bulk_copy_g2s(Wt, w.w_tile, K * 256 * sizeof(float), &wbar);
        if (!RAW_H) bulk_copy_g2s(WlT, w.lin_t, H * C * sizeof(float), &wbar);
    }
    float bias[8];
    {
        const float4 b0 = __ldg(reinterpret_cast<const float4*>(w.b_tile) + lane);
        const float4 b1 = __ldg(reinterpret_cast<const float4*>(w.b_tile + 128) + lane);
        bias[0] = b0.x; bias[1] = b0.y; bias[2] = b0.z; bias[3] = b0.w;
        bias[4] = b1.x; bias[5] = b1.y; bias[6] = b1.z; bias[7] = b1.w;
    }
    pdl_trigger();
#ifdef SB_EMU
    __syncthreads();                        // the host-emulated copy above is an ordinary store by thread 0
#endif
    bulk_wait(&wbar, 0);                    // the weights have landed; no block-wide barrier in the step loop either
    pdl_wait();

    const int row0 = (blockIdx.x * (blockDim.x >> 5) + warp) * RW;
    if (row0 >= a.n_rows) return;           // whole warp leaves; no barrier follows

    float* A = WlT + H * C + warp * Cfg::warp_floats;      // [RW][AS]: cols 0..C-1 = LN(x_s), C..K-1 = h_{s-1}
    float* res = A + RW * AS;                               // [2][RW][RS] x' kept for the residual

    // ---- load / LayerNorm mapping: lane -> (row lr, channel group q): channels 4*LPR*v + 4*q .. +3 ---------------
    const int lr = lane / LPR, q = lane % LPR;
    const int lrow = min(row0 + lr, a.n_rows - 1);
    const long long lbase = row_base(a, lrow) + 4 * q;
    const long long fbase = (long long)(lrow / a.film_row_div) * S * C + 4 * q;
    float4 g4[NV], b4[NV];
#pragma unroll
    for (int v = 0; v < NV; ++v) {
        g4[v] = __ldg(reinterpret_cast<const float4*>(w.ln_g + 4 * LPR * v) + q);
        b4[v] = __ldg(reinterpret_cast<const float4*>(w.ln_b + 4 * LPR * v) + q);
    }
    auto load_x = [&](int step, float4 (&xv)[NV]) {
        const int pos = dir ? S - 1 - step : step;
        const long long off = lbase + (long long)pos * a.stride_pos;
#pragma unroll
        for (int v = 0; v < NV; ++v) {
            float4 t = ldg4_stream(a.x0 + off + 4 * LPR * v);
            if (a.x1) {
                const float4 u = ldg4_stream(a.x1 + off + 4 * LPR * v);
                t.x += u.x; t.y += u.y; t.z += u.z; t.w += u.w;
            }
            if (a.film_scale) {
                const float4 fs = __ldg(reinterpret_cast<const float4*>(a.film_scale + fbase + (long long)pos * C + 4 * LPR * v));
                const float4 fb = __ldg(reinterpret_cast<const float4*>(a.film_shift + fbase + (long long)pos * C + 4 * LPR * v));
                t.x = fmaf(t.x, fs.x, fb.x); t.y = fmaf(t.y, fs.y, fb.y);
                t.z = fmaf(t.z, fs.z, fb.z); t.w = fmaf(t.w, fs.w, fb.w);
            }
            xv[v] = t;
        }
    };
    auto ln_store = [&](const float4 (&xv)[NV], int slot) {
        float s1 = 0.f;
#pragma unroll
        for (int v = 0; v < NV; ++v) s1 += (xv[v].x + xv[v].y) + (xv[v].z + xv[v].w);
        const float mean = group_sum<LPR>(s1) * (1.0f / C);
        float s2 = 0.f;
#pragma unroll
        for (int v = 0; v < NV; ++v) {
            const float dx = xv[v].x - mean, dy = xv[v].y - mean, dz = xv[v].z - mean, dw = xv[v].w - mean;
            s2 += (dx * dx + dy * dy) + (dz * dz + dw * dw);
        }
        const float rstd = rsqrtf(group_sum<LPR>(s2) * (1.0f / C) + kLnEps);
#pragma unroll
        for (int v = 0; v < NV; ++v) {
            float4 n;
            n.x = fmaf((xv[v].x - mean) * rstd, g4[v].x, b4[v].x);
            n.y = fmaf((xv[v].y - mean) * rstd, g4[v].y, b4[v].y);
            n.z = fmaf((xv[v].z - mean) * rstd, g4[v].z, b4[v].z);
            n.w = fmaf((xv[v].w - mean) * rstd, g4[v].w, b4[v].w);
            st4(A + lr * AS + 4 * LPR * v + 4 * q, n);
            if (!RAW_H) st4(res + (slot * RW + lr) * RS + 4 * LPR * v + 4 * q, xv[v]);
        }
    };

    // ---- projection / output mapping: lane -> channel oc, rows orow0 .. orow0 + ORW - 1 -------------------------
    const int oc = lane % C;
    const bool ohi = (ORW != RW) && (lane >= C);
    const int orow0 = ohi ? ORW : 0;
    long long obase[ORW];
    bool ovalid[ORW];
#pragma unroll
    for (int i = 0; i < ORW; ++i) {
        const int gr = row0 + orow0 + i;
        ovalid[i] = gr < a.n_rows;
        obase[i] = row_base(a, min(gr, a.n_rows - 1)) + oc;
    }
    const float blin = (!RAW_H && dir == 0) ? __ldg(w.lin_b + oc) : 0.0f;
    float* const outp = a.out[dir];

    auto emit = [&](int step, const float (&accp)[ORW]) {
        const int pos = dir ? S - 1 - step : step;
        const float* rs = res + ((step & 1) * RW + orow0) * RS + oc;
#pragma unroll
        for (int i = 0; i < ORW; ++i) {
            float v = accp[i];
            if (dir == 0) v += blin + rs[i * RS];
            if (ovalid[i]) outp[obase[i] + (long long)pos * a.stride_pos] = v;
        }
    };

    // ---- initial state: lane owns hidden units 2*lane, 2*lane+1 of all RW rows ------------------------------------
    float c[RW][2];
#pragma unroll
    for (int r = 0; r < RW; ++r) {
        const int gr = row0 + r;
        float2 hv = make_float2(0.f, 0.f), cv = make_float2(0.f, 0.f);
        if (a.h0 && gr < a.n_rows) {       // plain loads: hN / cN may alias h0 / c0
            const float* hp = a.h0 + (long long)gr * H + 2 * lane;
            const float* cp = a.c0 + (long long)gr * H + 2 * lane;
            hv = make_float2(ld_plain(hp), ld_plain(hp + 1));
            cv = make_float2(ld_plain(cp), ld_plain(cp + 1));
        }
        c[r][0] = cv.x; c[r][1] = cv.y;
        st2(A + r * AS + C + 2 * lane, hv);
    }
    {
        float4 x0v[NV];
        load_x(0, x0v);
        ln_store(x0v, 0);
    }
    __syncwarp();

    const float* Wl = Wt + 4 * lane;
    float hn[RW][2];
    for (int s = 0; s < S; ++s) {
        float4 xnext[NV];
        if (s + 1 < S) load_x(s + 1, xnext);

        float2 acc[RW][4];                  // {i, f, g, o} x (unit 2*lane, unit 2*lane+1): packed-FFMA2 accumulators
#pragma unroll
        for (int r = 0; r < RW; ++r)
#pragma unroll
            for (int j = 0; j < 4; ++j) acc[r][j] = make_float2(bias[2 * j], bias[2 * j + 1]);
        float accp[ORW];
#pragma unroll
        for (int i = 0; i < ORW; ++i) accp[i] = 0.f;

        // one block of four k: 8 broadcast LDS.128 (A) + 8 LDS.128 (W) feed 256 FFMA (+ the projection of h_{s-1})
        auto kblock = [&](int k4, auto with_proj) {
            float av[RW][4];
#pragma unroll
            for (int r = 0; r < RW; ++r) {
                const float4 t = ld4(A + r * AS + 4 * k4);
                av[r][0] = t.x; av[r][1] = t.y; av[r][2] = t.z; av[r][3] = t.w;
            }
#pragma unroll
            for (int kk = 0; kk < 4; ++kk) {
                const int k = 4 * k4 + kk;
                const float4 w0 = ld4(Wl + k * 256);
                const float4 w1 = ld4(Wl + k * 256 + 128);
#pragma unroll
                for (int r = 0; r < RW; ++r) {
                    const float x = av[r][kk];
                    ffma2(acc[r][0], make_float2(w0.x, w0.y), x);
                    ffma2(acc[r][1], make_float2(w0.z, w0.w), x);
                    ffma2(acc[r][2], make_float2(w1.x, w1.y), x);
                    ffma2(acc[r][3], make_float2(w1.z, w1.w), x);
                }
                if constexpr (decltype(with_proj)::value) {
                    const float wl = WlT[(k - C) * C + oc];
#pragma unroll
                    for (int i = 0; i < ORW; ++i) {
                        const float x = (ORW == RW) ? av[i][kk] : (ohi ? av[(i + ORW) % RW][kk] : av[i][kk]);
                        accp[i] = fmaf(x, wl, accp[i]);
                    }
                }
            }
        };
#pragma unroll 2
        for (int k4 = 0; k4 < C / 4; ++k4) kblock(k4, std::false_type{});
#pragma unroll 2
        for (int k4 = C / 4; k4 < K / 4; ++k4) kblock(k4, std::integral_constant<bool, !RAW_H>{});
        if (!RAW_H && s > 0) emit(s - 1, accp);

        // gates: acc[r] = {i, f, g, o} pairs over the lane's two hidden units  (PyTorch order i, f, g, o)
#pragma unroll
        for (int r = 0; r < RW; ++r) {
            {
                const float ig = sigmoid_f(acc[r][0].x), fg = sigmoid_f(acc[r][1].x);
                const float gg = tanh_f(acc[r][2].x), og = sigmoid_f(acc[r][3].x);
                c[r][0] = fmaf(fg, c[r][0], ig * gg);
                hn[r][0] = og * tanh_f(c[r][0]);
            }
            {
                const float ig = sigmoid_f(acc[r][0].y), fg = sigmoid_f(acc[r][1].y);
                const float gg = tanh_f(acc[r][2].y), og = sigmoid_f(acc[r][3].y);
                c[r][1] = fmaf(fg, c[r][1], ig * gg);
                hn[r][1] = og * tanh_f(c[r][1]);
            }
        }
        __syncwarp();                       // every lane is done reading A and res[(s-1)&1]
#pragma unroll
        for (int r = 0; r < RW; ++r) st2(A + r * AS + C + 2 * lane, make_float2(hn[r][0], hn[r][1]));
        if (RAW_H) {
            const int pos = dir ? S - 1 - s : s;
#pragma unroll
            for (int r = 0; r < RW; ++r)
                if (row0 + r < a.n_rows)
                    st2(outp + ((long long)(row0 + r) * S + pos) * H + 2 * lane, make_float2(hn[r][0], hn[r][1]));
        }
        if (s + 1 < S) ln_store(xnext, (s + 1) & 1);
        __syncwarp();
    }

    if (!RAW_H) {   // drain: projection of the last step
        float accp[ORW];
#pragma unroll
        for (int i = 0; i < ORW; ++i) accp[i] = 0.f;
#pragma unroll 4
        for (int k = 0; k < H; ++k) {
            const float wl = WlT[k * C + oc];
#pragma unroll
            for (int i = 0; i < ORW; ++i) accp[i] = fmaf(A[(orow0 + i) * AS + C + k], wl, accp[i]);
        }
        emit(S - 1, accp);
    }
    if (a.hN) {
#pragma unroll
        for (int r = 0; r < RW; ++r)
            if (row0 + r < a.n_rows) {
                st2(a.hN + (long long)(row0 + r) * H + 2 * lane, make_float2(hn[r][0], hn[r][1]));
                st2(a.cN + (long long)(row0 + r) * H + 2 * lane, make_float2(c[r][0], c[r][1]));
            }
    }
}

// =============================================================================================================
// lane kernel
// =============================================================================================================
template <int C, int RL, bool RAW_H>
__global__ void __launch_bounds__(256, 1) lstm_lane_kernel(const SeqArgs a) {
    constexpr int H = 64;
    constexpr int LPP = C / 4;            // lanes per (step,row) pair in the load/LayerNorm mapping
    constexpr int PB = 256 / LPP;         // pairs per block of steps
    constexpr int SB = PB / RL;           // steps per block
    constexpr int NP = H * C / 256;       // projection terms per thread
    constexpr int LPO = H / NP;           // lanes per projection output
    static_assert(PB * C == 1024, "block buffers are 1024 floats");
    static_assert(SB >= 2, "phase_c needs two steps of slack");
    __shared__ __align__(16) float xn[1024];
    __shared__ __align__(16) float res[2][1024];
    __shared__ __align__(16) float outb[2][1024];
    __shared__ __align__(16) float hs[2][RL * H];

    const int tid = threadIdx.x, lane = tid & 31;
    const int dir = blockIdx.y;
    const sb_lstm_dir& w = a.w[dir];
    const int S = a.n_steps;
    const int row0 = blockIdx.x * RL;
    const int nblk = (S + SB - 1) / SB;

    // gate-column role: slot tid <-> hidden unit u = tid / 4, gate g = tid % 4 (weight row g*H + u)
    const int g = lane & 3;
    const int u = tid >> 2;
    float wih[C], whh[H];
    {
        const float4* src = reinterpret_cast<const float4*>(w.w_lane) + tid;
#pragma unroll
        for (int qq = 0; qq < C / 4; ++qq) {
            const float4 v = __ldg(src + qq * 256);
            wih[4 * qq] = v.x; wih[4 * qq + 1] = v.y; wih[4 * qq + 2] = v.z; wih[4 * qq + 3] = v.w;
        }
#pragma unroll
        for (int qq = 0; qq < H / 4; ++qq) {
            const float4 v = __ldg(src + (C / 4 + qq) * 256);
            whh[4 * qq] = v.x; whh[4 * qq + 1] = v.y; whh[4 * qq + 2] = v.z; whh[4 * qq + 3] = v.w;
        }
    }
    const float bias = __ldg(w.b_lane + tid);
    const float act_in = (g == 2) ? 2.0f : 1.0f;        // tanh(x) = 2 sigmoid(2x) - 1 for the cell gate
    const float act_mul = (g == 2) ? 2.0f : 1.0f;
    const float act_add = (g == 2) ? -1.0f : 0.0f;
    const int quad = lane & ~3;

    // projection role: output channel po, k-slice pk
    const int po = tid / LPO, pk = tid % LPO;
    float wlin[NP];
#pragma unroll
    for (int j = 0; j < NP; ++j) wlin[j] = RAW_H ? 0.f : __ldg(w.lin_n + NP * tid + j);

    // load / LayerNorm role: pair pq (step-in-block, row), channel quad c4
    const int pq = tid / LPP, c4 = tid % LPP;
    const int p_sb = pq / RL, p_r = pq % RL;
    const bool p_rowok = row0 + p_r < a.n_rows;
    const int p_row = min(row0 + p_r, a.n_rows - 1);
    const long long p_base = row_base(a, p_row) + 4 * c4;
    const long long p_fbase = (long long)(p_row / a.film_row_div) * S * C + 4 * c4;
    const float4 g4 = __ldg(reinterpret_cast<const float4*>(w.ln_g) + c4);
    const float4 b4 = __ldg(reinterpret_cast<const float4*>(w.ln_b) + c4);
    float4 blin4 = make_float4(0, 0, 0, 0);
    if (!RAW_H && dir == 0) blin4 = __ldg(reinterpret_cast<const float4*>(w.lin_b) + c4);
    float* const outp = a.out[dir];

    pdl_trigger();
    pdl_wait();

    auto prefetch = [&](int blk) -> float4 {
        const int s = blk * SB + p_sb;
        float4 v = make_float4(0, 0, 0, 0);
        if (s < S && p_rowok) {
            const int pos = dir ? S - 1 - s : s;
            const long long off = p_base + (long long)pos * a.stride_pos;
            v = ldg4_stream(a.x0 + off);
            if (a.x1) {
                const float4 t = ldg4_stream(a.x1 + off);
                v.x += t.x; v.y += t.y; v.z += t.z; v.w += t.w;
            }
            if (a.film_scale) {
                const float4 fs = __ldg(reinterpret_cast<const float4*>(a.film_scale + p_fbase + (long long)pos * C));
                const float4 fb = __ldg(reinterpret_cast<const float4*>(a.film_shift + p_fbase + (long long)pos * C));
                v.x = fmaf(v.x, fs.x, fb.x); v.y = fmaf(v.y, fs.y, fb.y);
                v.z = fmaf(v.z, fs.z, fb.z); v.w = fmaf(v.w, fs.w, fb.w);
            }
        }
        return v;
    };
    auto phase_a = [&](int blk, const float4 v) {
        const float mean = group_sum<LPP>((v.x + v.y) + (v.z + v.w)) * (1.0f / C);
        const float dx = v.x - mean, dy = v.y - mean, dz = v.z - mean, dw = v.w - mean;
        const float var = group_sum<LPP>((dx * dx + dy * dy) + (dz * dz + dw * dw)) * (1.0f / C);
        const float rstd = rsqrtf(var + kLnEps);
        st4(xn + pq * C + 4 * c4, make_float4(fmaf(dx * rstd, g4.x, b4.x), fmaf(dy * rstd, g4.y, b4.y),
                                              fmaf(dz * rstd, g4.z, b4.z), fmaf(dw * rstd, g4.w, b4.w)));
        st4(res[blk & 1] + pq * C + 4 * c4, v);
    };
    auto phase_c = [&](int blk) {
        if (RAW_H) return;
        const int s = blk * SB + p_sb;
        if (s < S && p_rowok) {
            float4 v = ld4(outb[blk & 1] + pq * C + 4 * c4);
            if (dir == 0) {
                const float4 r = ld4(res[blk & 1] + pq * C + 4 * c4);
                v.x += blin4.x + r.x; v.y += blin4.y + r.y; v.z += blin4.z + r.z; v.w += blin4.w + r.w;
            }
            const int pos = dir ? S - 1 - s : s;
            st4(outp + p_base + (long long)pos * a.stride_pos, v);
        }
    };

    // ---- initial state -------------------------------------------------------------------------------------
    float c[RL], hlast[RL];
#pragma unroll
    for (int r = 0; r < RL; ++r) {
        const bool ok = a.h0 && row0 + r < a.n_rows;
        hlast[r] = ok ? ld_plain(a.h0 + (long long)(row0 + r) * H + u) : 0.0f;
        c[r] = ok ? ld_plain(a.c0 + (long long)(row0 + r) * H + u) : 0.0f;
        if (g == 0) hs[0][r * H + u] = hlast[r];
    }

    float4 xpre = prefetch(0);
    int cur = 0;
    for (int s = 0; s <= S; ++s) {
        const int blk = s / SB, sb = s - blk * SB;
        if (sb == 0 && s < S) {
            phase_a(blk, xpre);
            __syncthreads();
            xpre = prefetch(blk + 1);
        }
        float accx[RL];
        if (s < S) {
#pragma unroll
            for (int r = 0; r < RL; ++r) {
                const float* xr = xn + (sb * RL + r) * C;
                float a0 = bias, a1 = 0.f, a2 = 0.f, a3 = 0.f;
#pragma unroll
                for (int qq = 0; qq < C / 4; ++qq) {
                    const float4 v = ld4(xr + 4 * qq);
                    a0 = fmaf(wih[4 * qq], v.x, a0); a1 = fmaf(wih[4 * qq + 1], v.y, a1);
                    a2 = fmaf(wih[4 * qq + 2], v.z, a2); a3 = fmaf(wih[4 * qq + 3], v.w, a3);
                }
                accx[r] = (a0 + a1) + (a2 + a3);
            }
        }
        __syncthreads();                               // h_{s-1} of every column is in hs[cur]
        if (sb == 1 && blk > 0) phase_c(blk - 1);

#pragma unroll
        for (int r = 0; r < RL; ++r) {
            const float* hrow = hs[cur] + r * H;
            if (!RAW_H && s > 0) {                      // projection of step s-1 (needs the full h vector)
                float pp = 0.f;
#pragma unroll
                for (int j = 0; j < NP; j += 4) {
                    const float4 v = ld4(hrow + NP * pk + j);
                    pp = fmaf(wlin[j], v.x, pp); pp = fmaf(wlin[j + 1], v.y, pp);
                    pp = fmaf(wlin[j + 2], v.z, pp); pp = fmaf(wlin[j + 3], v.w, pp);
                }
                pp = group_sum<LPO>(pp);
                if (pk == 0) {
                    const int sp = s - 1;
                    const int bp = sp / SB;
                    outb[bp & 1][((sp - bp * SB) * RL + r) * C + po] = pp;
                }
            }
            if (s < S) {
                float a0 = accx[r], a1 = 0.f, a2 = 0.f, a3 = 0.f;
#pragma unroll
                for (int qq = 0; qq < H / 4; ++qq) {
                    const float4 v = ld4(hrow + 4 * qq);
                    a0 = fmaf(whh[4 * qq], v.x, a0); a1 = fmaf(whh[4 * qq + 1], v.y, a1);
                    a2 = fmaf(whh[4 * qq + 2], v.z, a2); a3 = fmaf(whh[4 * qq + 3], v.w, a3);
                }
                const float pre = (a0 + a1) + (a2 + a3);
                const float act = fmaf(sigmoid_f(pre * act_in), act_mul, act_add);
                const float ai = __shfl_sync(0xffffffffu, act, quad + 0);
                const float af = __shfl_sync(0xffffffffu, act, quad + 1);
                const float ag = __shfl_sync(0xffffffffu, act, quad + 2);
                const float ao = __shfl_sync(0xffffffffu, act, quad + 3);
                c[r] = fmaf(af, c[r], ai * ag);
                hlast[r] = ao * tanh_f(c[r]);
                if (g == 0) {
                    hs[cur ^ 1][r * H + u] = hlast[r];
                    if (RAW_H && row0 + r < a.n_rows) {
                        const int pos = dir ? S - 1 - s : s;
                        outp[((long long)(row0 + r) * S + pos) * H + u] = hlast[r];
                    }
                }
            }
        }
        cur ^= 1;
    }
    __syncthreads();
    phase_c(nblk - 1);

    if (g == 0 && a.hN) {
#pragma unroll
        for (int r = 0; r < RL; ++r) {
            if (row0 + r < a.n_rows) {
                a.hN[(long long)(row0 + r) * H + u] = hlast[r];
                a.cN[(long long)(row0 + r) * H + u] = c[r];
            }
        }
    }
}

// =============================================================================================================
// warp-specialised single-sequence kernel (streaming latency path)
// =============================================================================================================
// One sequence (and direction) per CTA, 256 threads.  What bounds a one-sequence LSTM step on an SM is not FMA issue
// but BYTES INTO REGISTERS: every gate column needs all 64 h values, and shared memory delivers 128 B/clk.  So:
//   warps 4-7 (recurrence): thread (ur, kq) owns all four gates of TWO hidden units (ur, ur+32) over a quarter of the
//     hidden state (k = 16kq..16kq+15): 128 weights in registers, 4 LDS.128 feed 64 packed FFMA2 (each loaded h value
//     is used 8 times -> 8 KB per step instead of 64 KB for one column per thread).  A two-stage shuffle
//     reduce-scatter over the quad leaves (i,f) of one unit on the even lane and (g,o) on the odd lane; two more
//     shuffles and the even lane updates c and h.  The projection of step s-1 reuses the h slice already in
//     registers (16 FMA, no loads).  One 128-thread named barrier per step.
//   warps 0-3 (helpers), a GROUP of 4 steps at a time, off the critical path: global loads + FiLM + LayerNorm per
//     block of SB steps, the input part of the gates (x W_ih^T + b) for the next group with the same mapping, and
//     the bias / residual / store of finished blocks.
// The two sides meet at one full barrier per group; gate inputs and hidden states travel through 8-deep rings.
template <int C, int NQ>
struct WsCfg {
    static constexpr int H = 64, LPP = C / 4, SB = 256 / LPP, XK = C / 4, HS = 20, G = 4, RING = 8;
    static constexpr int NPL = 32 / C;                      // projection planes (threads per output channel and quad)
    // per sequence of the CTA:
    static constexpr int xn_off = 0, res_off = 2 * SB * C, outp_off = 4 * SB * C;
    static constexpr int gx_off = outp_off + 2 * NPL * SB * C, hb_off = gx_off + RING * 256;
    static constexpr int seq_floats = hb_off + RING * 4 * HS;
    static constexpr int smem_floats = NQ * seq_floats;
    static_assert(SB >= 32 && SB % (2 * G) == 0 && XK % 4 == 0, "block / slice sizes");
};

// NQ sequences per CTA (rows NQ*blockIdx.x ..): the recurrence warps run the step of every sequence with the SAME weights
// in registers, phase by phase, so the shuffles, MUFU round trips and stores of one sequence sit in the shadow of the
// other's FMAs.  NQ = 1 is the lowest-latency shape (one sequence per SM); NQ = 2 (SB_ALGO_WS2) costs 1.6x the latency
// for 0.81x the SM-time per sequence, which is what a saturated pipelined session pays for.
template <int C, bool RAW_H, int NQ>
__global__ void __launch_bounds__(256, 1) lstm_ws_kernel(const SeqArgs a) {
    using Cfg = WsCfg<C, NQ>;
    constexpr int H = Cfg::H, LPP = Cfg::LPP, SB = Cfg::SB, XK = Cfg::XK, HS = Cfg::HS, NPL = Cfg::NPL;
    constexpr int G = Cfg::G, RING = Cfg::RING, SQ = Cfg::seq_floats;
    SB_DYN_SMEM(float, smem);
    float* xn = smem + Cfg::xn_off;         // [2][SB][C]        LayerNorm(x') of the current / next block   (+ q * SQ)
    float* res = smem + Cfg::res_off;       // [2][SB][C]        x' (residual)
    float* outp = smem + Cfg::outp_off;     // [2][NPL][SB][C]   (partial) projections of finished steps
    float* gx = smem + Cfg::gx_off;         // [RING][64][4]     x W_ih^T + b of step s in slot s % RING: unit, gate
    float* hb = smem + Cfg::hb_off;         // [RING][4][HS]     h_{s-1} in slot s % RING: four 16-float slices padded to 20

    const int tid = threadIdx.x;
    const int dir = blockIdx.y;
    const sb_lstm_dir& w = a.w[dir];
    const int S = a.n_steps;
    const int nblk = (S + SB - 1) / SB;
    const int ngrp = (S + G - 1) / G;
    const bool recur = tid >= 128;                       // warp-uniform role; the recurrence gets the higher warp ids,
                                                         // which the issue arbiter favours when both sides are ready
    const int t7 = tid & 127;
    const int ur = t7 >> 2, kq = t7 & 3;
    const bool hi = (kq & 2) != 0, odd = (kq & 1) != 0;
    const int ux = hi ? ur + 32 : ur;                     // the unit this lane ends up with after the reduce-scatter
    const int gsl = 4 * ux + (odd ? 2 : 0);               // its two gates: (i,f) on even lanes, (g,o) on odd lanes
    int row[NQ];
    bool rok[NQ];
#pragma unroll
    for (int q = 0; q < NQ; ++q) {
        rok[q] = blockIdx.x * NQ + q < a.n_rows;          // a ragged last CTA computes its missing row on a copy
        row[q] = min(blockIdx.x * NQ + q, a.n_rows - 1);  // of the last one and stores nothing for it
    }

    if (recur) {
        // ---------------------------------------------------------------------------------------- recurrence warps
        float4 wA[16], wB[16];
#pragma unroll
        for (int k = 0; k < 16; ++k) {
            wA[k] = __ldg(reinterpret_cast<const float4*>(w.w_rec) + (2 * k) * 128 + t7);
            wB[k] = __ldg(reinterpret_cast<const float4*>(w.w_rec) + (2 * k + 1) * 128 + t7);
        }
        float wp[16];
        if (!RAW_H) {
#pragma unroll
            for (int j4 = 0; j4 < 4; ++j4) {
                const float4 v = __ldg(reinterpret_cast<const float4*>(w.w_prj) + j4 * 128 + t7);
                wp[4 * j4] = v.x; wp[4 * j4 + 1] = v.y; wp[4 * j4 + 2] = v.z; wp[4 * j4 + 3] = v.w;
            }
        }
        const int pc = ur % C, ppl = ur / C;             // projection: output channel, plane
        pdl_trigger();
        pdl_wait();
        const bool has0 = a.h0 != nullptr;
        const int hslot = (ux >> 4) * HS + (ux & 15);
        float c[NQ], hlast[NQ];                           // meaningful on even lanes
#pragma unroll
        for (int q = 0; q < NQ; ++q) {
            c[q] = has0 ? ld_plain(a.c0 + (long long)row[q] * H + ux) : 0.0f;
            hlast[q] = has0 ? ld_plain(a.h0 + (long long)row[q] * H + ux) : 0.0f;
            if (!odd) hb[q * SQ + hslot] = hlast[q];
        }
        float* const outr = a.out[dir];
        const float act_in = odd ? 2.0f : 1.0f;           // odd lanes: value 0 is the cell gate -> tanh(x) = 2 sigmoid(2x) - 1
        const float act_mul = odd ? 2.0f : 1.0f, act_add = odd ? -1.0f : 0.0f;
        __syncthreads();                                  // (P) pairs with the helpers' prologue barriers
        __syncthreads();                                  // (Q)

        // projection of step sp from the h slice already in registers: 16 FMA, a quad reduction, one predicated store
        auto project = [&](const float (&hr)[16], int q, int sp, bool store) {
            float p0 = wp[0] * hr[0], p1 = wp[1] * hr[1], p2 = wp[2] * hr[2], p3 = wp[3] * hr[3];
#pragma unroll
            for (int k = 4; k < 16; k += 4) {
                p0 = fmaf(wp[k], hr[k], p0); p1 = fmaf(wp[k + 1], hr[k + 1], p1);
                p2 = fmaf(wp[k + 2], hr[k + 2], p2); p3 = fmaf(wp[k + 3], hr[k + 3], p3);
            }
            float pp = (p0 + p1) + (p2 + p3);
            pp += __shfl_xor_sync(0xffffffffu, pp, 1);
            pp += __shfl_xor_sync(0xffffffffu, pp, 2);
            const int spc = sp < 0 ? 0 : sp;
            const int bp = spc / SB;
            float* dst = outp + q * SQ + (((bp & 1) * NPL + ppl) * SB + (spc - bp * SB)) * C + pc;
            if (store && kq == 0) *dst = pp;
        };
        auto load_h = [&](int s, int q, float (&hr)[16]) {
            const float* hs = hb + q * SQ + (s & (RING - 1)) * 4 * HS + kq * HS;
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const float4 v = ld4(hs + 4 * i);
                hr[4 * i] = v.x; hr[4 * i + 1] = v.y; hr[4 * i + 2] = v.z; hr[4 * i + 3] = v.w;
            }
        };

        for (int g = 0; g < ngrp; ++g) {
            __syncthreads();                              // group barrier: gx of this group is ready
#pragma unroll 1
            for (int j = 0; j < G; ++j) {
                const int s = g * G + j;
                if (s >= S) break;
                if (j > 0) bar_sync(1, 128);              // h_{s-1} of every unit is in the ring
                // the step of the NQ sequences, phase by phase
                float hr[NQ][16];
                float2 g2[NQ];
                float mine0[NQ], mine1[NQ], r1a[NQ], r1b[NQ], r2a[NQ], r2b[NQ], r3a[NQ], r3b[NQ];
#pragma unroll
                for (int q = 0; q < NQ; ++q) {
                    load_h(s, q, hr[q]);
                    g2[q] = ld2(gx + q * SQ + (s & (RING - 1)) * 256 + gsl);
                }
#pragma unroll
                for (int q = 0; q < NQ; ++q) {
                    float2 aA01[2], aA23[2], aB01[2], aB23[2];
#pragma unroll
                    for (int e = 0; e < 2; ++e) {
                        aA01[e] = make_float2(0.f, 0.f); aA23[e] = aA01[e]; aB01[e] = aA01[e]; aB23[e] = aA01[e];
                    }
#pragma unroll
                    for (int k = 0; k < 16; ++k) {
                        ffma2(aA01[k & 1], make_float2(wA[k].x, wA[k].y), hr[q][k]);
                        ffma2(aA23[k & 1], make_float2(wA[k].z, wA[k].w), hr[q][k]);
                        ffma2(aB01[k & 1], make_float2(wB[k].x, wB[k].y), hr[q][k]);
                        ffma2(aB23[k & 1], make_float2(wB[k].z, wB[k].w), hr[q][k]);
                    }
                    const float A0 = aA01[0].x + aA01[1].x, A1 = aA01[0].y + aA01[1].y;
                    const float A2 = aA23[0].x + aA23[1].x, A3 = aA23[0].y + aA23[1].y;
                    const float B0 = aB01[0].x + aB01[1].x, B1 = aB01[0].y + aB01[1].y;
                    const float B2 = aB23[0].x + aB23[1].x, B3 = aB23[0].y + aB23[1].y;
                    // one-level reduce-scatter over the quad: this lane finishes (unit hi ? B : A, gates odd ? (g,o) : (i,f)); every
                    // partner sends the pair its reader wants: xor 1 = same unit / other pair, xor 2 = other unit / same pair,
                    // xor 3 = other unit / other pair.  Six independent shuffles, one latency level.
                    const float m0 = hi ? B0 : A0, m1 = hi ? B1 : A1, m2 = hi ? B2 : A2, m3 = hi ? B3 : A3;     // my unit
                    const float o0 = hi ? A0 : B0, o1 = hi ? A1 : B1, o2 = hi ? A2 : B2, o3 = hi ? A3 : B3;     // the other unit
                    mine0[q] = odd ? m2 : m0; mine1[q] = odd ? m3 : m1;
                    r1a[q] = __shfl_xor_sync(0xffffffffu, odd ? m0 : m2, 1);
                    r1b[q] = __shfl_xor_sync(0xffffffffu, odd ? m1 : m3, 1);
                    r2a[q] = __shfl_xor_sync(0xffffffffu, odd ? o2 : o0, 2);
                    r2b[q] = __shfl_xor_sync(0xffffffffu, odd ? o3 : o1, 2);
                    r3a[q] = __shfl_xor_sync(0xffffffffu, odd ? o0 : o2, 3);
                    r3b[q] = __shfl_xor_sync(0xffffffffu, odd ? o1 : o3, 3);
                }
                if (!RAW_H) {                             // rides in the shadow of the gate shuffles
#pragma unroll
                    for (int q = 0; q < NQ; ++q) project(hr[q], q, s - 1, s > 0);
                }
                float act0[NQ], act1[NQ], tg[NQ], so[NQ];
#pragma unroll
                for (int q = 0; q < NQ; ++q) {
                    const float v0 = (mine0[q] + r1a[q]) + (r2a[q] + r3a[q]);
                    const float v1 = (mine1[q] + r1b[q]) + (r2b[q] + r3b[q]);
                    act0[q] = fmaf(sigmoid_f((v0 + g2[q].x) * act_in), act_mul, act_add);   // sigma(i) | tanh(g)
                    act1[q] = sigmoid_f(v1 + g2[q].y);                                      // sigma(f) | sigma(o)
                }
#pragma unroll
                for (int q = 0; q < NQ; ++q) {
                    tg[q] = __shfl_xor_sync(0xffffffffu, act0[q], 1);            // even lanes receive tanh(g), sigma(o)
                    so[q] = __shfl_xor_sync(0xffffffffu, act1[q], 1);
                }
#pragma unroll
                for (int q = 0; q < NQ; ++q) {
                    c[q] = fmaf(act1[q], c[q], act0[q] * tg[q]);                 // even: c = sigma(f) c + sigma(i) tanh(g)
                    hlast[q] = so[q] * tanh_f(c[q]);
                }
#pragma unroll
                for (int q = 0; q < NQ; ++q) {
                    if (!odd) hb[q * SQ + ((s + 1) & (RING - 1)) * 4 * HS + hslot] = hlast[q];
                    if (RAW_H && !odd && rok[q]) {
                        const int pos = dir ? S - 1 - s : s;
                        outr[((long long)row[q] * S + pos) * H + ux] = hlast[q];
                    }
                }
            }
        }
        __syncthreads();                                  // h_{S-1} is in the ring (pairs with the helpers' last group barrier)
        if (!RAW_H) {
#pragma unroll
            for (int q = 0; q < NQ; ++q) {
                float hr[16];
                load_h(S, q, hr);
                project(hr, q, S - 1, true);
            }
        }
        __syncthreads();                                  // pairs with the helpers' closing barrier
        if (!odd && a.hN) {
#pragma unroll
            for (int q = 0; q < NQ; ++q)
                if (rok[q]) {
                    a.hN[(long long)row[q] * H + ux] = hlast[q];
                    a.cN[(long long)row[q] * H + ux] = c[q];
                }
        }
        return;
    }

    // -------------------------------------------------------------------------------------------------- helper warps
    float4 xA[XK], xB[XK];
#pragma unroll
    for (int k = 0; k < XK; ++k) {
        xA[k] = __ldg(reinterpret_cast<const float4*>(w.w_xp) + (2 * k) * 128 + t7);
        xB[k] = __ldg(reinterpret_cast<const float4*>(w.w_xp) + (2 * k + 1) * 128 + t7);
    }
    const float2 gbias = __ldg(reinterpret_cast<const float2*>(w.b_lane + gsl));
    // load / LayerNorm / store role: each helper thread serves steps pq and pq + SB/2 of a block, channels 4c4..4c4+3
    const int pq = t7 / LPP, c4 = t7 % LPP;
    const float4 g4 = __ldg(reinterpret_cast<const float4*>(w.ln_g) + c4);
    const float4 b4 = __ldg(reinterpret_cast<const float4*>(w.ln_b) + c4);
    float4 blin4 = make_float4(0, 0, 0, 0);
    if (!RAW_H && dir == 0) blin4 = __ldg(reinterpret_cast<const float4*>(w.lin_b) + c4);
    long long p_base[NQ], p_fbase[NQ];
#pragma unroll
    for (int q = 0; q < NQ; ++q) {
        p_base[q] = row_base(a, row[q]) + 4 * c4;
        p_fbase[q] = (long long)(row[q] / a.film_row_div) * S * C + 4 * c4;
    }
    float* const outp_g = a.out[dir];
    pdl_trigger();
    pdl_wait();

    auto prefetch1 = [&](int q, int s) -> float4 {
        float4 v = make_float4(0, 0, 0, 0);
        if (s < S) {
            const int pos = dir ? S - 1 - s : s;
            const long long off = p_base[q] + (long long)pos * a.stride_pos;
            v = ldg4_stream(a.x0 + off);
            if (a.x1) {
                const float4 t = ldg4_stream(a.x1 + off);
                v.x += t.x; v.y += t.y; v.z += t.z; v.w += t.w;
            }
            if (a.film_scale) {
                const float4 fs = __ldg(reinterpret_cast<const float4*>(a.film_scale + p_fbase[q] + (long long)pos * C));
                const float4 fb = __ldg(reinterpret_cast<const float4*>(a.film_shift + p_fbase[q] + (long long)pos * C));
                v.x = fmaf(v.x, fs.x, fb.x); v.y = fmaf(v.y, fs.y, fb.y);
                v.z = fmaf(v.z, fs.z, fb.z); v.w = fmaf(v.w, fs.w, fb.w);
            }
        }
        return v;
    };
    float4 xpre[NQ][2];
    auto prefetch = [&](int blk) {
#pragma unroll
        for (int q = 0; q < NQ; ++q) {
            xpre[q][0] = prefetch1(q, blk * SB + pq);
            xpre[q][1] = prefetch1(q, blk * SB + pq + SB / 2);
        }
    };
    auto phase_a = [&](int blk) {                         // LayerNorm(C) of one step by LPP adjacent lanes
#pragma unroll
        for (int q = 0; q < NQ; ++q)
#pragma unroll
            for (int e = 0; e < 2; ++e) {
                const float4 v = xpre[q][e];
                const int st = pq + e * (SB / 2);
                const float mean = group_sum<LPP>((v.x + v.y) + (v.z + v.w)) * (1.0f / C);
                const float dx = v.x - mean, dy = v.y - mean, dz = v.z - mean, dw = v.w - mean;
                const float var = group_sum<LPP>((dx * dx + dy * dy) + (dz * dz + dw * dw)) * (1.0f / C);
                const float rstd = rsqrtf(var + kLnEps);
                st4(xn + q * SQ + ((blk & 1) * SB + st) * C + 4 * c4,
                    make_float4(fmaf(dx * rstd, g4.x, b4.x), fmaf(dy * rstd, g4.y, b4.y),
                                fmaf(dz * rstd, g4.z, b4.z), fmaf(dw * rstd, g4.w, b4.w)));
                if (!RAW_H) st4(res + q * SQ + ((blk & 1) * SB + st) * C + 4 * c4, v);
            }
    };
    auto phase_c = [&](int blk) {                         // finished block: projection + bias + residual -> global
        if (RAW_H) return;
#pragma unroll
        for (int q = 0; q < NQ; ++q)
#pragma unroll
            for (int e = 0; e < 2; ++e) {
                const int st = pq + e * (SB / 2);
                const int s = blk * SB + st;
                if (s < S && rok[q]) {
                    float4 v = make_float4(0, 0, 0, 0);
#pragma unroll
                    for (int pl = 0; pl < NPL; ++pl) {
                        const float4 t = ld4(outp + q * SQ + (((blk & 1) * NPL + pl) * SB + st) * C + 4 * c4);
                        v.x += t.x; v.y += t.y; v.z += t.z; v.w += t.w;
                    }
                    if (dir == 0) {
                        const float4 r = ld4(res + q * SQ + ((blk & 1) * SB + st) * C + 4 * c4);
                        v.x += blin4.x + r.x; v.y += blin4.y + r.y; v.z += blin4.z + r.z; v.w += blin4.w + r.w;
                    }
                    const int pos = dir ? S - 1 - s : s;
                    st4(outp_g + p_base[q] + (long long)pos * a.stride_pos, v);
                }
            }
    };
    // gx[s] = LN(x_s) W_ih^T + b for the G steps of a group: same (two units, K quarter) mapping as the recurrence
    auto x_part_group = [&](int s0) {
        const int blk = s0 / SB;
#pragma unroll
        for (int q = 0; q < NQ; ++q) {
            const float* xr = xn + q * SQ + ((blk & 1) * SB + (s0 - blk * SB)) * C + XK * kq;
#pragma unroll
            for (int j = 0; j < G; ++j) {
                float xv[XK];
#pragma unroll
                for (int i = 0; i < XK / 4; ++i) {
                    const float4 v = ld4(xr + j * C + 4 * i);
                    xv[4 * i] = v.x; xv[4 * i + 1] = v.y; xv[4 * i + 2] = v.z; xv[4 * i + 3] = v.w;
                }
                float2 aA01 = make_float2(0.f, 0.f), aA23 = aA01, aB01 = aA01, aB23 = aA01;
#pragma unroll
                for (int k = 0; k < XK; ++k) {
                    ffma2(aA01, make_float2(xA[k].x, xA[k].y), xv[k]);
                    ffma2(aA23, make_float2(xA[k].z, xA[k].w), xv[k]);
                    ffma2(aB01, make_float2(xB[k].x, xB[k].y), xv[k]);
                    ffma2(aB23, make_float2(xB[k].z, xB[k].w), xv[k]);
                }
                float k0 = hi ? aB01.x : aA01.x, k1 = hi ? aB01.y : aA01.y, k2 = hi ? aB23.x : aA23.x, k3 = hi ? aB23.y : aA23.y;
                k0 += __shfl_xor_sync(0xffffffffu, hi ? aA01.x : aB01.x, 2);
                k1 += __shfl_xor_sync(0xffffffffu, hi ? aA01.y : aB01.y, 2);
                k2 += __shfl_xor_sync(0xffffffffu, hi ? aA23.x : aB23.x, 2);
                k3 += __shfl_xor_sync(0xffffffffu, hi ? aA23.y : aB23.y, 2);
                float v0 = odd ? k2 : k0, v1 = odd ? k3 : k1;
                v0 += __shfl_xor_sync(0xffffffffu, odd ? k0 : k2, 1);
                v1 += __shfl_xor_sync(0xffffffffu, odd ? k1 : k3, 1);
                if (s0 + j < S) st2(gx + q * SQ + ((s0 + j) & (RING - 1)) * 256 + gsl, make_float2(v0 + gbias.x, v1 + gbias.y));
            }
        }
    };

    prefetch(0);
    phase_a(0);
    __syncthreads();                                      // (P) xn of block 0 is complete
    x_part_group(0);
    if (nblk > 1) prefetch(1);
    __syncthreads();                                      // (Q) gx of group 0 is complete
    int next_c = 0;
    constexpr int GPB = SB / G;                           // groups per block
    for (int g = 0; g <= ngrp; ++g) {
        __syncthreads();                                  // group barrier
        const int blk = g / GPB, gb = g - blk * GPB;
        if (g + 1 < ngrp) x_part_group((g + 1) * G);      // needs xn of block (g+1)/GPB: written >= one group barrier ago
        if (gb == GPB / 2 - 1 && blk + 1 < nblk) phase_a(blk + 1);
        if (gb == GPB / 2 && blk + 2 < nblk) prefetch(blk + 2);
        if (gb == 2 && blk > 0) { phase_c(blk - 1); next_c = blk; }
    }
    __syncthreads();                                      // every projection is in outp
    for (; next_c < nblk; ++next_c) phase_c(next_c);
}

// =============================================================================================================
// host side
// =============================================================================================================
// Cost model in SM cycles per recurrent step, measured on B200 at 1965 MHz (profiles/r01_lstm_bench.txt):
//   tile : one warp (8 sequences) per SMSP needs ~13.2k cycles per step, two warps per SMSP ~20.7k (the launch gives a
//          CTA ceil(tasks / SMs) <= 8 warps); more tasks than 8 x SMs run in rounds
//   ws   : ~600-790 (one sequence per CTA, warp-specialised);  lane1/2/4 : 1360 / 2530 / 4950 (1/2/4 sequences per CTA)
//   ws2  : ~1260 for the two sequences of a CTA
//   tc   : ~15.5k per step for up to one wave of 128-sequence CTAs (tcgen05 gate GEMM + 256-thread cell update), ~10.8k
//          per wave once several waves keep every SM busy; wins when the tile family needs several rounds (offline)
// plus a launch + prologue constant; the cheapest family for (rows, dirs, steps) wins.
static int pick_algo(int n_rows, int n_dirs, int S, int sms, bool tc_ok, bool single_addend) {
    const int tasks = ceil_div(n_rows, 8) * n_dirs;
    const int resident = sms * 8;
    const double rounds = tasks <= resident ? 1.0 : (double)ceil_div(tasks, resident);
    const int warps_per_cta = tasks >= resident ? 8 : ceil_div(tasks, sms);
    const double tile_step = warps_per_cta <= 4 ? 13200.0 : 20700.0;
    const double tile = rounds * S * tile_step + 30000.0;
    auto per_cta = [&](int rl, double step, double fixed) {
        return (double)ceil_div(ceil_div(n_rows, rl) * n_dirs, sms) * (S * step + fixed);
    };
    const double cost[6] = {0.0, tile, per_cta(1, 1360.0, 8000.0), per_cta(2, 2530.0, 8000.0), per_cta(4, 4950.0, 8000.0),
                            per_cta(1, 600.0, 12000.0)};
    int best = SB_ALGO_TILE;
    for (int k = SB_ALGO_LANE1; k <= SB_ALGO_WS; ++k)
        if (cost[k] < cost[best]) best = k;
    double best_cost = cost[best];
    const double ws2 = per_cta(2, 1260.0, 14000.0);        // two sequences per CTA: wins when it saves a round of CTAs
    if (ws2 < best_cost) { best = SB_ALGO_WS2; best_cost = ws2; }
    if (tc_ok) {
        const double waves = (double)(ceil_div(n_rows, 128) * n_dirs) / sms;
        // per step and round of CTAs (one CTA per SM): lstm_tcr_kernel 3.45 us = 6.8k cycles (single-addend calls, profiles/
        // r02_tcr_ab.txt), lstm_tcp_kernel 4.36 us = 8.6k cycles
        const double step = single_addend && tc_pipe_enabled() && tc_cell7_enabled() ? 6800.0 : 8600.0;
        const double tc = tc_v1_enabled() ? S * (waves <= 1.0 ? 15500.0 : waves * 10800.0) + 60000.0
                                          : S * ceil(waves) * step + 40000.0;
        if (tc < best_cost) return SB_ALGO_TC;
    }
    return best;
}

template <int C, bool RAW_H>
static int run_seq_c(const SeqArgs& a, int algo, cudaStream_t st) {
    const int sms = sm_count();
#ifdef SB_EMU
    constexpr bool tc_ok = false;           // tensor-core instructions cannot be emulated on the host
#else
    constexpr bool tc_ok = C == 32 && !RAW_H;
#endif
    if (algo == SB_ALGO_AUTO) algo = pick_algo(a.n_rows, a.n_dirs, a.n_steps, sms, tc_ok, a.x1 == nullptr);
    switch (algo) {
        case SB_ALGO_TC:
            if constexpr (C == 32 && !RAW_H)
                return (!tc_v1_enabled() && seq_tcp_supported(a)) ? run_seq_tcp(a, st) : run_seq_tc(a, st);
            break;
        case SB_ALGO_TILE4:
        case SB_ALGO_TILE: {
            // 8 sequences per warp; 4 when the call is a step or two long (streaming inter path) and rows are few enough
            // that halving a warp's serial work beats the lower FMA : shared-memory-load ratio
            const bool short_call = a.n_steps <= 2 && ceil_div(a.n_rows, 4) * a.n_dirs <= 8 * sms;
            const bool small = C == 32 && (algo == SB_ALGO_TILE4 || short_call);
            const int rw = small ? 4 : 8;
            const int tasks = ceil_div(a.n_rows, rw);
            int nw = ceil_div(tasks * a.n_dirs, sms);
            nw = nw < 1 ? 1 : (nw > 8 ? 8 : nw);
            dim3 grid(ceil_div(tasks, nw), a.n_dirs);
            if constexpr (C == 32) {
                if (small)
                    return launch("lstm_tile4", lstm_tile_kernel<C, RAW_H, 4>, grid, dim3(32 * nw),
                                  TileCfg<C, 4>::smem_floats(nw) * sizeof(float), st, a);
            }
            return launch("lstm_tile", lstm_tile_kernel<C, RAW_H, 8>, grid, dim3(32 * nw),
                          TileCfg<C, 8>::smem_floats(nw) * sizeof(float), st, a);
        }
        case SB_ALGO_LANE1:
            return launch("lstm_lane1", lstm_lane_kernel<C, 1, RAW_H>, dim3(a.n_rows, a.n_dirs), dim3(256), 0, st, a);
        case SB_ALGO_LANE2:
            return launch("lstm_lane2", lstm_lane_kernel<C, 2, RAW_H>, dim3(ceil_div(a.n_rows, 2), a.n_dirs), dim3(256), 0, st, a);
        case SB_ALGO_LANE4:
            return launch("lstm_lane4", lstm_lane_kernel<C, 4, RAW_H>, dim3(ceil_div(a.n_rows, 4), a.n_dirs), dim3(256), 0, st, a);
        case SB_ALGO_WS:
            return launch("lstm_ws", lstm_ws_kernel<C, RAW_H, 1>, dim3(a.n_rows, a.n_dirs), dim3(256),
                          WsCfg<C, 1>::smem_floats * sizeof(float), st, a);
        case SB_ALGO_WS2:
            return launch("lstm_ws2", lstm_ws_kernel<C, RAW_H, 2>, dim3(ceil_div(a.n_rows, 2), a.n_dirs), dim3(256),
                          WsCfg<C, 2>::smem_floats * sizeof(float), st, a);
        default: break;
    }
    set_error("unknown LSTM algo %d", algo);
    return SB_E_BADARG;
}

// SeqArgs::sum_dirs is implemented by lstm_tcr_kernel alone: would this call end up there?
static bool seq_sum_supported(const SeqArgs& a, int C, int H, bool raw_h, int algo) {
#ifdef SB_EMU
    (void)a; (void)C; (void)H; (void)raw_h; (void)algo;
    return false;
#else
    if (C != 32 || H != 64 || raw_h || a.n_rows <= 0 || a.n_steps <= 0) return false;
    if (algo == SB_ALGO_AUTO) algo = pick_algo(a.n_rows, a.n_dirs, a.n_steps, sm_count(), true, a.x1 == nullptr);
    if (algo == SB_ALGO_TCP) return seq_tcr_selected(a);
    return algo == SB_ALGO_TC && !tc_v1_enabled() && seq_tcr_selected(a);
#endif
}

int run_seq(const SeqArgs& a, int C, int H, bool raw_h, int algo, cudaStream_t st) {
    SB_REQUIRE(!a.sum_dirs || seq_sum_supported(a, C, H, raw_h, algo), SB_E_UNSUPP,
               "summed directions (y_bwd == y_fwd) need the pipelined tensor-core LSTM kernel, which this call does not select");
    SB_REQUIRE(H == 64, SB_E_UNSUPP, "LSTM kernels are instantiated for H=64 only (got H=%d)", H);
    SB_REQUIRE(C == 32 || C == 16, SB_E_UNSUPP, "LSTM kernels are instantiated for C in {16, 32} (got C=%d)", C);
    SB_REQUIRE(a.n_rows > 0 && a.n_steps > 0, SB_E_BADARG, "empty LSTM problem (%d rows, %d steps)", a.n_rows, a.n_steps);
    if (algo == SB_ALGO_TC || algo == SB_ALGO_TCP || algo == SB_ALGO_TCQ) {
        SB_REQUIRE(C == 32 && !raw_h, SB_E_UNSUPP, "SB_ALGO_TC / SB_ALGO_TCP / SB_ALGO_TCQ support C=32 in projected mode only");
        if (algo == SB_ALGO_TCP) return run_seq_tcp(a, st);
        if (algo == SB_ALGO_TCQ) return run_seq_tcq(a, st);
        return (!tc_v1_enabled() && seq_tcp_supported(a)) ? run_seq_tcp(a, st) : run_seq_tc(a, st);
    }
    if (C == 32) return raw_h ? run_seq_c<32, true>(a, algo, st) : run_seq_c<32, false>(a, algo, st);
    return raw_h ? run_seq_c<16, true>(a, algo, st) : run_seq_c<16, false>(a, algo, st);
}

}  // namespace sb

extern "C" int sb_intra_lstm_fwd(const sb_intra_args* p, void* stream) {
    using namespace sb;
    SB_REQUIRE(p && p->x && p->y_fwd && p->y_bwd, SB_E_BADARG, "sb_intra_lstm_fwd: null pointer");
    SB_REQUIRE(p->B > 0 && p->T > 0 && p->F > 0, SB_E_BADARG, "sb_intra_lstm_fwd: bad sizes");
    SB_REQUIRE((p->film_scale == nullptr) == (p->film_shift == nullptr), SB_E_BADARG, "film scale/shift must come together");
    SeqArgs a{};
    a.x0 = p->x; a.x1 = nullptr;
    a.film_scale = p->film_scale; a.film_shift = p->film_shift;
    a.out[0] = p->y_fwd; a.out[1] = p->y_bwd;
    a.w[0] = p->dir[0]; a.w[1] = p->dir[1];
    a.n_rows = p->B * p->T; a.n_steps = p->F; a.n_dirs = 2;
    a.rows_inner = a.n_rows;                               // row (b,t) -> row * F * C
    a.stride_outer = 0; a.stride_inner = (long long)p->F * p->C; a.stride_pos = p->C;
    a.film_row_div = p->T;                                 // film[b][f][c]
    a.sum_dirs = p->y_bwd == p->y_fwd;
    if (a.sum_dirs) {
        SB_REQUIRE(seq_sum_supported(a, p->C, p->H, false, p->algo), SB_E_UNSUPP,
                   "sb_intra_lstm_fwd: y_bwd == y_fwd (summed directions) is not available for this call, see sb_intra_sum_supported");
#ifndef SB_EMU
        const cudaError_t e = cudaMemsetAsync(p->y_fwd, 0, (size_t)a.n_rows * p->F * p->C * sizeof(float), (cudaStream_t)stream);
        SB_REQUIRE(e == cudaSuccess, (int)e, "sb_intra_lstm_fwd: cudaMemsetAsync failed: %s", cudaGetErrorString(e));
#endif
    }
    return run_seq(a, p->C, p->H, false, p->algo, (cudaStream_t)stream);
}

extern "C" int sb_intra_sum_supported(const sb_intra_args* p) {
    using namespace sb;
    if (!p || !p->x || !p->y_fwd || p->B <= 0 || p->T <= 0 || p->F <= 0) return 0;
    SeqArgs a{};
    a.x0 = p->x;
    a.out[0] = p->y_fwd; a.out[1] = p->y_fwd;
    a.n_rows = p->B * p->T; a.n_steps = p->F; a.n_dirs = 2;
    a.rows_inner = a.n_rows;
    a.stride_outer = 0; a.stride_inner = (long long)p->F * p->C; a.stride_pos = p->C;
    a.film_row_div = p->T;
    a.sum_dirs = 1;
    return seq_sum_supported(a, p->C, p->H, false, p->algo) ? 1 : 0;
}

extern "C" int sb_inter_lstm_fwd(const sb_inter_args* p, void* stream) {
    using namespace sb;
    SB_REQUIRE(p && p->x0 && p->y, SB_E_BADARG, "sb_inter_lstm_fwd: null pointer");
    SB_REQUIRE(p->B > 0 && p->T > 0 && p->F > 0, SB_E_BADARG, "sb_inter_lstm_fwd: bad sizes");
    SB_REQUIRE((p->h0 == nullptr) == (p->c0 == nullptr), SB_E_BADARG, "h0/c0 must come together");
    SB_REQUIRE((p->hN == nullptr) == (p->cN == nullptr), SB_E_BADARG, "hN/cN must come together");
    SeqArgs a{};
    a.x0 = p->x0; a.x1 = p->x1;
    a.out[0] = p->y; a.out[1] = nullptr;
    a.w[0] = p->dir; a.w[1] = p->dir;
    a.h0 = p->h0; a.c0 = p->c0; a.hN = p->hN; a.cN = p->cN;
    a.n_rows = p->B * p->F; a.n_steps = p->T; a.n_dirs = 1;
    a.rows_inner = p->F;                                   // row b*F + f  (DE3:833)
    a.stride_outer = (long long)p->T * p->F * p->C; a.stride_inner = p->C; a.stride_pos = (long long)p->F * p->C;
    a.film_row_div = 1;
    return run_seq(a, p->C, p->H, false, p->algo, (cudaStream_t)stream);
}
