// Tensor-core (tcgen05 / TMEM) variant of the fused sequence kernel for large batches of sequences (offline calls).
//
// Same arithmetic as lstm_tile_kernel (sb_lstm.cu): [FiLM] -> LayerNorm(C) -> x W_ih^T + h W_hh^T + b -> gates ->
// Linear -> residual.  Reference: GridNetBlock.forward intra / inter branches,
// src/models/tfgridnet_realtime_clean_dis_embd3/tfgridnet_causal.py:794-849.
//
// One CTA owns 128 sequences for the whole recurrence.  Per step:
//   A_s = [LN(x_s) | h_{s-1}]  (128 x 96)  lives in shared memory as bf16 hi + bf16 lo in the canonical no-swizzle
//   K-major UMMA layout; the gate matrix (256 x 96, rows n = 4u + g) and the projection (32 x 64) sit next to it the
//   same way for the whole kernel.  fp32 parity needs more than bf16: every product is issued as the three-term split
//   A_hi W_hi + A_hi W_lo + A_lo W_hi (relative error ~2^-17, fp32 accumulation in TMEM).
//   One elected thread issues per step: 12 small MMAs (projection of h_{s-1}, N = 32), then 18 MMAs (N = 128) for
//   units 0..31 + commit, 18 MMAs for units 32..63 + commit.  Warps 0-3 (units 0..31) start their cell update from
//   TMEM while the second half is still in the tensor pipe.
//   The 256 threads are the epilogue: thread (row, half) pulls its 32 units x 4 gates from TMEM (tcgen05.ld), adds the
//   bias, evaluates the cell (state c in registers for all S steps), splits h into bf16 hi/lo and writes the h part
//   of A_{s+1}; it also LayerNorms its row of x_{s+1} into the x part, and stores step s-1's projected + residual
//   output.  generic-proxy writes are fenced (fence.proxy.async) before the barrier that precedes the next MMAs.
#include "sb_common.cuh"
#include "sb_lstm.cuh"

#ifndef SB_EMU
#include <cuda_bf16.h>
#endif

namespace sb {

#ifndef SB_EMU

namespace tc {

constexpr int kRows = 128, kC = 32, kH = 64, kK = kC + kH, kN = 4 * kH;
constexpr int kAChunkBytes = (kRows / 8) * 128;            // LBO of A: 2048
constexpr int kWChunkBytes = (kN / 8) * 128;               // LBO of the gate matrix: 4096
constexpr int kPChunkBytes = (kC / 8) * 128;               // LBO of the projection: 512
constexpr int kABytes = kRows * kK * 2, kWBytes = kN * kK * 2, kPBytes = kC * kH * 2;
constexpr int kStageStride = kH + 4;                       // floats per row of the (h | c) staging tile: conflict-free LDS.128
constexpr int kStageBytes = kRows * kStageStride * 4;
constexpr int kSlabStride = kC + 4;                         // floats per row of the x' slab of one step: conflict-free LDS.128
constexpr int kSlabBytes = kRows * kSlabStride * 4;
constexpr int kSmemBytes = 2 * kWBytes + 2 * kPBytes + 2 * kABytes + kN * 4 + 64 + 2 * kSlabBytes;  // + bias + barriers + in/out slabs
static_assert(kStageBytes <= 2 * kABytes && 2 * kStageBytes <= 2 * kWBytes, "state staging tiles alias the operand images");
constexpr uint32_t kTmemCols = 512;                        // 256 gate columns + 32 projection columns -> next power of 2

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    uint64_t d = 0;                                         // SmemDescriptor: no swizzle, K-major
    d |= (uint64_t)((saddr & 0x3FFFFu) >> 4);               // start address, 16-byte units
    d |= (uint64_t)(lbo_bytes >> 4) << 16;                  // leading byte offset: between the two 8-element k chunks
    d |= (uint64_t)(sbo_bytes >> 4) << 32;                  // stride byte offset: between 8-row groups
    d |= (uint64_t)1 << 46;                                 // descriptor version (Blackwell)
    return d;
}
__host__ __device__ constexpr uint32_t make_idesc(int M, int N) {
    return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);   // f32 acc, bf16 x bf16
}
__device__ __forceinline__ void umma(uint32_t tmem_c, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(tmem_c), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
// bounded wait: a wrong descriptor must end in a trap, not in a hung GPU
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    const uint32_t a = smem_u32(bar);
    for (long long spin = 0; spin < (1ll << 26); ++spin) {
        uint32_t ok;
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}\n"
                     : "=r"(ok) : "r"(a), "r"(parity) : "memory");
        if (ok) return;
    }
    asm volatile("trap;");
}
__device__ __forceinline__ void fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tmem_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

__device__ __forceinline__ void tmem_ld32(uint32_t taddr, float (&v)[32]) {
    uint32_t r[32];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];\n"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
          "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]),
          "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]),
          "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr)
        : "memory");
    tmem_wait_ld();
#pragma unroll
    for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float (&v)[16]) {
    uint32_t r[16];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];\n"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
          "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr)
        : "memory");
    tmem_wait_ld();
#pragma unroll
    for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}

// 8 consecutive k of one row -> one 16-byte core-matrix row in the hi image and one in the lo image
__device__ __forceinline__ void store_split8(unsigned char* a_hi, unsigned char* a_lo, int row, int chunk, const float (&v)[8]) {
    uint32_t hi[4], lo[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const __nv_bfloat162 h2 = __floats2bfloat162_rn(v[2 * i], v[2 * i + 1]);
        const float r0 = v[2 * i] - __bfloat162float(h2.x), r1 = v[2 * i + 1] - __bfloat162float(h2.y);
        const __nv_bfloat162 l2 = __floats2bfloat162_rn(r0, r1);
        hi[i] = *reinterpret_cast<const uint32_t*>(&h2);
        lo[i] = *reinterpret_cast<const uint32_t*>(&l2);
    }
    const int off = ((chunk * (kRows / 8) + (row >> 3)) * 8 + (row & 7)) * 16;
    *reinterpret_cast<uint4*>(a_hi + off) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
    *reinterpret_cast<uint4*>(a_lo + off) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
}

}  // namespace tc

__global__ void __launch_bounds__(256, 1) lstm_tc_kernel(const SeqArgs a) {
    using namespace tc;
    extern __shared__ __align__(128) unsigned char sm[];
    unsigned char* w_hi = sm;                               // gate matrix images
    unsigned char* w_lo = w_hi + kWBytes;
    unsigned char* p_hi = w_lo + kWBytes;                   // projection images
    unsigned char* p_lo = p_hi + kPBytes;
    unsigned char* a_hi = p_lo + kPBytes;                   // A operand images
    unsigned char* a_lo = a_hi + kABytes;
    float* bias_s = reinterpret_cast<float*>(a_lo + kABytes);
    uint64_t* mbar = reinterpret_cast<uint64_t*>(bias_s + kN);         // [2]: units 0..31 (+ projection), units 32..63
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(mbar + 3);
    // [kRows][kStageStride] staging tiles carry the state between its global layout (row-contiguous, coalesced) and the
    // (row, half) threads.  They alias operand images that are idle at that moment (a bigger carve-out would cost L1):
    float* slab = reinterpret_cast<float*>(mbar + 8);       // [kRows][kSlabStride]: x' of the next step, loaded coalesced
    float* oslab = slab + kRows * kSlabStride;              // [kRows][kSlabStride]: output of the previous step, stored coalesced
    float* stage_in_buf = reinterpret_cast<float*>(a_hi);   // prologue: before anything is written into A
    float* stage_h = reinterpret_cast<float*>(w_hi);        // epilogue: after the last MMA has completed
    float* stage_c = stage_h + kRows * kStageStride;

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int dir = blockIdx.y;
    const sb_lstm_dir& w = a.w[dir];
    const int S = a.n_steps;
    const int q = warp & 3, hf = warp >> 2;                 // TMEM lane quarter, unit half
    const int r = 32 * q + lane;                            // row of the tile

    // ---- constant operands: async copies of the packed images (hi/lo gate matrix, hi/lo projection), bias ----------
    {
        for (int i = tid; i < kN; i += 256) bias_s[i] = __ldg(w.tc_b + i);
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(kTmemCols) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    BulkBarrier* wbar = reinterpret_cast<BulkBarrier*>(mbar + 2);
    if (tid == 0) {
        mbar_init(mbar + 0, 1);
        mbar_init(mbar + 1, 1);
        bulk_barrier_init(wbar);
        // the packed operand images (hi / lo gate matrix, hi / lo projection: 106 KB) in one TMA bulk copy
        bulk_expect(wbar, 2 * kWBytes + 2 * kPBytes);
        bulk_copy_g2s(reinterpret_cast<float*>(sm), w.tc_w, 2 * kWBytes + 2 * kPBytes, wbar);
    }
    pdl_trigger();
    pdl_wait();

    // ---- per-thread state -----------------------------------------------------------------------------------------
    float c[32];
    // The tile's rows are consecutive in [n_rows][H]: the CTA moves them as one contiguous block (coalesced float4) through
    // the staging tile; thread (row, hf) then owns units 32*hf .. 32*hf + 31 of its row.
    const int tile_row0 = blockIdx.x * kRows;
    auto stage_out = [&](const float* stage, float* dst) {
        for (int i = tid; i < kRows * (kH / 4); i += 256) {
            const int rr = i / (kH / 4), c4 = i % (kH / 4);
            if (tile_row0 + rr < a.n_rows)
                st4(dst + (long long)(tile_row0 + rr) * kH + 4 * c4, ld4(stage + rr * kStageStride + 4 * c4));
        }
    };
    if (a.h0) {
        // both tiles in flight at once (plain loads: hN / cN may alias h0 / c0), then through the one staging tile in turn
        float4 cin[8], hin[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            const int i = tid + 256 * k, rr = i / (kH / 4), c4 = i % (kH / 4);
            const long long off = (long long)min(tile_row0 + rr, a.n_rows - 1) * kH + 4 * c4;
            cin[k] = ld_plain4(a.c0 + off);
            hin[k] = ld_plain4(a.h0 + off);
        }
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            const int i = tid + 256 * k;
            st4(stage_in_buf + (i / (kH / 4)) * kStageStride + 4 * (i % (kH / 4)), cin[k]);
        }
        __syncthreads();
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const float4 v = ld4(stage_in_buf + r * kStageStride + 32 * hf + 4 * i);
            c[4 * i] = v.x; c[4 * i + 1] = v.y; c[4 * i + 2] = v.z; c[4 * i + 3] = v.w;
        }
        __syncthreads();
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            const int i = tid + 256 * k;
            st4(stage_in_buf + (i / (kH / 4)) * kStageStride + 4 * (i % (kH / 4)), hin[k]);
        }
        __syncthreads();
        float4 hv[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) hv[i] = ld4(stage_in_buf + r * kStageStride + 32 * hf + 4 * i);
        __syncthreads();                                    // the staging tile is dead: A may be written
#pragma unroll
        for (int ch = 0; ch < 4; ++ch) {
            const float4 v0 = hv[2 * ch], v1 = hv[2 * ch + 1];
            const float h8[8] = {v0.x, v0.y, v0.z, v0.w, v1.x, v1.y, v1.z, v1.w};
            store_split8(a_hi, a_lo, r, 4 + 4 * hf + ch, h8);
        }
    } else {
#pragma unroll
        for (int j = 0; j < 32; ++j) c[j] = 0.0f;
        const float h8[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
#pragma unroll
        for (int ch = 0; ch < 4; ++ch) store_split8(a_hi, a_lo, r, 4 + 4 * hf + ch, h8);
    }
    float4 g4[4], b4[4];                                    // LayerNorm gain / bias of this thread's 16 channels
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        g4[i] = __ldg(reinterpret_cast<const float4*>(w.ln_g) + 4 * hf + i);
        b4[i] = __ldg(reinterpret_cast<const float4*>(w.ln_b) + 4 * hf + i);
    }
    float4 blin[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) blin[i] = dir == 0 ? __ldg(reinterpret_cast<const float4*>(w.lin_b) + 4 * hf + i) : make_float4(0, 0, 0, 0);

    // x' of one step for the whole tile, COALESCED: item i = tid + 256 k is (row i / 8, channel quad i % 8), so a warp
    // instruction touches 4 rows x 128 bytes instead of 32 rows x 16 bytes.  FiLM and the second addend are applied on
    // arrival; the slab then goes through shared memory to the (row, half) threads, which need whole rows for LayerNorm.
    long long sbase[4], sfilm[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const int i = tid + 256 * k, rr = i >> 3, c4 = i & 7;
        const int gr = min(tile_row0 + rr, a.n_rows - 1);
        sbase[k] = row_base(a, gr) + 4 * c4;
        sfilm[k] = (long long)(gr / a.film_row_div) * S * kC + 4 * c4;
    }
    auto load_slab = [&](int step, float4 (&xs)[4]) {
        const int pos = dir ? S - 1 - step : step;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const long long off = sbase[k] + (long long)pos * a.stride_pos;
            float4 t = ldg4_stream(a.x0 + off);
            if (a.x1) {
                const float4 u4 = ldg4_stream(a.x1 + off);
                t.x += u4.x; t.y += u4.y; t.z += u4.z; t.w += u4.w;
            }
            if (a.film_scale) {
                const float4 fs = __ldg(reinterpret_cast<const float4*>(a.film_scale + sfilm[k] + (long long)pos * kC));
                const float4 fb = __ldg(reinterpret_cast<const float4*>(a.film_shift + sfilm[k] + (long long)pos * kC));
                t.x = fmaf(t.x, fs.x, fb.x); t.y = fmaf(t.y, fs.y, fb.y); t.z = fmaf(t.z, fs.z, fb.z); t.w = fmaf(t.w, fs.w, fb.w);
            }
            xs[k] = t;
        }
    };
    auto store_slab = [&](const float4 (&xs)[4]) {
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const int i = tid + 256 * k;
            st4(slab + (i >> 3) * kSlabStride + 4 * (i & 7), xs[k]);
        }
    };
    auto read_row = [&](float4 (&xv)[8]) {
#pragma unroll
        for (int i = 0; i < 8; ++i) xv[i] = ld4(slab + r * kSlabStride + 4 * i);
    };
    // LayerNorm the row, write this thread's 16 channels (k = 16*hf ..) into the x part of A, keep them for the residual
    auto ln_store = [&](const float4 (&xv)[8], float4 (&keep)[4]) {
        float s1 = 0.f;
#pragma unroll
        for (int i = 0; i < 8; ++i) s1 += (xv[i].x + xv[i].y) + (xv[i].z + xv[i].w);
        const float mean = s1 * (1.0f / kC);
        float s2 = 0.f;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const float dx = xv[i].x - mean, dy = xv[i].y - mean, dz = xv[i].z - mean, dw = xv[i].w - mean;
            s2 += (dx * dx + dy * dy) + (dz * dz + dw * dw);
        }
        const float rstd = rsqrtf(s2 * (1.0f / kC) + kLnEps);
#pragma unroll
        for (int ch = 0; ch < 2; ++ch) {
            float v[8];
#pragma unroll
            for (int i = 0; i < 2; ++i) {
                const float4 t = xv[4 * hf + 2 * ch + i];
                const float4 g = g4[2 * ch + i], b = b4[2 * ch + i];
                v[4 * i + 0] = fmaf((t.x - mean) * rstd, g.x, b.x); v[4 * i + 1] = fmaf((t.y - mean) * rstd, g.y, b.y);
                v[4 * i + 2] = fmaf((t.z - mean) * rstd, g.z, b.z); v[4 * i + 3] = fmaf((t.w - mean) * rstd, g.w, b.w);
            }
            store_split8(a_hi, a_lo, r, 2 * hf + ch, v);
        }
#pragma unroll
        for (int i = 0; i < 4; ++i) keep[i] = xv[4 * hf + i];
    };

    float4 res_cur[4], res_prev[4];                         // x' channels 16*hf .. of step s and s-1 (residual)
    {
        float4 xs[4], xv[8];
        load_slab(0, xs);
        store_slab(xs);
        __syncthreads();
        read_row(xv);
        ln_store(xv, res_cur);
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) res_prev[i] = res_cur[i];
    bulk_wait(wbar, 0);                                     // the operand images have landed (async proxy wrote them)
    fence_async_smem();
    fence_before();
    __syncthreads();
    fence_after();
    const uint32_t tmem = *tmem_slot;

    // ---- MMA issue (one thread) -------------------------------------------------------------------------------------
    const uint32_t a_hi_s = smem_u32(a_hi), a_lo_s = smem_u32(a_lo), w_hi_s = smem_u32(w_hi), w_lo_s = smem_u32(w_lo);
    const uint32_t p_hi_s = smem_u32(p_hi), p_lo_s = smem_u32(p_lo);
    constexpr uint32_t idesc_g = make_idesc(128, 128), idesc_p = make_idesc(128, 32);
    auto issue_proj = [&]() {                               // proj[128 x 32] = h_{s-1} (k chunks 4..11 of A) . lin^T
        uint32_t acc = 0;
#pragma unroll
        for (int pass = 0; pass < 3; ++pass) {
            const uint32_t ab = pass == 2 ? a_lo_s : a_hi_s, pb = pass == 1 ? p_lo_s : p_hi_s;
#pragma unroll
            for (int ks = 0; ks < kH / 16; ++ks) {
                umma(tmem + 256, make_desc(ab + (4 + 2 * ks) * kAChunkBytes, kAChunkBytes, 128),
                     make_desc(pb + 2 * ks * kPChunkBytes, kPChunkBytes, 128), idesc_p, acc);
                acc = 1;
            }
        }
    };
    auto issue_gates = [&](int half) {                      // gates[128 x 128] for units 32*half .. : rows n of the image
        uint32_t acc = 0;
#pragma unroll
        for (int pass = 0; pass < 3; ++pass) {
            const uint32_t ab = pass == 2 ? a_lo_s : a_hi_s, wb = (pass == 1 ? w_lo_s : w_hi_s) + half * 16 * 128;
#pragma unroll
            for (int ks = 0; ks < kK / 16; ++ks) {
                umma(tmem + 128 * half, make_desc(ab + 2 * ks * kAChunkBytes, kAChunkBytes, 128),
                     make_desc(wb + 2 * ks * kWChunkBytes, kWChunkBytes, 128), idesc_g, acc);
                acc = 1;
            }
        }
    };

    const uint32_t lane_base = (uint32_t)(32 * q) << 16;    // TMEM address = lane << 16 | column
    float* const outp = a.out[dir];
    // projected + bias + residual of one step: 16 channels per thread into the output slab; flush_out() then stores the
    // slab coalesced (a warp instruction writes 4 rows x 128 bytes) once a barrier has published it
    auto emit = [&](const float4 (&resv)[4]) {
        float pv[16];
        tmem_ld16(tmem + lane_base + 256 + 16 * hf, pv);
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            float4 o = make_float4(pv[4 * i], pv[4 * i + 1], pv[4 * i + 2], pv[4 * i + 3]);
            if (dir == 0) {
                o.x += blin[i].x + resv[i].x; o.y += blin[i].y + resv[i].y;
                o.z += blin[i].z + resv[i].z; o.w += blin[i].w + resv[i].w;
            }
            st4(oslab + r * kSlabStride + 16 * hf + 4 * i, o);
        }
    };
    auto flush_out = [&](int step) {
        const int pos = dir ? S - 1 - step : step;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const int i = tid + 256 * k, rr = i >> 3;
            if (tile_row0 + rr < a.n_rows)
                st4(outp + sbase[k] + (long long)pos * a.stride_pos, ld4(oslab + rr * kSlabStride + 4 * (i & 7)));
        }
    };

    float hlast[32];
    for (int s = 0; s < S; ++s) {
        if (tid == 0) {
            fence_after();
            if (s > 0) issue_proj();
            issue_gates(0);
            umma_commit(mbar + 0);
            issue_gates(1);
            umma_commit(mbar + 1);
        }
        float4 xnext[4];
        if (s + 1 < S) load_slab(s + 1, xnext);             // in flight while the tensor pipe works
        mbar_wait(mbar + hf, s & 1);
        fence_after();
        if (s > 0) emit(res_prev);
#pragma unroll
        for (int ch = 0; ch < 4; ++ch) {                    // 8 units = 32 TMEM columns at a time
            float gv[32];
            tmem_ld32(tmem + lane_base + 128 * hf + 32 * ch, gv);
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const float4 bb = ld4(bias_s + 4 * (32 * hf + 8 * ch + j));
                const float ig = sigmoid_f(gv[4 * j + 0] + bb.x), fg = sigmoid_f(gv[4 * j + 1] + bb.y);
                const float gg = tanh_f(gv[4 * j + 2] + bb.z), og = sigmoid_f(gv[4 * j + 3] + bb.w);
                c[8 * ch + j] = fmaf(fg, c[8 * ch + j], ig * gg);
                hlast[8 * ch + j] = og * tanh_f(c[8 * ch + j]);
            }
        }
        // The first half's warps got here while the second half's MMAs may still be READING A: nothing may be written
        // into A (h part or x part) before mbar[1] has fired.
        if (hf == 0) mbar_wait(mbar + 1, s & 1);
#pragma unroll
        for (int ch = 0; ch < 4; ++ch) {
            float h8[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) h8[j] = hlast[8 * ch + j];
            store_split8(a_hi, a_lo, r, 4 + 4 * hf + ch, h8);
        }
#pragma unroll
        for (int i = 0; i < 4; ++i) res_prev[i] = res_cur[i];
        if (s + 1 < S) store_slab(xnext);                   // the slab's previous readers finished before the last barrier
        __syncthreads();                                    // publishes both slabs
        if (s > 0) flush_out(s - 1);
        if (s + 1 < S) {
            float4 xv[8];
            read_row(xv);
            ln_store(xv, res_cur);
        }
        fence_async_smem();
        fence_before();
        __syncthreads();
    }
    // ---- drain: projection of the last step ---------------------------------------------------------------------------
    if (tid == 0) {
        fence_after();
        issue_proj();
        umma_commit(mbar + 0);
    }
    mbar_wait(mbar + 0, S & 1);
    fence_after();
    emit(res_prev);
    __syncthreads();
    flush_out(S - 1);
    if (a.hN) {                                             // same staging tiles, the other way (the initial-state reads
                                                            // were over before the first step's barrier)
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            st4(stage_h + r * kStageStride + 32 * hf + 4 * i, make_float4(hlast[4 * i], hlast[4 * i + 1], hlast[4 * i + 2], hlast[4 * i + 3]));
            st4(stage_c + r * kStageStride + 32 * hf + 4 * i, make_float4(c[4 * i], c[4 * i + 1], c[4 * i + 2], c[4 * i + 3]));
        }
        __syncthreads();
        stage_out(stage_h, a.hN);
        stage_out(stage_c, a.cN);
    }
    fence_before();
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(kTmemCols) : "memory");
}

int run_seq_tc(const SeqArgs& a, cudaStream_t st) {
    dim3 grid(ceil_div(a.n_rows, tc::kRows), a.n_dirs);
    return launch("lstm_tc", lstm_tc_kernel, grid, dim3(256), (size_t)tc::kSmemBytes, st, a);
}

#else   // SB_EMU: tensor-core instructions cannot be emulated on the host

int run_seq_tc(const SeqArgs&, cudaStream_t) {
    set_error("SB_ALGO_TC (tcgen05) is not available in the host-emulated test build");
    return SB_E_UNSUPP;
}

#endif

}  // namespace sb
