// Sliding-window attention core on the tensor cores (tcgen05 / TMEM) for whole-utterance and time-slice calls (T >= 64).
//
// Same arithmetic as attn_core_kernel (sb_attn.cu): for every frame softmax(q K_win^T / sqrt(F*E)) V_win over the last W
// frames, the zero-initialised history NOT masked (DE3:722-744, 875-888).  Only the two contractions run on the tensor
// cores; softmax, masking and normalisation stay in fp32 registers.
//
// CTA = (tile of 128 queries t0 .. t0+127, head-row bl).  Query t0+r attends the concatenated key rows t0+r .. t0+r+W-1, so
// the tile needs the 256 key rows t0 .. t0+255 (W <= 129): a banded 128 x 256 score matrix.
//   phase 1   S[128 x 256] = Q_tile K_tile^T over DK (<= 320) in chunks of 64: both operands are row-major with the
//             contraction index contiguous, i.e. K-major UMMA operands after the fp32 -> (bf16 hi, bf16 lo) split; four
//             MMAs per k-step (hi*hi + hi*lo + lo*hi + lo*lo, fp32 accumulation in TMEM columns 0..255).
//   phase 2   thread (row, column half) reads its scores from TMEM, applies the band mask and the scale, computes max and
//             exp (two passes over TMEM instead of 128 live registers), and writes the UNNORMALISED probabilities as the
//             K-major A operand of phase 3 (hi / lo images, 8 keys = one 16-byte core-matrix row per store).
//   phase 3   O[128 x DV] = P V in N-tiles of 256 columns and key chunks of 64: V^T is the K-major B operand, built on the
//             fly (a thread reads 8 consecutive key rows of one column: coalesced across the warp, one 16-byte row per
//             image), accumulators in TMEM columns 256..511; the epilogue scales by 1 / rowsum and stores fp32.
// Operand staging is double-buffered (round 2): eight warps convert chunk g + 1 into the second pair of images while a ninth
// warp, which does nothing else, has the MMAs of chunk g in flight; `staged[b]` (8 warp arrivals) hands a buffer to the issue
// warp, `done[b]` (tcgen05.commit) hands it back.  Round 1 staged, issued and waited chunk by chunk (tensor pipe 12 % active,
// long scoreboard 64 % of the samples: profiles/r01_prof_attn_tc.txt).
// Operand layouts and descriptors are those of lstm_tc_kernel.
#include "sb_common.cuh"

#ifndef SB_EMU
#include <cuda_bf16.h>
#endif

namespace sb {

#ifndef SB_EMU
namespace atc {

constexpr int kMQ = 128, kNK = 256, kDKC = 64, kKC = 32, kNT = 256;      // kKC = 32: two V^T chunk buffers fit the 64 KB one had
constexpr int kThreads = 256 + 32;                          // 8 staging / softmax / read-back warps + the MMA-issue warp
constexpr int kAChunk = (kMQ / 8) * 128;                    // bytes between 8-element k chunks of a 128-row A image: 2048
constexpr int kBChunk = (kNK / 8) * 128;                    // ... of a 256-row B image: 4096
constexpr int kPBytes = kMQ * kNK * 2;                      // one P image (128 x 256 bf16): 65536
constexpr int kVBytes = kNT * kKC * 2;                      // one V^T chunk image (256 x 32 bf16): 16384
constexpr int kKBytes = kNK * kDKC * 2, kQBytes = kMQ * kDKC * 2;      // one K / Q chunk image of phase 1: 32768 / 16384
constexpr int kSmemBytes = 2 * kPBytes + 4 * kVBytes + 2 * kMQ * 4 * 2 + 64;     // + (max, sum) exchange + barriers
constexpr uint32_t kTmemCols = 512;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    uint64_t d = 0;                                         // no swizzle, K-major
    d |= (uint64_t)((saddr & 0x3FFFFu) >> 4);
    d |= (uint64_t)(lbo_bytes >> 4) << 16;
    d |= (uint64_t)(sbo_bytes >> 4) << 32;
    d |= (uint64_t)1 << 46;
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
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void bar_sync_256() { asm volatile("bar.sync 1, 256;" ::: "memory"); }     // the eight working warps
__device__ __forceinline__ bool elect_one() {
    uint32_t pred;
    asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}\n" : "=r"(pred));
    return pred != 0;
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {      // bounded: a bad descriptor must trap, not hang
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
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
}

// 8 consecutive k of one row -> one 16-byte core-matrix row in the hi image and one in the lo image.
// image layout: [k / 8][rows / 8][8 rows][8 k] bf16, rows_per_image = 128 (A) or 256 (B)
template <int ROWS>
__device__ __forceinline__ void store_split8(unsigned char* hi_img, unsigned char* lo_img, int row, int kchunk, const float (&v)[8]) {
    uint32_t hi[4], lo[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const __nv_bfloat162 h2 = __floats2bfloat162_rn(v[2 * i], v[2 * i + 1]);
        const float r0 = v[2 * i] - __bfloat162float(h2.x), r1 = v[2 * i + 1] - __bfloat162float(h2.y);
        const __nv_bfloat162 l2 = __floats2bfloat162_rn(r0, r1);
        hi[i] = *reinterpret_cast<const uint32_t*>(&h2);
        lo[i] = *reinterpret_cast<const uint32_t*>(&l2);
    }
    const int off = ((kchunk * (ROWS / 8) + (row >> 3)) * 8 + (row & 7)) * 16;
    *reinterpret_cast<uint4*>(hi_img + off) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
    *reinterpret_cast<uint4*>(lo_img + off) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
}

}  // namespace atc

__global__ void __launch_bounds__(atc::kThreads, 1) attn_core_tc_kernel(const sb_attn_args a, const float* Q, const float* Kc, const float* Vc,
                                                                       float* AO) {
    using namespace atc;
    extern __shared__ __align__(128) unsigned char sm[];
    unsigned char* p_hi = sm;                               // phase 2/3: P images.  Phase 1: Q images of both buffers (4 x 16 KB)
    unsigned char* p_lo = p_hi + kPBytes;                   //                       phase 1: K images of buffer 1 (2 x 32 KB)
    unsigned char* v_reg = p_lo + kPBytes;                  // phase 3: V^T images of both buffers (4 x 16 KB); phase 1: K images of buffer 0
    float* xmax = reinterpret_cast<float*>(v_reg + 4 * kVBytes);     // [2][128] partial row maxima of the two column halves
    float* xsum = xmax + 2 * kMQ;                           // [2][128] partial row sums
    uint64_t* staged = reinterpret_cast<uint64_t*>(xsum + 2 * kMQ);  // [2] the images of a buffer are complete (8 warps)
    uint64_t* done = staged + 2;                            // [2] tcgen05.commit: the MMAs that read a buffer have executed
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(done + 2);

    const int F = a.F, L = a.L, E = a.E, Vd = a.C / L, W = a.W, T = a.T;
    const int DK = F * E, DV = F * Vd, TT = T + W - 1;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int q4 = warp & 3, hh = (warp >> 2) & 1;          // TMEM lane quarter, column half
    const int r = 32 * q4 + lane;                           // query row of the tile
    const int t0 = blockIdx.x * kMQ, bl = blockIdx.y;
    const int nq = min(kMQ, T - t0);
    const float* qbase = Q + ((size_t)bl * T + t0) * DK;
    const float* kbase = Kc + ((size_t)bl * TT + t0) * DK;
    const float* vbase = Vc + ((size_t)bl * TT + t0) * DV;
    const int nkeys = min(kNK, TT - t0);                    // key rows of the tile that exist

    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(kTmemCols) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    if (tid == 0) {
        for (int b = 0; b < 2; ++b) { mbar_init(staged + b, 8); mbar_init(done + b, 1); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    pdl_trigger();
    pdl_wait();
    fence_before();
    __syncthreads();
    fence_after();
    const uint32_t tmem = *tmem_slot;
    const uint32_t lane_base = (uint32_t)(32 * q4) << 16;
    constexpr uint32_t idesc_s = make_idesc(kMQ, kNK);

    const int n_dkc = (DK + kDKC - 1) / kDKC;               // chunks of phase 1
    const int DVp = (DV + 15) & ~15;
    const int n_nt = (DVp + kNT - 1) / kNT;                 // N-tiles of phase 3, kNK / kKC chunks each
    // buffer b of chunk number g (counted through both phases): b = g & 1, its (g >> 1)-th use
    auto q_img = [&](int b, int lo) { return p_hi + (2 * b + lo) * kQBytes; };
    auto k_img = [&](int b, int lo) { return (b ? p_lo : v_reg) + lo * kKBytes; };
    auto v_img = [&](int b, int lo) { return v_reg + (2 * b + lo) * kVBytes; };

    if (warp == 8) {
        // ============================================================================================ MMA issue warp
        if (elect_one()) {
            int g = 0;
            for (int c = 0; c < n_dkc; ++c, ++g) {          // phase 1: S = Q K^T
                const int b = g & 1;
                mbar_wait(staged + b, (uint32_t)((g >> 1) & 1));
                fence_after();
#pragma unroll
                for (int pass = 0; pass < 4; ++pass) {      // hi*hi + hi*lo + lo*hi + lo*lo
                    const uint32_t ab = smem_u32(q_img(b, pass >= 2)), bb = smem_u32(k_img(b, pass & 1));
#pragma unroll
                    for (int ks = 0; ks < kDKC / 16; ++ks)
                        umma(tmem, make_desc(ab + 2 * ks * kAChunk, kAChunk, 128), make_desc(bb + 2 * ks * kBChunk, kBChunk, 128), idesc_s,
                             (c > 0 || pass > 0 || ks > 0) ? 1u : 0u);
                }
                umma_commit(done + b);
            }
            for (int nt = 0; nt < n_nt; ++nt) {             // phase 3: O = P V (the first `staged` of it also covers the P images)
                const int nw = min(kNT, DVp - kNT * nt);
                const uint32_t idesc_o = make_idesc(kMQ, nw);
                for (int kc = 0; kc < kNK / kKC; ++kc, ++g) {
                    const int b = g & 1;
                    mbar_wait(staged + b, (uint32_t)((g >> 1) & 1));
                    fence_after();
#pragma unroll
                    for (int pass = 0; pass < 4; ++pass) {
                        const uint32_t ab = smem_u32(pass >= 2 ? p_lo : p_hi) + (kKC / 8) * kc * kAChunk;
                        const uint32_t bb = smem_u32(v_img(b, pass & 1));
#pragma unroll
                        for (int ks = 0; ks < kKC / 16; ++ks)
                            umma(tmem + 256, make_desc(ab + 2 * ks * kAChunk, kAChunk, 128), make_desc(bb + 2 * ks * kBChunk, kBChunk, 128),
                                 idesc_o, (kc > 0 || pass > 0 || ks > 0) ? 1u : 0u);
                    }
                    umma_commit(done + b);
                }
            }
        }
        __syncwarp();
    } else {
        // ======================================================================== staging / softmax / read-back warps
        int g = 0;
        auto acquire = [&](int gg) {                        // the MMAs of chunk gg - 2 have read buffer gg & 1
            if (gg >= 2) {
                mbar_wait(done + (gg & 1), (uint32_t)(((gg >> 1) - 1) & 1));
                fence_after();
            }
        };
        auto release = [&](int gg) {                        // images of chunk gg complete: hand the buffer to the issue warp
            fence_async_smem();
            fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(staged + (gg & 1));
        };
        auto drain = [&](int gg) {                          // every chunk < gg has executed
            for (int k = gg - 2; k < gg; ++k)
                if (k >= 0) mbar_wait(done + (k & 1), (uint32_t)((k >> 1) & 1));
            fence_after();
        };
        // ------------------------------------------------------------------------------------------------ phase 1: S = Q K^T
        for (int c = 0; c < n_dkc; ++c, ++g) {
            const int b = g & 1;
            acquire(g);
            unsigned char* qh = q_img(b, 0); unsigned char* ql = q_img(b, 1);
            unsigned char* kh = k_img(b, 0); unsigned char* kl = k_img(b, 1);
            // stage: task = (row, k-group of 8); lanes run over the eight k-groups of a row first, so a warp reads four rows x
            // 256 contiguous bytes (with lanes over rows every load instruction touched 32 different sectors and this phase was
            // 30 % of the kernel's samples, all of them waiting for those loads)
#pragma unroll 6
            for (int i = tid; i < (kMQ + kNK) * 8; i += 256) {      // 12 tasks per thread: two rounds of 24 loads in flight
                const int gk = i & 7, row = i >> 3;
                const bool is_q = row < kMQ;
                const int rr = is_q ? row : row - kMQ;
                const int k0 = kDKC * c + 8 * gk;
                float v[8];
                const bool live = is_q ? rr < nq : rr < nkeys;
                const float* src = (is_q ? qbase : kbase) + (size_t)rr * DK + k0;
#pragma unroll
                for (int j = 0; j < 8; j += 2) {            // DK is even and k0 is a multiple of 8: float2 granularity
                    float2 t2 = make_float2(0.f, 0.f);
                    if (live && k0 + j < DK) t2 = ldg2_stream(src + j);
                    v[j] = t2.x; v[j + 1] = t2.y;
                }
                if (is_q) store_split8<kMQ>(qh, ql, rr, gk, v);
                else store_split8<kNK>(kh, kl, rr, gk, v);
            }
            release(g);
        }
        drain(g);                                           // S is complete; the phase-1 images may be overwritten by P

        // -------------------------------------------------------------------------------------- phase 2: banded softmax -> P
        const float scale = 1.0f / sqrtf((float)DK);
        const bool row_live = r < nq;
        float m = -INFINITY;
#pragma unroll 1
        for (int ch = 0; ch < 4; ++ch) {                    // pass A: row maximum over the band
            float sv[32];
            tmem_ld32(tmem + lane_base + 128 * hh + 32 * ch, sv);
#pragma unroll
            for (int j = 0; j < 32; ++j) {
                const int key = 128 * hh + 32 * ch + j;
                if (row_live && key >= r && key < r + W) m = fmaxf(m, sv[j] * scale);
            }
        }
        xmax[hh * kMQ + r] = m;
        bar_sync_256();
        m = fmaxf(xmax[r], xmax[kMQ + r]);
        float ssum = 0.f;
#pragma unroll 1
        for (int ch = 0; ch < 4; ++ch) {                    // pass B: exp, row sum, P images (zero outside the band)
            float sv[32];
            tmem_ld32(tmem + lane_base + 128 * hh + 32 * ch, sv);
#pragma unroll
            for (int gq = 0; gq < 4; ++gq) {
                float pv[8];
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    const int key = 128 * hh + 32 * ch + 8 * gq + j;
                    const bool in = row_live && key >= r && key < r + W;
                    pv[j] = in ? expf(sv[8 * gq + j] * scale - m) : 0.0f;
                    ssum += pv[j];
                }
                store_split8<kMQ>(p_hi, p_lo, r, (128 * hh + 32 * ch + 8 * gq) >> 3, pv);
            }
        }
        xsum[hh * kMQ + r] = ssum;
        fence_before();
        bar_sync_256();                                     // every P row and both partial sums are written
        fence_after();
        const float inv = row_live ? 1.0f / (xsum[r] + xsum[kMQ + r]) : 0.0f;

        // -------------------------------------------------------------------------------------------------- phase 3: O = P V
        // stage V^T: one task per thread = (four columns 4 n4 .. 4 n4 + 3, key-group of 8): eight 16-byte loads (DV % 4 == 0; a warp
        // reads 512 contiguous bytes of a key row), four core-matrix rows per image.  The loads of chunk ck + 2 are issued before
        // chunk ck is converted: the staging warps are the critical path of this phase and were waiting on their own loads.
        static_assert((kNT / 4) * (kKC / 8) == 256, "one task per staging thread");
        constexpr int kCk = kNK / kKC;                      // chunks per N-tile
        const int n4 = tid & (kNT / 4 - 1), gk = tid >> 6;
        auto load_chunk = [&](int ck, float4 (&t)[8]) {
            const int n0 = kNT * (ck / kCk), kc = ck % kCk;
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const int key = kKC * kc + 8 * gk + j;
                t[j] = (key < nkeys && n0 + 4 * n4 < DV) ? ldg4_stream(vbase + (size_t)key * DV + n0 + 4 * n4) : make_float4(0.f, 0.f, 0.f, 0.f);
            }
        };
        const int n_ck = n_nt * kCk;
        float4 cur[8], nx1[8], nx2[8];                      // the loads run two chunks ahead of their use
        load_chunk(0, cur);
        if (n_ck > 1) load_chunk(1, nx1);
        for (int ck = 0; ck < n_ck; ++ck, ++g) {
            const int nt = ck / kCk, kc = ck % kCk, b = g & 1;
            if (ck + 2 < n_ck) load_chunk(ck + 2, nx2);
            acquire(g);
            unsigned char* vh = v_img(b, 0); unsigned char* vl = v_img(b, 1);
            float v[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) v[j] = cur[j].x;
            store_split8<kNT>(vh, vl, 4 * n4 + 0, gk, v);
#pragma unroll
            for (int j = 0; j < 8; ++j) v[j] = cur[j].y;
            store_split8<kNT>(vh, vl, 4 * n4 + 1, gk, v);
#pragma unroll
            for (int j = 0; j < 8; ++j) v[j] = cur[j].z;
            store_split8<kNT>(vh, vl, 4 * n4 + 2, gk, v);
#pragma unroll
            for (int j = 0; j < 8; ++j) v[j] = cur[j].w;
            store_split8<kNT>(vh, vl, 4 * n4 + 3, gk, v);
            release(g);                                     // (the fence in it also publishes the P images to the tensor core)
#pragma unroll
            for (int j = 0; j < 8; ++j) { cur[j] = nx1[j]; nx1[j] = nx2[j]; }
            if (kc == kCk - 1) {
                drain(g + 1);                               // this N-tile is complete
                // read-back of this N-tile: thread (row, half) -> 128 columns
                const int n0 = kNT * nt;
                const int nw = min(kNT, DVp - n0);          // multiple of 16
                float* orow = AO + ((size_t)bl * T + t0 + r) * DV + n0 + 128 * hh;
#pragma unroll 1
                for (int ch = 0; ch < 4; ++ch) {
                    if (128 * hh + 32 * ch >= nw) break;    // warp-uniform
                    float ov[32];
                    tmem_ld32(tmem + lane_base + 256 + 128 * hh + 32 * ch, ov);
                    if (row_live) {
#pragma unroll
                        for (int j = 0; j < 32; j += 4) {
                            const int col = n0 + 128 * hh + 32 * ch + j;
                            if (col < DV) st4(orow + 32 * ch + j, make_float4(ov[j] * inv, ov[j + 1] * inv, ov[j + 2] * inv, ov[j + 3] * inv));
                        }
                    }
                }
                fence_before();                             // the next N-tile's first MMA overwrites these TMEM columns: it is
                bar_sync_256();                             // issued only after a `staged` that follows this barrier
                fence_after();
            }
        }
    }
    fence_before();
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(kTmemCols) : "memory");
}

// usable when the band fits one 256-key tile, DK fits five 64-wide chunks and the float2 / float4 accesses are aligned
bool attn_core_tc_supported(const sb_attn_args& a) {
    const int DK = a.F * a.E, DV = a.F * (a.C / a.L);
    return a.T >= 64 && a.W <= 129 && DK <= 320 && DK % 2 == 0 && DV % 4 == 0;
}

int attn_core_tc(const sb_attn_args& a, const float* Q, const float* Kc, const float* Vc, float* AO, cudaStream_t st) {
    dim3 grid(ceil_div(a.T, atc::kMQ), a.B * a.L);
    return launch("attn_core_tc", attn_core_tc_kernel, grid, dim3(atc::kThreads), (size_t)atc::kSmemBytes, st, a, Q, Kc, Vc, AO);
}

#else   // SB_EMU: tensor-core instructions cannot be emulated on the host

bool attn_core_tc_supported(const sb_attn_args&) { return false; }
int attn_core_tc(const sb_attn_args&, const float*, const float*, const float*, float*, cudaStream_t) {
    set_error("the tcgen05 attention core is not available in the host-emulated test build");
    return SB_E_UNSUPP;
}

#endif

}  // namespace sb
