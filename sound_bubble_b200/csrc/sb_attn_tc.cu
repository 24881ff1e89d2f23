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
// Operand layouts, descriptors and the single-thread issue / commit / mbarrier pattern are those of lstm_tc_kernel.
#include "sb_common.cuh"

#ifndef SB_EMU
#include <cuda_bf16.h>
#endif

namespace sb {

#ifndef SB_EMU
namespace atc {

constexpr int kMQ = 128, kNK = 256, kDKC = 64, kKC = 64, kNT = 256;
constexpr int kAChunk = (kMQ / 8) * 128;                    // bytes between 8-element k chunks of a 128-row A image: 2048
constexpr int kBChunk = (kNK / 8) * 128;                    // ... of a 256-row B image: 4096
constexpr int kPBytes = kMQ * kNK * 2;                      // one P image (128 x 256 bf16): 65536
constexpr int kVBytes = kNT * kKC * 2;                      // one V^T chunk image (256 x 64 bf16): 32768
constexpr int kSmemBytes = 2 * kPBytes + 2 * kVBytes + 2 * kMQ * 4 * 2 + 64;     // + (max, sum) exchange + barrier
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

__global__ void __launch_bounds__(256, 1) attn_core_tc_kernel(const sb_attn_args a, const float* Q, const float* Kc, const float* Vc,
                                                             float* AO) {
    using namespace atc;
    extern __shared__ __align__(128) unsigned char sm[];
    unsigned char* p_hi = sm;                               // phase 2/3: P images; phase 1: Q chunk images in their first 16 KB
    unsigned char* p_lo = p_hi + kPBytes;
    unsigned char* v_hi = p_lo + kPBytes;                   // phase 3: V^T chunk images; phase 1: K chunk images (same size)
    unsigned char* v_lo = v_hi + kVBytes;
    float* xmax = reinterpret_cast<float*>(v_lo + kVBytes); // [2][128] partial row maxima of the two column halves
    float* xsum = xmax + 2 * kMQ;                           // [2][128] partial row sums
    uint64_t* mbar = reinterpret_cast<uint64_t*>(xsum + 2 * kMQ);
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(mbar + 1);

    const int F = a.F, L = a.L, E = a.E, Vd = a.C / L, W = a.W, T = a.T;
    const int DK = F * E, DV = F * Vd, TT = T + W - 1;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int q4 = warp & 3, hh = warp >> 2;                // TMEM lane quarter, column half
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
        mbar_init(mbar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    pdl_trigger();
    pdl_wait();
    fence_before();
    __syncthreads();
    fence_after();
    const uint32_t tmem = *tmem_slot;
    const uint32_t lane_base = (uint32_t)(32 * q4) << 16;
    uint32_t phase = 0;
    constexpr uint32_t idesc_s = make_idesc(kMQ, kNK);

    // ---------------------------------------------------------------------------------------------------- phase 1: S = Q K^T
    unsigned char* q_hi = p_hi;                             // [8 k-chunks][16][8][8] = 16 KB each
    unsigned char* q_lo = p_hi + kMQ * kDKC * 2;
    const int n_dkc = (DK + kDKC - 1) / kDKC;
    for (int c = 0; c < n_dkc; ++c) {
        // stage: task = (row, k-group of 8); lanes run over rows (one 32-byte sector each from global, consecutive 16-byte
        // core-matrix rows in shared memory: no bank conflicts)
#pragma unroll 2
        for (int i = tid; i < (kMQ + kNK) * 8; i += 256) {
            const int row = i % (kMQ + kNK), g = i / (kMQ + kNK);
            const bool is_q = row < kMQ;
            const int rr = is_q ? row : row - kMQ;
            const int k0 = kDKC * c + 8 * g;
            float v[8];
            const bool live = is_q ? rr < nq : rr < nkeys;
            const float* src = (is_q ? qbase : kbase) + (size_t)rr * DK + k0;
#pragma unroll
            for (int j = 0; j < 8; j += 2) {                // DK is even and k0 is a multiple of 8: float2 granularity
                float2 t2 = make_float2(0.f, 0.f);
                if (live && k0 + j < DK) t2 = ldg2_stream(src + j);
                v[j] = t2.x; v[j + 1] = t2.y;
            }
            if (is_q) store_split8<kMQ>(q_hi, q_lo, rr, g, v);
            else store_split8<kNK>(v_hi, v_lo, rr, g, v);
        }
        fence_async_smem();
        fence_before();
        __syncthreads();
        if (tid == 0) {
            fence_after();
#pragma unroll
            for (int pass = 0; pass < 4; ++pass) {          // hi*hi + hi*lo + lo*hi + lo*lo
                const uint32_t ab = smem_u32(pass >= 2 ? q_lo : q_hi), bb = smem_u32((pass & 1) ? v_lo : v_hi);
#pragma unroll
                for (int ks = 0; ks < kDKC / 16; ++ks)
                    umma(tmem, make_desc(ab + 2 * ks * kAChunk, kAChunk, 128), make_desc(bb + 2 * ks * kBChunk, kBChunk, 128), idesc_s,
                         (c > 0 || pass > 0 || ks > 0) ? 1u : 0u);
            }
            umma_commit(mbar);
        }
        mbar_wait(mbar, phase & 1);
        ++phase;
        fence_after();
    }

    // ------------------------------------------------------------------------------------------ phase 2: banded softmax -> P
    const float scale = 1.0f / sqrtf((float)DK);
    const bool row_live = r < nq;
    float m = -INFINITY;
#pragma unroll 1
    for (int ch = 0; ch < 4; ++ch) {                        // pass A: row maximum over the band
        float sv[32];
        tmem_ld32(tmem + lane_base + 128 * hh + 32 * ch, sv);
#pragma unroll
        for (int j = 0; j < 32; ++j) {
            const int key = 128 * hh + 32 * ch + j;
            if (row_live && key >= r && key < r + W) m = fmaxf(m, sv[j] * scale);
        }
    }
    xmax[hh * kMQ + r] = m;
    __syncthreads();
    m = fmaxf(xmax[r], xmax[kMQ + r]);
    float ssum = 0.f;
#pragma unroll 1
    for (int ch = 0; ch < 4; ++ch) {                        // pass B: exp, row sum, P images (zero outside the band)
        float sv[32];
        tmem_ld32(tmem + lane_base + 128 * hh + 32 * ch, sv);
#pragma unroll
        for (int g = 0; g < 4; ++g) {
            float pv[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const int key = 128 * hh + 32 * ch + 8 * g + j;
                const bool in = row_live && key >= r && key < r + W;
                pv[j] = in ? expf(sv[8 * g + j] * scale - m) : 0.0f;
                ssum += pv[j];
            }
            store_split8<kMQ>(p_hi, p_lo, r, (128 * hh + 32 * ch + 8 * g) >> 3, pv);
        }
    }
    xsum[hh * kMQ + r] = ssum;
    fence_before();                                         // the TMEM reads above precede the MMAs that overwrite nothing of S,
    __syncthreads();                                        // but order them anyway before phase 3 is issued
    fence_after();
    const float inv = row_live ? 1.0f / (xsum[r] + xsum[kMQ + r]) : 0.0f;

    // ------------------------------------------------------------------------------------------------------ phase 3: O = P V
    const int DVp = (DV + 15) & ~15;
    const int n_nt = (DVp + kNT - 1) / kNT;
    for (int nt = 0; nt < n_nt; ++nt) {
        const int n0 = kNT * nt;
        const int nw = min(kNT, DVp - n0);                  // multiple of 16
        const uint32_t idesc_o = make_idesc(kMQ, nw);
        for (int kc = 0; kc < kNK / kKC; ++kc) {
            // stage V^T chunk: task = (column n, key-group g): 8 consecutive key rows of one column
#pragma unroll 2
            for (int i = tid; i < kNT * 8; i += 256) {
                const int n = i & (kNT - 1), g = i >> 8;    // lanes run over n: coalesced 128-byte rows
                float v[8];
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    const int key = kKC * kc + 8 * g + j;
                    v[j] = (key < nkeys && n0 + n < DV) ? ldg1_stream(vbase + (size_t)key * DV + n0 + n) : 0.0f;
                }
                store_split8<kNT>(v_hi, v_lo, n, g, v);
            }
            fence_async_smem();
            fence_before();
            __syncthreads();
            if (tid == 0) {
                fence_after();
#pragma unroll
                for (int pass = 0; pass < 4; ++pass) {
                    const uint32_t ab = smem_u32(pass >= 2 ? p_lo : p_hi) + (kKC / 8) * kc * kAChunk;
                    const uint32_t bb = smem_u32((pass & 1) ? v_lo : v_hi);
#pragma unroll
                    for (int ks = 0; ks < kKC / 16; ++ks)
                        umma(tmem + 256, make_desc(ab + 2 * ks * kAChunk, kAChunk, 128), make_desc(bb + 2 * ks * kBChunk, kBChunk, 128),
                             idesc_o, (kc > 0 || pass > 0 || ks > 0) ? 1u : 0u);
                }
                umma_commit(mbar);
            }
            mbar_wait(mbar, phase & 1);
            ++phase;
            fence_after();
        }
        // epilogue of this N-tile: thread (row, half) -> 128 columns
        float* orow = AO + ((size_t)bl * T + t0 + r) * DV + n0 + 128 * hh;
#pragma unroll 1
        for (int ch = 0; ch < 4; ++ch) {
            if (128 * hh + 32 * ch >= nw) break;            // warp-uniform
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
        fence_before();                                     // the next N-tile's first MMA overwrites these TMEM columns
        __syncthreads();
        fence_after();
    }
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(kTmemCols) : "memory");
}

// usable when the band fits one 256-key tile, DK fits five 64-wide chunks and the float2 / float4 accesses are aligned
bool attn_core_tc_supported(const sb_attn_args& a) {
    const int DK = a.F * a.E, DV = a.F * (a.C / a.L);
    return a.T >= 64 && a.W <= 129 && DK <= 320 && DK % 2 == 0 && DV % 4 == 0;
}

int attn_core_tc(const sb_attn_args& a, const float* Q, const float* Kc, const float* Vc, float* AO, cudaStream_t st) {
    dim3 grid(ceil_div(a.T, atc::kMQ), a.B * a.L);
    return launch("attn_core_tc", attn_core_tc_kernel, grid, dim3(256), (size_t)atc::kSmemBytes, st, a, Q, Kc, Vc, AO);
}

#else   // SB_EMU: tensor-core instructions cannot be emulated on the host

bool attn_core_tc_supported(const sb_attn_args&) { return false; }
int attn_core_tc(const sb_attn_args&, const float*, const float*, const float*, float*, cudaStream_t) {
    set_error("the tcgen05 attention core is not available in the host-emulated test build");
    return SB_E_UNSUPP;
}

#endif

}  // namespace sb
