// conv-in (Conv2d k = (3,3), pad (0,1), causal in time) + LayerNorm on the 5th-generation tensor cores (SB_OPT_FRONT_TC).
// Reference span: history cat DE3 tfgridnet_causal.py:504-505, Conv2d :332-347, LayerNormPermuted :219-231 - the same
// span conv_in_kernel (sb_frontend.cu) covers with fp32 FMAs; that kernel ran at 39 % of the fp32 FMA roof (82 us per group of
// 1024 frames) and, with the LSTMs down to 16 / 37 CTAs per launch, the front- and back-end had grown to 30 % of the grouped
// step's SM-time.
//
// Implicit GEMM without im2col.  Positions of an utterance are numbered linearly INCLUDING the two zero columns that pad a frame
// in frequency: u = (t + 2) * FP + (f + 1), FP = F + 2, frames -2, -1 = the carried history.  The staged input is ONE K-major
// operand [u][32 channels] (27 real, 5 zero) as bf16 hi / lo images in the canonical no-swizzle layout with SBO = 128 bytes,
// i.e. the 16-byte rows of a k chunk are simply contiguous in u.  Output position p = t * FP + j reads, for tap (kt, kf), row
// u = p + kt * FP + kf: every tap is the SAME image behind a start address that is kt * FP + kf rows further on, so the nine taps
// are nine descriptor offsets - nothing is copied or shifted.  The outputs of the two pad columns are computed and dropped
// (2 of 147).  Per 128-position tile: 9 taps x 2 k steps x 3 terms (hi.hi + hi.lo + lo.hi) = 54 tcgen05.mma (M = 128, N = 32,
// K = 16).  Dependent MMAs on one accumulator cost ~150 cycles each whatever N is, so the three tap rows kt accumulate into
// three separate 32-column accumulators that are issued round-robin (~25 ns per MMA) and summed at read-back.
// A CTA covers four tiles and is software-pipelined: 16 warps stage the rows tile 0 needs (128 + two frames of lookback), then
// 128 more rows per further tile; a 17th warp issues a tile's MMAs as soon as its rows are in place; after staging, warp
// (tile, lane quarter) waits for its tile's commit, reads the accumulators back (thread = position), adds the bias, normalises
// over the 32 channels in registers and stores 128 bytes per position.
#include "sb_common.cuh"

#ifndef SB_EMU
#include <cuda_bf16.h>
#include <type_traits>
#endif

namespace sb {

#ifndef SB_EMU
namespace cit {

constexpr int kC = 32;                                      // output channels = N
constexpr int kKPad = 32;                                   // input channels padded to two K = 16 steps
constexpr int kTiles = 4, kRowsOut = 128 * kTiles;          // output positions per CTA
constexpr int kStageThreads = 512, kThreads = kStageThreads + 32;      // 16 staging / read-back warps + the MMA-issue warp
constexpr int kWImgBytes = 9 * (kKPad / 8) * kC * 16;       // one bf16 image of the nine [N = 32][K = 32] tap matrices: 18 KB
constexpr uint32_t kTmemCols = 512;                         // 4 tiles x 3 tap-row accumulators x 32 columns = 384 -> 512

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    uint64_t d = 0;                                         // SmemDescriptor: no swizzle, K-major
    d |= (uint64_t)((saddr & 0x3FFFFu) >> 4);
    d |= (uint64_t)(lbo_bytes >> 4) << 16;
    d |= (uint64_t)(sbo_bytes >> 4) << 32;
    d |= (uint64_t)1 << 46;
    return d;
}
__host__ __device__ constexpr uint32_t make_idesc(int M, int N) {      // f32 accumulator, bf16 x bf16, both K-major
    return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
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
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {      // bounded: a protocol error traps
    const uint32_t a = smem_u32(bar);
#pragma unroll 1
    for (long long spin = 0; spin < (1ll << 26); ++spin) {
        uint32_t ok;
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}\n"
                     : "=r"(ok) : "r"(a), "r"(parity) : "memory");
        if (ok) return;
    }
    asm volatile("trap;");
}
__device__ __forceinline__ bool elect_one() {
    uint32_t pred;
    asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}\n" : "=r"(pred));
    return pred != 0;
}
__device__ __forceinline__ void fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&r)[32]) {
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
}
__device__ __forceinline__ void sts4(uint32_t saddr, uint32_t v) { asm volatile("st.shared.b32 [%0], %1;" ::"r"(saddr), "r"(v) : "memory"); }

// two fp32 values -> packed bf16 hi pair and the packed bf16 pair of what the hi rounding lost
__device__ __forceinline__ void split2(float v0, float v1, uint32_t& hi, uint32_t& lo) {
    const __nv_bfloat162 h2 = __floats2bfloat162_rn(v0, v1);
    const __nv_bfloat162 l2 = __floats2bfloat162_rn(v0 - __bfloat162float(h2.x), v1 - __bfloat162float(h2.y));
    hi = *reinterpret_cast<const uint32_t*>(&h2);
    lo = *reinterpret_cast<const uint32_t*>(&l2);
}

__device__ const float kZeroPair[2] = {0.0f, 0.0f};
__device__ __forceinline__ int ceil_div_dev(int a, int b) { return (a + b - 1) / b; }

// rows of the staged operand a CTA needs: 512 output positions + two frames of lookback + the kf reach
__host__ __device__ inline int rows_needed(int FP) { return kRowsOut + 2 * FP + 2; }
// padded so that consecutive k chunks start 8 banks apart: the staging stores of a warp (2 rows x 4 chunks x 4 words) then
// touch 32 different banks
__host__ __device__ inline int rows_padded(int FP) {
    int r = rows_needed(FP);
    while (r % 8 != 2) ++r;
    return r;
}

}  // namespace cit

// grid (ceil(T * FP / 512), B).  TRAIN: the training forward (sb_conv_in_train_fwd): weights as stored [o][c][kt][kf], zero history,
// no new history, no LayerNorm (ln_fwd_kernel keeps what the backward needs).
template <bool TRAIN>
__global__ void __launch_bounds__(cit::kThreads, 1) conv_in_tc_kernel(const sb_conv_in_args a) {
    using namespace cit;
    extern __shared__ unsigned char sm_raw[];
    unsigned char* sm = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(sm_raw) + 127) & ~uintptr_t(127));
    const int F = a.F, Cin = a.Cin, T = a.T, FP = F + 2;
    const int R = rows_padded(FP);
    const uint32_t lbo = (uint32_t)R * 16;                  // bytes between k chunks of an image
    const uint32_t img_bytes = 4 * lbo;
    unsigned char* img = sm;                                // [hi, lo][4 k chunks][R rows][8 channels] bf16
    unsigned char* wimg = img + 2 * img_bytes;              // [hi, lo][9 taps][4 k chunks][32 n][8 channels] bf16
    int* rowoff = reinterpret_cast<int*>(wimg + 2 * kWImgBytes);              // [R] where a staged row comes from
    float* par = reinterpret_cast<float*>(rowoff + R);      // bias, ln gain, ln bias: 3 x [32]
    uint64_t* done = reinterpret_cast<uint64_t*>(par + 3 * kC + ((R & 1) ? 1 : 0));   // [4] tcgen05.commit per tile (8-byte aligned)
    uint64_t* staged = done + kTiles;                       // [4] 16 warps: the rows tile k needs are in the images
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(staged + kTiles);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int b = blockIdx.y;
    const int p0 = blockIdx.x * kRowsOut;                   // first output position of the CTA
    // staged row s <-> input position u = p0 + s - 1 (u = (t + 2) * FP + j); tap (kt, kf) of output p = p0 + m reads
    // u = p + kt * FP + kf - 1, i.e. row s = m + kt * FP + kf

    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(kTmemCols) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    if (tid == 32) {
        for (int k = 0; k < kTiles; ++k) { mbar_init(done + k, 1); mbar_init(staged + k, kStageThreads / 32); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    // ---- constants: the nine tap matrices as K-major bf16 hi / lo images, bias, LayerNorm parameters -----------------------
    // w_pack[((kt * Cin + c) * 3 + kf) * C + n]; image: tap, k chunk c / 8, row n, element c % 8
    const uint32_t wimg_s = smem_u32(wimg);
    static_assert(9 * (kKPad / 2) * kC == 9 * kStageThreads, "nine fully unrolled passes: all 18 loads of a thread are in flight at once");
    if (tid < kStageThreads) {
#pragma unroll
        for (int it = 0; it < 9; ++it) {
            const int e = it * kStageThreads + tid;
            const int n = e & 31, c2 = (e >> 5) & 15, tap = e >> 9;
            const int kt = tap / 3, kf = tap - 3 * kt, c = 2 * c2;
            auto widx = [&](int cc) { return TRAIN ? ((size_t)n * Cin + cc) * 9 + tap : ((size_t)(kt * Cin + cc) * 3 + kf) * kC + n; };
            const float v0 = c < Cin ? __ldg(a.w_pack + widx(c)) : 0.0f;
            const float v1 = c + 1 < Cin ? __ldg(a.w_pack + widx(c + 1)) : 0.0f;
            uint32_t hi, lo;
            split2(v0, v1, hi, lo);
            const uint32_t off = (uint32_t)(((tap * 4 + (c2 >> 2)) * kC + n) * 16 + (c2 & 3) * 4);
            sts4(wimg_s + off, hi);
            sts4(wimg_s + kWImgBytes + off, lo);
        }
    }
    if (tid < kC) {
        par[tid] = __ldg(a.bias + tid);
        par[kC + tid] = a.ln_g ? __ldg(a.ln_g + tid) : 1.0f;
        par[2 * kC + tid] = a.ln_g ? __ldg(a.ln_b + tid) : 0.0f;
    }
    // where every staged row comes from: >= 0 float offset into feats; -1 zero row; <= -2: history, -2 - ((2 + t) * F + f)
    const int n_u = (T + 2) * FP;
    for (int s = tid; s < R; s += kThreads) {
        const int u = p0 + s - 1;
        int code = -1;
        if (u >= 0 && u < n_u) {
            const int fr = u / FP, j = u - fr * FP, f = j - 1, t = fr - 2;
            if (f >= 0 && f < F) code = t >= 0 ? ((b * T + t) * F + f) * Cin : (TRAIN ? -1 : -2 - ((2 + t) * F + f));
        }
        rowoff[s] = code;
    }
    pdl_trigger();
    pdl_wait();                                             // feats / conv_buf_in come from the predecessor
    fence_async_smem();                                     // the tap images are read by the tensor core
    fence_before();
    __syncthreads();
    fence_after();
    const uint32_t tmem = *tmem_slot;
    const uint32_t img_s = smem_u32(img);

    if (warp == kStageThreads / 32) {
        // ---- MMA issue: tile k as soon as its rows are staged; accumulator kt of a tile takes tap row kt ------------------
        if (elect_one()) {
            constexpr uint32_t idesc = make_idesc(128, kC);
#pragma unroll 1
            for (int tile = 0; tile < kTiles; ++tile) {
                mbar_wait(staged + tile, 0u);
                fence_after();
                const uint32_t arow = img_s + (uint32_t)tile * 128 * 16;
#pragma unroll
                for (int kf = 0; kf < 3; ++kf)
#pragma unroll
                    for (int ks = 0; ks < 2; ++ks)
#pragma unroll
                        for (int pass = 0; pass < 3; ++pass)
#pragma unroll
                            for (int kt = 0; kt < 3; ++kt) {
                                const uint32_t ab = arow + (pass == 2 ? img_bytes : 0) + (uint32_t)(2 * ks) * lbo + (uint32_t)(kt * FP + kf) * 16;
                                const uint32_t wb = wimg_s + (pass == 1 ? kWImgBytes : 0) + (uint32_t)(((kt * 3 + kf) * 4 + 2 * ks) * kC * 16);
                                umma(tmem + 32u * (3 * tile + kt), make_desc(ab, lbo, 128), make_desc(wb, kC * 16, 128), idesc,
                                     (kf | ks | pass) ? 1u : 0u);
                            }
                umma_commit(done + tile);
            }
        }
        __syncwarp();
    } else {
        // ---- stage the operand: one (row, channel pair) per thread and pass, coalesced along the channels of a position ---
        // Every load is unconditional - rows and channels that do not exist read a zero word - so that the U passes of a
        // round have their 2 U loads in flight together (with branches around the loads a pass cost a full L2 latency).
        const float* hist = TRAIN ? a.feats : a.conv_buf_in + (size_t)b * Cin * 2 * F;      // (TRAIN: no row refers to it)
        auto stage_rows = [&](int row0, int row1, auto unroll_tag) {
            constexpr int U = decltype(unroll_tag)::value;
            const int first = row0 * 16, n_pairs = (row1 - row0) * 16;
            for (int base = 0; base < n_pairs; base += U * kStageThreads) {
                float v0[U], v1[U];
#pragma unroll
                for (int k = 0; k < U; ++k) {
                    const int rel = base + k * kStageThreads + tid, idx = first + rel;
                    const int c = 2 * (idx & 15);
                    const int code = rel < n_pairs ? rowoff[idx >> 4] : -1;
                    const float* q0 = kZeroPair;
                    const float* q1 = kZeroPair;
                    if (code >= 0) {
                        if (c < Cin) q0 = a.feats + code + c;
                        if (c + 1 < Cin) q1 = a.feats + code + c + 1;
                    } else if (code <= -2) {
                        if (c < Cin) q0 = hist + (size_t)c * 2 * F + (-2 - code);
                        if (c + 1 < Cin) q1 = hist + (size_t)(c + 1) * 2 * F + (-2 - code);
                    }
                    v0[k] = __ldg(q0);
                    v1[k] = __ldg(q1);
                }
#pragma unroll
                for (int k = 0; k < U; ++k) {
                    const int rel = base + k * kStageThreads + tid, idx = first + rel;
                    if (rel < n_pairs) {
                        const int s = idx >> 4, c2 = idx & 15;
                        uint32_t hi, lo;
                        split2(v0[k], v1[k], hi, lo);
                        const uint32_t off = (uint32_t)(c2 >> 2) * lbo + (uint32_t)s * 16 + (uint32_t)(c2 & 3) * 4;
                        sts4(img_s + off, hi);
                        sts4(img_s + img_bytes + off, lo);
                    }
                }
            }
        };
        int row0 = 0;
        for (int k = 0; k < kTiles; ++k) {
            const int row1 = min(R, 128 * (k + 1) + 2 * FP + 2);         // tile k reads rows 128 k .. 128 k + 127 + 2 FP + 2
            if (k == 0) stage_rows(row0, row1, std::integral_constant<int, 7>{});
            else stage_rows(row0, row1, std::integral_constant<int, 4>{});
            row0 = row1;
            fence_async_smem();
            __syncwarp();
            if (lane == 0) mbar_arrive(staged + k);
        }

        // ---- the new history (fp32, exact): last two frames of [history ; feats] = frames T - 2, T - 1 (T >= 4 on this path) --
        if (!TRAIN && blockIdx.x == 0) {
            float* dst = a.conv_buf_out + (size_t)b * Cin * 2 * F;
#pragma unroll 8
            for (int i = tid; i < 2 * F * Cin; i += kStageThreads) {
                const int c = i % Cin, jf = i / Cin, j = jf / F, f = jf - j * F;
                dst[(size_t)c * 2 * F + (size_t)j * F + f] = __ldg(a.feats + ((size_t)(b * T + T - 2 + j) * F + f) * Cin + c);
            }
        }

        // ---- read-back: warp = (tile, TMEM lane quarter), thread = output position ---------------------------------------
        const int tile = warp >> 2, q = warp & 3;
        mbar_wait(done + tile, 0u);
        fence_after();
        uint32_t acc[32], part[32];
        const uint32_t taddr = tmem + ((uint32_t)(32 * q) << 16) + 32u * (3 * tile);
        tmem_ld32(taddr, acc);
        tmem_ld32(taddr + 32, part);
        float o[32];
#pragma unroll
        for (int n = 0; n < 32; ++n) o[n] = __uint_as_float(acc[n]) + __uint_as_float(part[n]);
        tmem_ld32(taddr + 64, part);
        const int p = p0 + tile * 128 + 32 * q + lane;
        const int t = p / FP, j = p - t * FP;
        if (t < T && j >= 1 && j <= F) {
            float s1 = 0.f;
#pragma unroll
            for (int n = 0; n < 32; ++n) { o[n] = (o[n] + __uint_as_float(part[n])) + par[n]; s1 += o[n]; }
            if (a.ln_g) {
                const float mean = s1 * (1.0f / kC);
                float s2 = 0.f;
#pragma unroll
                for (int n = 0; n < 32; ++n) { o[n] -= mean; s2 = fmaf(o[n], o[n], s2); }
                const float rstd = rsqrtf(s2 * (1.0f / kC) + kLnEps);
#pragma unroll
                for (int n = 0; n < 32; ++n) o[n] = fmaf(o[n] * rstd, par[kC + n], par[2 * kC + n]);
            }
            float* dst = a.x + ((size_t)(b * T + t) * F + (j - 1)) * kC;
#pragma unroll
            for (int n = 0; n < 32; n += 4) st4(dst + n, make_float4(o[n], o[n + 1], o[n + 2], o[n + 3]));
        }
    }
    fence_before();
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(kTmemCols) : "memory");
}

static size_t conv_in_tc_smem(int F) {
    const int R = cit::rows_padded(F + 2);
    return (size_t)2 * 4 * R * 16 + 2 * cit::kWImgBytes + (size_t)(R + 3 * cit::kC + 1) * 4 + 2 * cit::kTiles * 8 + 16 + 128;
}

bool conv_in_tc_supported(const sb_conv_in_args& p) {
    if (p.C != 32 || p.Cin > cit::kKPad || p.T < 4) return false;
    if ((long long)p.B * p.T * p.F * p.Cin >= (1ll << 31)) return false;                  // row offsets are ints
    return conv_in_tc_smem(p.F) <= 227 * 1024;
}

int run_conv_in_tc(const sb_conv_in_args& p, cudaStream_t st) {
    const int FP = p.F + 2;
    dim3 grid(ceil_div(p.T * FP, cit::kRowsOut), p.B);
    return launch("conv_in_tc", conv_in_tc_kernel<false>, grid, dim3(cit::kThreads), conv_in_tc_smem(p.F), st, p);
}

// the training forward on the same kernel: raw[B][T][F][C] = conv(feats from zero history) + bias, weights as stored
bool conv_in_tc_train_supported(int B, int T, int F, int Cin, int C) {
    if (C != 32 || Cin > cit::kKPad || T < 1) return false;
    if ((long long)B * T * F * Cin >= (1ll << 31)) return false;
    return conv_in_tc_smem(F) <= 227 * 1024;
}
int run_conv_in_tc_train(const float* feats, const float* w, const float* bias, float* raw, int B, int T, int F, int Cin, cudaStream_t st) {
    sb_conv_in_args p{};
    p.feats = feats; p.w_pack = w; p.bias = bias; p.x = raw; p.B = B; p.T = T; p.F = F; p.Cin = Cin; p.C = 32;
    dim3 grid(ceil_div(T * (F + 2), cit::kRowsOut), B);
    return launch("conv_in_tc_train", conv_in_tc_kernel<true>, grid, dim3(cit::kThreads), conv_in_tc_smem(F), st, p);
}

#else   // SB_EMU: tensor-core instructions cannot be emulated on the host

bool conv_in_tc_supported(const sb_conv_in_args&) { return false; }
bool conv_in_tc_train_supported(int, int, int, int, int) { return false; }
int run_conv_in_tc_train(const float*, const float*, const float*, float*, int, int, int, int, cudaStream_t) {
    set_error("conv_in_tc: not available in the host-emulated test build");
    return SB_E_UNSUPP;
}
int run_conv_in_tc(const sb_conv_in_args&, cudaStream_t) {
    set_error("conv_in_tc: not available in the host-emulated test build");
    return SB_E_UNSUPP;
}

#endif

}  // namespace sb
