// Weight gradients of one LSTM direction on the 5th-generation tensor cores (SB_OPT_TRAIN_TC).
//
//   dW_ih[j][c] = sum_n dz[n][j] LN(x)[n][c]      dW_hh[j][u] = sum_n dz[n][j] h_prev[n][u]      db[j] = sum_n dz[n][j]
//
// over the N = rows x steps (row, step) pairs of a path (725 000 per direction at batch 8 x 5 s): the reduction GEMM
// D[4H = 256][C + H = 96] = dz^T [LN(x) | h_prev] that autograd runs for torch.nn.LSTM's weights (reference caller:
// PLModule._step -> loss.backward(), src/hl_modules/distance_based_hl_module.py:303-330; span DE3 tfgridnet_causal.py:794-849).
// The SIMT form (outer_kernel<256,96,8,12>, sb_train.cu) ran at 37 TFLOP/s = half of the fp32 FMA roof and was 23 % of the
// training step (profiles/r01_launches_train.txt).  Here:
//
//   * a slab = 32 consecutive (row, step) pairs = three CONTIGUOUS pieces of HBM (32 KB of dz, 4 KB of LN(x), 8 KB of h shifted
//     by one step): one thread pulls them with three TMA bulk copies (cp.async.bulk + mbarrier byte counting) into a ring of
//     three raw fp32 stages, two slabs ahead of their use, so the loads never wait for the arithmetic;
//   * both operands are contraction-major in HBM ([n][j] and [n][k], n outermost), i.e. MN-major for UMMA: the 512 threads turn
//     a raw stage into bf16 hi / lo images in the canonical no-swizzle MN-major layout - lane = 8-column group, so a warp reads
//     one raw row contiguously; 16-byte rows of the 8 x 8 core matrices; core matrices along M / N are SBO = 144 bytes apart
//     (16 bytes of padding make the stores conflict-free), along the contraction LBO apart - no transposition anywhere, and the
//     instruction descriptor carries a_major = b_major = MN;
//   * fp32 parity through the three-term split dz_hi X_hi + dz_hi X_lo + dz_lo X_hi (bf16 x bf16 products are exact in the
//     fp32 accumulator; what is dropped is dz_lo X_lo, 2^-16 relative), 12 tcgen05.mma (M = 128, N = 96, K = 16) per slab
//     into two TMEM accumulators (gate rows 0-127 | 128-255) that live for the CTA's whole range of n;
//   * db comes from the converting threads (a lane always holds the same 8 gate columns), the TMEM accumulators leave
//     through tcgen05.ld + one fp32 atomic per entry per CTA (one wave of CTAs, <= 1 per SM).
// HBM-bound by construction: 1 408 bytes per (row, step) read once.
#include "sb_common.cuh"

#ifndef SB_EMU
#include <cuda_bf16.h>
#endif

namespace sb {

#ifndef SB_EMU
__device__ __forceinline__ void atomic_add(float* p, float v) { atomicAdd(p, v); }

namespace wtc {

constexpr int kJ = 256, kC = 32, kH = 64, kX = kC + kH;      // gate rows, LN(x) columns, h columns, N of the MMA
constexpr int kSlab = 32;                                    // (row, step) pairs per slab = K of a slab
constexpr int kStages = 3;                                   // raw fp32 ring
constexpr int kThreads = 512;
constexpr int kSbo = 144;                                    // bytes between core matrices along M / N (128 + 16 of padding)
constexpr int kALbo = (kJ / 8) * kSbo, kBLbo = (kX / 8) * kSbo;                 // bytes between 8-deep contraction groups
constexpr int kABytes = (kSlab / 8) * kALbo, kBBytes = (kSlab / 8) * kBLbo;     // one bf16 image: 18 KB / 6.75 KB
constexpr int kImgBytes = 2 * kABytes + 2 * kBBytes;        // hi + lo of both operands
constexpr int kRawDz = kSlab * kJ * 4, kRawX = kSlab * kC * 4, kRawH = kSlab * kH * 4;
constexpr int kRawBytes = kRawDz + kRawX + kRawH;           // 44 KB per stage
constexpr int kOffImg = kStages * kRawBytes;
constexpr int kOffBar = kOffImg + kImgBytes;
constexpr int kSmemBytes = kOffBar + 64 + 128;              // + barriers + slack for the 128-byte alignment
constexpr uint32_t kTmemCols = 256;                         // 2 x 96 accumulator columns -> next power of two
constexpr int kPartFloats = kJ * kX + kJ;                   // one CTA's partial sums: D [256][96] + db [256]

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    uint64_t d = 0;                                         // SmemDescriptor, no swizzle
    d |= (uint64_t)((saddr & 0x3FFFFu) >> 4);
    d |= (uint64_t)(lbo_bytes >> 4) << 16;
    d |= (uint64_t)(sbo_bytes >> 4) << 32;
    d |= (uint64_t)1 << 46;
    return d;
}
// f32 accumulator, bf16 x bf16, A and B both MN-major (bits 15 / 16)
__host__ __device__ constexpr uint32_t make_idesc_mn(int M, int N) {
    return (1u << 4) | (1u << 7) | (1u << 10) | (1u << 15) | (1u << 16) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
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
__device__ __forceinline__ void fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void sts16(uint32_t saddr, uint4 v) {
    asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(saddr), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}
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

// 8 fp32 values -> 8 bf16 (hi) + 8 bf16 (lo = what the hi rounding lost)
__device__ __forceinline__ void split8(const float4 a, const float4 b, uint4& hi4, uint4& lo4) {
    const float v[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
    uint32_t hi[4], lo[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const __nv_bfloat162 h2 = __floats2bfloat162_rn(v[2 * i], v[2 * i + 1]);
        const float r0 = v[2 * i] - __bfloat162float(h2.x), r1 = v[2 * i + 1] - __bfloat162float(h2.y);
        const __nv_bfloat162 l2 = __floats2bfloat162_rn(r0, r1);
        hi[i] = *reinterpret_cast<const uint32_t*>(&h2);
        lo[i] = *reinterpret_cast<const uint32_t*>(&l2);
    }
    hi4 = make_uint4(hi[0], hi[1], hi[2], hi[3]);
    lo4 = make_uint4(lo[0], lo[1], lo[2], lo[3]);
}

}  // namespace wtc

__device__ __forceinline__ float4 lds4(uint32_t saddr) {
    float4 v;
    asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(saddr) : "memory");
    return v;
}

__global__ void __launch_bounds__(wtc::kThreads, 1) lstm_wgrad_tc_kernel(const WgradTc w) {
    using namespace wtc;
    extern __shared__ unsigned char sm_raw[];
    unsigned char* sm = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(sm_raw) + 127) & ~uintptr_t(127));
    BulkBarrier* full = reinterpret_cast<BulkBarrier*>(sm + kOffBar);             // [kStages] bytes of a raw stage have landed
    uint64_t* done = reinterpret_cast<uint64_t*>(sm + kOffBar) + kStages;          // tcgen05.commit: the MMAs have read the images
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(done + 1);
    const uint32_t sm_s = smem_u32(sm);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(kTmemCols) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    if (tid == 32) {
        for (int i = 0; i < kStages; ++i) bulk_barrier_init(full + i);
        mbar_init(done, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    pdl_trigger();
    pdl_wait();
    fence_before();
    __syncthreads();
    fence_after();
    const uint32_t tmem = *tmem_slot;

    const long long n_begin = (long long)blockIdx.x * w.rows_per_cta;
    const long long n_end = n_begin + w.rows_per_cta < w.N ? n_begin + w.rows_per_cta : w.N;
    const int n_slabs = n_begin < n_end ? (int)((n_end - n_begin + kSlab - 1) / kSlab) : 0;

    // one thread: the three contiguous pieces of slab i -> raw stage i % kStages.  Rows of h that are not copied (before the
    // first / after the last (row, step) pair of the whole problem) are sequence starts: the converters never read them.
    auto issue_loads = [&](int i) {
        const int st = i % kStages;
        const long long n0 = n_begin + (long long)i * kSlab;
        const int nv = (int)(n_end - n0 < kSlab ? n_end - n0 : kSlab);
        long long h0 = w.reverse ? n0 + 1 : n0 - 1;
        int hc = nv, hskip = 0;
        if (h0 < 0) { hskip = 1; h0 = 0; hc -= 1; }
        if (h0 + hc > w.N) hc = (int)(w.N - h0);
        if (hc < 0) hc = 0;
        unsigned char* raw = sm + st * kRawBytes;
        bulk_expect(full + st, (unsigned)(nv * (kJ + kC) * 4 + hc * kH * 4));
        bulk_copy_g2s(reinterpret_cast<float*>(raw), w.dz + n0 * kJ, (unsigned)(nv * kJ * 4), full + st);
        bulk_copy_g2s(reinterpret_cast<float*>(raw + kRawDz), w.xn + n0 * kC, (unsigned)(nv * kC * 4), full + st);
        if (hc > 0)
            bulk_copy_g2s(reinterpret_cast<float*>(raw + kRawDz + kRawX + hskip * kH * 4), w.h + h0 * kH, (unsigned)(hc * kH * 4), full + st);
    };
    if (tid == 0)
        for (int i = 0; i < kStages && i < n_slabs; ++i) issue_loads(i);

    float dbacc[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) dbacc[i] = 0.f;
    constexpr uint32_t idesc = make_idesc_mn(128, kX);
    const uint32_t a_hi = sm_s + kOffImg, a_lo = a_hi + kABytes, b_hi = a_lo + kABytes, b_lo = b_hi + kBBytes;

    for (int i = 0; i < n_slabs; ++i) {
        const int st = i % kStages;
        const long long n0 = n_begin + (long long)i * kSlab;
        const int nv = (int)(n_end - n0 < kSlab ? n_end - n0 : kSlab);
        const uint32_t raw = sm_s + st * kRawBytes;
        bulk_wait(full + st, (unsigned)((i / kStages) & 1));
        if (i >= 1) mbar_wait(done, (uint32_t)((i - 1) & 1));       // the MMAs of slab i - 1 have read the images
        // ---- dz: row = 32 lanes x 8 columns; warp w takes rows w and w + 16 --------------------------------------------------
#pragma unroll
        for (int it = 0; it < 2; ++it) {
            const int row = warp + 16 * it;
            float4 v0 = make_float4(0.f, 0.f, 0.f, 0.f), v1 = v0;
            if (row < nv) {
                v0 = lds4(raw + row * (kJ * 4) + lane * 32);
                v1 = lds4(raw + row * (kJ * 4) + lane * 32 + 16);
            }
            dbacc[0] += v0.x; dbacc[1] += v0.y; dbacc[2] += v0.z; dbacc[3] += v0.w;
            dbacc[4] += v1.x; dbacc[5] += v1.y; dbacc[6] += v1.z; dbacc[7] += v1.w;
            uint4 hi4, lo4;
            split8(v0, v1, hi4, lo4);
            const uint32_t off = (uint32_t)((row >> 3) * kALbo + lane * kSbo + (row & 7) * 16);
            sts16(a_hi + off, hi4);
            sts16(a_lo + off, lo4);
        }
        // ---- [LN(x) | h_prev]: 32 rows x 12 column groups = 384 tasks -------------------------------------------------------
        if (tid < kSlab * (kX / 8)) {
            const int row = tid / (kX / 8), g = tid - row * (kX / 8);
            float4 v0 = make_float4(0.f, 0.f, 0.f, 0.f), v1 = v0;
            if (row < nv) {
                if (g < kC / 8) {
                    v0 = lds4(raw + kRawDz + row * (kC * 4) + g * 32);
                    v1 = lds4(raw + kRawDz + row * (kC * 4) + g * 32 + 16);
                } else {                                    // h of the previous step in processing order; zero row at the sequence start
                    const int s = (int)((n0 + row) % w.S);
                    const bool zero = w.reverse ? s == w.S - 1 : s == 0;
                    if (!zero) {
                        v0 = lds4(raw + kRawDz + kRawX + row * (kH * 4) + (g - kC / 8) * 32);
                        v1 = lds4(raw + kRawDz + kRawX + row * (kH * 4) + (g - kC / 8) * 32 + 16);
                    }
                }
            }
            uint4 hi4, lo4;
            split8(v0, v1, hi4, lo4);
            const uint32_t off = (uint32_t)((row >> 3) * kBLbo + g * kSbo + (row & 7) * 16);
            sts16(b_hi + off, hi4);
            sts16(b_lo + off, lo4);
        }
        fence_async_smem();                                 // generic-proxy stores -> visible to the tensor core's async proxy
        fence_before();
        __syncthreads();                                    // images complete; raw stage st consumed by every thread
        if (tid == 0) {
            fence_after();
#pragma unroll
            for (int term = 0; term < 3; ++term) {
                const uint32_t ab = term == 2 ? a_lo : a_hi, bb = term == 1 ? b_lo : b_hi;
#pragma unroll
                for (int ks = 0; ks < kSlab / 16; ++ks) {
                    const uint64_t bdesc = make_desc(bb + ks * 2 * kBLbo, kBLbo, kSbo);
#pragma unroll
                    for (int mt = 0; mt < 2; ++mt)
                        umma(tmem + mt * kX, make_desc(ab + ks * 2 * kALbo + mt * 16 * kSbo, kALbo, kSbo), bdesc, idesc,
                             (i > 0 || term > 0 || ks > 0) ? 1u : 0u);
                }
            }
            umma_commit(done);
            if (i + kStages < n_slabs) issue_loads(i + kStages);
        }
    }

    // ---- db: lane = column group, the 16 warps hold different rows: combine them through the (now idle) raw ring ---------------
    float* red = reinterpret_cast<float*>(sm);              // [16 warps][256]
    if (n_slabs > 0) {
        mbar_wait(done, (uint32_t)((n_slabs - 1) & 1));     // every MMA has finished (and with it every read of shared memory)
        fence_after();
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) red[warp * kJ + 8 * lane + i] = dbacc[i];
    __syncthreads();
    float* part = w.scratch ? w.scratch + (size_t)blockIdx.x * kPartFloats : nullptr;
    if (tid < kJ && n_slabs > 0) {
        float v = 0.f;
#pragma unroll
        for (int k = 0; k < kThreads / 32; ++k) v += red[k * kJ + tid];
        if (part) part[kJ * kX + tid] = v;
        else {
            if (w.db) atomic_add(w.db + tid, v);
            if (w.db2) atomic_add(w.db2 + tid, v);
        }
    }
    // ---- accumulators: warp q of a warpgroup reads TMEM lanes 32q .. 32q + 31 = gate rows; warpgroup = accumulator.  With a
    // scratch buffer the CTA writes its partial sums with plain stores and wgrad_reduce_kernel adds the CTAs up: the atomic
    // epilogue (148 CTAs x 24 832 adds onto the same addresses) was a third of the kernel (profiles/r02_prof_wgrad_tc.txt) ----
    if (n_slabs > 0 && warp < 8) {
        const int mt = warp >> 2, q = warp & 3;
        const int j = 128 * mt + 32 * q + lane;
        const uint32_t taddr = tmem + ((uint32_t)(32 * q) << 16) + mt * kX;
        uint32_t r[32];
#pragma unroll
        for (int blk = 0; blk < 3; ++blk) {
            tmem_ld32(taddr + 32 * blk, r);
            if (part) {
                float* dst = part + (size_t)j * kX + 32 * blk;
#pragma unroll
                for (int c = 0; c < 32; c += 4)
                    st4(dst + c, make_float4(__uint_as_float(r[c]), __uint_as_float(r[c + 1]), __uint_as_float(r[c + 2]), __uint_as_float(r[c + 3])));
            } else {
#pragma unroll
                for (int c = 0; c < 32; ++c) {
                    if (blk == 0) atomic_add(w.dW_ih + (size_t)j * kC + c, __uint_as_float(r[c]));
                    else atomic_add(w.dW_hh + (size_t)j * kH + 32 * (blk - 1) + c, __uint_as_float(r[c]));
                }
            }
        }
    }
    fence_before();
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(kTmemCols) : "memory");
}

// dW_ih / dW_hh / db += sum over the CTAs' partial sums; thread = one entry, consecutive threads = consecutive entries
__global__ void __launch_bounds__(256) wgrad_reduce_kernel(const WgradTc w, const int n_parts) {
    using namespace wtc;
    pdl_wait();
    const int e = blockIdx.x * 256 + threadIdx.x;
    if (e >= kPartFloats) return;
    float v = 0.f;
    for (int p = 0; p < n_parts; ++p) v += w.scratch[(size_t)p * kPartFloats + e];
    if (e < kJ * kX) {
        const int j = e / kX, k = e - j * kX;
        if (k < kC) w.dW_ih[(size_t)j * kC + k] += v;
        else w.dW_hh[(size_t)j * kH + (k - kC)] += v;
    } else {
        const int j = e - kJ * kX;
        if (w.db) w.db[j] += v;
        if (w.db2) w.db2[j] += v;
    }
}

int run_wgrad_tc(const WgradTc& w0, cudaStream_t st) {
    using namespace wtc;
    WgradTc w = w0;
    const long long ctas = sm_count();                      // one wave, one CTA per SM (182 KB of shared memory each)
    long long rows = ceil_div_ll(w.N, ctas);
    rows = ceil_div_ll(rows, kSlab) * kSlab;
    w.rows_per_cta = rows < kSlab ? kSlab : rows;
    const int grid = (int)ceil_div_ll(w.N, w.rows_per_cta);
    if (!w.scratch || w.scratch_floats < (long long)grid * kPartFloats) w.scratch = nullptr;     // small problems: atomics
    SB_CHECK(launch("lstm_wgrad_tc", lstm_wgrad_tc_kernel, dim3((unsigned)grid), dim3(kThreads), (size_t)kSmemBytes, st, w));
    if (!w.scratch) return 0;
    return launch("lstm_wgrad_reduce", wgrad_reduce_kernel, dim3((unsigned)ceil_div(kPartFloats, 256)), dim3(256), 0, st, w, grid);
}

#else   // SB_EMU: tensor-core instructions cannot be emulated on the host

int run_wgrad_tc(const WgradTc&, cudaStream_t) {
    set_error("lstm_wgrad_tc (tcgen05) is not available in the host-emulated test build");
    return SB_E_UNSUPP;
}

#endif

}  // namespace sb
