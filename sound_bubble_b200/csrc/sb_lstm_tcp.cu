// Warp-specialised tcgen05 / TMEM / TMA sequence kernel (SB_ALGO_TCP): the pipelined successor of lstm_tc_kernel.
//
// Same arithmetic, operand images and summation order as lstm_tc_kernel (sb_lstm_tc.cu): [FiLM] -> LayerNorm(C) ->
// [LN(x) | h] W^T + b as the three-term bf16 hi/lo split on the 5th-generation tensor cores -> cell -> Linear -> residual.
// Reference: GridNetBlock.forward intra / inter branches, DE3 tfgridnet_causal.py:794-849.
//
// What changed is WHO does what, so that the 145 (intra) / T (inter) dependent steps carry as little as possible:
//
//   warps 0-3  "stream" group, thread = tile row:
//        * lane 0 of warp 0 is the TMA producer: the [128 rows x C] fp32 slab of every step is pulled from the TF grid
//          with cp.async.bulk.tensor (4-D tensor map over [outer][pos|row][row|pos][C], 128-byte swizzle, SASS UTMALDG)
//          into a ring of 16 KB stages, several steps ahead of its use; completion is counted by an mbarrier.
//        * every thread turns its row of the NEXT step into LayerNorm(FiLM(x0 [+ x1])) as bf16 hi/lo in the x part of the
//          A operand while the cell update of the current step runs, keeps x' for the residual, and - once the projection
//          of h_{s-1} (issued with step s) has landed in TMEM - stores y_{s-1} = x' + b + lin h_{s-1} as full 128-byte rows.
//        * lane 0 of warp 1 issues the tcgen05.mma of a step the moment the 8 cell-update warps have published h:
//          12 MMAs (N = 256: the h part of the contraction, three terms x four k steps) + commit, 12 (N = 32) for the
//          projection + a commit that means "every MMA of this step has finished"; the 6 MMAs of the x part do not depend on
//          h and were issued a step earlier into the other of two TMEM gate buffers, while the cell warps were busy.
//          Measured (profiles/r02_tcp_*.txt, us per 145 steps): two N = 128 halves with a commit each 674, + x part ahead 681,
//          one N = 256 MMA per k step 688, + x part ahead 641: per-MMA overhead, not tensor throughput, sits on the serial path.  (Four N = 64 quarters, so that all eight cell warps start on the
//          first commit, were measured SLOWER - 858 vs 722 us per 145 steps: every MMA re-reads its 4 KB slice of A from
//          shared memory whatever N is, and at N = 64 that read, not the math, sets the MMA time.)
//   warps 4-11 cell-update group, thread = (row, half of the units): tcgen05.ld of its 128 gate columns in four chunks
//          (the load of chunk k+1 in flight while chunk k is evaluated), bias, cell, c in registers for the whole
//          sequence, h back into the A operand as bf16 hi/lo one chunk late (A may only be rewritten once every MMA has
//          read it), fence.proxy.async, ONE mbarrier arrive per warp.
//
// No __syncthreads in the step loop: full[] (TMA bytes), gates[2] + alldone (tcgen05.commit) and hready (8 warp arrivals) are
// mbarriers; the stream group uses one 128-thread named barrier per step.  Every wait is bounded and traps.
//
// Rows are tiled so that a TMA box never straddles an outer index (utterance): per outer index floor(rows_inner / 128)
// full tiles, and the remaining rows_inner % 128 rows of floor(128 / tail) consecutive outer indices share one tile
// (F = 145: one full tile per utterance + 17-row tails of 7 utterances per tile; 37 tiles at batch 32, as before).
#include "sb_common.cuh"
#include "sb_lstm.cuh"

#ifndef SB_EMU
#include <cuda.h>
#include <cuda_bf16.h>
#include <cstring>
#include <mutex>
#endif

namespace sb {

#ifndef SB_EMU

namespace tcp {

constexpr int kRows = 128, kC = 32, kH = 64, kK = kC + kH, kN = 4 * kH;
constexpr int kThreads = 384;
constexpr bool kPreX = true;                                // x part of the contraction issued one step ahead (688 -> 641 us per 145 steps)
constexpr int kGateN = 256;                                 // gate columns per MMA: 256 = one MMA per k step, 128 = two halves with a commit each
constexpr int kAChunkBytes = (kRows / 8) * 128;            // LBO of A: 2048
constexpr int kWChunkBytes = (kN / 8) * 128;               // LBO of the gate matrix: 4096
constexpr int kPChunkBytes = (kC / 8) * 128;               // LBO of the projection: 512
constexpr int kABytes = kRows * kK * 2, kWBytes = kN * kK * 2, kPBytes = kC * kH * 2;
constexpr int kSlabBytes = kRows * kC * 4;                 // one [128][32] fp32 stage = one TMA box
constexpr int kSlabs = 4;
constexpr int kOffW = kSlabs * kSlabBytes;                 // packed operand images: gate hi, gate lo, proj hi, proj lo
constexpr int kOffA = kOffW + 2 * kWBytes + 2 * kPBytes;   // A hi, A lo
constexpr int kOffBias = kOffA + 2 * kABytes;
constexpr int kOffLn = kOffBias + kN * 4;                  // LayerNorm gain, bias, projection bias: 3 x [C]
constexpr int kOffBar = kOffLn + 3 * kC * 4;
constexpr int kSmemBytes = kOffBar + 256 + 1024;           // + barriers + slack for the 1024-byte alignment of the stages
constexpr uint32_t kTmemCols = 512;                        // 256 gate columns + 32 projection columns -> next power of 2

struct Geom {
    int mode;           // 0: pos is dim 1, rows are dim 2 (intra: rows (b,t), steps f); 1: rows dim 1, pos dim 2 (inter)
    int n_outer;        // n_rows / rows_inner
    int nfull;          // full 128-row tiles per outer index
    int tail;           // rows_inner % 128
    int P;              // outer indices whose tails share a tile
    int n_full_tiles;   // n_outer * nfull
};

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    uint64_t d = 0;                                         // SmemDescriptor: no swizzle, K-major
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
__device__ __forceinline__ void mbar_expect(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
// bounded wait: a protocol error must end in a trap, not in a hung GPU
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
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
// One lane of a CONVERGED warp.  MMA issue sits behind `warp == 1 && elect_one()`, a warp-uniform branch plus an elected lane:
// behind a per-thread predicate (`tid == 32`) the compiler wraps every tcgen05.mma in an ELECT / R2UR.BROADCAST / BRA.U.ANY
// loop over the active lanes, ~150 cycles per MMA, and MMA ISSUE - not execution - was the serial part of a step.
__device__ __forceinline__ bool elect_one() {
    uint32_t pred;
    asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}\n" : "=r"(pred));
    return pred != 0;
}
__device__ __forceinline__ void fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tmem_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// issue only; the registers are valid after tmem_wait_ld() + pin()
__device__ __forceinline__ void tmem_ld32_issue(uint32_t taddr, uint32_t (&r)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];\n"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
          "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]),
          "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]),
          "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr)
        : "memory");
}
__device__ __forceinline__ void tmem_ld16_issue(uint32_t taddr, uint32_t (&r)[16]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];\n"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
          "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr)
        : "memory");
}
__device__ __forceinline__ void pin16(uint32_t (&r)[16]) {
#pragma unroll
    for (int i = 0; i < 16; ++i) asm volatile("" : "+r"(r[i]));
}
// keeps the compiler from using registers of an asynchronous tcgen05.ld before the wait that precedes this call
__device__ __forceinline__ void pin(uint32_t (&r)[32]) {
#pragma unroll
    for (int i = 0; i < 32; ++i) asm volatile("" : "+r"(r[i]));
}

// Explicit shared-space accesses.  The working set sits behind a pointer that was aligned by integer arithmetic, so the
// compiler cannot prove the address space and emits generic LD.E / ST.E for plain dereferences (r02 SASS: the bias loads of
// the cell update and the operand stores were generic, their consumers waited on the long scoreboard).
__device__ __forceinline__ float4 lds4(const void* p) {
    float4 v;
    asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(smem_u32(p)) : "memory");
    return v;
}
// Constants written in the prologue (biases, LayerNorm gain / bias): volatile keeps the load below the prologue's barrier and
// in program order (a free-floating load was hoisted across whole chunks and spilled), no memory clobber so that arithmetic
// may move around it; callers issue it one cell ahead of its use.
__device__ __forceinline__ float4 lds4_ro(const void* p) {
    float4 v;
    asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(smem_u32(p)));
    return v;
}
__device__ __forceinline__ void sts16(void* p, uint4 v) {
    asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(smem_u32(p)), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}

// 8 consecutive k of one row -> one 16-byte core-matrix row in the hi image and one in the lo image
__device__ __forceinline__ void split8(const float (&v)[8], uint4& hi4, uint4& lo4) {
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
__device__ __forceinline__ void store_pair(unsigned char* a_hi, unsigned char* a_lo, int row, int chunk, const uint4 hi4, const uint4 lo4) {
    const int off = ((chunk * (kRows / 8) + (row >> 3)) * 8 + (row & 7)) * 16;
    sts16(a_hi + off, hi4);
    sts16(a_lo + off, lo4);
}
__device__ __forceinline__ void store_split8(unsigned char* a_hi, unsigned char* a_lo, int row, int chunk, const float (&v)[8]) {
    uint4 hi4, lo4;
    split8(v, hi4, lo4);
    store_pair(a_hi, a_lo, row, chunk, hi4, lo4);
}

__device__ __forceinline__ void tma_load_4d(void* dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1, int c2, int c3) {
    asm volatile(
        "cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
        ::"r"(smem_u32(dst)), "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
        : "memory");
}

__device__ __forceinline__ void tma_store_4d(const CUtensorMap* map, const void* src, int c0, int c1, int c2, int c3) {
    asm volatile("cp.async.bulk.tensor.4d.global.shared::cta.tile.bulk_group [%0, {%2, %3, %4, %5}], [%1];"
                 ::"l"(map), "r"(smem_u32(src)), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
                 : "memory");
}
__device__ __forceinline__ void tma_reduce_add_4d(const CUtensorMap* map, const void* src, int c0, int c1, int c2, int c3) {
    asm volatile("cp.reduce.async.bulk.tensor.4d.global.shared::cta.add.tile.bulk_group [%0, {%2, %3, %4, %5}], [%1];"
                 ::"l"(map), "r"(smem_u32(src)), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
                 : "memory");
}
__device__ __forceinline__ void tma_store_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void tma_store_wait_read() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }

// LSTM cell with shared reciprocals.  With e_x = 2^(-x log2 e):  sigmoid(x) = 1 / (1 + e_x),  tanh(x) = (1 - e_2x) / (1 + e_2x), so
//   f     = (1 + e_i)(1 + e_2g) / D,   i * tanh(g) = (1 + e_f)(1 - e_2g) / D,   D = (1 + e_i)(1 + e_f)(1 + e_2g)
//   h     = o * tanh(c) = (1 - e_2c) / ((1 + e_o)(1 + e_2c))
// i.e. 5 ex2 + 2 rcp on the XU pipe instead of 5 ex2 + 5 rcp: the cell update is bound by that pipe (profiles/r02_prof_tcp.txt:
// XU 54 % of active cycles at 10 MUFU per cell).  Exponent arguments are clamped (40 / 60) so that the products stay finite: the
// clamped gates differ from the exact ones by < 1e-12.  zs = the gate pre-activations WITHOUT bias, nb = the biases pre-multiplied
// by -log2 e (-2 log2 e for g).  Relative error ~4 ulp, the same order as the one-reciprocal-per-gate form.
constexpr float kLog2e = 1.4426950408889634f;
__device__ __forceinline__ float cell7(float zi, float zf, float zg, float zo, const float4 nb, float& c) {
    const float ei = fast_ex2(fminf(fmaf(zi, -kLog2e, nb.x), 40.0f));
    const float ef = fast_ex2(fminf(fmaf(zf, -kLog2e, nb.y), 40.0f));
    const float eg = fast_ex2(fminf(fmaf(zg, -2.0f * kLog2e, nb.z), 40.0f));
    const float eo = fast_ex2(fminf(fmaf(zo, -kLog2e, nb.w), 60.0f));
    const float pi = 1.0f + ei, pf = 1.0f + ef, pg = 1.0f + eg;
    const float r = fast_rcp(pi * pf * pg);
    const float f = r * (pi * pg);
    const float ig = r * (pf * (1.0f - eg));
    c = fmaf(f, c, ig);
    const float ec = fast_ex2(fminf(c * (-2.0f * kLog2e), 60.0f));
    return (1.0f - ec) * fast_rcp((1.0f + eo) * (1.0f + ec));
}

}  // namespace tcp

// CELL7: the cell update with shared reciprocals (7 MUFU operations per cell instead of 10, see cell7 below)
template <bool CELL7>
__global__ void __launch_bounds__(tcp::kThreads, 1)
lstm_tcp_kernel(const SeqArgs a, const tcp::Geom g, const __grid_constant__ CUtensorMap map_x0,
                const __grid_constant__ CUtensorMap map_x0_tail, const __grid_constant__ CUtensorMap map_x1,
                const __grid_constant__ CUtensorMap map_x1_tail) {
    using namespace tcp;
    extern __shared__ unsigned char sm_raw[];
    unsigned char* sm = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(sm_raw) + 1023) & ~uintptr_t(1023));
    unsigned char* slabs = sm;
    unsigned char* w_hi = sm + kOffW;
    unsigned char* w_lo = w_hi + kWBytes;
    unsigned char* p_hi = w_lo + kWBytes;
    unsigned char* p_lo = p_hi + kPBytes;
    unsigned char* a_hi = sm + kOffA;
    unsigned char* a_lo = a_hi + kABytes;
    float* bias_s = reinterpret_cast<float*>(sm + kOffBias);
    float* ln_s = reinterpret_cast<float*>(sm + kOffLn);
    uint64_t* bars = reinterpret_cast<uint64_t*>(sm + kOffBar);
    uint64_t* full = bars;                                  // [kSlabs]  TMA bytes of a step's stage(s)
    uint64_t* gates = bars + kSlabs;                        // [2]       tcgen05.commit: the gate columns of units 32h .. 32h + 31
    uint64_t* alldone = bars + kSlabs + 4;                  //           tcgen05.commit: every MMA of the step (incl. the projection)
    uint64_t* hready = bars + kSlabs + 5;                   //           8 cell-update warps have published h
    BulkBarrier* wbar = reinterpret_cast<BulkBarrier*>(bars + kSlabs + 6);
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + kSlabs + 7);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int dir = blockIdx.y;
    const sb_lstm_dir& w = a.w[dir];
    const int S = a.n_steps;
    const int q = warp & 3;                                 // TMEM lane quarter this warp may read
    const int r = 32 * q + lane;                            // tile row of this thread
    const int nld = a.x1 ? 2 : 1;                           // stages per step
    const int nslots = kSlabs / nld;

    // ---- tile geometry: which (outer, inner) row this thread owns ----------------------------------------------------
    const int tile = blockIdx.x;
    const bool tail_tile = tile >= g.n_full_tiles;
    int outer0, inner0;                                     // first row of the tile (full tile) / first outer index (tail tile)
    int o_row, i_row;
    bool valid;
    if (!tail_tile) {
        outer0 = tile / g.nfull;
        inner0 = (tile - outer0 * g.nfull) * kRows;
        o_row = outer0; i_row = inner0 + r; valid = true;
    } else {
        outer0 = (tile - g.n_full_tiles) * g.P;
        inner0 = g.nfull * kRows;
        const int qq = r / g.tail;
        o_row = outer0 + qq; i_row = inner0 + (r - qq * g.tail);
        valid = qq < g.P && o_row < g.n_outer;
    }
    const int grow = valid ? o_row * a.rows_inner + i_row : 0;                 // global row: state and FiLM index
    const long long rbase = valid ? (long long)o_row * a.stride_outer + (long long)i_row * a.stride_inner : 0;

    // ---- one-time setup ---------------------------------------------------------------------------------------------
    // CELL7 folds the bias into the exponent argument: columns are (i, f, g, o) per unit, the g column feeds tanh
    for (int i = tid; i < kN; i += kThreads)
        bias_s[i] = CELL7 ? __ldg(w.tc_b + i) * ((i & 3) == 2 ? -2.0f * kLog2e : -kLog2e) : __ldg(w.tc_b + i);
    if (tid < kC) {
        ln_s[tid] = __ldg(w.ln_g + tid);
        ln_s[kC + tid] = __ldg(w.ln_b + tid);
        ln_s[2 * kC + tid] = dir == 0 ? __ldg(w.lin_b + tid) : 0.0f;       // direction 0 adds the projection bias
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(kTmemCols) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    if (tid == 32) {
        for (int i = 0; i < kSlabs; ++i) mbar_init(full + i, 1);
        for (int i = 0; i < 2; ++i) mbar_init(gates + i, 1);
        mbar_init(alldone, 1);
        mbar_init(hready, 8);
        bulk_barrier_init(wbar);                            // includes fence.mbarrier_init
        // the packed operand images (hi / lo gate matrix, hi / lo projection: 106 KB) in one TMA bulk copy (constant data)
        bulk_expect(wbar, 2 * kWBytes + 2 * kPBytes);
        bulk_copy_g2s(reinterpret_cast<float*>(w_hi), w.tc_w, 2 * kWBytes + 2 * kPBytes, wbar);
    }
    pdl_trigger();
    pdl_wait();                                             // nothing a predecessor wrote is touched before this line
    fence_before();
    __syncthreads();
    fence_after();
    const uint32_t tmem = *tmem_slot;
    const uint32_t lane_base = (uint32_t)(32 * q) << 16;    // TMEM address = lane << 16 | column

    if (warp < 4) {
        // =============================================================================================================
        // stream group: TMA producer, LayerNorm / operand builder, MMA issuer, output writer
        // =============================================================================================================
        auto issue_loads = [&](int step) {                  // one thread
            const int pos = dir ? S - 1 - step : step;
            const int slot = step % nslots;
            unsigned char* dst = slabs + (size_t)slot * nld * kSlabBytes;
            uint64_t* bar = full + slot;
            if (!tail_tile) {
                mbar_expect(bar, (uint32_t)(nld * kSlabBytes));
                const int c1 = g.mode == 0 ? pos : inner0, c2 = g.mode == 0 ? inner0 : pos;
                tma_load_4d(dst, &map_x0, bar, 0, c1, c2, outer0);
                if (nld == 2) tma_load_4d(dst + kSlabBytes, &map_x1, bar, 0, c1, c2, outer0);
            } else {
                int nq = g.n_outer - outer0;
                nq = nq < g.P ? nq : g.P;
                mbar_expect(bar, (uint32_t)(nld * nq * g.tail * kC * 4));
                const int c1 = g.mode == 0 ? pos : inner0, c2 = g.mode == 0 ? inner0 : pos;
                for (int qq = 0; qq < nq; ++qq) {
                    tma_load_4d(dst + (size_t)qq * g.tail * kC * 4, &map_x0_tail, bar, 0, c1, c2, outer0 + qq);
                    if (nld == 2) tma_load_4d(dst + kSlabBytes + (size_t)qq * g.tail * kC * 4, &map_x1_tail, bar, 0, c1, c2, outer0 + qq);
                }
            }
        };
        if (tid == 0) {
            asm volatile("prefetch.tensormap [%0];" ::"l"(&map_x0) : "memory");
            const int n0 = S < nslots ? S : nslots;
            for (int j = 0; j < n0; ++j) issue_loads(j);
        }

        const long long film_row = (long long)(grow / a.film_row_div) * S * kC;

        // the row of step `step` out of its stage (128-byte swizzle: 16-byte chunk j of row r sits at chunk j ^ (r & 7)),
        // second addend and FiLM on the way; LayerNorm; bf16 hi/lo into the x part of A; x' (+ projection bias) kept
        auto build = [&](int step, float4 (&keep)[8]) {
            const int pos = dir ? S - 1 - step : step;
            const int slot = step % nslots;
            mbar_wait(full + slot, (uint32_t)((step / nslots) & 1));
            const unsigned char* base = slabs + (size_t)slot * nld * kSlabBytes + (size_t)r * (kC * 4);
            float4 xv[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) xv[i] = lds4(base + ((i ^ (r & 7)) << 4));
            if (nld == 2) {
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    const float4 u4 = lds4(base + kSlabBytes + ((i ^ (r & 7)) << 4));
                    xv[i].x += u4.x; xv[i].y += u4.y; xv[i].z += u4.z; xv[i].w += u4.w;
                }
            }
            if (a.film_scale) {
                const float4* fs = reinterpret_cast<const float4*>(a.film_scale + film_row + (long long)pos * kC);
                const float4* fb = reinterpret_cast<const float4*>(a.film_shift + film_row + (long long)pos * kC);
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    const float4 s4 = __ldg(fs + i), h4 = __ldg(fb + i);
                    xv[i].x = fmaf(xv[i].x, s4.x, h4.x); xv[i].y = fmaf(xv[i].y, s4.y, h4.y);
                    xv[i].z = fmaf(xv[i].z, s4.z, h4.z); xv[i].w = fmaf(xv[i].w, s4.w, h4.w);
                }
            }
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
            for (int ch = 0; ch < 4; ++ch) {
                float v[8];
#pragma unroll
                for (int i = 0; i < 2; ++i) {
                    const float4 t = xv[2 * ch + i];
                    const float4 gg = lds4_ro(ln_s + 4 * (2 * ch + i)), bb = lds4_ro(ln_s + kC + 4 * (2 * ch + i));
                    v[4 * i + 0] = fmaf((t.x - mean) * rstd, gg.x, bb.x); v[4 * i + 1] = fmaf((t.y - mean) * rstd, gg.y, bb.y);
                    v[4 * i + 2] = fmaf((t.z - mean) * rstd, gg.z, bb.z); v[4 * i + 3] = fmaf((t.w - mean) * rstd, gg.w, bb.w);
                }
                store_split8(a_hi, a_lo, r, ch, v);
            }
#pragma unroll
            for (int i = 0; i < 8; ++i) keep[i] = xv[i];
        };

        // ---- MMA issue (one thread) ---------------------------------------------------------------------------------
        const uint32_t a_hi_s = smem_u32(a_hi), a_lo_s = smem_u32(a_lo), w_hi_s = smem_u32(w_hi), w_lo_s = smem_u32(w_lo);
        const uint32_t p_hi_s = smem_u32(p_hi), p_lo_s = smem_u32(p_lo);
        constexpr uint32_t idesc_g = make_idesc(128, kGateN), idesc_p = make_idesc(128, 32);
        // TMEM: two gate buffers of 256 columns (step s accumulates into buffer s & 1); the projection of h_{s-1}, issued with
        // step s, lands in columns 0..31 of the OTHER buffer - the one step s - 1 used, which every cell warp has finished
        // reading - and is read back by this group before the x part of step s + 1 overwrites that buffer.
        // (Dependent MMAs on one accumulator cost ~150 cycles each whatever N is - 12 of them: 0.93 us by globaltimer stamps -
        // but splitting the three terms over three accumulators that are summed at read-back was slower, 658 vs 635 us per 145
        // steps: the projection is not on the serial path of this kernel, the extra tcgen05.ld are.)
        auto issue_proj = [&](uint32_t col) {               // proj[128 x 32] = h (k chunks 4..11 of A) . lin^T
            uint32_t acc = 0;
#pragma unroll
            for (int pass = 0; pass < 3; ++pass) {
                const uint32_t ab = pass == 2 ? a_lo_s : a_hi_s, pb = pass == 1 ? p_lo_s : p_hi_s;
#pragma unroll
                for (int ks = 0; ks < kH / 16; ++ks) {
                    umma(tmem + col, make_desc(ab + (4 + 2 * ks) * kAChunkBytes, kAChunkBytes, 128),
                         make_desc(pb + 2 * ks * kPChunkBytes, kPChunkBytes, 128), idesc_p, acc);
                    acc = 1;
                }
            }
        };
        // gates[128 x 128] for units 32q .. 32q + 31 (rows 128q .. of the image), k steps [ks0, ks1) of the K = 96 contraction:
        // the x part (k steps 0, 1) does not depend on h and is issued one step ahead, the h part (2..5) accumulates onto it
        auto issue_gates = [&](uint32_t col, int q, int ks0, int ks1, bool fresh) {
            uint32_t acc = fresh ? 0u : 1u;
#pragma unroll
            for (int pass = 0; pass < 3; ++pass) {
                const uint32_t ab = pass == 2 ? a_lo_s : a_hi_s, wb = (pass == 1 ? w_lo_s : w_hi_s) + q * 16 * 128;
                for (int ks = ks0; ks < ks1; ++ks) {
                    umma(tmem + col + 128 * q, make_desc(ab + 2 * ks * kAChunkBytes, kAChunkBytes, 128),
                         make_desc(wb + 2 * ks * kWChunkBytes, kWChunkBytes, 128), idesc_g, acc);
                    acc = 1;
                }
            }
        };
        constexpr int kXs = kPreX ? kC / 16 : 0, kKs = kK / 16;  // k steps issued ahead (the x part) / of the whole contraction

        float* const outp = a.out[dir] + rbase;
        auto emit = [&](int step, const float4 (&resv)[8], uint32_t col) {    // y_step = lin h_step [+ b + x'_step]
            uint32_t pr[32];
            tmem_ld32_issue(tmem + lane_base + col, pr);
            tmem_wait_ld();
            pin(pr);
            if (!valid) return;
            const int pos = dir ? S - 1 - step : step;
            float* dst = outp + (long long)pos * a.stride_pos;
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                float4 o = make_float4(__uint_as_float(pr[4 * i]), __uint_as_float(pr[4 * i + 1]), __uint_as_float(pr[4 * i + 2]),
                                       __uint_as_float(pr[4 * i + 3]));
                if (dir == 0) {                             // direction 0 adds bias and residual
                    const float4 bl = lds4_ro(ln_s + 2 * kC + 4 * i);
                    o.x += bl.x + resv[i].x; o.y += bl.y + resv[i].y;
                    o.z += bl.z + resv[i].z; o.w += bl.w + resv[i].w;
                }
                st4(dst + 4 * i, o);
            }
        };

        float4 res_prev[8], res_cur[8];
        build(0, res_cur);
#pragma unroll
        for (int i = 0; i < 8; ++i) res_prev[i] = res_cur[i];
        fence_async_smem();
        for (int s = 0; s < S; ++s) {
            const uint32_t gbuf = 256u * (uint32_t)(s & 1), obuf = 256u - gbuf;
            fence_before();
            bar_sync(1, 128);                               // x part of step s complete, its stage read by all four warps
            if (tid == 0 && s + nslots < S) issue_loads(s + nslots);
            if (warp == 1 && elect_one()) {
                if (s == 0) {
                    bulk_wait(wbar, 0);                     // the operand images have landed
                    fence_after();
                    if (kPreX)
                        for (int q = 0; q < 256 / kGateN; ++q) issue_gates(gbuf, q, 0, kXs, true);     // x part of step 0 (later: one step ahead)
                }
                mbar_wait(hready, (uint32_t)(s & 1));       // h_{s-1} (or h0) is in A, the gate columns have been read
                fence_after();
#pragma unroll
                for (int q = 0; q < 256 / kGateN; ++q) {
                    issue_gates(gbuf, q, kXs, kKs, !kPreX);
                    umma_commit(gates + q);
                }
                if (kGateN == 256) umma_commit(gates + 1);
                if (s > 0) issue_proj(obuf);
                umma_commit(alldone);
            }
            __syncwarp();
            mbar_wait(alldone, (uint32_t)(s & 1));          // every MMA of step s is done: A may be rewritten, proj is ready
            fence_after();
            if (s > 0) emit(s - 1, res_prev, obuf);
#pragma unroll
            for (int i = 0; i < 8; ++i) res_prev[i] = res_cur[i];
            if (s + 1 < S) {
                build(s + 1, res_cur);
                fence_async_smem();
                fence_before();
                bar_sync(2, 128);                           // x part of step s + 1 in A; every warp has read the projection
                if (kPreX && warp == 1 && elect_one()) {    // ... so the other buffer may take the x part of step s + 1 now,
                    fence_after();                          // while the cell warps are still busy with step s
                    for (int q = 0; q < 256 / kGateN; ++q) issue_gates(obuf, q, 0, kXs, true);
                }
            }
        }
        // ---- drain: projection of the last step ------------------------------------------------------------------------
        const uint32_t dbuf = 256u - 256u * (uint32_t)(S & 1);
        fence_before();
        bar_sync(1, 128);
        if (warp == 1 && elect_one()) {
            mbar_wait(hready, (uint32_t)(S & 1));
            fence_after();
            issue_proj(dbuf);
            umma_commit(alldone);
        }
        __syncwarp();
        mbar_wait(alldone, (uint32_t)(S & 1));
        fence_after();
        emit(S - 1, res_prev, dbuf);
    } else {
        // =============================================================================================================
        // cell-update group: thread (row, hf) owns units 32hf + 8k .. + 7 for k = 0..3 (register c[8k + j])
        // =============================================================================================================
        const int hf = (warp - 4) >> 2;
        float c[32];
        if (a.h0 && valid) {
            const float* cp = a.c0 + (long long)grow * kH + 32 * hf;
            const float* hp = a.h0 + (long long)grow * kH + 32 * hf;
#pragma unroll
            for (int ch = 0; ch < 4; ++ch) {                // plain loads: hN / cN may alias h0 / c0
                const float4 c0 = ld_plain4(cp + 8 * ch), c1 = ld_plain4(cp + 8 * ch + 4);
                c[8 * ch] = c0.x; c[8 * ch + 1] = c0.y; c[8 * ch + 2] = c0.z; c[8 * ch + 3] = c0.w;
                c[8 * ch + 4] = c1.x; c[8 * ch + 5] = c1.y; c[8 * ch + 6] = c1.z; c[8 * ch + 7] = c1.w;
                const float4 v0 = ld_plain4(hp + 8 * ch), v1 = ld_plain4(hp + 8 * ch + 4);
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
        fence_async_smem();
        __syncwarp();
        if (lane == 0) mbar_arrive(hready);

        float* const hN = (a.hN && valid) ? a.hN + (long long)grow * kH + 32 * hf : nullptr;
        for (int s = 0; s < S; ++s) {
            const uint32_t par = (uint32_t)(s & 1);
            const uint32_t gcol = tmem + lane_base + 256u * par + 128 * hf;       // gate buffer of this step
            uint4 p_hi = make_uint4(0u, 0u, 0u, 0u), p_lo = p_hi;    // chunk k - 1 as operand rows, stored one chunk late
            mbar_wait(gates + hf, par);
            fence_after();
            uint32_t ga[32], gb[32];
            tmem_ld32_issue(gcol, ga);
            tmem_wait_ld();
            pin(ga);
#pragma unroll
            for (int ch = 0; ch < 4; ++ch) {                // 8 units = 32 TMEM columns at a time
                uint32_t (&cur)[32] = (ch & 1) ? gb : ga;
                uint32_t (&nxt)[32] = (ch & 1) ? ga : gb;
                if (ch < 3) tmem_ld32_issue(gcol + 32 * (ch + 1), nxt);
                const float* bp = bias_s + 4 * (32 * hf + 8 * ch);
                float4 nb_next = lds4_ro(bp);
                float h8[8];
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    const float4 nb = nb_next;
                    if (j < 7) nb_next = lds4_ro(bp + 4 * (j + 1));
                    if (CELL7) {
                        h8[j] = cell7(__uint_as_float(cur[4 * j + 0]), __uint_as_float(cur[4 * j + 1]), __uint_as_float(cur[4 * j + 2]),
                                      __uint_as_float(cur[4 * j + 3]), nb, c[8 * ch + j]);
                    } else {
                        const float ig = sigmoid_f(__uint_as_float(cur[4 * j + 0]) + nb.x), fg = sigmoid_f(__uint_as_float(cur[4 * j + 1]) + nb.y);
                        const float gg = tanh_f(__uint_as_float(cur[4 * j + 2]) + nb.z), og = sigmoid_f(__uint_as_float(cur[4 * j + 3]) + nb.w);
                        c[8 * ch + j] = fmaf(fg, c[8 * ch + j], ig * gg);
                        h8[j] = og * tanh_f(c[8 * ch + j]);
                    }
                }
                if (s == S - 1 && hN) {
                    st4(hN + 8 * ch, make_float4(h8[0], h8[1], h8[2], h8[3]));
                    st4(hN + 8 * ch + 4, make_float4(h8[4], h8[5], h8[6], h8[7]));
                }
                // Nothing may be written into A before every MMA of the step has read it (the first half's warps get here while
                // the second half's MMAs may still be reading A): chunk k goes out after chunk k + 1 has been evaluated.
                if (ch == 1) mbar_wait(alldone, par);
                if (ch > 0) store_pair(a_hi, a_lo, r, 4 + 4 * hf + ch - 1, p_hi, p_lo);
                split8(h8, p_hi, p_lo);
                if (ch < 3) {
                    tmem_wait_ld();
                    pin(nxt);
                }
            }
            store_pair(a_hi, a_lo, r, 4 + 4 * hf + 3, p_hi, p_lo);
            fence_before();
            fence_async_smem();
            __syncwarp();
            if (lane == 0) mbar_arrive(hready);
        }
        if (a.cN && valid) {
            float* cp = a.cN + (long long)grow * kH + 32 * hf;
#pragma unroll
            for (int ch = 0; ch < 4; ++ch) {
                st4(cp + 8 * ch, make_float4(c[8 * ch], c[8 * ch + 1], c[8 * ch + 2], c[8 * ch + 3]));
                st4(cp + 8 * ch + 4, make_float4(c[8 * ch + 4], c[8 * ch + 5], c[8 * ch + 6], c[8 * ch + 7]));
            }
        }
    }
    fence_before();
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(kTmemCols) : "memory");
}

// =================================================================================================================
// lstm_tcr_kernel: lstm_tcp_kernel with the recurrence pipelined at the granularity of one k step (single-addend calls:
// the intra-frame LSTMs, 78 % of the headline's SM time).
//
// In lstm_tcp_kernel a step is a chain  [12 h-part MMAs, ~1.0 us] -> [cell update of 64 units, ~3.0 us] -> handshake.  But k step
// j of the h part only reads units 16j .. 16j + 15 of h, so here
//   * a cell thread (row, half) works through its units in the order 16k + 8 half + (0..7), k = 0..3, and publishes every
//     chunk at once: after chunk k of both halves the 16 units of k step k are in A (mbarrier hk[k], 8 warps);
//   * a 13th warp only issues MMAs: it waits for hk[k] and issues the three MMAs (hi.hi, hi.lo, lo.hi) of gate k step k of
//     step s + 1 into the OTHER gate buffer, plus the three of projection k step k of h_s into columns 0..31 of the buffer
//     the cell warps are reading (free as soon as hk[0] says chunk 0 = columns 0..63 has been loaded).  Only the three gate
//     MMAs of k step 3 remain between the end of a cell update and the start of the next;
//   * the stream warps keep x' (input + FiLM, the residual of direction 0) in the TMA stage it came from instead of in
//     registers, prepare the bf16 images of step s + 2 in registers while step s runs, and at the start of a step only
//     read the projection back, write the output rows and store the prepared images, so the x part of the next step is
//     in TMEM long before its h part arrives.
// The accumulation order of a gate differs from lstm_tcp_kernel (k-step-major instead of term-major), so results agree
// with it to rounding (1e-6), not bitwise.
// =================================================================================================================
// globaltimer stamps of one CTA's roles (debug builds only, -DSB_TCQ_DEBUG; tools/tcq_timeline.py)
#ifdef SB_TCQ_DEBUG
// one slot per (step < 8, warp < 16, event < 32), written with plain stores: an atomic counter would stall the stamping lane for
// a global round trip (~0.5 us) and distort the very chain that is being looked at
__device__ long long g_tcq_dbg[16384];
#define TCQ_STAMP(role, ev, X, step)                                                                       \
    do {                                                                                                   \
        if (blockIdx.x == 0 && blockIdx.y == 0 && (step) >= 0 && (step) < 8 && (threadIdx.x & 31) == 0 && (threadIdx.x >> 5) < 16) { \
            long long t_;                                                                                  \
            asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t_));                                         \
            const int i_ = (((step) * 16 + (threadIdx.x >> 5)) * 32 + (ev));                               \
            g_tcq_dbg[4 * i_] = t_; g_tcq_dbg[4 * i_ + 1] = (role) * 100 + (ev); g_tcq_dbg[4 * i_ + 2] = (X); g_tcq_dbg[4 * i_ + 3] = (step) * 100 + (threadIdx.x >> 5); \
        }                                                                                                  \
    } while (0)
#else
#define TCQ_STAMP(role, ev, X, step) do {} while (0)
#endif

namespace tcr {
// 4 stream warps, NCW cell-update warps, 1 MMA-issue warp.  NCW = 16: a thread owns a QUARTER of a row's units (4 units = 16 gate
// columns per chunk), four cell warps per scheduler; six warpgroups (the last one holds the issue warp and three idle warps)
// compiled for 80 registers, setmaxnreg moves the issue group's to the stream group (24 / 128).
constexpr int threads(int ncw) { return ncw == 8 ? 416 : 768; }
}

__device__ __forceinline__ void sts8(void* p, uint32_t x, uint32_t y) {
    asm volatile("st.shared.v2.b32 [%0], {%1, %2};" ::"r"(tcp::smem_u32(p)), "r"(x), "r"(y) : "memory");
}

template <int NCW>
__global__ void __launch_bounds__(tcr::threads(NCW), 1)
lstm_tcr_kernel(const SeqArgs a, const tcp::Geom g, const __grid_constant__ CUtensorMap map_x0,
                const __grid_constant__ CUtensorMap map_x0_tail, const __grid_constant__ CUtensorMap map_o0,
                const __grid_constant__ CUtensorMap map_o0_tail, const __grid_constant__ CUtensorMap map_o1,
                const __grid_constant__ CUtensorMap map_o1_tail) {
    using namespace tcp;
    extern __shared__ unsigned char sm_raw[];
    unsigned char* sm = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(sm_raw) + 1023) & ~uintptr_t(1023));
    unsigned char* slabs = sm;
    unsigned char* w_hi = sm + kOffW;
    unsigned char* w_lo = w_hi + kWBytes;
    unsigned char* p_hi = w_lo + kWBytes;
    unsigned char* p_lo = p_hi + kPBytes;
    unsigned char* a_hi = sm + kOffA;
    unsigned char* a_lo = a_hi + kABytes;
    float* bias_s = reinterpret_cast<float*>(sm + kOffBias);
    float* ln_s = reinterpret_cast<float*>(sm + kOffLn);
    uint64_t* bars = reinterpret_cast<uint64_t*>(sm + kOffBar);
    uint64_t* full = bars;                                  // [kSlabs]  TMA bytes of a step's stage
    uint64_t* gates = bars + kSlabs;                        //           tcgen05.commit: the gates of a step (and every MMA before them)
    uint64_t* pdone = bars + kSlabs + 1;                    //           tcgen05.commit: the projection of a step's h
    uint64_t* hk = bars + kSlabs + 2;                       // [4]       8 cell warps have published units 16k .. 16k + 15 of h
    uint64_t* xready = bars + kSlabs + 6;                   //           stream group: projection read back, next x part in A
    BulkBarrier* wbar = reinterpret_cast<BulkBarrier*>(bars + kSlabs + 7);
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + kSlabs + 8);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int dir = blockIdx.y;
    const sb_lstm_dir& w = a.w[dir];
    const int S = a.n_steps;
    const int q = warp & 3;                                 // TMEM lane quarter this warp may read
    const int r = 32 * q + lane;                            // tile row of this thread

    const int tile = blockIdx.x;
    const bool tail_tile = tile >= g.n_full_tiles;
    int outer0, inner0, o_row, i_row;
    bool valid;
    if (!tail_tile) {
        outer0 = tile / g.nfull;
        inner0 = (tile - outer0 * g.nfull) * kRows;
        o_row = outer0; i_row = inner0 + r; valid = true;
    } else {
        outer0 = (tile - g.n_full_tiles) * g.P;
        inner0 = g.nfull * kRows;
        const int qq = r / g.tail;
        o_row = outer0 + qq; i_row = inner0 + (r - qq * g.tail);
        valid = qq < g.P && o_row < g.n_outer;
    }
    const int grow = valid ? o_row * a.rows_inner + i_row : 0;
    const long long rbase = valid ? (long long)o_row * a.stride_outer + (long long)i_row * a.stride_inner : 0;

    for (int i = tid; i < kN; i += tcr::threads(NCW))          // bias folded into the exponent argument, see cell7
        bias_s[i] = __ldg(w.tc_b + i) * ((i & 3) == 2 ? -2.0f * kLog2e : -kLog2e);
    if (tid < kC) {
        ln_s[tid] = __ldg(w.ln_g + tid);
        ln_s[kC + tid] = __ldg(w.ln_b + tid);
        ln_s[2 * kC + tid] = dir == 0 ? __ldg(w.lin_b + tid) : 0.0f;
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(kTmemCols) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    if (tid == 32) {
        for (int i = 0; i < kSlabs; ++i) mbar_init(full + i, 1);
        mbar_init(gates, 1);
        mbar_init(pdone, 1);
        for (int k = 0; k < 4; ++k) mbar_init(hk + k, NCW);
        mbar_init(xready, 1);
        bulk_barrier_init(wbar);
        bulk_expect(wbar, 2 * kWBytes + 2 * kPBytes);
        bulk_copy_g2s(reinterpret_cast<float*>(w_hi), w.tc_w, 2 * kWBytes + 2 * kPBytes, wbar);
    }
    pdl_trigger();
    pdl_wait();
    fence_before();
    __syncthreads();
    fence_after();
    const uint32_t tmem = *tmem_slot;
    const uint32_t lane_base = (uint32_t)(32 * q) << 16;

    if (warp < 4) {
        // =============================================================================================================
        // stream group: TMA producer, LayerNorm / operand builder, output writer
        // =============================================================================================================
        if (NCW == 16) asm volatile("setmaxnreg.inc.sync.aligned.u32 120;");
        auto issue_loads = [&](int step) {                  // one thread
            const int pos = dir ? S - 1 - step : step;
            unsigned char* dst = slabs + (size_t)(step % kSlabs) * kSlabBytes;
            uint64_t* bar = full + step % kSlabs;
            const int c1 = g.mode == 0 ? pos : inner0, c2 = g.mode == 0 ? inner0 : pos;
            if (!tail_tile) {
                mbar_expect(bar, (uint32_t)kSlabBytes);
                tma_load_4d(dst, &map_x0, bar, 0, c1, c2, outer0);
            } else {
                int nq = g.n_outer - outer0;
                nq = nq < g.P ? nq : g.P;
                mbar_expect(bar, (uint32_t)(nq * g.tail * kC * 4));
                for (int qq = 0; qq < nq; ++qq) tma_load_4d(dst + (size_t)qq * g.tail * kC * 4, &map_x0_tail, bar, 0, c1, c2, outer0 + qq);
            }
        };
        if (tid == 0) {
            asm volatile("prefetch.tensormap [%0];" ::"l"(&map_x0) : "memory");
            const int n0 = S < kSlabs ? S : kSlabs;
            for (int j = 0; j < n0; ++j) issue_loads(j);
        }
        const long long film_row = (long long)(grow / a.film_row_div) * S * kC;
        const int swz = r & 7;                              // 128-byte swizzle: 16-byte chunk j of row r sits at chunk j ^ (r & 7)

        // the row of step `step` out of its stage, FiLM on the way; x' goes BACK into the stage (direction 0 adds it to the
        // output one step after the step has run); LayerNorm; bf16 hi / lo of the four k chunks stay in registers
        uint4 xh[4], xl[4];
        auto prepare = [&](int step) {
            const int pos = dir ? S - 1 - step : step;
            const int slot = step % kSlabs;
            mbar_wait(full + slot, (uint32_t)((step / kSlabs) & 1));
            unsigned char* base = slabs + (size_t)slot * kSlabBytes + (size_t)r * (kC * 4);
            float4 xv[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) xv[i] = lds4(base + ((i ^ swz) << 4));
            if (a.film_scale) {
                const float4* fs = reinterpret_cast<const float4*>(a.film_scale + film_row + (long long)pos * kC);
                const float4* fb = reinterpret_cast<const float4*>(a.film_shift + film_row + (long long)pos * kC);
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    const float4 s4 = __ldg(fs + i), h4 = __ldg(fb + i);
                    xv[i].x = fmaf(xv[i].x, s4.x, h4.x); xv[i].y = fmaf(xv[i].y, s4.y, h4.y);
                    xv[i].z = fmaf(xv[i].z, s4.z, h4.z); xv[i].w = fmaf(xv[i].w, s4.w, h4.w);
                }
                if (dir == 0) {
#pragma unroll
                    for (int i = 0; i < 8; ++i)
                        sts16(base + ((i ^ swz) << 4), make_uint4(__float_as_uint(xv[i].x), __float_as_uint(xv[i].y),
                                                                  __float_as_uint(xv[i].z), __float_as_uint(xv[i].w)));
                }
            }
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
            for (int ch = 0; ch < 4; ++ch) {
                float v[8];
#pragma unroll
                for (int i = 0; i < 2; ++i) {
                    const float4 t = xv[2 * ch + i];
                    const float4 gg = lds4_ro(ln_s + 4 * (2 * ch + i)), bb = lds4_ro(ln_s + kC + 4 * (2 * ch + i));
                    v[4 * i + 0] = fmaf((t.x - mean) * rstd, gg.x, bb.x); v[4 * i + 1] = fmaf((t.y - mean) * rstd, gg.y, bb.y);
                    v[4 * i + 2] = fmaf((t.z - mean) * rstd, gg.z, bb.z); v[4 * i + 3] = fmaf((t.w - mean) * rstd, gg.w, bb.w);
                }
                split8(v, xh[ch], xl[ch]);
            }
        };
        auto publish = [&]() {
#pragma unroll
            for (int ch = 0; ch < 4; ++ch) store_pair(a_hi, a_lo, r, ch, xh[ch], xl[ch]);
        };
        // y_step = lin h_step [+ b + x'_step] goes into the stage x'_step sits in (same swizzled position, own row) and leaves
        // as ONE TMA tensor store per step: written from registers, the rows of a warp are 32 different 128-byte lines per
        // st.global.v4 - 1024 L1 wavefronts per step that the cell warps' LDS / STS / tcgen05.ld queued behind (globaltimer
        // timeline: a chunk of the cell update took 1.1 us while these stores were in flight, 0.62 us otherwise)
        auto emit = [&](int step, uint32_t col) {
            uint32_t pr[32];
            tmem_ld32_issue(tmem + lane_base + col, pr);
            tmem_wait_ld();
            pin(pr);
            unsigned char* base = slabs + (size_t)(step % kSlabs) * kSlabBytes + (size_t)r * (kC * 4);
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                float4 o = make_float4(__uint_as_float(pr[4 * i]), __uint_as_float(pr[4 * i + 1]), __uint_as_float(pr[4 * i + 2]),
                                       __uint_as_float(pr[4 * i + 3]));
                if (dir == 0) {
                    const float4 bl = lds4_ro(ln_s + 2 * kC + 4 * i), xr = lds4(base + ((i ^ swz) << 4));
                    o.x += bl.x + xr.x; o.y += bl.y + xr.y;
                    o.z += bl.z + xr.z; o.w += bl.w + xr.w;
                }
                sts16(base + ((i ^ swz) << 4), make_uint4(__float_as_uint(o.x), __float_as_uint(o.y), __float_as_uint(o.z), __float_as_uint(o.w)));
            }
        };
        auto store_out = [&](int step) {                     // one thread, after the barrier that follows emit(step)
            const int pos = dir ? S - 1 - step : step;
            const unsigned char* src = slabs + (size_t)(step % kSlabs) * kSlabBytes;
            const int c1 = g.mode == 0 ? pos : inner0, c2 = g.mode == 0 ? inner0 : pos;
            // SeqArgs::sum_dirs: both directions ADD into the (zeroed) buffer of direction 0 - 0 + a + b does not depend on the order
            const CUtensorMap* mo = tail_tile ? ((dir && !a.sum_dirs) ? &map_o1_tail : &map_o0_tail) : ((dir && !a.sum_dirs) ? &map_o1 : &map_o0);
            int nq = 1;
            if (tail_tile) {
                nq = g.n_outer - outer0;
                nq = nq < g.P ? nq : g.P;
            }
            const size_t box_bytes = tail_tile ? (size_t)g.tail * kC * 4 : 0;
            for (int qq = 0; qq < nq; ++qq) {
                if (a.sum_dirs) tma_reduce_add_4d(mo, src + qq * box_bytes, 0, c1, c2, outer0 + qq);
                else tma_store_4d(mo, src + qq * box_bytes, 0, c1, c2, outer0 + qq);
            }
            tma_store_commit();
        };

        prepare(0);
        publish();
        fence_async_smem();
        bar_sync(1, 128);
        if (tid == 0) mbar_arrive(xready);                  // phase 0: x_0 is in A
        if (S > 1) prepare(1);
        for (int i = 0; i < S; ++i) {
            // gates(i) complete = the x-part MMAs of step i have read A; pdone(i - 1) is committed after gates(i)
            TCQ_STAMP(4, 0, 0, i);
            if (i == 0) mbar_wait(gates, 0u);
            else mbar_wait(pdone, (uint32_t)((i - 1) & 1));
            fence_after();
            TCQ_STAMP(4, 1, 0, i);
            if (i > 0) emit(i - 1, 256u * (uint32_t)((i - 1) & 1));
            TCQ_STAMP(4, 2, 0, i);
            if (i + 1 < S) publish();
            fence_async_smem();
            fence_before();
            bar_sync(1, 128);                               // every row: y_{i-1} in its stage, x_{i+1} in A
            if (tid == 0) {
                mbar_arrive(xready);                        // phase i + 1
                if (i > 0) store_out(i - 1);
            }
            TCQ_STAMP(4, 3, 0, i);
            if (i + 2 < S) prepare(i + 2);
            if (tid == 0 && i > 0) {
                tma_store_wait_read();                      // the stage of step i - 1 has been read by the store: it takes step i + 3
                if (i + 3 < S) issue_loads(i + 3);
            }
            TCQ_STAMP(4, 4, 0, i);
        }
        mbar_wait(pdone, (uint32_t)((S - 1) & 1));
        fence_after();
        emit(S - 1, 256u * (uint32_t)((S - 1) & 1));
        fence_async_smem();
        bar_sync(1, 128);
        if (tid == 0) {
            store_out(S - 1);
            tma_store_wait_read();
        }
    } else if (warp >= 4 + NCW) {
        // =============================================================================================================
        // MMA issue: iteration i (while the cell warps update step i) builds the gates of step i + 1 and the projection of h_i
        // =============================================================================================================
        if (NCW == 16) asm volatile("setmaxnreg.dec.sync.aligned.u32 40;");
        if (warp == 4 + NCW && elect_one()) {
            uint32_t a_hi_s = smem_u32(a_hi), a_lo_s = smem_u32(a_lo), w_hi_s = smem_u32(w_hi), w_lo_s = smem_u32(w_lo);
            uint32_t p_hi_s = smem_u32(p_hi), p_lo_s = smem_u32(p_lo);
            constexpr uint32_t idesc_g = make_idesc(128, 256), idesc_p = make_idesc(128, 32);
            bulk_wait(wbar, 0);                             // the operand images have landed
            fence_after();
            for (int i = -1; i < S; ++i) {
                // opaque bases: the ~70 descriptors of an iteration are rebuilt from six registers every time instead of being
                // hoisted out of the loop (this warp has registers to spare only in the 8-cell-warp form, and nothing else to do)
                asm volatile("" : "+r"(a_hi_s), "+r"(a_lo_s), "+r"(w_hi_s), "+r"(w_lo_s), "+r"(p_hi_s), "+r"(p_lo_s));
                const uint32_t par = (uint32_t)((i + 1) & 1);
                const uint32_t nbuf = tmem + 256u * par, pbuf = tmem + 256u - 256u * par;       // buffer of step i + 1 / of step i
                const bool more = i + 1 < S;
                mbar_wait(xready, par);                     // x_{i+1} in A; projection of h_{i-1} read out of nbuf
                fence_after();
                TCQ_STAMP(6, 0, 0, i + 1);
                if (more) {
#pragma unroll
                    for (int ks = 0; ks < kC / 16; ++ks)
#pragma unroll
                        for (int pass = 0; pass < 3; ++pass)
                            umma(nbuf, make_desc((pass == 2 ? a_lo_s : a_hi_s) + 2 * ks * kAChunkBytes, kAChunkBytes, 128),
                                 make_desc((pass == 1 ? w_lo_s : w_hi_s) + 2 * ks * kWChunkBytes, kWChunkBytes, 128), idesc_g,
                                 (ks | pass) ? 1u : 0u);
                }
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    if (k == 0) TCQ_STAMP(6, 1, 0, i + 1);
                    mbar_wait(hk + k, par);                 // units 16k .. 16k + 15 of h_i (i = -1: the initial state) are in A
                    fence_after();
                    TCQ_STAMP(6, 10 + k, 0, i + 1);
                    if (more) {
#pragma unroll
                        for (int pass = 0; pass < 3; ++pass)
                            umma(nbuf, make_desc((pass == 2 ? a_lo_s : a_hi_s) + (4 + 2 * k) * kAChunkBytes, kAChunkBytes, 128),
                                 make_desc((pass == 1 ? w_lo_s : w_hi_s) + (4 + 2 * k) * kWChunkBytes, kWChunkBytes, 128), idesc_g, 1u);
                        if (k == 3) umma_commit(gates);
                    }
                    if (i >= 0) {
#pragma unroll
                        for (int pass = 0; pass < 3; ++pass)
                            umma(pbuf, make_desc((pass == 2 ? a_lo_s : a_hi_s) + (4 + 2 * k) * kAChunkBytes, kAChunkBytes, 128),
                                 make_desc((pass == 1 ? p_lo_s : p_hi_s) + 2 * k * kPChunkBytes, kPChunkBytes, 128), idesc_p,
                                 (k | pass) ? 1u : 0u);
                    }
                }
                if (i >= 0) umma_commit(pdone);
                TCQ_STAMP(6, 30, 0, i + 1);
            }
        }
        __syncwarp();
    } else if constexpr (NCW == 8) {
        // =============================================================================================================
        // cell-update group: thread (row, half), chunk k = units 16k + 8 half .. + 7 (register c[8k + j])
        // =============================================================================================================
        const int half = (warp - 4) >> 2;
        float c[32];
        if (a.h0 && valid) {
            const float* cp = a.c0 + (long long)grow * kH + 8 * half;
            const float* hp = a.h0 + (long long)grow * kH + 8 * half;
#pragma unroll
            for (int ch = 0; ch < 4; ++ch) {                // plain loads: hN / cN may alias h0 / c0
                const float4 c0 = ld_plain4(cp + 16 * ch), c1 = ld_plain4(cp + 16 * ch + 4);
                c[8 * ch] = c0.x; c[8 * ch + 1] = c0.y; c[8 * ch + 2] = c0.z; c[8 * ch + 3] = c0.w;
                c[8 * ch + 4] = c1.x; c[8 * ch + 5] = c1.y; c[8 * ch + 6] = c1.z; c[8 * ch + 7] = c1.w;
                const float4 v0 = ld_plain4(hp + 16 * ch), v1 = ld_plain4(hp + 16 * ch + 4);
                const float h8[8] = {v0.x, v0.y, v0.z, v0.w, v1.x, v1.y, v1.z, v1.w};
                store_split8(a_hi, a_lo, r, 4 + 2 * ch + half, h8);
            }
        } else {
#pragma unroll
            for (int j = 0; j < 32; ++j) c[j] = 0.0f;
            const float h8[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
#pragma unroll
            for (int ch = 0; ch < 4; ++ch) store_split8(a_hi, a_lo, r, 4 + 2 * ch + half, h8);
        }
        fence_async_smem();
        __syncwarp();
        if (lane == 0)
            for (int k = 0; k < 4; ++k) mbar_arrive(hk + k);

        float* const hN = (a.hN && valid) ? a.hN + (long long)grow * kH + 8 * half : nullptr;
        for (int s = 0; s < S; ++s) {
            const uint32_t par = (uint32_t)(s & 1);
            const uint32_t gcol = tmem + lane_base + 256u * par + 32u * half;     // chunk k: 32 columns at + 64 k
            uint32_t ga[32], gb[32];
            mbar_wait(gates, par);
            fence_after();
            TCQ_STAMP(5, 0, 0, s);
            tmem_ld32_issue(gcol, ga);
            float4 nb_next = lds4_ro(bias_s + 4 * (8 * half));
            tmem_wait_ld();
            pin(ga);
            TCQ_STAMP(5, 10, 0, s);
            // One straight line of 32 cells.  The asynchronous pieces are placed so that no chunk boundary drains the MUFU
            // pipeline: the columns of chunk k + 1 are requested at the top of chunk k and waited for in its middle, the first
            // bias of chunk k + 1 is loaded before chunk k is stored, the fence + arrive of chunk k sit in the middle of k + 1.
#pragma unroll
            for (int ch = 0; ch < 4; ++ch) {
                uint32_t (&cur)[32] = (ch & 1) ? gb : ga;
                uint32_t (&nxt)[32] = (ch & 1) ? ga : gb;
                if (ch < 3) tmem_ld32_issue(gcol + 64 * (ch + 1), nxt);
                const float* bp = bias_s + 4 * (16 * ch + 8 * half);
                float h8[8];
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    const float4 nb = nb_next;
                    if (j < 7) nb_next = lds4_ro(bp + 4 * (j + 1));
                    else if (ch < 3) nb_next = lds4_ro(bp + 4 * 16);
                    h8[j] = cell7(__uint_as_float(cur[4 * j + 0]), __uint_as_float(cur[4 * j + 1]), __uint_as_float(cur[4 * j + 2]),
                                  __uint_as_float(cur[4 * j + 3]), nb, c[8 * ch + j]);
                    if (j == 3 && ch > 0) {
                        // chunk ch - 1 was stored before this chunk began (only hk[3] is on the serial path of a step)
                        fence_before();
                        fence_async_smem();
                        __syncwarp();
                        TCQ_STAMP(5, ch, 0, s);
                        if (lane == 0) mbar_arrive(hk + ch - 1);
                    }
                    if (j == 4 && ch < 3) {
                        tmem_wait_ld();
                        pin(nxt);
                    }
                }
                if (s == S - 1 && hN) {
                    st4(hN + 16 * ch, make_float4(h8[0], h8[1], h8[2], h8[3]));
                    st4(hN + 16 * ch + 4, make_float4(h8[4], h8[5], h8[6], h8[7]));
                }
                TCQ_STAMP(5, 11 + 3 * ch, 0, s);
                // the last projection k step of h_{s-1} reads k chunks 10, 11 of A (every other reader of h_{s-1} is older than
                // gates(s)); it was issued right after the gates of this step, 3 us ago
                if (ch == 3 && s > 0) mbar_wait(pdone, (uint32_t)((s - 1) & 1));
                store_split8(a_hi, a_lo, r, 4 + 2 * ch + half, h8);
                TCQ_STAMP(5, 12 + 3 * ch, 0, s);
            }
            fence_before();
            fence_async_smem();
            __syncwarp();
            TCQ_STAMP(5, 4, 0, s);
            if (lane == 0) mbar_arrive(hk + 3);             // all of h_s is in A, every gate column of the step has been read
        }
        if (a.cN && valid) {
            float* cp = a.cN + (long long)grow * kH + 8 * half;
#pragma unroll
            for (int ch = 0; ch < 4; ++ch) {
                st4(cp + 16 * ch, make_float4(c[8 * ch], c[8 * ch + 1], c[8 * ch + 2], c[8 * ch + 3]));
                st4(cp + 16 * ch + 4, make_float4(c[8 * ch + 4], c[8 * ch + 5], c[8 * ch + 6], c[8 * ch + 7]));
            }
        }
    } else {
        // =============================================================================================================
        // cell-update group, 16 warps: thread (row, quarter), chunk k = units 16k + 4 quarter .. + 3 (register c[4k + j])
        // =============================================================================================================
        const int qt = (warp - 4) >> 2;
        float c[16];
        auto store4 = [&](int ch, const float (&h4)[4]) {    // 4 units = half a 16-byte core-matrix row of the hi and the lo image
            const __nv_bfloat162 a0 = __floats2bfloat162_rn(h4[0], h4[1]), a1 = __floats2bfloat162_rn(h4[2], h4[3]);
            const __nv_bfloat162 l0 = __floats2bfloat162_rn(h4[0] - __bfloat162float(a0.x), h4[1] - __bfloat162float(a0.y));
            const __nv_bfloat162 l1 = __floats2bfloat162_rn(h4[2] - __bfloat162float(a1.x), h4[3] - __bfloat162float(a1.y));
            const int chunk = 4 + 2 * ch + (qt >> 1);
            const int off = ((chunk * (kRows / 8) + (r >> 3)) * 8 + (r & 7)) * 16 + 8 * (qt & 1);
            sts8(a_hi + off, *reinterpret_cast<const uint32_t*>(&a0), *reinterpret_cast<const uint32_t*>(&a1));
            sts8(a_lo + off, *reinterpret_cast<const uint32_t*>(&l0), *reinterpret_cast<const uint32_t*>(&l1));
        };
        if (a.h0 && valid) {
            const float* cp = a.c0 + (long long)grow * kH + 4 * qt;
            const float* hp = a.h0 + (long long)grow * kH + 4 * qt;
#pragma unroll
            for (int ch = 0; ch < 4; ++ch) {                // plain loads: hN / cN may alias h0 / c0
                const float4 c0 = ld_plain4(cp + 16 * ch), v0 = ld_plain4(hp + 16 * ch);
                c[4 * ch] = c0.x; c[4 * ch + 1] = c0.y; c[4 * ch + 2] = c0.z; c[4 * ch + 3] = c0.w;
                const float h4[4] = {v0.x, v0.y, v0.z, v0.w};
                store4(ch, h4);
            }
        } else {
#pragma unroll
            for (int j = 0; j < 16; ++j) c[j] = 0.0f;
            const float h4[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
            for (int ch = 0; ch < 4; ++ch) store4(ch, h4);
        }
        fence_async_smem();
        __syncwarp();
        if (lane == 0)
            for (int k = 0; k < 4; ++k) mbar_arrive(hk + k);

        float* const hN = (a.hN && valid) ? a.hN + (long long)grow * kH + 4 * qt : nullptr;
        for (int s = 0; s < S; ++s) {
            const uint32_t par = (uint32_t)(s & 1);
            const uint32_t gcol = tmem + lane_base + 256u * par + 16u * qt;       // chunk k: 16 columns at + 64 k
            mbar_wait(gates, par);
            fence_after();
#pragma unroll
            for (int ch = 0; ch < 4; ++ch) {
                uint32_t cur[16];
                tmem_ld16_issue(gcol + 64 * ch, cur);
                const float* bp = bias_s + 4 * (16 * ch + 4 * qt);
                const float4 nb0 = lds4_ro(bp), nb1 = lds4_ro(bp + 4), nb2 = lds4_ro(bp + 8), nb3 = lds4_ro(bp + 12);
                tmem_wait_ld();
                pin16(cur);
                float h4[4];
                h4[0] = cell7(__uint_as_float(cur[0]), __uint_as_float(cur[1]), __uint_as_float(cur[2]), __uint_as_float(cur[3]), nb0, c[4 * ch]);
                h4[1] = cell7(__uint_as_float(cur[4]), __uint_as_float(cur[5]), __uint_as_float(cur[6]), __uint_as_float(cur[7]), nb1, c[4 * ch + 1]);
                h4[2] = cell7(__uint_as_float(cur[8]), __uint_as_float(cur[9]), __uint_as_float(cur[10]), __uint_as_float(cur[11]), nb2, c[4 * ch + 2]);
                h4[3] = cell7(__uint_as_float(cur[12]), __uint_as_float(cur[13]), __uint_as_float(cur[14]), __uint_as_float(cur[15]), nb3, c[4 * ch + 3]);
                if (s == S - 1 && hN) st4(hN + 16 * ch, make_float4(h4[0], h4[1], h4[2], h4[3]));
                if (ch == 3 && s > 0) mbar_wait(pdone, (uint32_t)((s - 1) & 1));     // see the 8-warp form
                store4(ch, h4);
                fence_before();
                fence_async_smem();
                __syncwarp();
                if (lane == 0) mbar_arrive(hk + ch);
            }
        }
        if (a.cN && valid) {
            float* cp = a.cN + (long long)grow * kH + 4 * qt;
#pragma unroll
            for (int ch = 0; ch < 4; ++ch) st4(cp + 16 * ch, make_float4(c[4 * ch], c[4 * ch + 1], c[4 * ch + 2], c[4 * ch + 3]));
        }
    }
    fence_before();
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(kTmemCols) : "memory");
}

// =================================================================================================================
// lstm_tcq_kernel: TWO 128-row tiles per CTA in ping-pong (SB_ALGO_TCQ; chosen for SB_ALGO_TC when a direction has at least
// two tiles).  In lstm_tcp_kernel a step is [MMAs that need the complete h: 0.8 us] + [handshakes: 0.6 us] + [cell update:
// 3 us] in series, and the tensor pipe idles while the cell warps work.  Here the eight cell warps alternate between tile A
// and tile B and never wait: while they update the cells of one tile, the stream group and the tensor pipe prepare the other
// (projection of the previous h -> output rows, x part of the next step, gate MMAs), all of which fits in one cell phase.
//
// What had to give for two tiles to fit one SM:
//   * shared memory (227 KB): the gate / projection images are shared (104 KB), each tile owns its A operand (2 x 48 KB), and
//     the TMA ring shrinks to ONE 16 KB stage the tiles take turns on (a build every 3 us leaves the next load ample time);
//     calls with two addends (the inter-frame path) would need a 32 KB stage and stay on lstm_tcp_kernel;
//   * TMEM (512 columns): 256 gate columns per tile and no room for the 32 projection columns, so the projection of h_{s-1}
//     goes into columns 0..31 of the tile's own gate buffer right after the cell warps have released it, the stream group
//     reads it back, and only then the gate MMAs of the next step overwrite the buffer (no x part ahead of time);
//   * registers: a cell thread carries c of both tiles (64 registers), so the gate columns are loaded single-buffered; the
//     stream threads re-read x' for the residual (FiLM / second addend re-applied) instead of holding it across a step.
// Same arithmetic and operand images as lstm_tcp_kernel<true> (the terms of the split are accumulated pass-major here: results
// agree to 1e-6, tools/tcp_check.py tcq).  Build index j = 2 * step + tile orders everything the stream group does; every
// wait is bounded and traps.
//
// STATUS (round 2): correct on the B200 (oracle parity 2e-6, state hand-over, odd tile counts, tail tiles; tests/test_gpu_parity.py)
// but NOT faster than lstm_tcp_kernel, so nothing selects it by default (SB_ALGO_TCQ only).  Measured, us per 145 steps for two
// tiles (2 x 635 on lstm_tcp_kernel):
//   1 680  MMA issue by a lane of the stream group.  The globaltimer timeline of one CTA (tools/tcq_timeline.py, build with
//          -DSB_TCQ_DEBUG; profiles/r02_tcq_timeline.txt) shows the cell phases at 3.0 us as planned, but the issuing lane blocks
//          for as long as the MMAs EXECUTE (the queue is a few instructions deep; 18 gate MMAs: 1.5 us; 12 dependent N = 32
//          projection MMAs: 0.93 us, ~150 cycles each whatever N is) and the other stream warps meet it at the next barrier.
//   1 403  this version: a fourth warpgroup whose one working warp only issues MMAs, stream warps and MMA warp talk through
//          mbarriers (xready, pbar, pfree, gbar), and setmaxnreg moves the registers to where they are needed (512 threads
//          compiled for 128 registers; issue warpgroup 40, stream 104, cell 184; each setmaxnreg must be the first statement of
//          its role branch or ptxas budgets nothing).  Now the four stream warps are the bottleneck: 5.7 us per tile-step
//          (x part 2.0 us, projection read-back, residual re-read + output rows 2.7 us) against the 3.0 us of a cell phase.
// What it needs next: the stream work spread over more warps (two threads per row) or the FiLM / residual taken off it.
//

namespace tcq {
using namespace tcp;
constexpr int kStageBytes = kSlabBytes;                    // ONE 16 KB TMA stage the tiles take turns on (single-addend calls only)
constexpr int kOffW2 = kStageBytes;
constexpr int kAxBytes = kRows * kC * 2, kAhBytes = kRows * kH * 2;       // one bf16 image of the x part / of the h part of a tile
constexpr int kOffAx = kOffW2 + 2 * kWBytes + 2 * kPBytes;  // x part: tile 0 hi, lo, tile 1 hi, lo
constexpr int kOffAh = kOffAx + 4 * kAxBytes;               // h part: tile 0 hi, lo, tile 1 hi, lo
constexpr int kOffBias2 = kOffAh + 4 * kAhBytes;
constexpr int kOffLn2 = kOffBias2 + kN * 4;
constexpr int kOffBar2 = kOffLn2 + 3 * kC * 4;
constexpr int kSmemBytes2 = kOffBar2 + 256 + 1024;
constexpr int kThreads2 = 512;                             // warpgroups: 0 stream (rows), 1-2 cell update, 3 = one MMA-issue warp + 3 idle
struct TileGeo {
    bool active, tail_tile, valid;
    int outer0, inner0, grow;
    long long rbase;
};
__device__ __forceinline__ TileGeo tile_geo(const SeqArgs& a, const Geom& g, int tile, int n_tiles, int r) {
    TileGeo t{};
    t.active = tile < n_tiles;
    if (!t.active) return t;
    t.tail_tile = tile >= g.n_full_tiles;
    int o_row, i_row;
    if (!t.tail_tile) {
        t.outer0 = tile / g.nfull;
        t.inner0 = (tile - t.outer0 * g.nfull) * kRows;
        o_row = t.outer0; i_row = t.inner0 + r; t.valid = true;
    } else {
        t.outer0 = (tile - g.n_full_tiles) * g.P;
        t.inner0 = g.nfull * kRows;
        const int qq = r / g.tail;
        o_row = t.outer0 + qq; i_row = t.inner0 + (r - qq * g.tail);
        t.valid = qq < g.P && o_row < g.n_outer;
    }
    t.grow = t.valid ? o_row * a.rows_inner + i_row : 0;
    t.rbase = t.valid ? (long long)o_row * a.stride_outer + (long long)i_row * a.stride_inner : 0;
    return t;
}
// operand rows of the h part: chunk = 0..7 within the tile's own image
__device__ __forceinline__ void store_pair_h(unsigned char* hi, unsigned char* lo, int row, int chunk, const uint4 hi4, const uint4 lo4) {
    const int off = ((chunk * (kRows / 8) + (row >> 3)) * 8 + (row & 7)) * 16;
    sts16(hi + off, hi4);
    sts16(lo + off, lo4);
}
}  // namespace tcq

__global__ void __launch_bounds__(tcq::kThreads2, 1)
lstm_tcq_kernel(const SeqArgs a, const tcp::Geom g, const int n_tiles, const __grid_constant__ CUtensorMap map_x0,
                const __grid_constant__ CUtensorMap map_x0_tail, const __grid_constant__ CUtensorMap map_x1,
                const __grid_constant__ CUtensorMap map_x1_tail) {
    using namespace tcq;
    extern __shared__ unsigned char sm_raw[];
    unsigned char* sm = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(sm_raw) + 1023) & ~uintptr_t(1023));
    unsigned char* stages = sm;
    unsigned char* w_hi = sm + kOffW2;
    unsigned char* w_lo = w_hi + kWBytes;
    unsigned char* p_hi = w_lo + kWBytes;
    unsigned char* p_lo = p_hi + kPBytes;
    unsigned char* ax = sm + kOffAx;                        // tile X: hi at ax + 2 X kAxBytes, lo right behind
    unsigned char* ah = sm + kOffAh;                        // tile X: hi at ah + 2 X kAhBytes, lo right behind
    float* bias_s = reinterpret_cast<float*>(sm + kOffBias2);
    float* ln_s = reinterpret_cast<float*>(sm + kOffLn2);
    uint64_t* bars = reinterpret_cast<uint64_t*>(sm + kOffBar2);
    uint64_t* full = bars;                                  // [2] TMA bytes of a stage
    uint64_t* gbar = bars + 2;                              // [2] tcgen05.commit: gate MMAs of tile X (A's x part free, gates ready)
    uint64_t* pbar = bars + 4;                              // [2] tcgen05.commit: projection of tile X
    uint64_t* hready = bars + 6;                            // [2] 8 cell-update warps have published h of tile X
    uint64_t* xready = bars + 8;                            // [2] 4 stream warps have written the x part of tile X
    uint64_t* pfree = bars + 10;                            // [2] 4 stream warps have read the projection columns of tile X
    BulkBarrier* wbar = reinterpret_cast<BulkBarrier*>(bars + 12);
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 13);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int dir = blockIdx.y;
    const sb_lstm_dir& w = a.w[dir];
    const int S = a.n_steps;
    const int q = warp & 3;
    const int r = 32 * q + lane;
    constexpr int nld = 1, nst = 1;                         // one addend (the host sends two-addend calls to lstm_tcp_kernel), one stage
    const TileGeo t0 = tile_geo(a, g, 2 * blockIdx.x, n_tiles, r), t1 = tile_geo(a, g, 2 * blockIdx.x + 1, n_tiles, r);
    const int nact = t1.active ? 2 : 1;                     // tiles of this CTA; builds are numbered j = nact * step + tile

    for (int i = tid; i < kN; i += kThreads2) bias_s[i] = __ldg(w.tc_b + i) * ((i & 3) == 2 ? -2.0f * kLog2e : -kLog2e);
    if (tid < kC) {
        ln_s[tid] = __ldg(w.ln_g + tid);
        ln_s[kC + tid] = __ldg(w.ln_b + tid);
        ln_s[2 * kC + tid] = dir == 0 ? __ldg(w.lin_b + tid) : 0.0f;
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(kTmemCols) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    if (tid == 32) {
        for (int i = 0; i < 2; ++i) {
            mbar_init(full + i, 1);
            mbar_init(gbar + i, 1);
            mbar_init(pbar + i, 1);
            mbar_init(hready + i, 8);
            mbar_init(xready + i, 4);
            mbar_init(pfree + i, 4);
        }
        bulk_barrier_init(wbar);
        bulk_expect(wbar, 2 * kWBytes + 2 * kPBytes);
        bulk_copy_g2s(reinterpret_cast<float*>(w_hi), w.tc_w, 2 * kWBytes + 2 * kPBytes, wbar);
    }
    pdl_trigger();
    pdl_wait();
    fence_before();
    __syncthreads();
    fence_after();
    const uint32_t tmem = *tmem_slot;
    const uint32_t lane_base = (uint32_t)(32 * q) << 16;
    const int n_builds = nact * S;

    // Register budgets (setmaxnreg, warpgroup-wide): the kernel is compiled for 128 registers per thread (512 threads); the
    // fourth warpgroup gives its registers up (its one working warp only issues MMAs), the two cell warpgroups take them.
    // (each setmaxnreg is the first statement of its role branch: ptxas budgets the code a setmaxnreg dominates)

    if (warp >= 12) {
        // =============================================================================================================
        // MMA issue: one lane of warp 12; it may block for as long as the tensor pipe is busy without holding anybody else up
        // =============================================================================================================
        asm volatile("setmaxnreg.dec.sync.aligned.u32 40;");
        if (warp == 12 && elect_one()) {
            // The shared-memory base is made opaque HERE so that the compiler cannot hoist the descriptor arithmetic of this
            // region above the role branch, where it would stay live through the stream and cell regions (it did: 3.7 KB of spills).
            uint32_t sm_base = smem_u32(sm);
            asm volatile("" : "+r"(sm_base));
            const uint32_t ax_s = sm_base + kOffAx, w_hi_s = sm_base + kOffW2, w_lo_s = w_hi_s + kWBytes;
            const uint32_t p_hi_s = w_lo_s + kWBytes, p_lo_s = p_hi_s + kPBytes, ah_s = sm_base + kOffAh;
            constexpr uint32_t idesc_g = make_idesc(128, 256), idesc_p = make_idesc(128, 32);
            // rolled loops: this warpgroup runs on a small register budget, the descriptors are computed as they are needed
            auto issue_proj = [&](int X) {                  // proj[128 x 32] = h_X . lin^T -> columns 0..31 of tile X's gate buffer
                uint32_t acc = 0;
#pragma unroll 1
                for (int pass = 0; pass < 3; ++pass) {
                    const uint32_t ab = ah_s + (2 * X + (pass == 2 ? 1 : 0)) * kAhBytes, pb = pass == 1 ? p_lo_s : p_hi_s;
#pragma unroll 1
                    for (int ks = 0; ks < kH / 16; ++ks) {
                        umma(tmem + 256 * X, make_desc(ab + 2 * ks * kAChunkBytes, kAChunkBytes, 128),
                             make_desc(pb + 2 * ks * kPChunkBytes, kPChunkBytes, 128), idesc_p, acc);
                        acc = 1;
                    }
                }
            };
            auto issue_gates = [&](int X) {                 // gates[128 x 256] of tile X = [x part | h] of the tile . W^T, three terms
                uint32_t acc = 0;
#pragma unroll 1
                for (int pass = 0; pass < 3; ++pass) {
                    const uint32_t axb = ax_s + (2 * X + (pass == 2 ? 1 : 0)) * kAxBytes, ahb = ah_s + (2 * X + (pass == 2 ? 1 : 0)) * kAhBytes;
                    const uint32_t wb = pass == 1 ? w_lo_s : w_hi_s;
#pragma unroll 1
                    for (int ks = 0; ks < kK / 16; ++ks) {
                        const uint32_t ab = ks < kC / 16 ? axb + 2 * ks * kAChunkBytes : ahb + 2 * (ks - kC / 16) * kAChunkBytes;
                        umma(tmem + 256 * X, make_desc(ab, kAChunkBytes, 128), make_desc(wb + 2 * ks * kWChunkBytes, kWChunkBytes, 128), idesc_g, acc);
                        acc = 1;
                    }
                }
            };
            bulk_wait(wbar, 0);                             // the operand images have landed
#pragma unroll 1
            for (int step = 0; step <= S; ++step) {
#pragma unroll 1
                for (int X = 0; X < nact; ++X) {
                    TCQ_STAMP(3, 0, X, step);
                    mbar_wait(hready + X, (uint32_t)(step & 1));        // h_{step-1} of tile X is in A, its gate columns have been read
                    fence_after();
                    TCQ_STAMP(3, 1, X, step);
                    if (step > 0) {
                        issue_proj(X);
                        umma_commit(pbar + X);
                    }
                    TCQ_STAMP(3, 2, X, step);
                    if (step == S) continue;                // drain: only the projection of the last step
                    mbar_wait(xready + X, (uint32_t)(step & 1));        // the x part of (X, step) is in A
                    if (step > 0) mbar_wait(pfree + X, (uint32_t)((step - 1) & 1));     // the projection columns have been read back
                    fence_after();
                    TCQ_STAMP(3, 3, X, step);
                    issue_gates(X);
                    umma_commit(gbar + X);
                    TCQ_STAMP(3, 4, X, step);
                }
            }
        }
        __syncwarp();
    } else if (warp < 4) {
        // =============================================================================================================
        // stream group
        // =============================================================================================================
        asm volatile("setmaxnreg.dec.sync.aligned.u32 104;");
        auto issue_loads = [&](int j) {                     // one thread: the slab(s) of build j -> stage j % nst
            const int X = j % nact, step = j / nact;
            const TileGeo& t = X ? t1 : t0;
            const int pos = dir ? S - 1 - step : step;
            const int sg = j % nst;
            unsigned char* dst = stages + (size_t)sg * nld * kSlabBytes;
            uint64_t* bar = full + sg;
            const int c1 = g.mode == 0 ? pos : t.inner0, c2 = g.mode == 0 ? t.inner0 : pos;
            if (!t.tail_tile) {
                mbar_expect(bar, (uint32_t)(nld * kSlabBytes));
                tma_load_4d(dst, &map_x0, bar, 0, c1, c2, t.outer0);
            } else {
                int nq = g.n_outer - t.outer0;
                nq = nq < g.P ? nq : g.P;
                mbar_expect(bar, (uint32_t)(nld * nq * g.tail * kC * 4));
                for (int qq = 0; qq < nq; ++qq)
                    tma_load_4d(dst + (size_t)qq * g.tail * kC * 4, &map_x0_tail, bar, 0, c1, c2, t.outer0 + qq);
            }
        };
        if (tid == 0) {
            asm volatile("prefetch.tensormap [%0];" ::"l"(&map_x0) : "memory");
            for (int j = 0; j < nst && j < n_builds; ++j) issue_loads(j);
        }

        // x' of (tile, position) as the LSTM sees it before the LayerNorm: x0 [+ x1], FiLM
        auto film_apply = [&](float4 (&xv)[8], const TileGeo& t, int pos) {
            if (a.film_scale) {
                const long long film_row = (long long)(t.grow / a.film_row_div) * S * kC;
                const float4* fs = reinterpret_cast<const float4*>(a.film_scale + film_row + (long long)pos * kC);
                const float4* fb = reinterpret_cast<const float4*>(a.film_shift + film_row + (long long)pos * kC);
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    const float4 s4 = __ldg(fs + i), h4 = __ldg(fb + i);
                    xv[i].x = fmaf(xv[i].x, s4.x, h4.x); xv[i].y = fmaf(xv[i].y, s4.y, h4.y);
                    xv[i].z = fmaf(xv[i].z, s4.z, h4.z); xv[i].w = fmaf(xv[i].w, s4.w, h4.w);
                }
            }
        };
        auto build = [&](int j, float4 (&xv)[8]) {          // x part of build j into the shared region; x' returned for the residual
            const int X = j % nact, step = j / nact;
            const TileGeo& t = X ? t1 : t0;
            const int pos = dir ? S - 1 - step : step;
            const int sg = j % nst;
            mbar_wait(full + sg, (uint32_t)((j / nst) & 1));
            const unsigned char* base = stages + (size_t)sg * nld * kSlabBytes + (size_t)r * (kC * 4);
#pragma unroll
            for (int i = 0; i < 8; ++i) xv[i] = lds4(base + ((i ^ (r & 7)) << 4));
            film_apply(xv, t, pos);
            if (a.film_scale && step + 1 < S) {             // the FiLM rows of this tile's next step: into L1 while nobody waits for them
                const long long nxt = (long long)(t.grow / a.film_row_div) * S * kC + (long long)(dir ? pos - 1 : pos + 1) * kC;
                asm volatile("prefetch.global.L1 [%0];" ::"l"(a.film_scale + nxt));
                asm volatile("prefetch.global.L1 [%0];" ::"l"(a.film_shift + nxt));
            }
            unsigned char* ax_hi = ax + 2 * X * kAxBytes;
            unsigned char* ax_lo = ax_hi + kAxBytes;
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
            for (int ch = 0; ch < 4; ++ch) {
                float v[8];
#pragma unroll
                for (int i = 0; i < 2; ++i) {
                    const float4 tt = xv[2 * ch + i];
                    const float4 gg = lds4_ro(ln_s + 4 * (2 * ch + i)), bb = lds4_ro(ln_s + kC + 4 * (2 * ch + i));
                    v[4 * i + 0] = fmaf((tt.x - mean) * rstd, gg.x, bb.x); v[4 * i + 1] = fmaf((tt.y - mean) * rstd, gg.y, bb.y);
                    v[4 * i + 2] = fmaf((tt.z - mean) * rstd, gg.z, bb.z); v[4 * i + 3] = fmaf((tt.w - mean) * rstd, gg.w, bb.w);
                }
                uint4 hi4, lo4;
                split8(v, hi4, lo4);
                store_pair_h(ax_hi, ax_lo, r, ch, hi4, lo4);
            }
        };

        // y of (tile X, step) = lin h [+ b + x']: the projection is read back first (the gate MMAs may then overwrite the
        // buffer), the rows are stored after the MMAs have been issued
        auto proj_read = [&](int X, int step, uint32_t (&pr)[32]) {
            mbar_wait(pbar + X, (uint32_t)(step & 1));      // pbar[X] completes once per emitted step
            fence_after();
            tmem_ld32_issue(tmem + lane_base + 256 * X, pr);
            tmem_wait_ld();
            pin(pr);
        };
        // the rows are stored after the projection columns have been handed back; x' for the residual is re-read (and FiLM
        // re-applied) here, off the path that the MMA warp and the cell warps wait on
        // the rows are stored after the projection columns have been handed back; x' for the residual is re-read (and FiLM
        // re-applied) here, off the path that the MMA warp and the cell warps wait on.  (Requesting the row before the wait for
        // the projection was slower: 1 529 vs 1 403 us.)
        auto emit_store = [&](int X, int step, const uint32_t (&pr)[32]) {
            const TileGeo& t = X ? t1 : t0;
            if (!t.valid) return;
            const int pos = dir ? S - 1 - step : step;
            float4 xv[8];
            if (dir == 0) {
                const float* src = a.x0 + t.rbase + (long long)pos * a.stride_pos;
#pragma unroll
                for (int i = 0; i < 8; ++i) xv[i] = ld_plain4(src + 4 * i);
                film_apply(xv, t, pos);
            }
            float* dst = a.out[dir] + t.rbase + (long long)pos * a.stride_pos;
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                float4 o = make_float4(__uint_as_float(pr[4 * i]), __uint_as_float(pr[4 * i + 1]), __uint_as_float(pr[4 * i + 2]),
                                       __uint_as_float(pr[4 * i + 3]));
                if (dir == 0) {
                    const float4 bl = lds4_ro(ln_s + 2 * kC + 4 * i);
                    o.x += bl.x + xv[i].x; o.y += bl.y + xv[i].y;
                    o.z += bl.z + xv[i].z; o.w += bl.w + xv[i].w;
                }
                st4(dst + 4 * i, o);
            }
        };

        // One iteration per build j = (step, tile): x part -> hand it to the MMA warp -> read the projection of the previous step
        // back -> hand the columns back -> store the output rows.  Nothing here waits for MMAs other than that projection.
        for (int step = 0; step < S; ++step) {
#pragma unroll
            for (int X = 0; X < 2; ++X) {
                if (X >= nact) continue;
                const int j = nact * step + X;
                if (step > 0) {                             // the tile's x part is free once its previous gate MMAs have completed
                    mbar_wait(gbar + X, (uint32_t)((step - 1) & 1));
                    fence_after();
                }
                float4 xv[8];
                TCQ_STAMP(1, 0, X, step);
                build(j, xv);
                fence_async_smem();
                __syncwarp();
                if (lane == 0) mbar_arrive(xready + X);
                TCQ_STAMP(1, 1, X, step);
                bar_sync(1, 128);                           // the stage has been read by all four warps
                if (tid == 0 && j + nst < n_builds) issue_loads(j + nst);
                if (step > 0) {
                    uint32_t pr[32];
                    proj_read(X, step - 1, pr);
                    fence_before();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(pfree + X);
                    TCQ_STAMP(1, 4, X, step);
                    emit_store(X, step - 1, pr);
                    TCQ_STAMP(1, 6, X, step);
                }
            }
        }
        // ---- drain: projection of the last step of each tile --------------------------------------------------------
#pragma unroll
        for (int X = 0; X < 2; ++X) {
            if (X >= nact) continue;
            uint32_t pr[32];
            proj_read(X, S - 1, pr);
            emit_store(X, S - 1, pr);
        }
    } else if (warp < 12) {
        // =============================================================================================================
        // cell-update group: thread (row, hf) owns units 32hf + 8k .. + 7 of BOTH tiles (registers c[X][8k + j])
        // =============================================================================================================
        asm volatile("setmaxnreg.inc.sync.aligned.u32 184;");
        const int hf = (warp - 4) >> 2;
        float c[2][32];
#pragma unroll
        for (int X = 0; X < 2; ++X) {
            const TileGeo& t = X ? t1 : t0;
            unsigned char* h_hi = ah + 2 * X * kAhBytes;
            unsigned char* h_lo = h_hi + kAhBytes;
            if (X < nact) {
                if (a.h0 && t.valid) {
                    const float* cp = a.c0 + (long long)t.grow * kH + 32 * hf;
                    const float* hp = a.h0 + (long long)t.grow * kH + 32 * hf;
#pragma unroll
                    for (int ch = 0; ch < 4; ++ch) {        // plain loads: hN / cN may alias h0 / c0
                        const float4 c0 = ld_plain4(cp + 8 * ch), c1 = ld_plain4(cp + 8 * ch + 4);
                        c[X][8 * ch] = c0.x; c[X][8 * ch + 1] = c0.y; c[X][8 * ch + 2] = c0.z; c[X][8 * ch + 3] = c0.w;
                        c[X][8 * ch + 4] = c1.x; c[X][8 * ch + 5] = c1.y; c[X][8 * ch + 6] = c1.z; c[X][8 * ch + 7] = c1.w;
                        const float4 v0 = ld_plain4(hp + 8 * ch), v1 = ld_plain4(hp + 8 * ch + 4);
                        const float h8[8] = {v0.x, v0.y, v0.z, v0.w, v1.x, v1.y, v1.z, v1.w};
                        uint4 hi4, lo4;
                        split8(h8, hi4, lo4);
                        store_pair_h(h_hi, h_lo, r, 4 * hf + ch, hi4, lo4);
                    }
                } else {
#pragma unroll
                    for (int jj = 0; jj < 32; ++jj) c[X][jj] = 0.0f;
                    const uint4 z4 = make_uint4(0u, 0u, 0u, 0u);
#pragma unroll
                    for (int ch = 0; ch < 4; ++ch) store_pair_h(h_hi, h_lo, r, 4 * hf + ch, z4, z4);
                }
                fence_async_smem();
                __syncwarp();
                if (lane == 0) mbar_arrive(hready + X);
            } else {
#pragma unroll
                for (int jj = 0; jj < 32; ++jj) c[X][jj] = 0.0f;
            }
        }
        for (int s = 0; s < S; ++s) {
            const uint32_t par = (uint32_t)(s & 1);
#pragma unroll
            for (int X = 0; X < 2; ++X) {
                if (X >= nact) continue;
                const TileGeo& t = X ? t1 : t0;
                unsigned char* h_hi = ah + 2 * X * kAhBytes;
                unsigned char* h_lo = h_hi + kAhBytes;
                float* const hN = (a.hN && t.valid) ? a.hN + (long long)t.grow * kH + 32 * hf : nullptr;
                const uint32_t gcol = tmem + lane_base + 256u * X + 128 * hf;
                TCQ_STAMP(2, 0, X, s);
                mbar_wait(gbar + X, par);                   // gate MMAs of (X, s) complete: they have also finished reading h_{s-1}
                fence_after();
                TCQ_STAMP(2, 1, X, s);
#pragma unroll
                for (int ch = 0; ch < 4; ++ch) {
                    float h8[8];
#pragma unroll
                    for (int hh = 0; hh < 2; ++hh) {        // 4 units = 16 gate columns at a time: the thread also carries c of both tiles
                        uint32_t cur[16];
                        tmem_ld16_issue(gcol + 32 * ch + 16 * hh, cur);
                        const float* bp = bias_s + 4 * (32 * hf + 8 * ch + 4 * hh);
                        float4 nb_next = lds4_ro(bp);
                        tmem_wait_ld();
                        pin16(cur);
#pragma unroll
                        for (int jj = 0; jj < 4; ++jj) {
                            const float4 nb = nb_next;
                            if (jj < 3) nb_next = lds4_ro(bp + 4 * (jj + 1));
                            h8[4 * hh + jj] = cell7(__uint_as_float(cur[4 * jj + 0]), __uint_as_float(cur[4 * jj + 1]), __uint_as_float(cur[4 * jj + 2]),
                                                    __uint_as_float(cur[4 * jj + 3]), nb, c[X][8 * ch + 4 * hh + jj]);
                        }
                    }
                    if (s == S - 1 && hN) {
                        st4(hN + 8 * ch, make_float4(h8[0], h8[1], h8[2], h8[3]));
                        st4(hN + 8 * ch + 4, make_float4(h8[4], h8[5], h8[6], h8[7]));
                    }
                    uint4 hi4, lo4;
                    split8(h8, hi4, lo4);
                    store_pair_h(h_hi, h_lo, r, 4 * hf + ch, hi4, lo4);
                }
                fence_before();
                fence_async_smem();
                __syncwarp();
                if (lane == 0) mbar_arrive(hready + X);
                TCQ_STAMP(2, 2, X, s);
            }
        }
#pragma unroll
        for (int X = 0; X < 2; ++X) {
            const TileGeo& t = X ? t1 : t0;
            if (X < nact && a.cN && t.valid) {
                float* cp = a.cN + (long long)t.grow * kH + 32 * hf;
#pragma unroll
                for (int ch = 0; ch < 4; ++ch) {
                    st4(cp + 8 * ch, make_float4(c[X][8 * ch], c[X][8 * ch + 1], c[X][8 * ch + 2], c[X][8 * ch + 3]));
                    st4(cp + 8 * ch + 4, make_float4(c[X][8 * ch + 4], c[X][8 * ch + 5], c[X][8 * ch + 6], c[X][8 * ch + 7]));
                }
            }
        }
    }
    fence_before();
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(kTmemCols) : "memory");
}

// ---- host: tensor maps ------------------------------------------------------------------------------------------
namespace tcp {

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn encode_fn() {
    static EncodeTiledFn fn = nullptr;
    static std::once_flag once;
    std::call_once(once, [] {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess &&
            qres == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiledFn>(p);
        (void)cudaGetLastError();
    });
    return fn;
}

// 4-D fp32 map over the activation grid as this call addresses it; the box is [C][1][rows][1] (mode 0) or [C][rows][1][1]
static int make_map(CUtensorMap* m, const float* base, const SeqArgs& a, const Geom& g, int box_rows) {
    EncodeTiledFn enc = encode_fn();
    SB_REQUIRE(enc, SB_E_UNSUPP, "SB_ALGO_TCP: cuTensorMapEncodeTiled is not available from this driver");
    const cuuint64_t outer_stride = g.n_outer > 1 ? (cuuint64_t)a.stride_outer : (cuuint64_t)a.rows_inner * a.stride_inner;
    cuuint64_t dims[4], strides[3];
    cuuint32_t box[4], estr[4] = {1, 1, 1, 1};
    dims[0] = kC;
    if (g.mode == 0) {
        dims[1] = (cuuint64_t)a.n_steps; dims[2] = (cuuint64_t)a.rows_inner;
        strides[0] = (cuuint64_t)a.stride_pos * 4; strides[1] = (cuuint64_t)a.stride_inner * 4;
        box[0] = kC; box[1] = 1; box[2] = (cuuint32_t)box_rows; box[3] = 1;
    } else {
        dims[1] = (cuuint64_t)a.rows_inner; dims[2] = (cuuint64_t)a.n_steps;
        strides[0] = (cuuint64_t)a.stride_inner * 4; strides[1] = (cuuint64_t)a.stride_pos * 4;
        box[0] = kC; box[1] = (cuuint32_t)box_rows; box[2] = 1; box[3] = 1;
    }
    dims[3] = (cuuint64_t)g.n_outer;
    strides[2] = outer_stride * 4;
    const CUresult rc = enc(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, const_cast<float*>(base), dims, strides, box, estr,
                            CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                            CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    SB_REQUIRE(rc == CUDA_SUCCESS, SB_E_BADARG, "SB_ALGO_TCP: cuTensorMapEncodeTiled failed (%d): base %p, strides %lld / %lld / %lld",
               (int)rc, (const void*)base, (long long)strides[0], (long long)strides[1], (long long)strides[2]);
    return 0;
}

}  // namespace tcp

bool seq_tcp_supported(const SeqArgs& a) {
    if (a.n_rows % a.rows_inner != 0) return false;
    const bool al = ((uintptr_t)a.x0 & 15) == 0 && (!a.x1 || ((uintptr_t)a.x1 & 15) == 0);
    const bool strides = a.stride_pos % 4 == 0 && a.stride_inner % 4 == 0 && a.stride_outer % 4 == 0;
    const bool ordered = a.stride_pos < a.stride_inner ? a.stride_pos >= tcp::kC : a.stride_inner >= tcp::kC;
    return al && strides && ordered && tcp::encode_fn() != nullptr;
}

static int tcp_setup(const SeqArgs& a, tcp::Geom& g, int& n_tiles, CUtensorMap (&m)[4]) {
    using namespace tcp;
    SB_REQUIRE(seq_tcp_supported(a), SB_E_UNSUPP, "SB_ALGO_TCP: activation layout not addressable by a TMA tensor map");
    g = Geom{};
    g.mode = a.stride_pos < a.stride_inner ? 0 : 1;
    g.n_outer = a.n_rows / a.rows_inner;
    g.nfull = a.rows_inner / kRows;
    g.tail = a.rows_inner % kRows;
    g.P = g.tail ? kRows / g.tail : 1;
    g.n_full_tiles = g.n_outer * g.nfull;
    n_tiles = g.n_full_tiles + (g.tail ? ceil_div(g.n_outer, g.P) : 0);
    SB_CHECK(make_map(&m[0], a.x0, a, g, kRows));
    SB_CHECK(make_map(&m[1], a.x0, a, g, g.tail ? g.tail : kRows));
    SB_CHECK(make_map(&m[2], a.x1 ? a.x1 : a.x0, a, g, kRows));
    SB_CHECK(make_map(&m[3], a.x1 ? a.x1 : a.x0, a, g, g.tail ? g.tail : kRows));
    return 0;
}

static bool tcr_outputs_ok(const SeqArgs& a) {
    for (int d = 0; d < a.n_dirs; ++d)
        if (!a.out[d] || ((uintptr_t)a.out[d] & 15)) return false;
    return true;
}

bool seq_tcr_selected(const SeqArgs& a) {
    return seq_tcp_supported(a) && !a.x1 && tc_cell7_enabled() && tc_pipe_enabled() && tcr_outputs_ok(a);
}

int run_seq_tcp(const SeqArgs& a, cudaStream_t st) {
    using namespace tcp;
    Geom g;
    int n_tiles;
    CUtensorMap m[4];
    SB_CHECK(tcp_setup(a, g, n_tiles, m));
    dim3 grid(n_tiles, a.n_dirs);
    SB_REQUIRE(!a.sum_dirs || seq_tcr_selected(a), SB_E_UNSUPP, "SB_ALGO_TCP: summed directions need lstm_tcr_kernel (single addend, SB_OPT_TC_PIPE, SB_OPT_TC_CELL7)");
    if (seq_tcr_selected(a)) {
        CUtensorMap mo[4];                                  // the outputs leave through TMA tensor stores: same geometry as x0
        for (int d = 0; d < 2; ++d) {
            float* base = a.out[d < a.n_dirs ? d : 0];
            SB_CHECK(make_map(&mo[2 * d], base, a, g, kRows));
            SB_CHECK(make_map(&mo[2 * d + 1], base, a, g, g.tail ? g.tail : kRows));
        }
        if (tc_cw16_enabled())
            return launch("lstm_tcr16", lstm_tcr_kernel<16>, grid, dim3(tcr::threads(16)), (size_t)kSmemBytes, st, a, g, m[0], m[1], mo[0], mo[1], mo[2], mo[3]);
        return launch("lstm_tcr", lstm_tcr_kernel<8>, grid, dim3(tcr::threads(8)), (size_t)kSmemBytes, st, a, g, m[0], m[1], mo[0], mo[1], mo[2], mo[3]);
    }
    if (tc_cell7_enabled())
        return launch("lstm_tcp", lstm_tcp_kernel<true>, grid, dim3(kThreads), (size_t)kSmemBytes, st, a, g, m[0], m[1], m[2], m[3]);
    return launch("lstm_tcp", lstm_tcp_kernel<false>, grid, dim3(kThreads), (size_t)kSmemBytes, st, a, g, m[0], m[1], m[2], m[3]);
}

#ifdef SB_TCQ_DEBUG
extern "C" int sb_tcq_debug_read(long long* out, int max_events) {
    static long long host[16384];
    cudaMemcpyFromSymbol(host, g_tcq_dbg, sizeof(host));
    int n = 0;
    for (int i = 0; i < 4096 && n < max_events; ++i)
        if (host[4 * i]) {
            for (int j = 0; j < 4; ++j) out[4 * n + j] = host[4 * i + j];
            ++n;
        }
    std::memset(host, 0, sizeof(host));
    cudaMemcpyToSymbol(g_tcq_dbg, host, sizeof(host));
    return n;
}
#endif

// two tiles per CTA in ping-pong; falls back to one tile per CTA when there is only one
int run_seq_tcq(const SeqArgs& a, cudaStream_t st) {
    using namespace tcp;
    Geom g;
    int n_tiles;
    CUtensorMap m[4];
    SB_CHECK(tcp_setup(a, g, n_tiles, m));
    if (n_tiles < 2 || a.x1) return run_seq_tcp(a, st);       // one tile, or two addends (no room for a 32 KB stage)
    dim3 grid(ceil_div(n_tiles, 2), a.n_dirs);
    return launch("lstm_tcq", lstm_tcq_kernel, grid, dim3(tcq::kThreads2), (size_t)tcq::kSmemBytes2, st, a, g, n_tiles, m[0], m[1], m[2], m[3]);
}

#else   // SB_EMU: tensor-core / TMA instructions cannot be emulated on the host

bool seq_tcp_supported(const SeqArgs&) { return false; }

int run_seq_tcp(const SeqArgs&, cudaStream_t) {
    set_error("SB_ALGO_TCP (tcgen05 + TMA) is not available in the host-emulated test build");
    return SB_E_UNSUPP;
}
int run_seq_tcq(const SeqArgs&, cudaStream_t) {
    set_error("SB_ALGO_TCQ (tcgen05 + TMA) is not available in the host-emulated test build");
    return SB_E_UNSUPP;
}

#endif

}  // namespace sb
