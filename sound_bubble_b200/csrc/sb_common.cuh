// Shared device/host helpers for libsoundbubble_sm100a.so (sm_100a only).
//
// -DSB_EMU builds the same sources with g++ against tests/emu/cuda_emu.h; that build exists only so the tests can
// check kernel logic in a GPU-less container and is never part of the product library.
#pragma once

#ifdef SB_EMU
#include "cuda_emu.h"
#else
#include <cuda_runtime.h>
#endif
#include <stdint.h>
#include <stdio.h>

#include "soundbubble.h"

namespace sb {

// ------------------------------------------------------------------------------------------------------------
// host side: error reporting, launch accounting, launch helper
// ------------------------------------------------------------------------------------------------------------
void set_error(const char* fmt, ...);
void count_launch();
int  check_launch(const char* what);        // cudaGetLastError() -> 0 / cudaError_t, records the message
bool pdl_enabled();                         // programmatic dependent launch (sb_set_option(SB_OPT_PDL, 1))
bool attn_tc_enabled();                     // tcgen05 attention core for T >= 64 (sb_set_option(SB_OPT_ATTN_TC, 0) turns it off)
bool train_one_row_enabled();                // training LSTM kernels with one gate row / one W_hh column per thread (sb_set_option(SB_OPT_TRAIN_ONE_ROW, 1))
bool train_tc_enabled();                     // LSTM weight gradients on tcgen05 (sb_train_tc.cu; sb_set_option(SB_OPT_TRAIN_TC, v))
bool tc_cw16_enabled();                      // lstm_tcr_kernel with 16 cell-update warps (sb_set_option(SB_OPT_TC_CW16, v))
bool front_tc_enabled();                     // conv-in on the tensor cores (sb_set_option(SB_OPT_FRONT_TC, v))
bool conv_in_tc_supported(const sb_conv_in_args& p);
int run_conv_in_tc(const sb_conv_in_args& p, cudaStream_t st);   // sb_frontend_tc.cu
bool conv_in_tc_train_supported(int B, int T, int F, int Cin, int C);
int run_conv_in_tc_train(const float* feats, const float* w, const float* bias, float* raw, int B, int T, int F, int Cin, cudaStream_t st);
bool tc_pipe_enabled();                      // single-addend SB_ALGO_TC calls on lstm_tcr_kernel (sb_set_option(SB_OPT_TC_PIPE, v))
bool tc_cell7_enabled();                     // shared-reciprocal cell update in lstm_tcp_kernel (sb_set_option(SB_OPT_TC_CELL7, v))
bool tc_v1_enabled();                        // SB_ALGO_TC on the first tcgen05 kernel instead of the TMA pipeline (sb_set_option(SB_OPT_TC_V1, 1))
bool train_ffma2_enabled();                  // packed FFMA2 in the training GEMM kernels (sb_set_option(SB_OPT_TRAIN_FFMA2, v))
int  sm_count();
int  ensure_smem(const void* func, size_t bytes, const char* name);

#define SB_REQUIRE(cond, code, ...)                 \
    do {                                            \
        if (!(cond)) {                              \
            sb::set_error(__VA_ARGS__);             \
            return (code);                          \
        }                                           \
    } while (0)

#define SB_CHECK(expr)                              \
    do {                                            \
        const int sb_rc_ = (expr);                  \
        if (sb_rc_ != 0) return sb_rc_;             \
    } while (0)

inline int ceil_div(int a, int b) { return (a + b - 1) / b; }
inline long long ceil_div_ll(long long a, long long b) { return (a + b - 1) / b; }

// sb_train_tc.cu: dW_ih / dW_hh / db of one LSTM direction as one tcgen05 reduction GEMM over the n = (row, step) pairs
struct WgradTc {
    const float* dz;        // [N][4H]
    const float* xn;        // [N][C]   LayerNorm output (the LSTM's input)
    const float* h;         // [N][H]   hidden states; row n pairs with h[n - 1] (n + 1 for the reverse direction)
    float* dW_ih;           // [4H][C]  accumulated with atomics
    float* dW_hh;           // [4H][H]
    float* db;              // [4H] or NULL
    float* db2;             // [4H] or NULL (b_hh receives the same sum as b_ih)
    int S, reverse;
    long long N, rows_per_cta;
    float* scratch;         // optional: per-CTA partial sums [ctas][4H * (C + H) + 4H]; a second kernel adds them up (no atomics)
    long long scratch_floats;
};
int run_wgrad_tc(const WgradTc& w, cudaStream_t st);

// One place through which every kernel of the library is launched: opt-in shared memory, optional programmatic
// dependent launch (the kernel then calls pdl_wait() before it touches anything a predecessor wrote), accounting.
template <typename... KArgs, typename... Args>
int launch(const char* name, void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st,
           Args... args) {
#ifdef SB_EMU
    (void)st;
    emu::launch_impl(grid, block, smem, [=]() { kern(args...); });
    count_launch();
    return 0;
#else
    if (smem > 48 * 1024) SB_CHECK(ensure_smem(reinterpret_cast<const void*>(kern), smem, name));
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = grid;
    cfg.blockDim = block;
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    unsigned n_attr = 0;
    if (pdl_enabled()) {
        attr[n_attr].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        attr[n_attr].val.programmaticStreamSerializationAllowed = 1;
        ++n_attr;
    }
    cfg.attrs = attr;
    cfg.numAttrs = n_attr;
    cudaError_t e = cudaLaunchKernelEx(&cfg, kern, static_cast<KArgs>(args)...);
    count_launch();
    if (e != cudaSuccess) {
        set_error("%s: launch failed: %s", name, cudaGetErrorString(e));
        return (int)e;
    }
    return check_launch(name);
#endif
}

// ------------------------------------------------------------------------------------------------------------
// device side
// ------------------------------------------------------------------------------------------------------------
#ifdef SB_EMU
#define SB_DYN_SMEM(type, name) \
    type* name = reinterpret_cast<type*>((reinterpret_cast<uintptr_t>(emu::g_block->dyn.data()) + 63) & ~uintptr_t(63))
#else
#define SB_DYN_SMEM(type, name)                                   \
    extern __shared__ __align__(16) unsigned char sb_dyn_smem_[]; \
    type* name = reinterpret_cast<type*>(sb_dyn_smem_)
#endif

// Programmatic dependent launch: everything before pdl_wait() (weight staging, index math) may overlap the tail of the
// previous kernel in the stream; nothing a predecessor wrote may be read before it.  No-ops when launched normally.
__device__ __forceinline__ void pdl_wait() {
#ifndef SB_EMU
    asm volatile("griddepcontrol.wait;" ::: "memory");
#endif
}
__device__ __forceinline__ void pdl_trigger() {
#ifndef SB_EMU
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
#endif
}

// named barrier over a subset of the CTA's warps (PTX bar.sync id, nthreads; id 0 is __syncthreads)
__device__ __forceinline__ void bar_sync(int id, int nthreads) {
#ifdef SB_EMU
    emu::named_barrier(id, nthreads);
#else
    asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
#endif
}

// sigmoid / tanh through one ex2.approx + one rcp.approx (2 MUFU ops, no slow paths): |abs err| ~ 1e-7, far inside the
// 1e-3 RMS waveform parity bar.  tanh.approx (2^-11 rel err) is NOT accurate enough here.
__device__ __forceinline__ float fast_ex2(float x) {
#ifdef SB_EMU
    return exp2f(x);
#else
    float r;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
#endif
}
__device__ __forceinline__ float fast_rcp(float x) {
#ifdef SB_EMU
    return 1.0f / x;
#else
    float r;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
#endif
}
__device__ __forceinline__ float sigmoid_f(float x) { return fast_rcp(1.0f + fast_ex2(-1.4426950408889634f * x)); }
__device__ __forceinline__ float tanh_f(float x) { return fmaf(2.0f, sigmoid_f(2.0f * x), -1.0f); }

// packed fp32x2 FMA (Blackwell FFMA2): d.{x,y} += a.{x,y} * b.  Same rounding as two fmaf, half the issue slots.
__device__ __forceinline__ void ffma2(float2& d, const float2 a, const float b) {
#ifdef SB_EMU
    d.x = fmaf(a.x, b, d.x);
    d.y = fmaf(a.y, b, d.y);
#else
    unsigned long long dd, aa, bb;
    asm("mov.b64 %0, {%1, %2};" : "=l"(dd) : "f"(d.x), "f"(d.y));
    asm("mov.b64 %0, {%1, %2};" : "=l"(aa) : "f"(a.x), "f"(a.y));
    asm("mov.b64 %0, {%1, %2};" : "=l"(bb) : "f"(b), "f"(b));
    asm("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(dd) : "l"(aa), "l"(bb));
    asm("mov.b64 {%0, %1}, %2;" : "=f"(d.x), "=f"(d.y) : "l"(dd));
#endif
}

__device__ __forceinline__ float4 ld4(const float* p) { return *reinterpret_cast<const float4*>(p); }
__device__ __forceinline__ float2 ld2(const float* p) { return *reinterpret_cast<const float2*>(p); }
__device__ __forceinline__ void st4(float* p, float4 v) { *reinterpret_cast<float4*>(p) = v; }
__device__ __forceinline__ void st2(float* p, float2 v) { *reinterpret_cast<float2*>(p) = v; }

// read-only, streaming (do not pollute L1): activations are touched once per kernel
__device__ __forceinline__ float4 ldg4_stream(const float* p) {
#ifdef SB_EMU
    return ld4(p);
#else
    float4 r;
    asm("ld.global.nc.L1::no_allocate.v4.f32 {%0, %1, %2, %3}, [%4];"
                 : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w)
                 : "l"(p));
    return r;
#endif
}
__device__ __forceinline__ float2 ldg2_stream(const float* p) {
#ifdef SB_EMU
    return ld2(p);
#else
    float2 r;
    asm("ld.global.nc.L1::no_allocate.v2.f32 {%0, %1}, [%2];" : "=f"(r.x), "=f"(r.y) : "l"(p));
    return r;
#endif
}
__device__ __forceinline__ float ldg1_stream(const float* p) {
#ifdef SB_EMU
    return *p;
#else
    float r;
    asm("ld.global.nc.L1::no_allocate.f32 %0, [%1];" : "=f"(r) : "l"(p));
    return r;
#endif
}
// plain (coherent) loads for buffers an aliasing output of the same launch may point at
__device__ __forceinline__ float ld_plain(const float* p) { return *reinterpret_cast<const volatile float*>(p); }
__device__ __forceinline__ float4 ld_plain4(const float* p) {
#ifdef SB_EMU
    return ld4(p);
#else
    float4 r;
    asm volatile("ld.volatile.global.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "l"(p));
    return r;
#endif
}

// Asynchronous global -> shared copies (LDGSTS): every copy of a staging loop is in flight at once instead of one L2
// round trip per loop iteration; cp_async_wait_all() before the __syncthreads() that publishes the tile.
__device__ __forceinline__ void cp_async_16(float* smem_dst, const float* gsrc) {
#ifdef SB_EMU
    *reinterpret_cast<float4*>(smem_dst) = *reinterpret_cast<const float4*>(gsrc);
#else
    const unsigned sa = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(sa), "l"(gsrc) : "memory");
#endif
}
__device__ __forceinline__ void cp_async_4(float* smem_dst, const float* gsrc) {
#ifdef SB_EMU
    *smem_dst = *gsrc;
#else
    const unsigned sa = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(sa), "l"(gsrc) : "memory");
#endif
}
__device__ __forceinline__ void cp_async_wait_all() {
#ifndef SB_EMU
    asm volatile("cp.async.commit_group;\n\tcp.async.wait_group 0;" ::: "memory");
#endif
}
// cooperative copy of n4 float4 (16-byte aligned both sides); complete after cp_async_wait_all() + a barrier
__device__ __forceinline__ void stage_f4(float* dst, const float* src, int n4, int tid, int nthreads) {
    for (int i = tid; i < n4; i += nthreads) cp_async_16(dst + 4 * i, src + 4 * i);
}

// ------------------------------------------------------------------------------------------------------------
// TMA bulk copies (cp.async.bulk, SASS UBLKCP): one thread hands a contiguous global -> shared copy of any multiple of
// 16 bytes to the copy engine and the consumers wait on an mbarrier that counts the arriving bytes.  Used to stage the
// packed weight images (up to 106 KB per CTA) with one instruction instead of thousands of 16-byte LDGSTS.
// ------------------------------------------------------------------------------------------------------------
struct alignas(8) BulkBarrier { unsigned long long v; };

__device__ __forceinline__ void bulk_barrier_init(BulkBarrier* bar) {           // one thread, then a block barrier
#ifndef SB_EMU
    const unsigned a = (unsigned)__cvta_generic_to_shared(bar);
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(a) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
#else
    bar->v = 0;
#endif
}
// one thread: announce `total_bytes`, then issue the copies (each a multiple of 16 bytes, 16-byte aligned on both sides)
__device__ __forceinline__ void bulk_expect(BulkBarrier* bar, unsigned total_bytes) {
#ifndef SB_EMU
    const unsigned a = (unsigned)__cvta_generic_to_shared(bar);
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(a), "r"(total_bytes) : "memory");
#else
    (void)bar; (void)total_bytes;
#endif
}
__device__ __forceinline__ void bulk_copy_g2s(float* smem_dst, const float* gsrc, unsigned bytes, BulkBarrier* bar) {
#ifndef SB_EMU
    const unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst), b = (unsigned)__cvta_generic_to_shared(bar);
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(d), "l"(gsrc), "r"(bytes), "r"(b) : "memory");
#else
    (void)bar;
    for (unsigned i = 0; i < bytes / 4; ++i) smem_dst[i] = gsrc[i];
#endif
}
// every consumer thread: wait for phase `parity` of the barrier (the copies have landed and are visible)
__device__ __forceinline__ void bulk_wait(BulkBarrier* bar, unsigned parity) {
#ifndef SB_EMU
    const unsigned a = (unsigned)__cvta_generic_to_shared(bar);
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "SB_BULK_WAIT_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@!p bra SB_BULK_WAIT_%=;\n\t}\n" ::"r"(a), "r"(parity) : "memory");
#else
    (void)bar; (void)parity;
#endif
}

template <int WIDTH>
__device__ __forceinline__ float group_sum(float v) {
#pragma unroll
    for (int o = WIDTH / 2; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

constexpr float kLnEps = 1e-5f;

}  // namespace sb
