// cuda_emu.h — TEST INFRASTRUCTURE ONLY.  A tiny host-side stand-in for the CUDA execution model, so that the SIMT
// kernels of sound_bubble_b200/csrc can be compiled with g++ (-DSB_EMU) and functionally checked against the oracle in
// this GPU-less container before they are sent to a B200.  It is never built into, loaded by or shipped with the
// product library (libsoundbubble_sm100a.so); `sound_bubble_b200` has no code path that reaches it.
//
// Model: one block at a time; every CUDA thread of the block is an OS thread; __syncthreads / __syncwarp / warp
// shuffles are std::barrier rendezvous.  Slow (fine for the tiny shapes the emu tests use) but faithful to the
// synchronisation structure, which is what the tests are after (indexing, layouts, barrier placement).
#pragma once
#ifndef SB_EMU
#error "cuda_emu.h is only for -DSB_EMU builds"
#endif

#include <algorithm>
#include <barrier>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <map>
#include <mutex>
#include <memory>
#include <thread>
#include <vector>

struct dim3 {
    unsigned x, y, z;
    dim3(unsigned x_ = 1, unsigned y_ = 1, unsigned z_ = 1) : x(x_), y(y_), z(z_) {}
};
struct alignas(8) float2 { float x, y; };
struct alignas(16) float4 { float x, y, z, w; };
struct alignas(8) int2 { int x, y; };
struct alignas(16) int4 { int x, y, z, w; };
struct alignas(16) uint4 { unsigned x, y, z, w; };
static inline float2 make_float2(float x, float y) { return float2{x, y}; }
static inline float4 make_float4(float x, float y, float z, float w) { return float4{x, y, z, w}; }

typedef void* cudaStream_t;
typedef int cudaError_t;
enum { cudaSuccess = 0 };
enum cudaFuncAttribute { cudaFuncAttributeMaxDynamicSharedMemorySize = 8 };
enum { cudaDevAttrMultiProcessorCount = 16 };
static inline cudaError_t cudaGetLastError() { return 0; }
static inline cudaError_t cudaPeekAtLastError() { return 0; }
static inline const char* cudaGetErrorString(cudaError_t) { return "emu"; }
template <class K> static inline cudaError_t cudaFuncSetAttribute(K, int, int) { return 0; }
static inline cudaError_t cudaGetDevice(int* d) { *d = 0; return 0; }
static inline cudaError_t cudaDeviceGetAttribute(int* v, int, int) { *v = 4; return 0; }     // pretend 4 SMs
static inline cudaError_t cudaMemsetAsync(void* p, int v, size_t n, cudaStream_t) { memset(p, v, n); return 0; }
static inline cudaError_t cudaMemcpyAsync(void* d, const void* s, size_t n, int, cudaStream_t) { memcpy(d, s, n); return 0; }
enum { cudaMemcpyDeviceToDevice = 3 };

#define __global__
#define __device__
#define __host__
#define __forceinline__ inline
#define __launch_bounds__(...)
#define __shared__ static
#define __align__(n) __attribute__((aligned(n)))
#define __restrict__

namespace emu {
struct Block {
    std::unique_ptr<std::barrier<>> bar;
    std::vector<std::unique_ptr<std::barrier<>>> wbar;
    std::vector<uint32_t> xch;                  // one exchange slot per thread (warp shuffles)
    std::vector<unsigned char> dyn;             // dynamic shared memory
    std::map<int, std::unique_ptr<std::barrier<>>> named;   // bar.sync id, count
    std::mutex mu;
};
extern Block* g_block;
extern thread_local unsigned t_lane, t_warp;
void launch_impl(dim3 grid, dim3 block, size_t smem, const std::function<void()>& body);
}   // namespace emu

extern thread_local dim3 threadIdx, blockIdx;
extern dim3 blockDim, gridDim;

static inline void __syncthreads() { emu::g_block->bar->arrive_and_wait(); }
static inline void __syncwarp(unsigned = 0xffffffffu) { emu::g_block->wbar[emu::t_warp]->arrive_and_wait(); }
namespace emu {
static inline void named_barrier(int id, int nthreads) {       // PTX bar.sync id, nthreads
    std::barrier<>* b;
    {
        std::lock_guard<std::mutex> lock(g_block->mu);
        auto& slot = g_block->named[id];
        if (!slot) slot = std::make_unique<std::barrier<>>(nthreads);
        b = slot.get();
    }
    b->arrive_and_wait();
}
}   // namespace emu

template <class T> static inline T emu_shfl_idx(T v, int src_lane) {
    static_assert(sizeof(T) == 4, "32-bit shuffles only");
    emu::Block* b = emu::g_block;
    uint32_t bits;
    memcpy(&bits, &v, 4);
    const unsigned base = emu::t_warp * 32;
    b->xch[base + emu::t_lane] = bits;
    b->wbar[emu::t_warp]->arrive_and_wait();
    uint32_t got = b->xch[base + (unsigned)(src_lane & 31)];
    b->wbar[emu::t_warp]->arrive_and_wait();
    T r;
    memcpy(&r, &got, 4);
    return r;
}
template <class T> static inline T __shfl_sync(unsigned, T v, int src, int width = 32) {
    const int lane = (int)emu::t_lane;
    return emu_shfl_idx(v, (lane & ~(width - 1)) | (src & (width - 1)));
}
template <class T> static inline T __shfl_xor_sync(unsigned, T v, int m, int width = 32) {
    (void)width;
    return emu_shfl_idx(v, (int)emu::t_lane ^ m);
}
template <class T> static inline T __shfl_down_sync(unsigned, T v, int d, int width = 32) {
    const int lane = (int)emu::t_lane;
    const int src = ((lane & (width - 1)) + d < width) ? lane + d : lane;
    return emu_shfl_idx(v, src);
}
template <class T> static inline T __shfl_up_sync(unsigned, T v, int d, int width = 32) {
    const int lane = (int)emu::t_lane;
    const int src = ((lane & (width - 1)) - d >= 0) ? lane - d : lane;
    return emu_shfl_idx(v, src);
}

template <class T> static inline T __ldg(const T* p) { return *p; }
namespace emu { extern std::mutex g_atomic_mu; }
static inline void emu_atomic_add(float* p, float v) { std::lock_guard<std::mutex> lock(emu::g_atomic_mu); *p += v; }
#define __expf(x) expf(x)
static inline float __frcp_rn(float x) { return 1.0f / x; }
static inline float __fdividef(float a, float b) { return a / b; }
static inline float rsqrtf(float x) { return 1.0f / sqrtf(x); }
static inline float __fsqrt_rn(float x) { return sqrtf(x); }
static inline int min(int a, int b) { return a < b ? a : b; }
static inline int max(int a, int b) { return a > b ? a : b; }

#define SB_EMU_LAUNCH(kern, grid, block, smem, ...) \
    emu::launch_impl((grid), (block), (smem), [=]() { kern(__VA_ARGS__); })
