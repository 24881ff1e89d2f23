// Micro-benchmarks that calibrate the cost model of the SIMT kernels: FFMA vs packed FFMA2 issue rate,
// cost of warp-wide LDS.128 (distinct vs broadcast), SHFL, MUFU, barrier.  Build: nvcc -gencode arch=compute_100a,code=sm_100a
#include <cstdio>
#include <cuda_runtime.h>

#define CHECK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); return 1; } } while (0)

template <int MODE>
__global__ void __launch_bounds__(1024) k_bench(float* out, long long* cyc, int iters, float x, float y) {
    __shared__ __align__(16) float sm[8192];
    const int tid = threadIdx.x, lane = tid & 31;
    for (int i = tid; i < 8192; i += blockDim.x) sm[i] = 0.001f * i;
    __syncthreads();
    float a[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) a[i] = 0.001f * (tid + i);
    long long t0 = clock64();
    if (MODE == 0) {            // 16 independent FFMA chains
        for (int it = 0; it < iters; ++it) {
#pragma unroll
            for (int i = 0; i < 16; ++i) a[i] = fmaf(a[i], x, y);
        }
    } else if (MODE == 1) {     // packed f32x2 FMA: 8 independent chains of pairs
        unsigned long long p[8], xx, yy;
        asm volatile("mov.b64 %0, {%1, %2};" : "=l"(xx) : "f"(x), "f"(x));
        asm volatile("mov.b64 %0, {%1, %2};" : "=l"(yy) : "f"(y), "f"(y));
#pragma unroll
        for (int i = 0; i < 8; ++i) asm volatile("mov.b64 %0, {%1, %2};" : "=l"(p[i]) : "f"(a[2 * i]), "f"(a[2 * i + 1]));
        for (int it = 0; it < iters; ++it) {
#pragma unroll
            for (int i = 0; i < 8; ++i) asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(p[i]) : "l"(xx), "l"(yy));
        }
#pragma unroll
        for (int i = 0; i < 8; ++i) asm volatile("mov.b64 {%0, %1}, %2;" : "=f"(a[2 * i]), "=f"(a[2 * i + 1]) : "l"(p[i]));
    } else if (MODE == 2) {     // LDS.128, every lane its own 16 bytes (conflict-free)
        for (int it = 0; it < iters; ++it) {
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const float4 v = *reinterpret_cast<const float4*>(sm + ((4 * lane + 128 * i + 4 * it) & 8188));
                a[4 * i] += v.x; a[4 * i + 1] += v.y; a[4 * i + 2] += v.z; a[4 * i + 3] += v.w;
            }
        }
    } else if (MODE == 3) {     // LDS.128 broadcast: all lanes the same 16 bytes
        for (int it = 0; it < iters; ++it) {
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const float4 v = *reinterpret_cast<const float4*>(sm + ((128 * i + 4 * it) & 8188));
                a[4 * i] += v.x; a[4 * i + 1] += v.y; a[4 * i + 2] += v.z; a[4 * i + 3] += v.w;
            }
        }
    } else if (MODE == 4) {     // LDS.32 broadcast
        for (int it = 0; it < iters; ++it) {
#pragma unroll
            for (int i = 0; i < 16; ++i) a[i] += sm[(37 * i + it) & 8191];
        }
    } else if (MODE == 5) {     // SHFL
        for (int it = 0; it < iters; ++it) {
#pragma unroll
            for (int i = 0; i < 16; ++i) a[i] += __shfl_xor_sync(0xffffffffu, a[i], 1 + (i & 3));
        }
    } else if (MODE == 6) {     // MUFU ex2 + rcp (sigmoid)
        for (int it = 0; it < iters; ++it) {
#pragma unroll
            for (int i = 0; i < 16; ++i) {
                float e, r;
                asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(a[i]));
                asm volatile("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(1.0f + e));
                a[i] = r;
            }
        }
    } else if (MODE == 7) {     // __syncthreads latency
        for (int it = 0; it < iters; ++it) { __syncthreads(); a[0] += 1.0f; }
    } else if (MODE == 8) {     // dependent FFMA chain (latency)
        for (int it = 0; it < iters; ++it) {
#pragma unroll
            for (int i = 0; i < 16; ++i) a[0] = fmaf(a[0], x, y);
        }
    } else if (MODE == 9) {     // LDS.128 broadcast in 4 groups (4 distinct addresses per warp)
        for (int it = 0; it < iters; ++it) {
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const float4 v = *reinterpret_cast<const float4*>(sm + ((16 * (lane & 3) + 128 * i + 4 * it) & 8188));
                a[4 * i] += v.x; a[4 * i + 1] += v.y; a[4 * i + 2] += v.z; a[4 * i + 3] += v.w;
            }
        }
    } else if (MODE == 10) {    // dependent chain: STS -> barrier -> LDS (broadcast round trip through shared memory)
        for (int it = 0; it < iters; ++it) {
            if (lane == 0) sm[tid >> 5] = a[0];
            __syncthreads();
            a[0] += sm[(tid >> 5) ^ 1];
        }
    } else if (MODE == 11) {    // dependent SHFL chain (latency)
        for (int it = 0; it < iters; ++it) {
#pragma unroll
            for (int i = 0; i < 16; ++i) a[0] += __shfl_xor_sync(0xffffffffu, a[0], 1);
        }
    } else if (MODE == 12) {    // dependent sigmoid chain (latency of ex2 + add + rcp)
        for (int it = 0; it < iters; ++it) {
#pragma unroll
            for (int i = 0; i < 16; ++i) {
                float e, r;
                asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(a[0]));
                asm volatile("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(1.0f + e));
                a[0] = r;
            }
        }
    }
    long long t1 = clock64();
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < 16; ++i) s += a[i];
    out[blockIdx.x * blockDim.x + tid] = s;
    if (tid == 0) cyc[blockIdx.x] = t1 - t0;
}

// The recurrence warps' FMA section as it is compiled in lstm_ws_kernel: 128 weights in registers (never the same
// operand twice, so the register reuse cache cannot help), 16 h values from shared memory, 8 accumulator pairs.
//   MODE 0: fma.rn.f32x2 with a {h, h} pair built by mov.b64 (what ffma2() emits)   MODE 1: 128 scalar FFMA
//   MODE 2: fma.rn.f32x2 with the h pair packed once per k (4 FFMA2 share it)
template <int MODE>
__global__ void __launch_bounds__(256, 1) k_rec(float* out, long long* cyc, int iters, float x, float y) {
    __shared__ __align__(16) float sm[8192];
    const int tid = threadIdx.x, lane = tid & 31;
    for (int i = tid; i < 8192; i += blockDim.x) sm[i] = 0.001f * i;
    __syncthreads();
    float w[128];
#pragma unroll
    for (int i = 0; i < 128; ++i) w[i] = x + 0.01f * i + 0.001f * tid;
    float acc[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) acc[i] = 0.001f * (tid + i);
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
        float hr[16];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const float4 v = *reinterpret_cast<const float4*>(sm + ((16 * (lane & 3) + 128 * i + 4 * it) & 8188));
            hr[4 * i] = v.x; hr[4 * i + 1] = v.y; hr[4 * i + 2] = v.z; hr[4 * i + 3] = v.w;
        }
#pragma unroll
        for (int k = 0; k < 16; ++k) {
            if (MODE == 1) {
#pragma unroll
                for (int j = 0; j < 8; ++j) acc[8 * (k & 1) + j] = fmaf(w[8 * k + j], hr[k], acc[8 * (k & 1) + j]);
            } else {
                unsigned long long bb;
                if (MODE == 2) asm("mov.b64 %0, {%1, %1};" : "=l"(bb) : "f"(hr[k]));
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    unsigned long long dd, aa;
                    float& d0 = acc[8 * (k & 1) + 2 * j];
                    float& d1 = acc[8 * (k & 1) + 2 * j + 1];
                    asm("mov.b64 %0, {%1, %2};" : "=l"(dd) : "f"(d0), "f"(d1));
                    asm("mov.b64 %0, {%1, %2};" : "=l"(aa) : "f"(w[8 * k + 2 * j]), "f"(w[8 * k + 2 * j + 1]));
                    if (MODE == 0) asm("mov.b64 %0, {%1, %2};" : "=l"(bb) : "f"(hr[k]), "f"(hr[k]));
                    asm("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(dd) : "l"(aa), "l"(bb));
                    asm("mov.b64 {%0, %1}, %2;" : "=f"(d0), "=f"(d1) : "l"(dd));
                }
            }
        }
    }
    long long t1 = clock64();
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < 16; ++i) s += acc[i];
#pragma unroll
    for (int i = 0; i < 128; ++i) s += w[i] * 1e-9f;
    out[blockIdx.x * blockDim.x + tid] = s;
    if (tid == 0) cyc[blockIdx.x] = t1 - t0;
}

template <int MODE>
int run_rec(const char* name, int threads, int iters, double ops_per_iter_per_thread) {
    float* out; long long* cyc;
    CHECK(cudaMalloc(&out, 148 * 1024 * sizeof(float)));
    CHECK(cudaMalloc(&cyc, 148 * sizeof(long long)));
    k_rec<MODE><<<148, threads>>>(out, cyc, iters, 1.0001f, 0.9999f);
    k_rec<MODE><<<148, threads>>>(out, cyc, iters, 1.0001f, 0.9999f);
    CHECK(cudaDeviceSynchronize());
    long long h[148];
    CHECK(cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost));
    double c = (double)h[0];
    printf("%-52s threads=%4d  cycles/iter=%8.2f  warp-instr/cycle/SM=%6.3f\n", name, threads, c / iters,
           ops_per_iter_per_thread * (threads / 32) * iters / c);
    cudaFree(out); cudaFree(cyc);
    return 0;
}

template <int MODE>
int run(const char* name, int threads, int iters, double ops_per_iter_per_thread) {
    float* out; long long* cyc;
    CHECK(cudaMalloc(&out, 148 * 1024 * sizeof(float)));
    CHECK(cudaMalloc(&cyc, 148 * sizeof(long long)));
    k_bench<MODE><<<148, threads>>>(out, cyc, iters, 1.0001f, 0.9999f);
    k_bench<MODE><<<148, threads>>>(out, cyc, iters, 1.0001f, 0.9999f);
    CHECK(cudaDeviceSynchronize());
    long long h[148];
    CHECK(cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost));
    double c = (double)h[0];
    printf("%-44s threads=%4d  cycles/iter=%8.2f  warp-instr/cycle/SM=%6.3f\n", name, threads, c / iters,
           ops_per_iter_per_thread * (threads / 32) * iters / c);
    cudaFree(out); cudaFree(cyc);
    return 0;
}

int main() {
    for (int threads : {32, 128, 256, 512, 1024}) {
        run<0>("FFMA x16 independent", threads, 4096, 16);
        run<1>("FFMA2 (f32x2) x8 independent", threads, 4096, 8);
        run<2>("LDS.128 distinct x4", threads, 4096, 4);
        run<3>("LDS.128 broadcast x4", threads, 4096, 4);
        run<9>("LDS.128 4-address broadcast x4", threads, 4096, 4);
        run<4>("LDS.32 broadcast x16", threads, 4096, 16);
        run<5>("SHFL x16", threads, 4096, 16);
        run<6>("sigmoid (ex2+rcp) x16", threads, 4096, 16);
        run<7>("__syncthreads", threads, 4096, 1);
        run<10>("STS->bar->LDS round trip", threads, 4096, 1);
    }
    for (int threads : {32, 128, 256}) {
        run_rec<0>("rec FMA section: 64 FFMA2, 128 distinct weights", threads, 2048, 64);
        run_rec<1>("rec FMA section: 128 FFMA, 128 distinct weights", threads, 2048, 128);
        run_rec<2>("rec FMA section: 64 FFMA2, packed h operand", threads, 2048, 64);
    }
    run<8>("FFMA dependent chain x16", 32, 4096, 16);
    run<11>("SHFL dependent chain x16", 32, 4096, 16);
    run<12>("sigmoid dependent chain x16", 32, 4096, 16);
    return 0;
}
