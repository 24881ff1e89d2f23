// Batch assembly on the device: what the reference's Dataset.__getitem__ + collate do on 16 host
// workers per step (src/datasets/general_multisrc_dataset_dis_embed.py:112-218) for a batch whose PCM is already in HBM:
//   * int16 PCM -> float32 / 32768                       (utils.read_audio_file_torch, torchaudio's int16 normalisation)
//   * target = sum of the reference-microphone solo tracks of the voices inside the bubble              (:141-171)
//   * the cheap per-channel perturbations: SampleShiftPerturbation (torch.roll), ChannelGainPerturbation,
//     ChannelDropPerturbation (never channel 0) and PeakNormPerturbation (scale / (peak + 1e-6) on mixture and target);
//     the random draws (shifts, gains, drops, scales) stay on the host, only the sample-sized work moves here
//   * the one-hot radius embedding                        (:194-201: 1 m -> [0,0,1], 1.5 m -> [0,1,0], 2 m -> [1,0,0])
// Shift, gain and drop are per-channel and commute; the peak normalisation comes last, as in the shipped perturbation
// lists.  The target follows the reference channel (microphone 0): its shift, its gain and the peak scale.
// HBM-bound: (M + V_inside) * 2 bytes read and (M + 1) * 4 bytes written per sample; one pass (+ a cheap peak pass).
#include "sb_common.cuh"

namespace sb {

constexpr int kPrepThreads = 256, kPrepPerThread = 8, kPrepChunk = kPrepThreads * kPrepPerThread;      // samples per CTA

// partial |x| maxima: ws[B + (b * M + m) * n_chunks + chunk]   (ws[0..B) holds the final per-row scale)
__global__ void __launch_bounds__(kPrepThreads) prepare_peak_kernel(const sb_prepare_args a, int n_chunks) {
    __shared__ int red[kPrepThreads / 32];
    const int chunk = blockIdx.x, m = blockIdx.y, b = blockIdx.z;
    const int16_t* src = a.mix + ((size_t)b * a.M + m) * a.N;
    pdl_wait();                                              // the PCM may come from a predecessor in the stream
    int mx = 0;
    const int n0 = chunk * kPrepChunk + threadIdx.x * kPrepPerThread;
    if (n0 + kPrepPerThread <= a.N && (a.N & 7) == 0) {
        const int4 q = *reinterpret_cast<const int4*>(src + n0);
        const int w[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int lo = (int16_t)(w[i] & 0xffff), hi = (int16_t)((unsigned)w[i] >> 16);
            mx = max(mx, max(lo < 0 ? -lo : lo, hi < 0 ? -hi : hi));
        }
    } else {
        for (int i = 0; i < kPrepPerThread && n0 + i < a.N; ++i) {
            const int v = src[n0 + i];
            mx = max(mx, v < 0 ? -v : v);
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) mx = max(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = mx;
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int w = 1; w < kPrepThreads / 32; ++w) mx = max(mx, red[w]);
        a.peak_ws[a.B + ((size_t)b * a.M + m) * n_chunks + chunk] = (float)mx;
    }
}

// one CTA per batch row: peak of the perturbed mixture -> PeakNormPerturbation's scale / (peak + 1e-6) into ws[b]
__global__ void __launch_bounds__(kPrepThreads) prepare_scale_kernel(const sb_prepare_args a, int n_chunks) {
    __shared__ float red[kPrepThreads / 32];
    const int b = blockIdx.x;
    pdl_wait();
    float peak = 0.f;
    for (int i = threadIdx.x; i < a.M * n_chunks; i += kPrepThreads) {
        const int mm = i / n_chunks;
        if (a.drop && a.drop[b * a.M + mm]) continue;
        const float gn = fabsf(a.gain ? a.gain[b * a.M + mm] : 1.0f);
        peak = fmaxf(peak, gn * a.peak_ws[a.B + (size_t)b * a.M * n_chunks + i] * (1.0f / 32768.0f));
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) peak = fmaxf(peak, __shfl_xor_sync(0xffffffffu, peak, o));
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = peak;
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int w = 1; w < kPrepThreads / 32; ++w) peak = fmaxf(peak, red[w]);
        const float ps = a.peak_scale[b];
        a.peak_ws[b] = ps != 0.0f ? ps / (peak + 1e-6f) : 1.0f;
    }
}

// grid (chunks, M + 1, B): rows 0..M-1 = mixture channels, row M = target
__global__ void __launch_bounds__(kPrepThreads) prepare_batch_kernel(const sb_prepare_args a) {
    __shared__ float s_scale;
    const int chunk = blockIdx.x, row = blockIdx.y, b = blockIdx.z;
    const bool is_target = row == a.M;
    const int m = is_target ? 0 : row;                       // the target follows the reference channel
    const int N = a.N;
    pdl_wait();
    if (threadIdx.x == 0) {
        s_scale = a.peak_scale ? a.peak_ws[b] : 1.0f;
        if (chunk == 0 && row == 0) {
            const int r = a.radius_idx ? a.radius_idx[b] : 0;
            a.dis_embed[b * 3 + 0] = r == 2 ? 1.f : 0.f;
            a.dis_embed[b * 3 + 1] = r == 1 ? 1.f : 0.f;
            a.dis_embed[b * 3 + 2] = r == 0 ? 1.f : 0.f;
        }
    }
    __syncthreads();
    const float g = (a.gain ? a.gain[b * a.M + m] : 1.0f) * s_scale * (1.0f / 32768.0f);
    int sh = a.shift ? a.shift[b * a.M + m] % N : 0;         // torch.roll(x, sh): out[n] = x[(n - sh) mod N]
    if (sh < 0) sh += N;
    const bool dropped = !is_target && a.drop && a.drop[b * a.M + m];
    float* dst = is_target ? a.target + (size_t)b * N : a.mixture + ((size_t)b * a.M + m) * N;
    const int n0 = chunk * kPrepChunk + threadIdx.x * kPrepPerThread;
    if (n0 >= N) return;
    float acc[kPrepPerThread];
#pragma unroll
    for (int i = 0; i < kPrepPerThread; ++i) acc[i] = 0.f;
    if (!dropped) {
        const int n_src = is_target ? a.V : 1;
        for (int v = 0; v < n_src; ++v) {
            if (is_target && !a.inside[b * a.V + v]) continue;
            const int16_t* src = is_target ? a.voices + ((size_t)b * a.V + v) * N : a.mix + ((size_t)b * a.M + m) * N;
            int p = n0 - sh;
            if (p < 0) p += N;
            if (sh == 0 && n0 + kPrepPerThread <= N && ((N & 7) == 0)) {           // aligned: one 16-byte load
                const int4 q = *reinterpret_cast<const int4*>(src + n0);
                const int w[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    acc[2 * i] += (float)(int16_t)(w[i] & 0xffff);
                    acc[2 * i + 1] += (float)(int16_t)((unsigned)w[i] >> 16);
                }
            } else {
#pragma unroll
                for (int i = 0; i < kPrepPerThread; ++i) {
                    if (n0 + i < N) acc[i] += (float)src[p];
                    if (++p == N) p = 0;
                }
            }
        }
    }
    if (n0 + kPrepPerThread <= N && ((N & 3) == 0)) {
        st4(dst + n0, make_float4(acc[0] * g, acc[1] * g, acc[2] * g, acc[3] * g));
        st4(dst + n0 + 4, make_float4(acc[4] * g, acc[5] * g, acc[6] * g, acc[7] * g));
    } else {
        for (int i = 0; i < kPrepPerThread && n0 + i < N; ++i) dst[n0 + i] = acc[i] * g;
    }
}

}  // namespace sb

extern "C" size_t sb_prepare_workspace_floats(int B, int M, int N) {
    if (B <= 0 || M <= 0 || N <= 0) return 0;
    return (size_t)B + (size_t)B * M * sb::ceil_div(N, sb::kPrepChunk);
}

extern "C" int sb_prepare_batch_fwd(const sb_prepare_args* p, void* stream) {
    using namespace sb;
    SB_REQUIRE(p && p->mix && p->mixture && p->target && p->dis_embed, SB_E_BADARG, "sb_prepare_batch_fwd: null pointer");
    SB_REQUIRE(p->B > 0 && p->M > 0 && p->N > 0 && p->V >= 0, SB_E_BADARG, "sb_prepare_batch_fwd: bad sizes");
    SB_REQUIRE(p->V == 0 || (p->voices && p->inside), SB_E_BADARG, "sb_prepare_batch_fwd: voices / inside are required when V > 0");
    SB_REQUIRE(!p->peak_scale || p->peak_ws, SB_E_BADARG, "sb_prepare_batch_fwd: peak normalisation needs the workspace");
    cudaStream_t st = (cudaStream_t)stream;
    const int n_chunks = ceil_div(p->N, kPrepChunk);
    if (p->peak_scale) {
        SB_CHECK(launch("prepare_peak", prepare_peak_kernel, dim3(n_chunks, p->M, p->B), dim3(kPrepThreads), 0, st, *p, n_chunks));
        SB_CHECK(launch("prepare_scale", prepare_scale_kernel, dim3(p->B), dim3(kPrepThreads), 0, st, *p, n_chunks));
    }
    return launch("prepare_batch", prepare_batch_kernel, dim3(n_chunks, p->M + 1, p->B), dim3(kPrepThreads), 0, st, *p);
}
