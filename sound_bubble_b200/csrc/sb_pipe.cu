// Native runtime of the pipelined streaming session (throughput mode of the edge/causal_infer.py:28-47 protocol).
//
// One call per 8 ms chunk, state carried, but consecutive chunks overlap on the GPU: every state tensor of the
// reference belongs to exactly one unit of the launch sequence (conv_buf: front-end; h0/c0 [+K/V]: the inter-frame path
// of one GridNet block; deconv_buf/istft_buf: back-end; DE3:403-421, 696-720), so chunk t+1 depends on chunk t per unit only.  The pipe
// owns `depth` streams; chunk t runs on stream t % depth as a few CUDA graphs (one per unit range, captured here from
// sb_net_forward_range) and range j of chunk t waits for the event range j of chunk t-1 recorded on its own stream.
// Per chunk the host issues 2 copies + n_ranges x (wait, graph launch, record): a few microseconds, no Python.
// The intra-frame units carry no state (DE3:819-827 starts every frame's BiLSTM from zeros), so a range made of one
// intra unit does not wait for its predecessor: the intra work of all chunks in flight runs side by side.
//
// Grouped throughput mode (sb_pipe_feed_chunk): when the sb_net_io entries describe G > 1 frames, the caller still feeds
// ONE 8 ms window per call; the pipe gathers G consecutive windows into the slot's wave buffer (the windows of the
// rolling protocol overlap by n_fft - stride samples, edge/causal_infer.py:39-40) and launches the group as one T = G
// call, so that the intra-frame recurrences of G chunks share tcgen05 tiles (G x B rows per direction) and the
// inter-frame kernel walks G steps per launch.  A partial last group is run eagerly with T = the number of pending chunks.
//
// All device memory of the network (windows, results, workspaces, both state arenas) belongs to the caller and arrives inside the
// sb_net_io array: entry k describes chunk numbers t with t % n_ios == k (n_ios = lcm(depth, 2): slot t % depth for
// wave / wave_out / workspace, arena t % 2 for the state inputs, the other arena for the state outputs).  The pipe itself
// allocates only the staging of the grouped mode's host path: G windows + G results per slot.
#include <vector>

#include "sb_common.cuh"

struct sb_pipe {
    const sb_net_desc* desc = nullptr;
    int depth = 0, n_ranges = 0, n_ios = 0;
    size_t window_bytes = 0, out_bytes = 0;
    long long n_calls = 0;
    std::vector<sb_net_io> io;
    std::vector<int> first, last;
    std::vector<char> stateless;                // per range: 1 = a single intra-frame unit (no carried state)
    std::vector<const float*> pend_win;         // grouped mode: windows / results of the chunks gathered so far
    std::vector<float*> pend_out;
    int B = 0, G = 0;
#ifndef SB_EMU
    std::vector<cudaStream_t> streams;
    std::vector<cudaEvent_t> events;            // [depth][n_ranges]
    std::vector<cudaEvent_t> joins;             // [depth]
    std::vector<cudaGraphExec_t> graphs;        // [n_ios][n_ranges]
    cudaEvent_t fork = nullptr;
    // grouped mode, host windows / results that are consecutive pieces of ONE host array (a recording read into memory, the
    // bench's pinned clip): per slot, room for G contiguous windows and G contiguous results on the device - the only device
    // memory the pipe owns - so that a group costs one copy each way plus the gather / scatter kernels
    float* stage_in = nullptr;
    float* stage_out = nullptr;
    size_t win_floats = 0, res_floats = 0;      // one chunk's window [B][M][n_fft] / result [B][S][stride]
#endif
};

#ifndef SB_EMU
namespace sb {

constexpr int kMaxGather = 64;              // chunks of a group the gather / scatter kernels take (pointer table in the parameters)

static int cuda_fail(const char* what, cudaError_t e) {
    set_error("%s: %s", what, cudaGetErrorString(e));
    (void)cudaGetLastError();
    return (int)e;
}
#define SB_CUDA(call)                                           \
    do {                                                        \
        const cudaError_t sb_e_ = (call);                       \
        if (sb_e_ != cudaSuccess) return cuda_fail(#call, sb_e_); \
    } while (0)

static void destroy(sb_pipe* p) {
    if (!p) return;
    for (auto g : p->graphs) if (g) cudaGraphExecDestroy(g);
    for (auto e : p->events) if (e) cudaEventDestroy(e);
    for (auto e : p->joins) if (e) cudaEventDestroy(e);
    if (p->fork) cudaEventDestroy(p->fork);
    if (p->stage_in) cudaFree(p->stage_in);
    if (p->stage_out) cudaFree(p->stage_out);
    for (auto s : p->streams) if (s) cudaStreamDestroy(s);
    delete p;
}

static int capture_range(sb_pipe* p, int k, int j, cudaStream_t st, cudaGraphExec_t* out) {
    SB_CUDA(cudaStreamBeginCapture(st, cudaStreamCaptureModeThreadLocal));
    const int rc = sb_net_forward_range(p->desc, &p->io[k], p->first[j], p->last[j], st);
    cudaGraph_t g = nullptr;
    const cudaError_t e = cudaStreamEndCapture(st, &g);
    if (rc != 0) {
        if (g) cudaGraphDestroy(g);
        return rc;
    }
    if (e != cudaSuccess) return cuda_fail("cudaStreamEndCapture", e);
    const cudaError_t e2 = cudaGraphInstantiate(out, g, 0);
    cudaGraphDestroy(g);
    if (e2 != cudaSuccess) return cuda_fail("cudaGraphInstantiate", e2);
    return 0;
}

}  // namespace sb
#endif

extern "C" int sb_pipe_create(const sb_net_desc* d, const sb_net_io* ios, int n_ios, int depth, const int* range_first,
                              const int* range_last, int n_ranges, sb_pipe** out) {
    using namespace sb;
    SB_REQUIRE(d && ios && range_first && range_last && out, SB_E_BADARG, "sb_pipe_create: null pointer");
    SB_REQUIRE(depth >= 1 && depth <= 64, SB_E_BADARG, "sb_pipe_create: depth %d out of range", depth);
    const int period = depth % 2 == 0 ? depth : 2 * depth;
    SB_REQUIRE(n_ios == period, SB_E_BADARG, "sb_pipe_create: need %d sb_net_io entries for depth %d, got %d", period, depth, n_ios);
    SB_REQUIRE(n_ranges >= 1 && n_ranges <= 2 * SB_MAX_BLOCKS + 2, SB_E_BADARG, "sb_pipe_create: bad number of ranges %d", n_ranges);
    int next = 0;
    for (int j = 0; j < n_ranges; ++j) {
        SB_REQUIRE(range_first[j] == next && range_last[j] >= range_first[j], SB_E_BADARG,
                   "sb_pipe_create: ranges must cover units 0..%d in order", 2 * d->n_blocks + 1);
        next = range_last[j] + 1;
    }
    SB_REQUIRE(next == 2 * d->n_blocks + 2, SB_E_BADARG, "sb_pipe_create: ranges must cover units 0..%d in order", 2 * d->n_blocks + 1);
    for (int k = 0; k < n_ios; ++k) {
        SB_REQUIRE(ios[k].T >= 1 && ios[k].T == ios[0].T && ios[k].B == ios[0].B && ios[k].B > 0, SB_E_BADARG,
                   "sb_pipe_create: every sb_net_io must describe the same number of frames of the same batch");
        SB_REQUIRE(ios[k].wave == ios[k % depth].wave && ios[k].wave_out == ios[k % depth].wave_out &&
                   ios[k].workspace == ios[k % depth].workspace, SB_E_BADARG,
                   "sb_pipe_create: entries of one slot must share wave / wave_out / workspace");
    }
#ifdef SB_EMU
    set_error("sb_pipe_create: streams and graphs do not exist in the host-emulated test build");
    return SB_E_UNSUPP;
#else
    sb_pipe* p = new sb_pipe();
    p->desc = d; p->depth = depth; p->n_ranges = n_ranges; p->n_ios = n_ios;
    p->io.assign(ios, ios + n_ios);
    p->first.assign(range_first, range_first + n_ranges);
    p->last.assign(range_last, range_last + n_ranges);
    p->stateless.assign(n_ranges, 0);
    for (int j = 0; j < n_ranges; ++j)          // units 1 + 2i are the intra-frame paths
        p->stateless[j] = range_first[j] == range_last[j] && (range_first[j] & 1) && range_first[j] <= 2 * d->n_blocks - 1;
    p->B = ios[0].B; p->G = ios[0].T;
    p->window_bytes = sizeof(float) * (size_t)ios[0].B * d->M * ((size_t)d->stride * ios[0].T + d->n_fft - d->stride);
    p->out_bytes = sizeof(float) * (size_t)ios[0].B * d->n_src * d->stride * ios[0].T;
    p->streams.assign(depth, nullptr);
    p->events.assign((size_t)depth * n_ranges, nullptr);
    p->joins.assign(depth, nullptr);
    p->graphs.assign((size_t)n_ios * n_ranges, nullptr);
    auto fail = [&](int rc) { destroy(p); return rc; };
    cudaError_t e = cudaEventCreateWithFlags(&p->fork, cudaEventDisableTiming);
    if (e != cudaSuccess) return fail(cuda_fail("cudaEventCreate", e));
    p->win_floats = (size_t)p->B * d->M * d->n_fft;
    p->res_floats = (size_t)p->B * d->n_src * d->stride;
    if (p->G > 1 && p->G <= kMaxGather) {
        if ((e = cudaMalloc(&p->stage_in, sizeof(float) * p->win_floats * p->G * depth)) != cudaSuccess) return fail(cuda_fail("cudaMalloc (window staging)", e));
        if ((e = cudaMalloc(&p->stage_out, sizeof(float) * p->res_floats * p->G * depth)) != cudaSuccess) return fail(cuda_fail("cudaMalloc (result staging)", e));
    }
    for (int s = 0; s < depth; ++s) {
        if ((e = cudaStreamCreateWithFlags(&p->streams[s], cudaStreamNonBlocking)) != cudaSuccess) return fail(cuda_fail("cudaStreamCreate", e));
        if ((e = cudaEventCreateWithFlags(&p->joins[s], cudaEventDisableTiming)) != cudaSuccess) return fail(cuda_fail("cudaEventCreate", e));
        for (int j = 0; j < n_ranges; ++j)
            if ((e = cudaEventCreateWithFlags(&p->events[(size_t)s * n_ranges + j], cudaEventDisableTiming)) != cudaSuccess)
                return fail(cuda_fail("cudaEventCreate", e));
    }
    for (int k = 0; k < n_ios; ++k)
        for (int j = 0; j < n_ranges; ++j) {
            const int rc = capture_range(p, k, j, p->streams[k % depth], &p->graphs[(size_t)k * n_ranges + j]);
            if (rc != 0) return fail(rc);
        }
    *out = p;
    return 0;
#endif
}

extern "C" int sb_pipe_destroy(sb_pipe* p) {
#ifndef SB_EMU
    if (p) {
        for (auto s : p->streams) if (s) cudaStreamSynchronize(s);
        sb::destroy(p);
    }
#else
    delete p;
#endif
    return 0;
}

/* Orders every stream of the pipe after what `caller_stream` has enqueued so far (the windows, a state reset). */
extern "C" int sb_pipe_begin(sb_pipe* p, void* caller_stream) {
    using namespace sb;
    SB_REQUIRE(p, SB_E_BADARG, "sb_pipe_begin: null pipe");
#ifndef SB_EMU
    SB_CUDA(cudaEventRecord(p->fork, (cudaStream_t)caller_stream));
    for (auto s : p->streams) SB_CUDA(cudaStreamWaitEvent(s, p->fork, 0));
#endif
    return 0;
}

/* Restarts the chunk counter (the caller zeroes or reloads the state arena 0 on its own stream, then sb_pipe_begin). */
extern "C" int sb_pipe_reset(sb_pipe* p) {
    using namespace sb;
    SB_REQUIRE(p, SB_E_BADARG, "sb_pipe_reset: null pipe");
    p->n_calls = 0;
    p->pend_win.clear();
    p->pend_out.clear();
    return 0;
}

extern "C" long long sb_pipe_calls(const sb_pipe* p) { return p ? p->n_calls : -1; }

#ifndef SB_EMU
namespace sb {
// one call of the launch sequence on the slot's stream: per range (wait for the same range of the previous call unless
// it is stateless, run, record).  n_frames == 0: replay the captured graphs; otherwise an eager T = n_frames call.
static int run_ranges(sb_pipe* p, long long t, int n_frames) {
    const int slot = (int)(t % p->depth), k = (int)(t % p->n_ios), R = p->n_ranges;
    cudaStream_t st = p->streams[slot];
    const cudaEvent_t* prev = &p->events[(size_t)((t + p->depth - 1) % p->depth) * R];
    cudaEvent_t* mine = &p->events[(size_t)slot * R];
    sb_net_io tail = p->io[k];
    if (n_frames > 0) tail.T = n_frames;
    for (int j = 0; j < R; ++j) {
        if (t > 0 && p->depth > 1 && !p->stateless[j]) SB_CUDA(cudaStreamWaitEvent(st, prev[j], 0));
        if (n_frames > 0) SB_CHECK(sb_net_forward_range(p->desc, &tail, p->first[j], p->last[j], st));
        else SB_CUDA(cudaGraphLaunch(p->graphs[(size_t)k * R + j], st));
        if (!p->stateless[j]) SB_CUDA(cudaEventRecord(mine[j], st));
    }
    return 0;
}

// Device-resident windows / results of a group move with ONE kernel each way: as n pitched cudaMemcpy2DAsync calls they ran as
// n small copy kernels per direction that each had to find a free SM among the one-CTA-per-SM LSTM kernels of the other groups
// in flight (the device-resident pass of bench.py was 14 % slower than the pass from pinned host memory, whose copies use the
// copy engines).
struct ChunkPtrs { const float* src[kMaxGather]; float* dst[kMaxGather]; };

// window c = [rows][nfft] contiguous -> wave[row * pitch + c * hop + i]
__global__ void gather_windows_kernel(const ChunkPtrs ptrs, float* wave, int n, int rows, int nfft, int hop, long long pitch) {
    pdl_wait();
    const long long per = (long long)rows * nfft, total = per * n;
    for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (long long)gridDim.x * blockDim.x) {
        const int c = (int)(idx / per);
        const long long e = idx - (long long)c * per;
        const int row = (int)(e / nfft), i = (int)(e - (long long)row * nfft);
        wave[(long long)row * pitch + (long long)c * hop + i] = __ldg(ptrs.src[c] + e);
    }
}
// wave_out[row * pitch + c * hop + i] -> result c = [rows][hop] contiguous (NULL: not wanted)
__global__ void scatter_results_kernel(const ChunkPtrs ptrs, const float* wave_out, int n, int rows, int hop, long long pitch) {
    pdl_wait();
    const long long per = (long long)rows * hop, total = per * n;
    for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (long long)gridDim.x * blockDim.x) {
        const int c = (int)(idx / per);
        float* dst = ptrs.dst[c];
        if (!dst) continue;
        const long long e = idx - (long long)c * per;
        const int row = (int)(e / hop), i = (int)(e - (long long)row * hop);
        dst[e] = wave_out[(long long)row * pitch + (long long)c * hop + i];
    }
}

static bool on_device(const void* ptr) {
    cudaPointerAttributes at{};
    if (cudaPointerGetAttributes(&at, ptr) != cudaSuccess) {
        (void)cudaGetLastError();
        return false;
    }
    return at.type == cudaMemoryTypeDevice;
}

// launches the chunks gathered so far as one call of n = pend_win.size() frames (n == G: the captured graphs)
static int launch_group(sb_pipe* p) {
    const int n = (int)p->pend_win.size();
    if (n == 0) return 0;
    const long long t = p->n_calls;
    const int slot = (int)(t % p->depth), k = (int)(t % p->n_ios);
    cudaStream_t st = p->streams[slot];
    const sb_net_io& io = p->io[k];
    const sb_net_desc* d = p->desc;
    const size_t hop = (size_t)d->stride, nfft = (size_t)d->n_fft;
    const size_t in_pitch = sizeof(float) * (hop * n + nfft - hop), out_pitch = sizeof(float) * hop * n;
    // the memory kind of a group is taken from its first window / first wanted result (a group does not mix kinds)
    float* first_out = nullptr;
    for (int c = 0; c < n && !first_out; ++c) first_out = p->pend_out[c];
    const bool dev_in = n <= kMaxGather && on_device(p->pend_win[0]);
    const bool dev_out = n <= kMaxGather && first_out && on_device(first_out);
    // Host windows / results that are consecutive pieces of one array travel as ONE copy through the pipe's device staging
    // (64 pitched copies cost 0.6 ms before a group's first kernel can start - exposed at the start of a pass - and 64 API calls)
    auto consecutive = [&](auto get, size_t floats) {
        for (int c = 1; c < n; ++c)
            if (!get(c) || get(c) != get(c - 1) + floats) return false;
        return n > 1 && get(0) != nullptr;
    };
    const bool stage_in = !dev_in && p->stage_in && n <= p->G && consecutive([&](int c) { return p->pend_win[c]; }, p->win_floats);
    const bool stage_out = !dev_out && p->stage_out && n <= p->G && consecutive([&](int c) { return (const float*)p->pend_out[c]; }, p->res_floats);
    ChunkPtrs ptrs{};
    for (int c = 0; c < n && c < kMaxGather; ++c) {
        ptrs.src[c] = stage_in ? p->stage_in + ((size_t)slot * p->G + c) * p->win_floats : p->pend_win[c];
        ptrs.dst[c] = !p->pend_out[c] ? nullptr : stage_out ? p->stage_out + ((size_t)slot * p->G + c) * p->res_floats : p->pend_out[c];
    }
    if (stage_in)
        SB_CUDA(cudaMemcpyAsync(const_cast<float*>(ptrs.src[0]), p->pend_win[0], sizeof(float) * p->win_floats * n, cudaMemcpyDefault, st));
    if (dev_in || stage_in) {
        const long long total = (long long)n * p->B * d->M * (long long)nfft;
        SB_CHECK(launch("gather_windows", gather_windows_kernel, dim3((unsigned)ceil_div_ll(total, 256 * 8)), dim3(256), 0, st, ptrs,
                        const_cast<float*>(io.wave), n, p->B * d->M, (int)nfft, (int)hop, (long long)(in_pitch / sizeof(float))));
    } else {
        for (int c = 0; c < n; ++c)             // window c covers samples [c*hop, c*hop + n_fft) of the group's wave
            SB_CUDA(cudaMemcpy2DAsync(const_cast<float*>(io.wave) + c * hop, in_pitch, p->pend_win[c], sizeof(float) * nfft,
                                      sizeof(float) * nfft, (size_t)p->B * d->M, cudaMemcpyDefault, st));
    }
    SB_CHECK(run_ranges(p, t, n == p->G ? 0 : n));
    if (dev_out || stage_out) {
        const long long total = (long long)n * p->B * d->n_src * (long long)hop;
        SB_CHECK(launch("scatter_results", scatter_results_kernel, dim3((unsigned)ceil_div_ll(total, 256 * 8)), dim3(256), 0, st, ptrs,
                        (const float*)io.wave_out, n, p->B * d->n_src, (int)hop, (long long)(out_pitch / sizeof(float))));
        if (stage_out)
            SB_CUDA(cudaMemcpyAsync(p->pend_out[0], ptrs.dst[0], sizeof(float) * p->res_floats * n, cudaMemcpyDefault, st));
    } else {
        for (int c = 0; c < n; ++c)
            if (p->pend_out[c])
                SB_CUDA(cudaMemcpy2DAsync(p->pend_out[c], sizeof(float) * hop, io.wave_out + c * hop, out_pitch,
                                          sizeof(float) * hop, (size_t)p->B * d->n_src, cudaMemcpyDefault, st));
    }
    p->pend_win.clear();
    p->pend_out.clear();
    p->n_calls = t + 1;
    return 0;
}
}  // namespace sb
#endif

/* Grouped mode: one 8 ms window [B][M][n_fft] per call (host, pinned for asynchronous copies, or device); the result    */
/* [B][S][stride] lands in `out` (may be NULL) once the chunk's group of G = ios[].T frames has run.  The window must stay  */
/* untouched until its group has been launched (G calls later, or sb_pipe_flush / sb_pipe_end).                          */
extern "C" int sb_pipe_feed_chunk(sb_pipe* p, const float* window, float* out) {
    using namespace sb;
    SB_REQUIRE(p && window, SB_E_BADARG, "sb_pipe_feed_chunk: null pipe / window");
#ifndef SB_EMU
    p->pend_win.push_back(window);
    p->pend_out.push_back(out);
    if ((int)p->pend_win.size() == p->G) return launch_group(p);
    return 0;
#else
    (void)out;
    return SB_E_UNSUPP;
#endif
}

/* Runs the chunks of a partial group now (eager T = pending call on the slot's stream). */
extern "C" int sb_pipe_flush(sb_pipe* p) {
    using namespace sb;
    SB_REQUIRE(p, SB_E_BADARG, "sb_pipe_flush: null pipe");
#ifndef SB_EMU
    return launch_group(p);
#else
    return 0;
#endif
}

/* Enqueues one chunk of T frames and returns.  window: [B][M][stride*T + n_fft - stride] floats, host (pinned for a     */
/* truly asynchronous copy) or device, NULL = the slot's window buffer was filled by the caller; out: [B][S][stride*T],   */
/* host or device, NULL = leave the result in the slot's wave_out buffer (valid until `depth` calls later).              */
extern "C" int sb_pipe_feed(sb_pipe* p, const float* window, float* out) {
    using namespace sb;
    SB_REQUIRE(p, SB_E_BADARG, "sb_pipe_feed: null pipe");
#ifndef SB_EMU
    SB_REQUIRE(p->pend_win.empty(), SB_E_BADARG, "sb_pipe_feed: chunks of a group are pending (sb_pipe_feed_chunk); flush first");
    const long long t = p->n_calls;
    const int slot = (int)(t % p->depth), k = (int)(t % p->n_ios);
    cudaStream_t st = p->streams[slot];
    const sb_net_io& io = p->io[k];
    if (window) SB_CUDA(cudaMemcpyAsync(const_cast<float*>(io.wave), window, p->window_bytes, cudaMemcpyDefault, st));
    SB_CHECK(run_ranges(p, t, 0));
    if (out) SB_CUDA(cudaMemcpyAsync(out, io.wave_out, p->out_bytes, cudaMemcpyDefault, st));
    p->n_calls = t + 1;
    return 0;
#else
    (void)window; (void)out;
    return SB_E_UNSUPP;
#endif
}

/* Makes `caller_stream` wait for every chunk fed so far. */
extern "C" int sb_pipe_end(sb_pipe* p, void* caller_stream) {
    using namespace sb;
    SB_REQUIRE(p, SB_E_BADARG, "sb_pipe_end: null pipe");
#ifndef SB_EMU
    SB_CHECK(launch_group(p));
    for (int s = 0; s < p->depth; ++s) {
        SB_CUDA(cudaEventRecord(p->joins[s], p->streams[s]));
        SB_CUDA(cudaStreamWaitEvent((cudaStream_t)caller_stream, p->joins[s], 0));
    }
#endif
    return 0;
}
