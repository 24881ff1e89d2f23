// Host-side bookkeeping of libsoundbubble_sm100a.so: error strings, launch accounting, options, ABI self-description.
#include <stdarg.h>

#include <atomic>
#include <mutex>
#include <set>
#include <utility>

#include "sb_common.cuh"

namespace sb {

static thread_local char g_err[512] = "";
static std::atomic<uint64_t> g_launches{0};
static std::atomic<int> g_pdl{0};
static std::atomic<int> g_attn_tc{1};
static std::atomic<int> g_train_one_row{0};
static std::atomic<int> g_train_ffma2{1};
static std::atomic<int> g_tc_v1{0};
static std::atomic<int> g_tc_cell7{1};
static std::atomic<int> g_train_tc{1};
static std::atomic<int> g_tc_pipe{1};
static std::atomic<int> g_front_tc{1};
static std::atomic<int> g_tc_cw16{0};

void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

void count_launch() { g_launches.fetch_add(1, std::memory_order_relaxed); }

int check_launch(const char* what) {
    const cudaError_t e = cudaGetLastError();
    if (e == cudaSuccess) return 0;
    set_error("%s: %s", what, cudaGetErrorString(e));
    return (int)e;
}

bool pdl_enabled() { return g_pdl.load(std::memory_order_relaxed) != 0; }
bool attn_tc_enabled() { return g_attn_tc.load(std::memory_order_relaxed) != 0; }
bool train_one_row_enabled() { return g_train_one_row.load(std::memory_order_relaxed) != 0; }
bool train_ffma2_enabled() { return g_train_ffma2.load(std::memory_order_relaxed) != 0; }
bool tc_v1_enabled() { return g_tc_v1.load(std::memory_order_relaxed) != 0; }
bool tc_cell7_enabled() { return g_tc_cell7.load(std::memory_order_relaxed) != 0; }
bool train_tc_enabled() { return g_train_tc.load(std::memory_order_relaxed) != 0; }
bool tc_cw16_enabled() { return g_tc_cw16.load(std::memory_order_relaxed) != 0; }
bool front_tc_enabled() { return g_front_tc.load(std::memory_order_relaxed) != 0; }
bool tc_pipe_enabled() { return g_tc_pipe.load(std::memory_order_relaxed) != 0; }

int sm_count() {
    static thread_local int dev_cached = -1, sms = 0;
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev != dev_cached) {
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
        if (sms <= 0) sms = 148;
        dev_cached = dev;
    }
    return sms;
}

// cudaFuncSetAttribute once per (device, kernel); the set only grows (a handful of kernels).
int ensure_smem(const void* func, size_t bytes, const char* name) {
#ifdef SB_EMU
    (void)func; (void)bytes; (void)name;
    return 0;
#else
    static std::mutex mu;
    static std::set<std::pair<int, const void*>> done;
    int dev = 0;
    cudaGetDevice(&dev);
    std::lock_guard<std::mutex> lock(mu);
    if (done.count({dev, func})) return 0;
    cudaFuncAttributes fa{};
    size_t stat = 0;                        // static shared memory counts against the same 227 KB
    if (cudaFuncGetAttributes(&fa, func) == cudaSuccess) stat = fa.sharedSizeBytes;
    const cudaError_t e = cudaFuncSetAttribute(func, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(227 * 1024 - stat));
    if (e != cudaSuccess) {
        (void)cudaGetLastError();
        set_error("%s: cannot opt in to %zu bytes of shared memory: %s", name, bytes, cudaGetErrorString(e));
        return SB_E_SMEM;
    }
    done.insert({dev, func});
    return 0;
#endif
}

}  // namespace sb

extern "C" int sb_version(void) { return SB_VERSION; }
extern "C" const char* sb_last_error_string(void) { return sb::g_err; }
extern "C" uint64_t sb_launch_count(void) { return sb::g_launches.load(); }

extern "C" int sb_set_option(int option, int value) {
    if (option == SB_OPT_PDL) {
        sb::g_pdl.store(value ? 1 : 0);
        return 0;
    }
    if (option == SB_OPT_ATTN_TC) {
        sb::g_attn_tc.store(value ? 1 : 0);
        return 0;
    }
    if (option == SB_OPT_TRAIN_FFMA2) {
        sb::g_train_ffma2.store(value ? 1 : 0);
        return 0;
    }
    if (option == SB_OPT_TC_CW16) {
        sb::g_tc_cw16.store(value ? 1 : 0);
        return 0;
    }
    if (option == SB_OPT_FRONT_TC) {
        sb::g_front_tc.store(value ? 1 : 0);
        return 0;
    }
    if (option == SB_OPT_TC_PIPE) {
        sb::g_tc_pipe.store(value ? 1 : 0);
        return 0;
    }
    if (option == SB_OPT_TRAIN_TC) {
        sb::g_train_tc.store(value ? 1 : 0);
        return 0;
    }
    if (option == SB_OPT_TC_CELL7) {
        sb::g_tc_cell7.store(value ? 1 : 0);
        return 0;
    }
    if (option == SB_OPT_TC_V1) {
        sb::g_tc_v1.store(value ? 1 : 0);
        return 0;
    }
    if (option == SB_OPT_TRAIN_ONE_ROW) {
        sb::g_train_one_row.store(value ? 1 : 0);
        return 0;
    }
    sb::set_error("sb_set_option: unknown option %d", option);
    return SB_E_BADARG;
}

extern "C" int sb_abi_sizeof(int which) {
    switch (which) {
        case 0: return (int)sizeof(sb_lstm_dir);
        case 1: return (int)sizeof(sb_stft_args);
        case 2: return (int)sizeof(sb_conv_in_args);
        case 3: return (int)sizeof(sb_film_args);
        case 4: return (int)sizeof(sb_intra_args);
        case 5: return (int)sizeof(sb_inter_args);
        case 6: return (int)sizeof(sb_backend_args);
        case 7: return (int)sizeof(sb_net_desc);
        case 8: return (int)sizeof(sb_net_io);
        case 9: return (int)sizeof(sb_intra_conv_args);
        case 10: return (int)sizeof(sb_attn_proj);
        case 11: return (int)sizeof(sb_attn_args);
        case 12: return (int)sizeof(sb_block_desc);
        case 13: return (int)sizeof(sb_prepare_args);
        case 14: return (int)sizeof(sb_path_train_args);
        case 15: return (int)sizeof(sb_path_bwd_args);
        case 16: return (int)sizeof(sb_film_apply_args);
        case 17: return (int)sizeof(sb_film_bwd_args);
        case 18: return (int)sizeof(sb_conv_in_train_args);
        case 19: return (int)sizeof(sb_backend_bwd_args);
        case 20: return (int)sizeof(sb_convpath_train_args);
        case 21: return (int)sizeof(sb_convpath_bwd_args);
        case 22: return (int)sizeof(sb_attn_train_args);
        case 23: return (int)sizeof(sb_attn_proj_grad);
        case 24: return (int)sizeof(sb_attn_bwd_args);
        default: return -1;
    }
}
