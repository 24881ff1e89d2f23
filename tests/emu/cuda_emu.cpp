// cuda_emu.cpp — TEST INFRASTRUCTURE ONLY (see cuda_emu.h).
#include "cuda_emu.h"

thread_local dim3 threadIdx, blockIdx;
dim3 blockDim, gridDim;

namespace emu {
Block* g_block = nullptr;
std::mutex g_atomic_mu;
thread_local unsigned t_lane = 0, t_warp = 0;

// Runs the grid one block at a time on a pool of `blockDim` OS threads.  Kernels of this library never let a thread
// return while others of its block still have a barrier ahead, so a plain rendezvous closes each block.
void launch_impl(dim3 grid, dim3 block, size_t smem, const std::function<void()>& body) {
    const unsigned nthreads = block.x * block.y * block.z;
    if (nthreads == 0 || nthreads % 32 != 0) {
        fprintf(stderr, "emu: block size %u is not a positive multiple of 32\n", nthreads);
        abort();
    }
    blockDim = block;
    gridDim = grid;

    Block blk;
    blk.bar = std::make_unique<std::barrier<>>(nthreads);
    for (unsigned w = 0; w < nthreads / 32; ++w) blk.wbar.push_back(std::make_unique<std::barrier<>>(32));
    blk.xch.assign(nthreads, 0);
    blk.dyn.assign(smem + 64, 0);
    g_block = &blk;

    const unsigned long long nblocks = (unsigned long long)grid.x * grid.y * grid.z;
    std::vector<std::thread> pool;
    pool.reserve(nthreads);
    for (unsigned t = 0; t < nthreads; ++t) {
        pool.emplace_back([&, t]() {
            t_lane = t & 31;
            t_warp = t >> 5;
            threadIdx = dim3(t % block.x, (t / block.x) % block.y, t / (block.x * block.y));
            for (unsigned long long b = 0; b < nblocks; ++b) {
                blockIdx = dim3((unsigned)(b % grid.x), (unsigned)((b / grid.x) % grid.y),
                                (unsigned)(b / ((unsigned long long)grid.x * grid.y)));
                body();
                blk.bar->arrive_and_wait();
            }
        });
    }
    for (auto& th : pool) th.join();
    g_block = nullptr;
}
}   // namespace emu
