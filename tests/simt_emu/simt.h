// TEST INFRASTRUCTURE: a minimal SIMT emulator that runs CUDA kernel SOURCE on the host, one std::thread per CUDA
// thread, one block at a time: __syncthreads is a block barrier, __shfl_down_sync exchanges through a per-warp
// buffer between two warp barriers, atomicAdd is an atomic fetch_add, static and dynamic shared memory are process
// globals (only one block is alive at a time).  It executes the kernels of forte2_b200/csrc/rdm_kernels.cuh
// unchanged (tests/simt_emu/build.sh only rewrites the two shared-memory declarations) so that the CPU suite can
// check the kernel text itself -- indexing, reductions, barrier placement -- against the definitional RDMs before
// the kernels ever run on a GPU.  It does not model memory ordering, warp divergence or performance.
#pragma once
#include <atomic>
#include <barrier>
#include <cstdint>
#include <functional>
#include <memory>
#include <thread>
#include <vector>

struct Dim3 {
    unsigned x = 1, y = 1, z = 1;
};
inline thread_local Dim3 threadIdx;
inline Dim3 blockIdx, blockDim, gridDim;

#define __global__
#define __device__
#define __host__
#define __forceinline__ inline
#define __launch_bounds__(...)
#define __restrict__

namespace simt {
inline std::unique_ptr<std::barrier<>> block_barrier;
inline std::vector<std::unique_ptr<std::barrier<>>> warp_barrier;
inline std::vector<double> warp_buf; // [nwarps][32]
inline double dyn_smem[1 << 16];
} // namespace simt

// dynamic shared memory of the kernels under test
#define rdm_smem simt::dyn_smem

inline void __syncthreads() { simt::block_barrier->arrive_and_wait(); }

inline double __shfl_down_sync(unsigned, double v, int delta) {
    const unsigned warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    simt::warp_buf[warp * 32 + lane] = v;
    simt::warp_barrier[warp]->arrive_and_wait();
    const double r = lane + delta < 32 ? simt::warp_buf[warp * 32 + lane + delta] : v;
    simt::warp_barrier[warp]->arrive_and_wait();
    return r;
}

inline double atomicAdd(double* p, double v) { return std::atomic_ref<double>(*p).fetch_add(v); }

namespace simt {
// kernel<<<grid, block>>>(args...) with the blocks run one after the other
template <typename F> void launch(unsigned grid, unsigned block, F&& body) {
    gridDim.x = grid;
    blockDim.x = block;
    const unsigned nwarps = (block + 31) / 32;
    for (unsigned b = 0; b < grid; ++b) {
        blockIdx.x = b;
        block_barrier = std::make_unique<std::barrier<>>(block);
        warp_barrier.clear();
        for (unsigned w = 0; w < nwarps; ++w)
            warp_barrier.push_back(std::make_unique<std::barrier<>>(std::min(32u, block - 32 * w)));
        warp_buf.assign(static_cast<size_t>(nwarps) * 32, 0.0);
        std::vector<std::thread> threads;
        for (unsigned t = 0; t < block; ++t)
            threads.emplace_back([&, t] {
                threadIdx.x = t;
                body();
            });
        for (auto& th : threads)
            th.join();
    }
}
} // namespace simt
