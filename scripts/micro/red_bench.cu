// Microbenchmark: throughput of FP64 reductions into global memory on B200.
//   redg   : red.global.add.f64, one 8-byte element per lane, 32 consecutive elements per warp instruction
//   blkred : cp.reduce.async.bulk ... .add.f64 (TMA bulk reduce) of S bytes from shared memory
// Both hit pseudo-random S-byte aligned segments of a 1 GiB buffer.  Build: nvcc -O3 -gencode
// arch=compute_100a,code=sm_100a -o red_bench red_bench.cu ; run: ./red_bench
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t hash32(uint32_t x) {
    x ^= x >> 16; x *= 0x7feb352dU; x ^= x >> 15; x *= 0x846ca68bU; x ^= x >> 16; return x;
}

__global__ void redg_kernel(double* buf, uint64_t nseg, int seg_doubles, int iters) {
    const int lane = threadIdx.x & 31;
    const uint32_t wid = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    for (int it = 0; it < iters; ++it) {
        const uint64_t seg = hash32(wid * 7919u + it) % nseg;
        double* p = buf + seg * seg_doubles;
        for (int e = lane; e < seg_doubles; e += 32)
            atomicAdd(p + e, 1.0);
    }
}

__global__ void blkred_kernel(double* buf, uint64_t nseg, int seg_bytes, int iters) {
    extern __shared__ __align__(128) unsigned char smem[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    double* mine = reinterpret_cast<double*>(smem) + warp * (seg_bytes / 8);
    for (int e = lane; e < seg_bytes / 8; e += 32)
        mine[e] = 1.0;
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncwarp();
    const uint32_t wid = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (lane == 0) {
        const unsigned sa = (unsigned)__cvta_generic_to_shared(mine);
        for (int it = 0; it < iters; ++it) {
            const uint64_t seg = hash32(wid * 7919u + it) % nseg;
            unsigned char* p = reinterpret_cast<unsigned char*>(buf) + seg * (uint64_t)seg_bytes;
            asm volatile("cp.reduce.async.bulk.global.shared::cta.bulk_group.add.f64 [%0], [%1], %2;"
                         :: "l"(p), "r"(sa), "r"(seg_bytes) : "memory");
            if ((it & 7) == 7) {
                asm volatile("cp.async.bulk.commit_group;" ::: "memory");
                asm volatile("cp.async.bulk.wait_group.read 2;" ::: "memory");
            }
        }
        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
    }
}

// half of the warps of every CTA reduce with REDG, the other half with TMA bulk reductions
__global__ void mixed_kernel(double* buf, uint64_t nseg, int seg_bytes, int iters) {
    extern __shared__ __align__(128) unsigned char smem[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint32_t wid = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int sd = seg_bytes / 8;
    if (warp & 1) {
        for (int it = 0; it < iters; ++it) {
            const uint64_t seg = hash32(wid * 7919u + it) % nseg;
            double* p = buf + seg * sd;
            for (int e = lane; e < sd; e += 32)
                atomicAdd(p + e, 1.0);
        }
    } else {
        double* mine = reinterpret_cast<double*>(smem) + warp * sd;
        for (int e = lane; e < sd; e += 32)
            mine[e] = 1.0;
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        __syncwarp();
        if (lane == 0) {
            const unsigned sa = (unsigned)__cvta_generic_to_shared(mine);
            for (int it = 0; it < iters; ++it) {
                const uint64_t seg = hash32(wid * 7919u + it) % nseg;
                unsigned char* p = reinterpret_cast<unsigned char*>(buf) + seg * (uint64_t)seg_bytes;
                asm volatile("cp.reduce.async.bulk.global.shared::cta.bulk_group.add.f64 [%0], [%1], %2;"
                             :: "l"(p), "r"(sa), "r"(seg_bytes) : "memory");
                if ((it & 7) == 7) {
                    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
                    asm volatile("cp.async.bulk.wait_group.read 2;" ::: "memory");
                }
            }
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
            asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
        }
    }
}

int main() {
    const uint64_t bytes = 1ull << 30;
    double* buf;
    cudaMalloc(&buf, bytes);
    cudaMemset(buf, 0, bytes);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    const int grid = 148 * 2, block = 256, iters = 4096;
    for (int seg_bytes : {128, 256, 512, 1024, 4096}) {
        const uint64_t nseg = bytes / seg_bytes;
        const int sd = seg_bytes / 8;
        for (int rep = 0; rep < 2; ++rep) {
            cudaEventRecord(e0);
            redg_kernel<<<grid, block>>>(buf, nseg, sd, iters / (sd / 16 > 0 ? sd / 16 : 1));
            cudaEventRecord(e1); cudaEventSynchronize(e1);
            float ms; cudaEventElapsedTime(&ms, e0, e1);
            const double n = double(grid) * (block / 32) * (iters / (sd / 16 > 0 ? sd / 16 : 1)) * sd;
            if (rep) printf("redg   seg %5d B: %8.2f G elem/s  %7.1f GB/s  (%.3f ms)\n", seg_bytes, n / ms / 1e6, n * 8 / ms / 1e6, ms);
        }
        for (int rep = 0; rep < 2; ++rep) {
            cudaEventRecord(e0);
            blkred_kernel<<<grid, block, (block / 32) * seg_bytes>>>(buf, nseg, seg_bytes, iters);
            cudaEventRecord(e1); cudaEventSynchronize(e1);
            float ms; cudaEventElapsedTime(&ms, e0, e1);
            const double n = double(grid) * (block / 32) * iters * sd;
            if (rep) printf("blkred seg %5d B: %8.2f G elem/s  %7.1f GB/s  (%.3f ms)\n", seg_bytes, n / ms / 1e6, n * 8 / ms / 1e6, ms);
        }
    }
    // is the limit per SM (issue) or chip-wide (L2 atomic units)?  Same work per CTA on 37, 74, 148, 296 CTAs.
    {
        const int seg_bytes = 1024, sd = seg_bytes / 8;
        const uint64_t nseg = bytes / seg_bytes;
        for (int g : {37, 74, 148, 296}) {
            for (int rep = 0; rep < 2; ++rep) {
                cudaEventRecord(e0);
                redg_kernel<<<g, block>>>(buf, nseg, sd, iters / (sd / 16));
                cudaEventRecord(e1); cudaEventSynchronize(e1);
                float ms; cudaEventElapsedTime(&ms, e0, e1);
                const double n = double(g) * (block / 32) * (iters / (sd / 16)) * sd;
                if (rep) printf("redg   grid %3d: %8.2f G elem/s (%.3f ms)\n", g, n / ms / 1e6, ms);
            }
            for (int rep = 0; rep < 2; ++rep) {
                cudaEventRecord(e0);
                blkred_kernel<<<g, block, (block / 32) * seg_bytes>>>(buf, nseg, seg_bytes, iters / 8);
                cudaEventRecord(e1); cudaEventSynchronize(e1);
                float ms; cudaEventElapsedTime(&ms, e0, e1);
                const double n = double(g) * (block / 32) * (iters / 8) * sd;
                if (rep) printf("blkred grid %3d: %8.2f G elem/s (%.3f ms)\n", g, n / ms / 1e6, ms);
            }
            for (int rep = 0; rep < 2; ++rep) {
                cudaEventRecord(e0);
                mixed_kernel<<<g, block, (block / 32) * seg_bytes>>>(buf, nseg, seg_bytes, iters / 8);
                cudaEventRecord(e1); cudaEventSynchronize(e1);
                float ms; cudaEventElapsedTime(&ms, e0, e1);
                const double n = double(g) * (block / 32) * (iters / 8) * sd;
                if (rep) printf("mixed  grid %3d: %8.2f G elem/s (%.3f ms)\n", g, n / ms / 1e6, ms);
            }
        }
    }
    printf("last error: %s\n", cudaGetErrorString(cudaDeviceSynchronize()));
    return 0;
}
