// Reduced density matrices <L| ... |R> on the device, every spin block up to three-body, one formulation:
//
//   Gamma[(A B), (A' B')] = sum over K = (Ka, Kb) of  D_L[(A B), K] * D_R[(A' B'), K]
//   D_X[(A B), (Ka, Kb)]  = s_A s_B X[Ka + A, Kb + B]
//
// A / B = decreasing tuples of la alpha / lb beta orbitals, Ka / Kb the (N - la) / (N - lb) electron strings
// (HoleLists, strings.hpp), s = phase of removing the tuple.  (la, lb) = (1,0) (0,1): 1-RDMs; (2,0) (1,1) (0,2):
// 2-RDMs; (3,0) (2,1) (1,2) (0,3): 3-RDMs.  This is how the reference itself builds the mixed-spin 3-RDM blocks
// (forte2/ci/ci_sigma_builder_3rdm.cc:14-76 gather, :235-258 GEMM over bounded chunks of the composite hole index) and,
// with dot products instead of a GEMM, the others (ci_sigma_builder_1rdm.cc:9-63, ci_sigma_builder_2rdm.cc:45-201,
// ci_sigma_builder_3rdm.cc:86-162).  Two kernels:
//
//   rdm_gather_kernel   D for a chunk of columns: one CTA per (Ka, tile of Kb); the parents of Ka sit in shared memory,
//                       a thread owns one Kb and walks its parents.  D is written once (after a memset: most of a
//                       column is structurally zero).  HBM-bound on the writes of D.
//   rdm_gemm_kernel     Gamma += D_L D_R^T: FP64 tensor cores (DMMA m16n8k4), 128 x 128 x 16 tiles staged with cp.async in
//                       three stages, split over the chunk's columns across CTAs, partial tiles reduced into Gamma with
//                       red.global.add.f64.  Tensor-bound for every block with 128 or more rows.
#pragma once

#include <cstdint>
#include <cuda_runtime.h>

namespace f2 {

struct RdmGatherParams {
    const double* X;
    double* D;
    const uint32_t *ptrA, *ptrB; // hole lists (CSR over Ka / Kb)
    const uint64_t *entA, *entB;
    const uint32_t *a_cls, *a_loc, *b_cls, *b_loc, *nB_cls;
    const int64_t* blk_off; // [nclsA][nclsB] offset of the determinant block in the CI vector, -1 if none
    int nclsB;
    int Pb;           // number of beta operator tuples (row = opA * Pb + opB)
    uint32_t ka0, kb0; // first Ka / Kb of the chunk
    uint32_t nka, nkb; // chunk extent; column = (Ka - ka0) * nkb + (Kb - kb0)
    int64_t ldd;       // nka * nkb rounded up to a multiple of 2 (16-byte rows for cp.async)
};

constexpr int RDM_GATHER_THREADS = 256;
constexpr int RDM_MAXPA = 512; // parents of one Ka staged per pass

__global__ void __launch_bounds__(RDM_GATHER_THREADS) rdm_gather_kernel(const RdmGatherParams p) {
    __shared__ uint32_t s_cls[RDM_MAXPA], s_loc[RDM_MAXPA], s_row[RDM_MAXPA]; // row = opA * Pb | sign << 31
    const uint32_t Ka = p.ka0 + blockIdx.x;
    const uint32_t kb = blockIdx.y * RDM_GATHER_THREADS + threadIdx.x; // within the chunk
    const uint32_t a0 = p.ptrA[Ka], a1 = p.ptrA[Ka + 1];
    const int64_t col = static_cast<int64_t>(blockIdx.x) * p.nkb + kb;
    uint32_t b0 = 0, b1 = 0;
    if (kb < p.nkb) {
        b0 = p.ptrB[p.kb0 + kb];
        b1 = p.ptrB[p.kb0 + kb + 1];
    }
    for (uint32_t base = a0; base < a1; base += RDM_MAXPA) {
        const uint32_t cnt = a1 - base < RDM_MAXPA ? a1 - base : RDM_MAXPA;
        __syncthreads();
        for (uint32_t i = threadIdx.x; i < cnt; i += RDM_GATHER_THREADS) {
            const uint64_t e = p.entA[base + i];
            const uint32_t I = static_cast<uint32_t>(e);
            s_cls[i] = p.a_cls[I];
            s_loc[i] = p.a_loc[I];
            s_row[i] = (static_cast<uint32_t>((e >> 32) & 0xffff) * static_cast<uint32_t>(p.Pb)) |
                       (static_cast<uint32_t>((e >> 48) & 1) << 31);
        }
        __syncthreads();
        for (uint32_t eb = b0; eb < b1; ++eb) {
            const uint64_t e = p.entB[eb];
            const uint32_t J = static_cast<uint32_t>(e);
            const uint32_t cB = p.b_cls[J];
            const int64_t lJ = p.b_loc[J];
            const int64_t nB = p.nB_cls[cB];
            const uint32_t opB = static_cast<uint32_t>((e >> 32) & 0xffff);
            const uint32_t negB = static_cast<uint32_t>((e >> 48) & 1);
            for (uint32_t i = 0; i < cnt; ++i) {
                const int64_t off = p.blk_off[static_cast<int64_t>(s_cls[i]) * p.nclsB + cB];
                if (off < 0)
                    continue;
                const double v = p.X[off + static_cast<int64_t>(s_loc[i]) * nB + lJ];
                const uint32_t row = (s_row[i] & 0x7fffffffu) + opB;
                p.D[static_cast<int64_t>(row) * p.ldd + col] = ((s_row[i] >> 31) ^ negB) ? -v : v;
            }
        }
    }
}

// ---------------------------------------------------------------------------------------------
// C[M x M] += A[M x K] B[M x K]^T, both row-major with the K index contiguous (leading dimension ld, even).
constexpr int RG_BM = 128, RG_BN = 128, RG_BK = 16, RG_STAGES = 3, RG_THREADS = 256;
constexpr int RG_LDS = RG_BK + 4; // 20 doubles: rows g = 0..7 of a fragment load fall on disjoint banks
constexpr size_t rdm_gemm_smem_bytes() { return size_t(RG_STAGES) * (RG_BM + RG_BN) * RG_LDS * 8; }

struct RdmGemmParams {
    const double* A;
    const double* B;
    double* C;
    int M;          // rows of A and of B, C is M x M (leading dimension M)
    int64_t K, ld;  // columns of this chunk, leading dimension of A and B
    int ksplit;     // CTAs along K per tile
};

__global__ void __launch_bounds__(RG_THREADS) rdm_gemm_kernel(const RdmGemmParams p) {
    extern __shared__ __align__(16) unsigned char rg_smem[];
    double* As = reinterpret_cast<double*>(rg_smem);                 // [STAGES][BM][LDS]
    double* Bs = As + RG_STAGES * RG_BM * RG_LDS;                    // [STAGES][BN][LDS]
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int g = lane >> 2, t = lane & 3;
    const int tiles_n = (p.M + RG_BN - 1) / RG_BN;
    const int tile = blockIdx.x / p.ksplit, ks = blockIdx.x - tile * p.ksplit;
    const int m0 = (tile / tiles_n) * RG_BM, n0 = (tile % tiles_n) * RG_BN;
    // this CTA's share of the K range, in steps of BK
    const int64_t nsteps_all = (p.K + RG_BK - 1) / RG_BK;
    const int64_t s0 = nsteps_all * ks / p.ksplit, s1 = nsteps_all * (ks + 1) / p.ksplit;
    const int nsteps = static_cast<int>(s1 - s0);
    if (nsteps <= 0)
        return;
    const int wm = (warp >> 2) * 64, wn = (warp & 3) * 32; // warp tile 64 x 32

    auto load_stage = [&](int step, int stage) {
        // 256 rows (128 of A, 128 of B) x 16 doubles = 8 pieces of 16 bytes per row, 2048 pieces, 8 per thread
        const int64_t k0 = (s0 + step) * RG_BK;
#pragma unroll
        for (int it = 0; it < 8; ++it) {
            const int idx = tid + it * RG_THREADS;
            const int row = idx >> 3, piece = idx & 7;
            const bool isB = row >= RG_BM;
            const int r = isB ? row - RG_BM : row;
            const int grow = (isB ? n0 : m0) + r;
            const int64_t k = k0 + 2 * piece;
            int bytes = 0;
            if (grow < p.M && k < p.K)
                bytes = (p.K - k >= 2) ? 16 : 8;
            const double* src = (isB ? p.B : p.A) + (bytes ? static_cast<int64_t>(grow) * p.ld + k : 0);
            double* dst = (isB ? Bs : As) + (stage * (isB ? RG_BN : RG_BM) + r) * RG_LDS + 2 * piece;
            cp_async16(dst, src, bytes);
        }
        cp_async_commit();
    };

    double acc[4][4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j)
            acc[i][j][0] = acc[i][j][1] = acc[i][j][2] = acc[i][j][3] = 0.0;

#pragma unroll
    for (int s = 0; s < RG_STAGES - 1; ++s) {
        if (s < nsteps)
            load_stage(s, s);
        else
            cp_async_commit();
    }
    for (int step = 0; step < nsteps; ++step) {
        cp_async_wait<RG_STAGES - 2>();
        __syncthreads();
        if (step + RG_STAGES - 1 < nsteps)
            load_stage(step + RG_STAGES - 1, (step + RG_STAGES - 1) % RG_STAGES);
        else
            cp_async_commit();
        const double* Ast = As + (step % RG_STAGES) * RG_BM * RG_LDS;
        const double* Bst = Bs + (step % RG_STAGES) * RG_BN * RG_LDS;
#pragma unroll
        for (int kk = 0; kk < RG_BK; kk += 4) {
            double a0[4], a1[4], b[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                a0[i] = Ast[(wm + 16 * i + g) * RG_LDS + kk + t];
                a1[i] = Ast[(wm + 16 * i + g + 8) * RG_LDS + kk + t];
            }
#pragma unroll
            for (int j = 0; j < 4; ++j)
                b[j] = Bst[(wn + 8 * j + g) * RG_LDS + kk + t];
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j)
                    dmma_16x8x4(acc[i][j], a0[i], a1[i], b[j]);
        }
    }
    cp_async_wait<0>();
    // accumulator (tile i, j): rows wm + 16 i + g (+8), columns wn + 8 j + 2 t (+1)
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            const int row = m0 + wm + 16 * i + g + 8 * h;
            if (row >= p.M)
                continue;
#pragma unroll
            for (int j = 0; j < 4; ++j)
#pragma unroll
                for (int c = 0; c < 2; ++c) {
                    const int colc = n0 + wn + 8 * j + 2 * t + c;
                    if (colc < p.M)
                        atomicAdd(p.C + static_cast<int64_t>(row) * p.M + colc, acc[i][j][2 * h + c]);
                }
        }
}

// ---------------------------------------------------------------------------------------------
// Transition 1-RDM between two CI spaces (different symmetry / GAS restrictions, same orbitals and electron numbers):
// out[pq] += sign * sum_i L[offL + i strideL] R[offR + i strideR] over a host-built list of matching segments
// (CISigmaBuilder::compute_s_1trdm, forte2/ci/ci_sigma_builder_1trdm.cc:11-110: its VO string map x the map of the
// untouched spin's strings).  One warp per entry.
struct TrdmEntry {
    int64_t offL, offR;
    int32_t strideL, strideR;
    int32_t len;
    int32_t out; // p * norb + q, bit 31 = negative
};

__global__ void __launch_bounds__(256) trdm_kernel(const double* __restrict__ L, const double* __restrict__ R,
                                                   const TrdmEntry* __restrict__ ent, int64_t nent, double* out) {
    const int64_t w = (static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (w >= nent)
        return;
    const TrdmEntry e = ent[w];
    double acc = 0.0;
    for (int i = lane; i < e.len; i += 32)
        acc += L[e.offL + static_cast<int64_t>(i) * e.strideL] * R[e.offR + static_cast<int64_t>(i) * e.strideR];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1)
        acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if (lane == 0 && acc != 0.0)
        atomicAdd(out + (e.out & 0x7fffffff), (e.out < 0) ? -acc : acc);
}

} // namespace f2
