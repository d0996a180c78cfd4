// Reduced density matrices <L| ... |R> from two CI vectors on the device: first correct path (scalar FP64, one
// CTA per string, shared-memory accumulation, REDs into the small result).  Same lists as the sigma kernels:
// alpha single replacements (CSR by output string), beta single replacements (ELL by source string), two-hole
// members (CSR by (N-2)-electron string).  Conventions of the reference (forte2/ci/ci_sigma_builder_1rdm.cc:9-63,
// ci_sigma_builder_2rdm.cc:45-201):
//   a_1rdm[p, q]        = <L| a+_p a_q |R>                         (b_1rdm: beta)
//   aa_2rdm[pq, rs]     = <L| a+_p a+_q a_s a_r |R>,  p > q, r > s, pair index p (p - 1) / 2 + q
//   ab_2rdm[u, v, x, y] = <L| a+_u(alpha) a+_v(beta) a_y(beta) a_x(alpha) |R> = <L| E^a_ux E^b_vy |R>
// The tensor-core formulations that replace these kernels are laid out in DESIGN.md section 6a.
#pragma once

#include <cstdint>
#include <cuda_runtime.h>

namespace f2 {

struct RdmParams {
    const double* L;
    const double* R;
    double* out;
    const uint32_t *a_cls, *a_loc, *b_cls, *b_loc, *nA_cls, *nB_cls;
    const int64_t* blk_off; // [nclsA][nclsB] offset of the determinant block in the CI vector, -1 if none
    const uint32_t* as_ptr;
    const uint64_t* as_ent; // <Ia| E_kl |Ja> = sign: Ja | (k norb + l) << 32 | (sign < 0) << 48, CSR by Ia
    const uint64_t* bs_ent; // <Ib| E_ij |Jb> = sign: address(Ib) | class(Ib) << 28 | (i norb + j) << 40 | neg << 56,
                            // entry t of source Jb at [t nbs + Jb], all ones = padding
    const uint32_t* h2_ptr;
    const uint64_t* h2_ent; // a_p a_q |I> = sign |K>: I | pair(p > q) << 32 | neg << 48, CSR by K
    int norb, nclsA, nclsB, bs_width;
    int64_t nas, nbs;
};

__device__ __forceinline__ double rdm_block_sum(double v, double* red) {
    // sum over the CTA (blockDim.x a multiple of 32, at most 1024); the result is valid in thread 0
#pragma unroll
    for (int o = 16; o > 0; o >>= 1)
        v += __shfl_down_sync(0xffffffffu, v, o);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    __syncthreads(); // red may still be read from the previous call
    if (lane == 0)
        red[warp] = v;
    __syncthreads();
    double t = 0.0;
    if (warp == 0) {
        t = lane < (blockDim.x >> 5) ? red[lane] : 0.0;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1)
            t += __shfl_down_sync(0xffffffffu, t, o);
    }
    return t;
}

// alpha 1-RDM: one CTA per output alpha string Ia; g[k, l] += sign sum_Ib L[Ia, Ib] R[Ja, Ib]
__global__ void __launch_bounds__(256) rdm1a_kernel(const RdmParams p) {
    extern __shared__ double rdm_smem[];
    double* g = rdm_smem; // [norb^2]
    __shared__ double red[32];
    const int n2 = p.norb * p.norb;
    for (int i = threadIdx.x; i < n2; i += blockDim.x)
        g[i] = 0.0;
    __syncthreads();
    const uint32_t Ia = blockIdx.x;
    const int cI = static_cast<int>(p.a_cls[Ia]);
    const int64_t lI = p.a_loc[Ia];
    for (uint32_t e = p.as_ptr[Ia]; e < p.as_ptr[Ia + 1]; ++e) {
        const uint64_t ent = p.as_ent[e];
        const uint32_t Ja = static_cast<uint32_t>(ent);
        const int kl = static_cast<int>((ent >> 32) & 0xffff);
        const int cJ = static_cast<int>(p.a_cls[Ja]);
        const int64_t lJ = p.a_loc[Ja];
        double acc = 0.0;
        for (int c = 0; c < p.nclsB; ++c) {
            const int64_t oI = p.blk_off[cI * p.nclsB + c], oJ = p.blk_off[cJ * p.nclsB + c];
            const int64_t nB = p.nB_cls[c];
            if (oI < 0 || oJ < 0)
                continue;
            const double* l = p.L + oI + lI * nB;
            const double* r = p.R + oJ + lJ * nB;
            for (int64_t ib = threadIdx.x; ib < nB; ib += blockDim.x)
                acc += l[ib] * r[ib];
        }
        const double tot = rdm_block_sum(acc, red);
        if (threadIdx.x == 0)
            g[kl] += ((ent >> 48) & 1) ? -tot : tot;
    }
    __syncthreads();
    for (int i = threadIdx.x; i < n2; i += blockDim.x)
        if (g[i] != 0.0)
            atomicAdd(p.out + i, g[i]);
}

// beta 1-RDM: one CTA per alpha string Ia (a row of both vectors); g[i, j] += sign L[Ia, Ib] R[Ia, Jb]
__global__ void __launch_bounds__(256) rdm1b_kernel(const RdmParams p) {
    extern __shared__ double rdm_smem[];
    double* g = rdm_smem; // [norb^2]
    const int n2 = p.norb * p.norb;
    for (int i = threadIdx.x; i < n2; i += blockDim.x)
        g[i] = 0.0;
    __syncthreads();
    const uint32_t Ia = blockIdx.x;
    const int cI = static_cast<int>(p.a_cls[Ia]);
    const int64_t lI = p.a_loc[Ia];
    for (int64_t Jb = threadIdx.x; Jb < p.nbs; Jb += blockDim.x) {
        const int cJ = static_cast<int>(p.b_cls[Jb]);
        const int64_t oR = p.blk_off[cI * p.nclsB + cJ];
        if (oR < 0)
            continue;
        const double r = p.R[oR + lI * static_cast<int64_t>(p.nB_cls[cJ]) + p.b_loc[Jb]];
        for (int t = 0; t < p.bs_width; ++t) {
            const uint64_t ent = p.bs_ent[static_cast<int64_t>(t) * p.nbs + Jb];
            if (ent == ~0ull)
                break;
            const int ci = static_cast<int>((ent >> 28) & 0xfff);
            const int64_t oL = p.blk_off[cI * p.nclsB + ci];
            if (oL < 0)
                continue;
            const double l = p.L[oL + lI * static_cast<int64_t>(p.nB_cls[ci]) + static_cast<int64_t>(ent & 0xfffffff)];
            const double v = l * r;
            atomicAdd(g + static_cast<int>((ent >> 40) & 0xffff), ((ent >> 56) & 1) ? -v : v);
        }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < n2; i += blockDim.x)
        if (g[i] != 0.0)
            atomicAdd(p.out + i, g[i]);
}

// opposite-spin 2-RDM: one CTA per alpha string Ia; for every <Ia| E^a_ux |Ja> = s_a the tile
//   T[v, y] = s_a sum_{Jb, <Ib| E^b_vy |Jb> = s_b} s_b L[Ia, Ib] R[Ja, Jb]   is added to out[u, :, x, :]
__global__ void __launch_bounds__(256) rdm2ab_kernel(const RdmParams p) {
    extern __shared__ double rdm_smem[];
    double* tile = rdm_smem; // [norb^2]
    const int n = p.norb, n2 = n * n;
    const uint32_t Ia = blockIdx.x;
    const int cI = static_cast<int>(p.a_cls[Ia]);
    const int64_t lI = p.a_loc[Ia];
    for (uint32_t e = p.as_ptr[Ia]; e < p.as_ptr[Ia + 1]; ++e) {
        const uint64_t aent = p.as_ent[e];
        const uint32_t Ja = static_cast<uint32_t>(aent);
        const int ux = static_cast<int>((aent >> 32) & 0xffff);
        const int u = ux / n, x = ux - u * n;
        const bool aneg = (aent >> 48) & 1;
        const int cJ = static_cast<int>(p.a_cls[Ja]);
        const int64_t lJ = p.a_loc[Ja];
        for (int i = threadIdx.x; i < n2; i += blockDim.x)
            tile[i] = 0.0;
        __syncthreads();
        for (int64_t Jb = threadIdx.x; Jb < p.nbs; Jb += blockDim.x) {
            const int cb = static_cast<int>(p.b_cls[Jb]);
            const int64_t oR = p.blk_off[cJ * p.nclsB + cb];
            if (oR < 0)
                continue;
            const double r = p.R[oR + lJ * static_cast<int64_t>(p.nB_cls[cb]) + p.b_loc[Jb]];
            for (int t = 0; t < p.bs_width; ++t) {
                const uint64_t ent = p.bs_ent[static_cast<int64_t>(t) * p.nbs + Jb];
                if (ent == ~0ull)
                    break;
                const int ci = static_cast<int>((ent >> 28) & 0xfff);
                const int64_t oL = p.blk_off[cI * p.nclsB + ci];
                if (oL < 0)
                    continue;
                const double l =
                    p.L[oL + lI * static_cast<int64_t>(p.nB_cls[ci]) + static_cast<int64_t>(ent & 0xfffffff)];
                const double v = l * r;
                const bool neg = (((ent >> 56) & 1) != 0) != aneg;
                atomicAdd(tile + static_cast<int>((ent >> 40) & 0xffff), neg ? -v : v);
            }
        }
        __syncthreads();
        for (int i = threadIdx.x; i < n2; i += blockDim.x) {
            if (tile[i] != 0.0) {
                const int v = i / n, y = i - v * n;
                atomicAdd(p.out + ((static_cast<int64_t>(u) * n + v) * n + x) * n + y, tile[i]);
            }
        }
        __syncthreads();
    }
}

// same-spin 2-RDM: one CTA per (N-2)-electron string K; for every pair of parents (I_a, I_b) of K
//   out[pair_a, pair_b] += s_a s_b sum over the other spin's strings of L[I_a, .] R[I_b, .]
// SPIN 0: parents are alpha strings (rows), the sum runs over the columns of the blocks both rows have;
// SPIN 1: parents are beta strings (columns), the sum runs over the alpha rows.
template <int SPIN>
__global__ void __launch_bounds__(256) rdm2ss_kernel(const RdmParams p, const int npairs) {
    __shared__ double red[32];
    const uint32_t K = blockIdx.x;
    const uint32_t e0 = p.h2_ptr[K], e1 = p.h2_ptr[K + 1];
    const int m = static_cast<int>(e1 - e0);
    for (int ab = 0; ab < m * m; ++ab) {
        const int a = ab / m, b = ab - a * m;
        const uint64_t ea = p.h2_ent[e0 + a], eb = p.h2_ent[e0 + b];
        const uint32_t Sa = static_cast<uint32_t>(ea), Sb = static_cast<uint32_t>(eb);
        double acc = 0.0;
        if (SPIN == 0) {
            const int ca = static_cast<int>(p.a_cls[Sa]), cb = static_cast<int>(p.a_cls[Sb]);
            const int64_t la = p.a_loc[Sa], lb = p.a_loc[Sb];
            for (int c = 0; c < p.nclsB; ++c) {
                const int64_t oa = p.blk_off[ca * p.nclsB + c], ob = p.blk_off[cb * p.nclsB + c];
                const int64_t nB = p.nB_cls[c];
                if (oa < 0 || ob < 0)
                    continue;
                const double* l = p.L + oa + la * nB;
                const double* r = p.R + ob + lb * nB;
                for (int64_t i = threadIdx.x; i < nB; i += blockDim.x)
                    acc += l[i] * r[i];
            }
        } else {
            const int ca = static_cast<int>(p.b_cls[Sa]), cb = static_cast<int>(p.b_cls[Sb]);
            const int64_t la = p.b_loc[Sa], lb = p.b_loc[Sb];
            const int64_t na = p.nB_cls[ca], nb = p.nB_cls[cb];
            for (int c = 0; c < p.nclsA; ++c) {
                const int64_t oa = p.blk_off[c * p.nclsB + ca], ob = p.blk_off[c * p.nclsB + cb];
                const int64_t nA = p.nA_cls[c];
                if (oa < 0 || ob < 0)
                    continue;
                for (int64_t i = threadIdx.x; i < nA; i += blockDim.x)
                    acc += p.L[oa + i * na + la] * p.R[ob + i * nb + lb];
            }
        }
        const double tot = rdm_block_sum(acc, red);
        if (threadIdx.x == 0 && tot != 0.0) {
            const int pa = static_cast<int>((ea >> 32) & 0xffff), pb = static_cast<int>((eb >> 32) & 0xffff);
            const bool neg = (((ea >> 48) & 1) != 0) != (((eb >> 48) & 1) != 0);
            atomicAdd(p.out + static_cast<int64_t>(pa) * npairs + pb, neg ? -tot : tot);
        }
    }
}

} // namespace f2
