// Same-spin term, reduction form with the parent rows staged in shared memory (sm_100a).
//
// ssb_kernel does the work of ssg_kernel (kernels.cuh: per (N-2)-electron string K the m x m block of antisymmetrised
// integrals acts on the m parent rows of K, the result is reduced into Y with red.global.add.f64) with a different
// operand path.  ssg_kernel loads the B fragments (parent rows of X) from global memory into registers two k-steps ahead:
// the list look-up, the address arithmetic and the L2 latency of every k-step sit in the multiply loop (ncu: tensor pipe
// 65 % active, short- and long-scoreboard stalls 2.4 per issue).  Here every warp owns a ring of SSB_NSTG stages of
// 8 rows x 16 columns and keeps it SSB_NSTG - 1 stages ahead with cp.async, across chunk and string boundaries (the parent
// lists of all SSG_KG strings of the CTA are decoded before the first multiplication), so the multiply loop holds
// nothing but shared-memory fragment loads and DMMAs.  The phases of the parents go into two bit masks per string
// instead of per-k-step list reads.
// Reference counterpart: CISigmaBuilder::H2_hz_same_spin and H1_hz (ci_sigma_builder_harrison_zarrabian.cc:14-195).
#pragma once
#include "kernels.cuh"

namespace f2 {

#ifndef F2_SSB_NSTG
#define F2_SSB_NSTG 4
#endif
constexpr int SSB_NSTG = F2_SSB_NSTG; // stages of the per-warp ring (2 k-steps = 8 parent rows each)
constexpr int SSB_BSTR = 20;          // row stride of a stage in doubles (== 4 mod 16: conflict-free B fragments)
constexpr int SSB_STAGE = 8 * SSB_BSTR;

__host__ __device__ constexpr int ssb_astride(int rows) { // smallest stride >= rows that is 4 or 12 (mod 16)
    int s = rows;
    while (s % 16 != 4 && s % 16 != 12)
        ++s;
    return s;
}
__host__ __device__ constexpr size_t ssb_smem_bytes(int rows, int nwarps) {
    return 2 * size_t(rows) * ssb_astride(rows) * 8 + size_t(SSG_KG) * rows * (8 + 8 + 4) +
           size_t(nwarps) * SSB_NSTG * SSB_STAGE * 8 + 16;
}
// two CTAs of 8 warps per SM, which fit up to 56 parent rows per string.  One CTA of 16 warps for 57..80 rows (the
// shape of the 3-GAS / 20-orbital spaces) was measured and is slower than ssg_kernel there (13.5 against 10.6 ms per
// spin on the 22 M determinant twin, profiles/r3a_quick_ssb_16_warps.log): those shapes stay with ssg_kernel.
__host__ __device__ constexpr int ssb_warps(int) { return 8; }
__host__ __device__ constexpr bool ssb_fits(int rows) {
    return 2 * (ssb_smem_bytes(rows, ssb_warps(rows)) + 1024 + 64) <= size_t(227) * 1024;
}

template <int N> __device__ __forceinline__ void cp_async_wait_n() {
    asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}

// position of column c (0..15) of a chunk inside a stage row: the two middle column pairs of each group of eight are
// swapped, so that B-fragment lane g, which needs column (g >> 1) + 4 (g & 1) (the accumulator layout that gives every
// reduction instruction full 32-byte sectors, as in ssg_kernel), reads position ((g & 4) | (g & 1) << 1 | (g >> 1 & 1))
// and a half-warp touches sixteen different 8-byte banks
__device__ __forceinline__ int ssb_pos(int c) {
    const int pr = (c >> 1) & 3;
    return (c & 8) | (((pr & 1) << 1 | (pr >> 1)) << 1) | (c & 1);
}

template <int NM16, bool HAS8, bool AL16, bool GK>
__global__ void __launch_bounds__(32 * ssb_warps(NM16 * 16 + (HAS8 ? 8 : 0)), ssb_warps(NM16 * 16 + (HAS8 ? 8 : 0)) == 8 ? 2 : 1)
    ssb_kernel(const SSParams p, const SSGroupLists gk) {
    constexpr int ROWS = NM16 * 16 + (HAS8 ? 8 : 0);
    constexpr int S = ssb_astride(ROWS);
    constexpr int NW = ssb_warps(ROWS);
    extern __shared__ __align__(16) unsigned char smem_raw[];
    double* Abuf = reinterpret_cast<double*>(smem_raw);                        // [2][ROWS][S]
    double* Bring = Abuf + 2 * ROWS * S;                                       // [NW warps][SSB_NSTG][8][SSB_BSTR]
    long long* msrcK = reinterpret_cast<long long*>(Bring + NW * SSB_NSTG * SSB_STAGE); // [SSG_KG][ROWS]
    long long* mdstK = msrcK + SSG_KG * ROWS;
    unsigned* minfoK = reinterpret_cast<unsigned*>(mdstK + SSG_KG * ROWS);
    __shared__ int s_m[SSG_KG];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int g = lane >> 2, t = lane & 3;
    const int grp = p.tiles[2 * blockIdx.y];
    const int j0 = p.tiles[2 * blockIdx.y + 1];
    const int ncols = p.g_ncols[grp];
    const long long ldx = p.g_ldx[grp], ldy = p.g_ldy[grp];
    const int k_first = blockIdx.x * SSG_KG;
    const uint32_t* gklist = nullptr;
    int gnk = 0;
    if constexpr (GK) {
        gklist = gk.list + gk.ptr[grp];
        gnk = static_cast<int>(gk.ptr[grp + 1] - gk.ptr[grp]);
        if (k_first >= gnk) // the grid is sized for the longest group list
            return;
    }
    const int k_cnt = GK ? (gnk - k_first < SSG_KG ? gnk - k_first : SSG_KG)
                         : (p.nk - k_first < SSG_KG ? p.nk - k_first : SSG_KG);

    // parent lists of all strings of this CTA, one string per warp
    for (int j = warp; j < SSG_KG; j += NW) {
        long long* msrc = msrcK + j * ROWS;
        long long* mdst = mdstK + j * ROWS;
        unsigned* minfo = minfoK + j * ROWS;
        int cnt = 0;
        if (j < k_cnt) {
            const uint32_t K = GK ? gklist[k_first + j] : p.klist[k_first + j];
            const int e0 = p.h2_ptr[K], e1 = p.h2_ptr[K + 1];
            for (int base = e0; base < e1; base += 32) {
                const int e = base + lane;
                bool ok = false;
                long long src = -1, dst = -1;
                unsigned info = 0;
                if (e < e1) {
                    const uint64_t ent = p.h2_ent[e];
                    const uint32_t I = static_cast<uint32_t>(ent);
                    const int c = p.s_cls[I];
                    const long long xo = p.xoff[c * p.ngroups + grp];
                    if (xo >= 0) {
                        ok = true;
                        src = xo + static_cast<long long>(p.s_loc[I]) * ldx;
                        dst = p.yoff[c * p.ngroups + grp] + static_cast<long long>(p.s_loc[I]) * ldy;
                        info = static_cast<unsigned>((ent >> 32) & 0xffff) | (static_cast<unsigned>((ent >> 48) & 1) << 31);
                    }
                }
                const unsigned mask = __ballot_sync(0xffffffffu, ok);
                if (ok) {
                    const int pos = cnt + __popc(mask & ((1u << lane) - 1u));
                    if (pos < ROWS) {
                        msrc[pos] = src;
                        mdst[pos] = dst;
                        minfo[pos] = info;
                    }
                }
                cnt += __popc(mask);
            }
            cnt = cnt < ROWS ? cnt : ROWS;
        }
        for (int r = cnt + lane; r < ROWS; r += 32) {
            msrc[r] = 0;
            mdst[r] = -1;
            minfo[r] = 0;
        }
        if (lane == 0)
            s_m[j] = cnt;
    }
    __syncthreads();

    auto copy_block = [&](int j) { // all threads: W[pair_a, pair_b] of string j into buffer j % 2 (async)
        const unsigned* minfo = minfoK + j * ROWS;
        const int m = s_m[j];
        double* A = Abuf + (j & 1) * ROWS * S;
        const int rows_used = ((m + 15) >> 4) << 4; // only the 16-row tiles that hold parents are multiplied
        for (int idx = tid; idx < (rows_used < ROWS ? rows_used : ROWS) * S; idx += 32 * NW) {
            const int a = idx / S, b = idx - a * S;
            const bool in = a < m && b < m;
            const double* src = in ? p.W + static_cast<size_t>(minfo[a] & 0xffff) * p.P + (minfo[b] & 0xffff) : p.W;
            cp_async8(A + idx, src, in ? 8 : 0);
        }
        cp_async_commit();
    };

    const int jend = j0 + SSG_TW < ncols ? j0 + SSG_TW : ncols;
    const int nchunks = (jend - j0 + 15) >> 4;
    double* Bw = Bring + warp * SSB_NSTG * SSB_STAGE;

    // producer cursor: the stage (string pj, chunk pc, rows 8 ps ..) that is loaded next, into ring slot pslot
    int pj = 0, pc = warp, ps = 0, pslot = 0;
    auto skip_empty = [&]() { // strings without parents in this column group, warps without a chunk
        if (pc >= nchunks)
            pj = k_cnt;
        while (pj < k_cnt && s_m[pj] == 0)
            ++pj;
    };
    skip_empty();
    auto produce = [&]() {
        if (pj < k_cnt) {
            const int m = s_m[pj];
            const long long* msrc = msrcK + pj * ROWS + 8 * ps;
            const int n0 = j0 + (pc << 4);
            double* dst = Bw + pslot * SSB_STAGE;
            if constexpr (AL16) {
#pragma unroll
                for (int i = 0; i < 2; ++i) {
                    const int q = lane + 32 * i, r = q >> 3, c = (q & 7) << 1;
                    int bytes = (jend - (n0 + c)) * 8;
                    bytes = 8 * ps + r < m ? (bytes > 16 ? 16 : (bytes < 0 ? 0 : bytes)) : 0;
                    cp_async16(dst + r * SSB_BSTR + ssb_pos(c), p.X + msrc[r] + (bytes ? n0 + c : 0), bytes);
                }
            } else {
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const int q = lane + 32 * i, r = q >> 4, c = q & 15;
                    const bool in = 8 * ps + r < m && n0 + c < jend;
                    cp_async8(dst + r * SSB_BSTR + ssb_pos(c), p.X + msrc[r] + (in ? n0 + c : 0), in ? 8 : 0);
                }
            }
            // advance: next stage of the chunk, else the next chunk of the string, else the next string
            if (++ps >= (m + 7) >> 3) {
                ps = 0;
                pc += NW;
                if (pc >= nchunks) {
                    pc = warp;
                    ++pj;
                    skip_empty();
                }
            }
        }
        cp_async_commit(); // one group per call, empty or not: the consumer counts groups
        pslot = pslot + 1 == SSB_NSTG ? 0 : pslot + 1;
    };

    copy_block(0);
    cp_async_wait<0>();
    __syncthreads();
    if (k_cnt > 1)
        copy_block(1);
#pragma unroll
    for (int i = 0; i < SSB_NSTG - 1; ++i)
        produce();

    const int pg = (g & 4) | ((g & 1) << 1) | ((g >> 1) & 1); // position of this lane's B column in a stage row
    int cslot = 0;
    for (int j = 0; j < k_cnt; ++j) {
        if (j > 0 && j + 1 < k_cnt)
            copy_block(j + 1);
        const int m = s_m[j];
        const long long* mdst = mdstK + j * ROWS;
        const unsigned* minfo = minfoK + j * ROWS;
        const int nks = (m + 3) >> 2;
        const int nst = (m + 7) >> 3;
        // phases: bit ks of bsg = parent row 4 ks + t (B operand), bit 2 mt + h of esg = row 16 mt + g + 8 h (result)
        unsigned bsg = 0, esg = 0;
        if (m > 0) {
            for (int ks = 0; ks < nks; ++ks)
                bsg |= (minfo[4 * ks + t] >> 31) << ks;
#pragma unroll
            for (int i = 0; i < 2 * NM16 + (HAS8 ? 1 : 0); ++i)
                esg |= (minfo[8 * i + g] >> 31) << i;
        }
        const double* ap = Abuf + (j & 1) * ROWS * S + g * S + t;
        for (int chunk = warp; chunk < nchunks && m > 0; chunk += NW) {
            const int n0 = j0 + (chunk << 4);
            double acc[NM16 > 0 ? NM16 : 1][2][4];
            double acc8[2][2];
#pragma unroll
            for (int i = 0; i < NM16; ++i)
#pragma unroll
                for (int nt = 0; nt < 2; ++nt)
                    acc[i][nt][0] = acc[i][nt][1] = acc[i][nt][2] = acc[i][nt][3] = 0.0;
            acc8[0][0] = acc8[0][1] = acc8[1][0] = acc8[1][1] = 0.0;
            for (int s = 0; s < nst; ++s) {
                cp_async_wait_n<SSB_NSTG - 2>(); // this thread's pieces of the stage have landed
                __syncwarp();                    // ... and everybody else's; the slot read last is free again
                produce();
                const double* bs = Bw + cslot * SSB_STAGE + t * SSB_BSTR + pg;
                cslot = cslot + 1 == SSB_NSTG ? 0 : cslot + 1;
#pragma unroll
                for (int l = 0; l < 2; ++l) {
                    const int ks = 2 * s + l;
                    if (ks < nks) {
                        const unsigned sg = (bsg >> ks) << 31;
                        const double b0 = flip_sign(bs[4 * l * SSB_BSTR], sg), b1 = flip_sign(bs[4 * l * SSB_BSTR + 8], sg);
                        const double* a = ap + 4 * ks;
#pragma unroll
                        for (int mt = 0; mt < NM16; ++mt) {
                            if (16 * mt < m) {
                                const double a0 = a[(16 * mt) * S], a1 = a[(16 * mt + 8) * S];
                                dmma_16x8x4(acc[mt][0], a0, a1, b0);
                                dmma_16x8x4(acc[mt][1], a0, a1, b1);
                            }
                        }
                        if constexpr (HAS8) {
                            if (16 * NM16 < m) {
                                const double a0 = a[(16 * NM16) * S];
                                dmma_8x8x4(acc8[0], a0, b0);
                                dmma_8x8x4(acc8[1], a0, b1);
                            }
                        }
                    }
                }
            }
            // accumulator (row, tile nt, n index 2t + c) is column n0 + 8 nt + 4 c + t
            const int c0 = n0 + t;
#pragma unroll
            for (int mt = 0; mt < NM16; ++mt)
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    const int a = 16 * mt + g + 8 * h;
                    if (a < m) {
                        double* dst = p.Y + mdst[a] + c0;
                        const unsigned sg = (esg >> (2 * mt + h)) << 31;
#pragma unroll
                        for (int nt = 0; nt < 2; ++nt)
#pragma unroll
                            for (int c = 0; c < 2; ++c)
                                if (c0 + 8 * nt + 4 * c < jend)
                                    atomicAdd(dst + 8 * nt + 4 * c, flip_sign(acc[mt][nt][2 * h + c], sg));
                    }
                }
            if constexpr (HAS8) {
                const int a = 16 * NM16 + g;
                if (a < m) {
                    double* dst = p.Y + mdst[a] + c0;
                    const unsigned sg = (esg >> (2 * NM16)) << 31;
#pragma unroll
                    for (int nt = 0; nt < 2; ++nt)
#pragma unroll
                        for (int c = 0; c < 2; ++c)
                            if (c0 + 8 * nt + 4 * c < jend)
                                atomicAdd(dst + 8 * nt + 4 * c, flip_sign(acc8[nt][c], sg));
                }
            }
        }
        // the block of string j + 1 is older than the SSB_NSTG - 1 stages in flight, so the ring is not drained -
        // unless this warp committed fewer groups than that since the block copy (no parents, no chunk, short lists)
        if (m > 0 && warp < nchunks && nst * ((nchunks - warp + NW - 1) / NW) >= SSB_NSTG - 1)
            cp_async_wait_n<SSB_NSTG - 1>();
        else
            cp_async_wait<0>();
        __syncthreads();
    }
    cp_async_wait<0>();
}

} // namespace f2
