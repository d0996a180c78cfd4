// Device kernels of the sigma build (sm_100a).  See DESIGN.md for the algebra and the data layout.
//
//   t3_kernel   opposite-spin term: per output alpha row Ia, G[ij, Jb] = sum_r A_Ia[ij, r] C[Ja_r, Jb]
//               on the FP64 tensor pipe (DMMA), A = signed columns of the integral matrix held in
//               registers, B = rows of C staged in shared memory with cp.async; then the beta
//               replacement list scatters G into sigma[Ia, :].
//               Reference counterpart: CISigmaBuilder::H2_hz_opposite_spin
//               (forte2/ci/ci_sigma_builder_harrison_zarrabian.cc:197-308) and the alpha gather /
//               beta scatter halves of H2_kh (forte2/ci/ci_sigma_builder_knowles_handy.cc:88-171).
//   ss_kernel   same-spin term (one- and two-electron): per (N-2)-electron string K a dense
//               m x m block of antisymmetrised integrals acts on the m parent strings of K.
//               Reference counterpart: CISigmaBuilder::H2_hz_same_spin and H1_hz
//               (ci_sigma_builder_harrison_zarrabian.cc:14-195).
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

namespace f2 {

constexpr int NT = 64; // columns per tile
constexpr int BS = 72; // shared-memory row stride in doubles (== 8 mod 16: conflict-free fragments)

__device__ __forceinline__ void dmma_16x8x4(double (&c)[4], double a0, double a1, double b0) {
    asm volatile("mma.sync.aligned.m16n8k4.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5}, {%6}, "
                 "{%0,%1,%2,%3};"
                 : "+d"(c[0]), "+d"(c[1]), "+d"(c[2]), "+d"(c[3])
                 : "d"(a0), "d"(a1), "d"(b0));
}
__device__ __forceinline__ void cp_async8(void* smem, const void* gmem, int src_bytes) {
    const unsigned s = static_cast<unsigned>(__cvta_generic_to_shared(smem));
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;" ::"r"(s), "l"(gmem), "r"(src_bytes));
}
__device__ __forceinline__ void cp_async16(void* smem, const void* gmem, int src_bytes) {
    const unsigned s = static_cast<unsigned>(__cvta_generic_to_shared(smem));
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(s), "l"(gmem), "r"(src_bytes));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;"); }
template <int N> __device__ __forceinline__ void cp_async_wait() {
    asm volatile("cp.async.wait_group %0;" ::"n"(N));
}

// red.global.add.f64 executed only where the predicate holds (no branch, no return value)
__device__ __forceinline__ void red_add_if(double* addr, double v, bool ok) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.s32 p, %2, 0;\n\t@p red.global.add.f64 [%0], %1;\n\t}"
                 ::"l"(addr), "d"(v), "r"(static_cast<int>(ok))
                 : "memory");
}
// same, valid where the packed replacement target x is not the all-ones filler
__device__ __forceinline__ void red_add_target(double* addr, double v, uint32_t x) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.u32 p, %2, 0xffffffff;\n\t@p red.global.add.f64 [%0], %1;\n\t}"
                 ::"l"(addr), "d"(v), "r"(x)
                 : "memory");
}

struct T3Params {
    const double* C;
    double* sigma;
    const uint32_t* items; // 2 per CTA: alpha string (global index), beta class | m-split << 16
    const uint32_t *a_cls, *a_loc, *as_ptr;
    const uint64_t* as_ent;
    const uint32_t* tgt; // [M][nbs] beta replacement table: row m, source column Jb -> packed target
    int64_t nbs;
    const int64_t* blk_off;  // [nclsA*nclsB] offset of block (cA,cB) in C, -1 if not allowed
    const int64_t* sblk_off; // same for sigma (differs when sigma holds the owned rows only)
    const uint32_t* srow0;   // [nclsA] first owned row of each alpha class
    const uint32_t *nB_cls, *b_off;
    int nclsA, nclsB, norb;
    const double* Vx; // [M][norb*norb]  Vx[m(i,j)][k*norb+l] = <k i | l j>
    int M;
    int tgt_pair; // 1: table rows and class offsets are even, two targets per 8-byte load
    // alpha same-spin folding (complete alpha string space, one GPU): the accumulators hold
    // Z = G + 2 kk C[Ia] and every row m = (ij) is also added, halved, to sigma[E_ij Ia, :]
    int fold;
    const double* kk2;      // [M] 2 kk_ij / na, added to the diagonal (k = l) columns of A
    const uint16_t* rowmap; // [norb*norb] (i,j) -> m
};

// packed beta replacement target: bit 31 sign, bits 20..30 class of Ib, bits 0..19 address of Ib in
// its class; all ones = no replacement for this (row, column)
constexpr uint32_t kNoTarget = 0xffffffffu;
constexpr int TB = 68; // row stride of the staged C tile (== 4 mod 16: conflict-free B fragments)

__device__ __forceinline__ void dmma_8x8x4(double (&c)[2], double a0, double b0) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                 : "+d"(c[0]), "+d"(c[1])
                 : "d"(a0), "d"(b0));
}

constexpr size_t t3_smem_bytes(int ks, int ns, int nclsB, int rows) {
    return size_t(ns) * ks * 4 * TB * 8 + size_t(ks) * 4 * (8 + 8 + 4) + size_t(nclsB) * 8 + size_t(rows) * 2 + 64;
}

// Opposite-spin term.  One CTA per (output alpha string Ia, beta class of the intermediate column);
// warp w owns rows [16w, 16w+16) of the integral matrix: its A fragments (signed columns of Vx picked
// by the single replacements of Ia) stay in registers for the whole row, the B operand is the
// gathered rows C[Ja_r, tile] staged by cp.async, and every accumulator element is added straight
// from registers into sigma through the replacement table (red.global.add.f64).
// TAIL: the last (at most 8) rows of the integral matrix do not get a warp of their own -- that would
// leave one SM sub-partition with an extra warp of tensor work -- but are spread over the warps, each
// taking an 8-row x (64/nwarps)-column strip (m8n8k4).
// ONEB: a single beta class (no symmetry / GAS split of the beta strings) and paired table loads: the
// scatter target is sigma[Ia, :] itself, so an emit is mask + address + sign flip + predicated RED.
template <int KS, bool AL16, int NS, bool TAIL, bool ONEB>
__global__ void __launch_bounds__((TAIL ? 256 : (KS <= 24 ? 384 : 320)), 1) t3_kernel(const T3Params p) {
    constexpr int KP = KS * 4;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int tid = threadIdx.x, nthr = blockDim.x, warp = tid >> 5, lane = tid & 31;
    const int nwarps = nthr >> 5, g = lane >> 2, t = lane & 3;
    const int rows_cta = nwarps * 16 + (TAIL ? 8 : 0);
    double* Bs = reinterpret_cast<double*>(smem_raw);
    long long* ksrc = reinterpret_cast<long long*>(Bs + NS * KP * TB);
    long long* krow = ksrc + KP; // sigma row of Ja_r in the block paired with this column class
    long long* rowbase = krow + KP;
    unsigned* kinfo = reinterpret_cast<unsigned*>(rowbase + p.nclsB);
    short* arow = reinterpret_cast<short*>(kinfo + KP); // local row m -> replacement r with E_ij Ia = Ja_r
    __shared__ int s_Kn;

    const uint32_t Ia = p.items[2 * blockIdx.x];
    const uint32_t packed = p.items[2 * blockIdx.x + 1];
    const int cBp = packed & 0xffff, msp = packed >> 16;
    const int cA = p.a_cls[Ia];
    const uint32_t ia = p.a_loc[Ia];
    const int nBp = p.nB_cls[cBp];
    const int mbase = msp * nwarps * 16;

    if (warp == 0) {
        const int e0 = p.as_ptr[Ia], e1 = p.as_ptr[Ia + 1];
        int cnt = 0;
        for (int base = e0; base < e1; base += 32) {
            const int e = base + lane;
            bool ok = false;
            long long src = -1, dst = -1;
            unsigned info = 0;
            if (e < e1) {
                const uint64_t ent = p.as_ent[e];
                const uint32_t Ja = static_cast<uint32_t>(ent);
                const int cJ = p.a_cls[Ja];
                const long long off = p.blk_off[cJ * p.nclsB + cBp];
                if (off >= 0) {
                    ok = true;
                    src = off + static_cast<long long>(p.a_loc[Ja]) * nBp;
                    const long long so = p.sblk_off[cJ * p.nclsB + cBp];
                    dst = so >= 0 ? so + (static_cast<long long>(p.a_loc[Ja]) - p.srow0[cJ]) * nBp : -1;
                    info = static_cast<unsigned>((ent >> 32) & 0xffff) |
                           (static_cast<unsigned>((ent >> 48) & 1) << 31);
                }
            }
            const unsigned mask = __ballot_sync(0xffffffffu, ok);
            if (ok) {
                const int pos = cnt + __popc(mask & ((1u << lane) - 1u));
                ksrc[pos] = src;
                krow[pos] = dst;
                kinfo[pos] = info;
            }
            cnt += __popc(mask);
        }
        for (int r = cnt + lane; r < KP; r += 32) {
            ksrc[r] = -1;
            krow[r] = -1;
            kinfo[r] = 0;
        }
        if (lane == 0)
            s_Kn = cnt;
    }
    for (int c = tid; c < p.nclsB; c += nthr) {
        const long long off = p.sblk_off[cA * p.nclsB + c];
        rowbase[c] = off >= 0 ? off + static_cast<long long>(ia - p.srow0[cA]) * p.nB_cls[c] : -1;
    }
    for (int i = tid; i < rows_cta; i += nthr)
        arow[i] = -1;
    for (int i = tid; i < NS * KP * TB; i += nthr)
        Bs[i] = 0.0;
    __syncthreads();
    const int Kn = s_Kn;
    if (Kn == 0)
        return;
    const int nks = (Kn + 3) >> 2;
    if (p.fold) {
        // <Ia|E_kl|Ja_r> = s_r  <=>  E_lk Ia = s_r Ja_r : row m(l,k) of the accumulators goes to row Ja_r
        for (int r = tid; r < Kn; r += nthr) {
            const int kl = kinfo[r] & 0xffff;
            const int k = kl / p.norb, l = kl - k * p.norb;
            const int m = static_cast<int>(p.rowmap[l * p.norb + k]) - mbase;
            if (m >= 0 && m < rows_cta)
                arow[m] = static_cast<short>(r);
        }
        __syncthreads();
    }

    // A fragments: this warp owns rows [m0, m0+16) of the integral matrix for the whole row Ia
    const int m0 = mbase + warp * 16;
    const int mt = mbase + nwarps * 16 + g; // TAIL: row of the shared 8-row strip held by this lane
    const int ldv = p.norb * p.norb;
    double a[KS][2];
    double at[TAIL ? KS : 1];
#pragma unroll
    for (int ks = 0; ks < KS; ++ks) {
        a[ks][0] = 0.0;
        a[ks][1] = 0.0;
        if constexpr (TAIL)
            at[ks] = 0.0;
        const int r = 4 * ks + t;
        if (r < Kn) {
            const unsigned info = kinfo[r];
            const double sg = (info >> 31) ? -1.0 : 1.0;
            const int kl = info & 0xffff;
            const bool diag = p.fold && (kl / p.norb == kl % p.norb);
            if (m0 + g < p.M)
                a[ks][0] = sg * p.Vx[static_cast<size_t>(m0 + g) * ldv + kl] + (diag ? p.kk2[m0 + g] : 0.0);
            if (m0 + g + 8 < p.M)
                a[ks][1] = sg * p.Vx[static_cast<size_t>(m0 + g + 8) * ldv + kl] + (diag ? p.kk2[m0 + g + 8] : 0.0);
            if constexpr (TAIL)
                if (mt < p.M)
                    at[ks] = sg * p.Vx[static_cast<size_t>(mt) * ldv + kl] + (diag ? p.kk2[mt] : 0.0);
        }
    }

    const int ntiles = (nBp + NT - 1) / NT;
    auto load_tile = [&](int tile) {
        if (tile < ntiles) {
            const int j0 = tile * NT;
            double* dst = Bs + (tile % NS) * KP * TB;
            if (AL16) {
                for (int idx = tid; idx < Kn * (NT / 2); idx += nthr) {
                    const int r = idx / (NT / 2), c = idx % (NT / 2);
                    const int col = j0 + 2 * c;
                    int bytes = (nBp - col) * 8;
                    bytes = bytes < 0 ? 0 : (bytes > 16 ? 16 : bytes);
                    cp_async16(dst + r * TB + 2 * c, p.C + ksrc[r] + (bytes ? col : 0), bytes);
                }
            } else {
                for (int idx = tid; idx < Kn * NT; idx += nthr) {
                    const int r = idx / NT, c = idx % NT;
                    const int col = j0 + c;
                    const bool in = col < nBp;
                    cp_async8(dst + r * TB + c, p.C + ksrc[r] + (in ? col : 0), in ? 8 : 0);
                }
            }
        }
        cp_async_commit(); // one group per call, empty past the end, keeps the wait count uniform
    };

    // per accumulator row of this lane: replacement-table row (beta scatter) and, when folding, the
    // sigma row and sign of the alpha replacement E_ij Ia
    constexpr int NR = TAIL ? 3 : 2;
    const int rowm[3] = {m0 + g, m0 + g + 8, mt};
    bool row_ok[NR], aok[NR];
    const uint32_t* trow[NR];
    double* aptr[NR];
    double ahalf[NR];
#pragma unroll
    for (int h = 0; h < NR; ++h) {
        row_ok[h] = rowm[h] < p.M;
        trow[h] = p.tgt + static_cast<size_t>(row_ok[h] ? rowm[h] : 0) * p.nbs + p.b_off[cBp];
        aok[h] = false;
        aptr[h] = p.sigma;
        ahalf[h] = 0.5;
        if (p.fold && row_ok[h]) {
            const int r = arow[rowm[h] - mbase];
            if (r >= 0 && krow[r] >= 0) {
                aok[h] = true;
                aptr[h] = p.sigma + krow[r];
                ahalf[h] = (kinfo[r] >> 31) ? -0.5 : 0.5;
            }
        }
    }
    double* const srow = p.sigma + (ONEB ? rowbase[0] : 0); // ONEB: the host guarantees the block exists

    auto fetch = [&](uint32_t (&tg)[2], int h, int col) {
        tg[0] = kNoTarget;
        tg[1] = kNoTarget;
        if (row_ok[h]) {
            if (ONEB || p.tgt_pair) {
                if (col < nBp) { // nBp even: col + 1 is in range too
                    const uint2 v = *reinterpret_cast<const uint2*>(trow[h] + col);
                    tg[0] = v.x;
                    tg[1] = v.y;
                }
            } else {
                if (col < nBp)
                    tg[0] = trow[h][col];
                if (col + 1 < nBp)
                    tg[1] = trow[h][col + 1];
            }
        }
    };
    // beta replacement: sigma[Ia, Ib] += sign * Z[m, Jb]
    auto emit = [&](double v, uint32_t x) {
        if constexpr (ONEB) {
            // address = row + 8 * (x & 0xfffff) in one mad.wide; sign = bit 31 of x onto bit 63 of v
            asm volatile("{\n\t.reg .pred p;\n\t.reg .b32 t, lo, hi;\n\t.reg .b64 a;\n\t.reg .f64 sv;\n\t"
                         "setp.ne.u32 p, %2, 0xffffffff;\n\t"
                         "and.b32 t, %2, 0xfffff;\n\t"
                         "mad.wide.u32 a, t, 8, %0;\n\t"
                         "mov.b64 {lo, hi}, %1;\n\t"
                         "and.b32 t, %2, 0x80000000;\n\t"
                         "xor.b32 hi, hi, t;\n\t"
                         "mov.b64 sv, {lo, hi};\n\t"
                         "@p red.global.add.f64 [a], sv;\n\t}"
                         ::"l"(srow), "d"(v), "r"(x)
                         : "memory");
        } else {
            if (x != kNoTarget) {
                const long long base = rowbase[(x >> 20) & 0x7ff];
                if (base >= 0)
                    atomicAdd(p.sigma + base + (x & 0xfffffu), (x >> 31) ? -v : v);
            }
        }
    };
    // folded alpha replacement: sigma[E_ij Ia, Jb] += sign/2 * Z[m, Jb]
    auto emit_fold = [&](double v, int h, int col, bool full) {
        red_add_if(aptr[h] + col, ahalf[h] * v, aok[h] && (full || col < nBp));
    };

#pragma unroll
    for (int s = 0; s < NS - 1; ++s)
        load_tile(s);
    for (int tile = 0; tile < ntiles; ++tile) {
        cp_async_wait<NS - 2>();
        __syncthreads(); // tile landed for every thread; stage (tile-1)%NS is no longer being read
        load_tile(tile + NS - 1);
        const double* Bst = Bs + (tile % NS) * KP * TB;
        const int j0 = tile * NT;
        const bool full = j0 + NT <= nBp;
#pragma unroll 1
        for (int nh = 0; nh < 2; ++nh) { // two halves of the tile keep the accumulators at 16 regs
            const int c0 = j0 + nh * (NT / 2) + 2 * t;
            // fetch the replacement targets first: their latency hides behind the MMAs
            uint32_t tg[NT / 16][2][2];
#pragma unroll
            for (int nt = 0; nt < NT / 16; ++nt)
#pragma unroll
                for (int h = 0; h < 2; ++h)
                    fetch(tg[nt][h], h, c0 + nt * 8);
            double acc[NT / 16][4];
#pragma unroll
            for (int nt = 0; nt < NT / 16; ++nt)
                acc[nt][0] = acc[nt][1] = acc[nt][2] = acc[nt][3] = 0.0;
#pragma unroll
            for (int ks = 0; ks < KS; ++ks) {
                if (ks < nks) {
                    const double* bp = Bst + (4 * ks + t) * TB + nh * (NT / 2) + g;
#pragma unroll
                    for (int nt = 0; nt < NT / 16; ++nt)
                        dmma_16x8x4(acc[nt], a[ks][0], a[ks][1], bp[nt * 8]);
                }
            }
#pragma unroll
            for (int nt = 0; nt < NT / 16; ++nt)
#pragma unroll
                for (int h = 0; h < 2; ++h)
#pragma unroll
                    for (int c = 0; c < 2; ++c)
                        emit(acc[nt][2 * h + c], tg[nt][h][c]);
            if (p.fold) {
                if (full) { // one (divergent) branch per row, then unconditional REDs at fixed offsets
#pragma unroll
                    for (int h = 0; h < 2; ++h)
                        if (aok[h]) {
                            double* dst = aptr[h] + c0;
#pragma unroll
                            for (int nt = 0; nt < NT / 16; ++nt)
#pragma unroll
                                for (int c = 0; c < 2; ++c)
                                    atomicAdd(dst + nt * 8 + c, ahalf[h] * acc[nt][2 * h + c]);
                        }
                } else {
#pragma unroll
                    for (int nt = 0; nt < NT / 16; ++nt)
#pragma unroll
                        for (int h = 0; h < 2; ++h)
#pragma unroll
                            for (int c = 0; c < 2; ++c)
                                emit_fold(acc[nt][2 * h + c], h, c0 + nt * 8 + c, false);
                }
            }
        }
        if constexpr (TAIL) {
            // this warp's strip of the shared 8-row tail: columns [cw, cw + 64/nwarps)
            const int per = NT / 8 / nwarps; // n8 tiles per warp (1 for 8 warps, 2 for 4)
#pragma unroll 1
            for (int q = 0; q < per; ++q) {
                const int cw = (warp * per + q) * 8;
                const int col = j0 + cw + 2 * t;
                uint32_t tg[2];
                fetch(tg, 2, col);
                double acc[2] = {0.0, 0.0};
#pragma unroll
                for (int ks = 0; ks < KS; ++ks)
                    if (ks < nks)
                        dmma_8x8x4(acc, at[ks], Bst[(4 * ks + t) * TB + cw + g]);
                emit(acc[0], tg[0]);
                emit(acc[1], tg[1]);
                if (p.fold) {
                    emit_fold(acc[0], 2, col, full);
                    emit_fold(acc[1], 2, col + 1, full);
                }
            }
        }
    }
}

// ---------------------------------------------------------------------------------------------
// mbarrier / TMA bulk-copy helpers (t3g_kernel)
__device__ __forceinline__ unsigned smem_u32(const void* p) {
    return static_cast<unsigned>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ void mbar_init(unsigned long long* bar, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_arrive(unsigned long long* bar) {
    asm volatile("{\n\t.reg .b64 st;\n\tmbarrier.arrive.shared::cta.b64 st, [%0];\n\t}" ::"r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long* bar, unsigned bytes) {
    asm volatile("{\n\t.reg .b64 st;\n\tmbarrier.arrive.expect_tx.shared::cta.b64 st, [%0], %1;\n\t}" ::"r"(
                     smem_u32(bar)),
                 "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long* bar, unsigned parity) {
    asm volatile("{\n\t.reg .pred p;\n\t"
                 "WAIT_%=:\n\t"
                 "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
                 "@p bra DONE_%=;\n\t"
                 "bra WAIT_%=;\n\t"
                 "DONE_%=:\n\t}" ::"r"(smem_u32(bar)),
                 "r"(parity)
                 : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, unsigned bytes, unsigned long long* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_u32(dst)),
                 "l"(src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}


// ---------------------------------------------------------------------------------------------
// Opposite-spin term, gather form (t3g_kernel): one CTA per INPUT alpha string Ja.
//
//   D[m, Ib]   = <Ib|E^b_m|Jb> C[Ja, Jb]                 gathered from the row C[Ja, :] held in
//                                                        shared memory, straight into B fragments
//   F[r, Ib]   = sum_m A'[r, m] D[m, Ib]                 A'[r, m] = s_r V[k_r l_r, m], r = single
//                                                        replacements Ia_r = s_r E^a_{k_r l_r} Ja
//   sigma[Ia_r, Ib] += F[r, Ib]                          contiguous row segments, RED
//
// Compared with t3_kernel (one CTA per OUTPUT row, scatter through the beta table) the number of
// global reductions per determinant halves (K = na(n-na)+na instead of 2K) and every one of them is
// issued once per 2*M*16 flops of tensor work instead of once per 2*K*... ; the row of C is read from
// HBM/L2 exactly once per CTA.  The alpha same-spin term cannot ride along (it needs the scatter
// form), so the caller runs ss_kernel for both spins.
// Applicable when the CI space is a single determinant block whose rows fit in shared memory.
struct T3GParams {
    const double* C;
    double* sigma;
    const uint32_t* as_ptr;
    const uint64_t* as_ent;
    const uint32_t* gsrc; // [Mpad][gld]: bits 0..19 source column (nbs = the zero slot), bit 31 sign
    const double* VxT;    // [norb*norb][M]  VxT[k*norb+l][m(i,j)] = <k i | l j> (pair-symmetrised)
    int64_t nbs, gld;
    int M, Mpad, norb;
    int first; // first alpha string of this launch
    int astride;
    // sharded build: sigma row of alpha string Ia in the reduce-scatter layout (world slots of equal
    // size, slot q = the rows rank q owns); NULL = row Ia itself
    const uint32_t* rowpad;
    // columns of the row of C staged in shared memory; the rest (long rows, e.g. CAS(18,18): 389 KB) is
    // gathered from global memory, where the row stays L2-resident while its CTA runs
    int nsm;
};

constexpr size_t t3g_smem_bytes(int64_t nsm, int rows, int astride) {
    return size_t(nsm + 2) * 8 + size_t(rows) * astride * 8 + size_t(rows) * 8 + 64;
}

// LONG = false: the whole row is staged (nsm == nbs).  NT = n8 tiles (8 target columns each) per warp work item:
// 2 up to 88 output rows; 1 above (CAS(18,18): 90 rows = 6 row tiles), which halves the accumulators so that the
// kernel keeps 12 warps per SM -- the gathers of a long row mostly go to L2 and need the extra warps to hide it.
// NX = output rows beyond the row tiles that are multiplied on the FP64 FMA pipe instead (one DFMA per lane, row and
// n8 tile and k-step, summed over the four lanes of a column at the end): the DMMA and DFMA instructions share one
// FP64 datapath on B200 (scripts/micro/fp64_pipes.cu, profiles/r2e_fp64_pipes.log).  On paper a row then costs 1/8 of
// what it costs inside an 8-row tile; measured, a DFMA issued between DMMAs costs 6-15 pipe cycles instead of 2, so the
// gain is small: CAS(16,16) 64 rows in tiles + the merged diagonal row (was 72 rows in tiles) 0.9 %, CAS(18,18) 80 + 2
// (was 88) 2.7 % (profiles/r2j_quick.log, r2j_cas18.log).
template <int NM16, bool HAS8, bool LONG, int THREADS, int NT, int NX = 0>
__global__ void __launch_bounds__(THREADS, 1) t3g_kernel(const T3GParams p) {
    constexpr int ROWS_MMA = NM16 * 16 + (HAS8 ? 8 : 0);
    constexpr int ROWS = ROWS_MMA + NX;
    constexpr int CW = 8 * NT; // target columns per work item
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int g = lane >> 2, t = lane & 3;
    const int S = p.astride;
    const int nsm = p.nsm;
    double* row = reinterpret_cast<double*>(smem_raw);            // [nsm + 2], last two are zero
    double* Asm = row + nsm + 2;                                  // [ROWS][S]
    long long* rowoff = reinterpret_cast<long long*>(Asm + ROWS * S); // [ROWS] sigma offset of Ia_r, -1
    unsigned long long* bar = reinterpret_cast<unsigned long long*>(rowoff + ROWS);
    __shared__ int s_kl[ROWS]; // (l_r * norb + k_r) | sign << 31, -1 for padding rows

    const uint32_t Ja = p.first + blockIdx.x;
    const double* src = p.C + static_cast<int64_t>(Ja) * p.nbs;
    if (tid == 0) {
        mbar_init(bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        const unsigned total = static_cast<unsigned>(nsm) * 8u;
        mbar_expect_tx(bar, total);
        for (unsigned off = 0; off < total; off += 32768u) {
            const unsigned n = total - off < 32768u ? total - off : 32768u;
            bulk_g2s(reinterpret_cast<unsigned char*>(row) + off,
                     reinterpret_cast<const unsigned char*>(src) + off, n, bar);
        }
        row[nsm] = 0.0;
        row[nsm + 1] = 0.0;
    }
    // the single replacements of Ja, read as <Ia_r| E_{lk} |Ja> = s_r from the list of Ja as an output row.  The
    // diagonal ones (k = l, one per electron) all have Ia_r = Ja and phase +1: they share ONE row of A', the sum of
    // their integral rows, placed behind the others (CAS(18,18): 82 rows = 5 tiles + 8 instead of 90 = 6 tiles)
    __shared__ int s_diag[64]; // k norb + k of the occupied orbitals
    __shared__ int s_cnt[2];   // rows (merged one included), diagonal entries
    const int e0 = p.as_ptr[Ja], e1 = p.as_ptr[Ja + 1];
    if (warp == 0) {
        int nrow = 0, ndiag = 0;
        for (int base = e0; base < e1; base += 32) {
            const int e = base + lane;
            bool off = false, dia = false;
            uint64_t ent = 0;
            int k = 0, l = 0;
            if (e < e1) {
                ent = p.as_ent[e];
                const int kl = static_cast<int>((ent >> 32) & 0xffff);
                k = kl / p.norb;
                l = kl - k * p.norb;
                dia = k == l;
                off = !dia;
            }
            const unsigned moff = __ballot_sync(0xffffffffu, off), mdia = __ballot_sync(0xffffffffu, dia);
            const unsigned below = (1u << lane) - 1u;
            if (off) {
                const int r = nrow + __popc(moff & below);
                if (r < ROWS - 1) {
                    const uint32_t Iar = static_cast<uint32_t>(ent);
                    s_kl[r] = (l * p.norb + k) | (static_cast<int>((ent >> 48) & 1) << 30);
                    rowoff[r] = static_cast<long long>(p.rowpad ? p.rowpad[Iar] : Iar) * p.nbs;
                }
            }
            if (dia) {
                const int d = ndiag + __popc(mdia & below);
                if (d < 64)
                    s_diag[d] = k * p.norb + k;
            }
            nrow += __popc(moff);
            ndiag += __popc(mdia);
        }
        nrow = nrow < ROWS - 1 ? nrow : ROWS - 1;
        ndiag = ndiag < 64 ? ndiag : 64;
        if (ndiag > 0) {
            if (lane == 0) {
                s_kl[nrow] = -2; // the merged diagonal row
                rowoff[nrow] = static_cast<long long>(p.rowpad ? p.rowpad[Ja] : Ja) * p.nbs;
            }
            nrow += 1;
        }
        for (int r = nrow + lane; r < ROWS; r += 32) {
            s_kl[r] = -1;
            rowoff[r] = -1;
        }
        if (lane == 0) {
            s_cnt[0] = nrow;
            s_cnt[1] = ndiag;
        }
    }
    __syncthreads();
    // A'[r][m] = s_r VxT[lk_r][m]; columns m >= M and rows past the list are zero
    const int ndiag = s_cnt[1];
    for (int idx = tid; idx < ROWS * S; idx += THREADS) {
        const int r = idx / S, m = idx - r * S;
        double v = 0.0;
        const int e = s_kl[r];
        if (m < p.M) {
            if (e >= 0) {
                v = __ldg(p.VxT + static_cast<size_t>(e & 0xffff) * p.M + m);
                if (e & (1 << 30))
                    v = -v;
            } else if (e == -2) {
                for (int d = 0; d < ndiag; ++d)
                    v += __ldg(p.VxT + static_cast<size_t>(s_diag[d]) * p.M + m);
            }
        }
        Asm[idx] = v;
    }
    __syncthreads();
    mbar_wait(bar, 0);

    const int nks = p.Mpad >> 2;
    const int nchunks = static_cast<int>((p.nbs + CW - 1) / CW);
    const uint32_t nbs32 = static_cast<uint32_t>(p.nbs);
    for (int chunk = warp; chunk < nchunks; chunk += THREADS / 32) {
        const int n0 = chunk * CW;
        double acc[NM16][NT][4];
        double acc8[NT][2];
#pragma unroll
        for (int i = 0; i < NM16; ++i)
#pragma unroll
            for (int nt = 0; nt < NT; ++nt)
                acc[i][nt][0] = acc[i][nt][1] = acc[i][nt][2] = acc[i][nt][3] = 0.0;
#pragma unroll
        for (int nt = 0; nt < NT; ++nt)
            acc8[nt][0] = acc8[nt][1] = 0.0;
        double accx[NX > 0 ? NX : 1][NT];
#pragma unroll
        for (int x = 0; x < (NX > 0 ? NX : 1); ++x)
#pragma unroll
            for (int nt = 0; nt < NT; ++nt)
                accx[x][nt] = 0.0;
        // lane (g, t) feeds B[k = t][n = g] of every n8 tile; n index gg of tile nt is the actual column
        // n0 + 8 nt + (gg >> 1) + 4 (gg & 1), so that the four lanes of an accumulator row hold four
        // consecutive columns (one full 32-byte sector per reduction instruction and row)
        const uint32_t* gp = p.gsrc + static_cast<size_t>(t) * p.gld + n0 + (g >> 1) + 4 * (g & 1);
        const size_t gstep = static_cast<size_t>(4) * p.gld;
        struct Q {
            uint32_t e[NT];
        };
        auto table = [&](int ks) { // entries of k-step ks (clamped: never used past the end)
            const uint32_t* q = gp + (ks < nks ? ks : 0) * gstep;
            Q r;
#pragma unroll
            for (int nt = 0; nt < NT; ++nt)
                r.e[nt] = __ldg(q + 8 * nt);
            return r;
        };
        auto gather = [&](uint32_t e) {
            const uint32_t idx = e & 0xfffffu;
            double v;
            if constexpr (LONG) {
                // staged part from shared memory (index nsm = the zero slot), the rest of the row from L2
                v = row[idx < static_cast<uint32_t>(nsm) ? idx : static_cast<uint32_t>(nsm)];
                if (idx >= static_cast<uint32_t>(nsm) && idx < nbs32)
                    v = __ldg(src + idx);
            } else {
                v = row[idx];
            }
            return __hiloint2double(__double2hiint(v) ^ static_cast<int>(e & 0x80000000u), __double2loint(v));
        };
        // table entries run five k-steps ahead of the MMAs, gathered operands one (LONG: two) ahead
        Q q0 = table(2), q1 = table(3), q2 = table(4), q3 = table(5);
        double bn[NT], bm[NT];
        {
            const Q x0 = table(0), x1 = table(1);
#pragma unroll
            for (int nt = 0; nt < NT; ++nt) {
                bn[nt] = gather(x0.e[nt]);
                bm[nt] = gather(x1.e[nt]);
            }
        }
        const double* ap = Asm + g * S + t;
#pragma unroll 4
        for (int ks = 0; ks < nks; ++ks) {
            double b[NT];
#pragma unroll
            for (int nt = 0; nt < NT; ++nt) {
                b[nt] = bn[nt];
                bn[nt] = bm[nt];
                bm[nt] = gather(q0.e[nt]);
            }
            q0 = q1;
            q1 = q2;
            q2 = q3;
            q3 = table(ks + 6);
            const double* a = ap + 4 * ks;
#pragma unroll
            for (int i = 0; i < NM16; ++i) {
                const double a0 = a[(16 * i) * S], a1 = a[(16 * i + 8) * S];
#pragma unroll
                for (int nt = 0; nt < NT; ++nt)
                    dmma_16x8x4(acc[i][nt], a0, a1, b[nt]);
            }
            if constexpr (HAS8) {
                const double a0 = a[(16 * NM16) * S];
#pragma unroll
                for (int nt = 0; nt < NT; ++nt)
                    dmma_8x8x4(acc8[nt], a0, b[nt]);
            }
            if constexpr (NX > 0) {
                const double* ax = Asm + ROWS_MMA * S + 4 * ks + t; // b[nt] is D[4 ks + t][column of lane group g]
#pragma unroll
                for (int x = 0; x < NX; ++x) {
                    const double av = ax[x * S];
#pragma unroll
                    for (int nt = 0; nt < NT; ++nt)
                        accx[x][nt] = fma(av, b[nt], accx[x][nt]);
                }
            }
        }
        if constexpr (NX > 0) {
            // the four lanes t of a group hold the four quarters of the k sum of column n0 + 8 nt + (g >> 1) + 4 (g & 1)
            const int cx = n0 + (g >> 1) + 4 * (g & 1);
#pragma unroll
            for (int x = 0; x < NX; ++x) {
                const long long ro = rowoff[ROWS_MMA + x];
#pragma unroll
                for (int nt = 0; nt < NT; ++nt) {
                    double v = accx[x][nt];
                    v += __shfl_xor_sync(0xffffffffu, v, 1);
                    v += __shfl_xor_sync(0xffffffffu, v, 2);
                    if (t == nt && ro >= 0 && cx + 8 * nt < p.nbs)
                        atomicAdd(p.sigma + ro + cx + 8 * nt, v);
                }
            }
        }
        // accumulator (row, tile nt, n index 2t + c) is column n0 + 8 nt + 4 c + t
        const int c0 = n0 + t;
#pragma unroll
        for (int i = 0; i < NM16; ++i)
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const long long ro = rowoff[16 * i + g + 8 * h];
                if (ro >= 0) {
                    double* dst = p.sigma + ro + c0;
#pragma unroll
                    for (int nt = 0; nt < NT; ++nt)
#pragma unroll
                        for (int c = 0; c < 2; ++c)
                            if (c0 + 8 * nt + 4 * c < p.nbs)
                                atomicAdd(dst + 8 * nt + 4 * c, acc[i][nt][2 * h + c]);
                }
            }
        if constexpr (HAS8) {
            const long long ro = rowoff[16 * NM16 + g];
            if (ro >= 0) {
                double* dst = p.sigma + ro + c0;
#pragma unroll
                for (int nt = 0; nt < NT; ++nt)
#pragma unroll
                    for (int c = 0; c < 2; ++c)
                        if (c0 + 8 * nt + 4 * c < p.nbs)
                            atomicAdd(dst + 8 * nt + 4 * c, acc8[nt][c]);
            }
        }
    }
}

// ---------------------------------------------------------------------------------------------
// Opposite-spin term, gather form for spaces with several determinant blocks (point-group symmetry,
// GAS): t3gp_kernel.  Same algebra as t3g_kernel, one CTA per INPUT alpha string Ja, but block-aware:
// for a chunk of 16 target beta strings of class cT only
//   * the output rows r whose block (class(Ia_r), cT) exists are multiplied (compacted per class when
//     the CTA starts), and
//   * the integral rows m whose source class srccls(cT, m) has a block with class(Ja) are contracted
//     (a class-level list made on the host).
// With point-group symmetry that is one transfer irrep per (Ja, cT): about K/nirrep rows and M/nirrep
// k-steps instead of K and M.  The B operand is gathered straight from global memory (the rows of C of
// one alpha string stay L2-resident while its CTA runs); a table entry names its source as
// class << 20 | address.  Input (Ja, Jb) and output (Ia, Ib) determinants are members of the CI space
// by construction: P H P without any intermediate outside the space.
struct T3GMParams {
    const double* C;
    double* sigma;
    const uint32_t* as_ptr;
    const uint64_t* as_ent;
    const uint32_t* a_cls;
    const uint32_t* gsrc;   // [Mpad + 1][gld]: bits 0..19 source address, 20..30 source class, bit 31 sign;
                            // low 31 bits all ones = no source; row Mpad is all "no source"
    const uint32_t* gpos;   // [nclsB] first table column of each beta class (multiple of 16)
    const uint32_t* nB_cls; // [nclsB]
    const uint32_t* crow;   // [nas][nclsB] offset of C[Ja, class, 0], 0xffffffff if the block does not exist
    const uint32_t* srow;   // [nas][nclsB] the same for sigma
    const uint16_t* ml_ent; // [nclsA][nclsB][Mpad] integral rows to contract for (class(Ja), target class)
    const int32_t* ml_cnt;  // [nclsA][nclsB]
    const double* VxT;
    int64_t gld;
    int M, Mpad, norb, nclsB;
    int first, astride;
    int rows;              // rows of A' (multiple of 16)
    const uint32_t* jlist; // sharded build: the owned input alpha strings (NULL: first + blockIdx.x)
};

constexpr uint32_t kNoRow = 0xffffffffu;

// ---------------------------------------------------------------------------------------------
// t3gp_kernel is organised for latency, not for register reuse.  ncu of its predecessor (one warp per 16
// targets, all row tiles in registers, k-steps one after the other; profiles/r03e_t3gm_*.md) on symmetry /
// GAS spaces: after the per-class compaction a
// (Ja, target class) pair has about one 16-row tile and a short k list, so a warp spends its time in the
// dependent chain  k list -> table entry -> gathered operand -> DMMA  with two warps per scheduler
// (tensor pipe 15-19 % active, 10 % of the samples waiting for the chunk list).  Here
//   * work items are (target class, 16 targets, group of T3GP_TG row tiles), enumerated from a per-class
//     prefix in shared memory (no chunk list in global memory, classes without rows or k entries cost nothing);
//   * the k loop runs in batches of T3GP_KB k-steps through a three-stage register pipeline that crosses item
//     boundaries: table entries of batch j+2 and gathered operands of batch j+1 are in flight while batch j
//     multiplies (8 + 8 independent loads per lane instead of 2 + 2);
//   * accumulators are T3GP_TG tiles only (wider row sets re-gather the operand per tile group), which leaves
//     registers for 16 warps per SM.
// (the F2_T3GP_* macros are tuning hooks like F2_SSG_*; the defaults are the measured best)
#ifndef F2_T3GP_KB
#define F2_T3GP_KB 4
#endif
#ifndef F2_T3GP_TG
#define F2_T3GP_TG 2
#endif
#ifndef F2_T3GP_THREADS
#define F2_T3GP_THREADS 512
#endif
constexpr int T3GP_KB = F2_T3GP_KB;
constexpr int T3GP_TG = F2_T3GP_TG;
constexpr int T3GP_THREADS = F2_T3GP_THREADS;
constexpr int T3GP_MAXROWS = 128;

constexpr size_t t3gp_smem_bytes(int rows, int astride, int nclsB, int mpad) {
    return size_t(rows + 1) * astride * 8 + size_t(rows) * nclsB * 4 + size_t(nclsB) * 6 * 4 + 4 +
           size_t(nclsB) * rows + size_t(nclsB) * mpad * 2 + 64;
}

// KB (k-steps per batch) and TG (row tiles per work item) are template parameters: the best pair depends on the lengths
// of the per-class k lists and row sets of the space (profiles/r3d_quick_t3gp_shapes.log: (3, 2) is 5 % faster than
// (4, 2) on the 3-GAS / 20-orbital twin and 9 % slower on CAS(16,16) D2h, (4, 1) is 5 % faster on D2h and 18 % slower
// on C2v); the host tries three instances on the first builds of a builder and keeps the fastest (launch_t3gm).
template <int THREADS, int KB = T3GP_KB, int TG = T3GP_TG>
__global__ void __launch_bounds__(THREADS, 1) t3gp_kernel(const T3GMParams p) {
    constexpr int NW = THREADS / 32;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int g = lane >> 2, t = lane & 3;
    const int S = p.astride, NC = p.nclsB, MP = p.Mpad, ROWS = p.rows;
    double* Asm = reinterpret_cast<double*>(smem_raw);                    // [ROWS + 1][S], last row zero
    uint32_t* rowoff = reinterpret_cast<uint32_t*>(Asm + (ROWS + 1) * S); // [ROWS][NC] sigma offset or kNoRow
    uint32_t* srcbase = rowoff + ROWS * NC;                               // [NC]
    int* rcnt = reinterpret_cast<int*>(srcbase + NC);                     // [NC] output rows per target class
    int* mcnt = rcnt + NC;                                                // [NC] integral rows per target class
    int* nch = mcnt + NC;                                                 // [NC] chunks of 16 targets
    int* nbc = nch + NC;                                                  // [NC] strings of the class
    int* istart = nbc + NC;                                               // [NC + 1] first work item of the class
    uint16_t* mlist = reinterpret_cast<uint16_t*>(istart + NC + 1);       // [NC][MP]
    uint8_t* rlist = reinterpret_cast<uint8_t*>(mlist + NC * MP);         // [NC][ROWS]
    __shared__ int s_kl[T3GP_MAXROWS];
    __shared__ uint32_t s_ia[T3GP_MAXROWS];

    const uint32_t Ja = p.jlist ? p.jlist[blockIdx.x] : p.first + blockIdx.x;
    const int cJa = static_cast<int>(p.a_cls[Ja]);
    const int e0 = p.as_ptr[Ja], e1 = p.as_ptr[Ja + 1];
    const int Kn = (e1 - e0) < ROWS ? (e1 - e0) : ROWS;
    for (int r = tid; r < ROWS; r += THREADS) {
        int kl = -1;
        uint32_t Iar = 0;
        if (r < Kn) {
            const uint64_t ent = p.as_ent[e0 + r];
            Iar = static_cast<uint32_t>(ent);
            const int klr = static_cast<int>((ent >> 32) & 0xffff);
            const int k = klr / p.norb, l = klr - k * p.norb;
            kl = (l * p.norb + k) | (static_cast<int>((ent >> 48) & 1) << 30);
        }
        s_kl[r] = kl;
        s_ia[r] = Iar;
    }
    for (int c = tid; c < NC; c += THREADS) {
        srcbase[c] = p.crow[static_cast<size_t>(Ja) * NC + c];
        mcnt[c] = p.ml_cnt[cJa * NC + c];
        const int nb = static_cast<int>(p.nB_cls[c]);
        nbc[c] = nb;
        nch[c] = (nb + 15) >> 4;
    }
    for (int idx = tid; idx < NC * MP; idx += THREADS)
        mlist[idx] = p.ml_ent[static_cast<size_t>(cJa) * NC * MP + idx];
    __syncthreads();
    for (int idx = tid; idx < ROWS * NC; idx += THREADS) {
        const int r = idx / NC, c = idx - r * NC;
        rowoff[idx] = s_kl[r] >= 0 ? p.srow[static_cast<size_t>(s_ia[r]) * NC + c] : kNoRow;
    }
    // A'[r][m] = s_r VxT[lk_r][m]; columns m >= M, rows r >= Kn and the extra row are zero
    for (int idx = tid; idx < (ROWS + 1) * S; idx += THREADS) {
        const int r = idx / S, m = idx - r * S;
        double v = 0.0;
        const int e = r < ROWS ? s_kl[r] : -1;
        if (e >= 0 && m < p.M) {
            v = __ldg(p.VxT + static_cast<size_t>(e & 0xffff) * p.M + m);
            if (e & (1 << 30))
                v = -v;
        }
        Asm[idx] = v;
    }
    __syncthreads();
    // output rows of every target class, compacted in order (one warp per class)
    for (int c = warp; c < NC; c += NW) {
        int cnt = 0;
        for (int base = 0; base < ROWS; base += 32) {
            const int r = base + lane;
            const bool ok = r < ROWS && rowoff[r * NC + c] != kNoRow;
            const unsigned mask = __ballot_sync(0xffffffffu, ok);
            if (ok)
                rlist[c * ROWS + cnt + __popc(mask & ((1u << lane) - 1u))] = static_cast<uint8_t>(r);
            cnt += __popc(mask);
        }
        for (int r = cnt + lane; r < ROWS; r += 32)
            rlist[c * ROWS + r] = static_cast<uint8_t>(ROWS); // the zero row of A'
        if (lane == 0)
            rcnt[c] = cnt;
    }
    __syncthreads();
    if (tid == 0) {
        int at = 0;
        for (int c = 0; c < NC; ++c) {
            istart[c] = at;
            if (rcnt[c] > 0 && mcnt[c] > 0)
                at += nch[c] * ((((rcnt[c] + 15) >> 4) + TG - 1) / TG);
        }
        istart[NC] = at;
    }
    __syncthreads();
    const int total = istart[NC];

    // ---- stage T: k list and table entries of the next batch of the next work item
    int w = warp, wcls = 0; // cursor of stage T
    int t_kb = 0, t_nkb = 0, t_nm = 0, t_desc = -1;
    const uint32_t* t_gp = p.gsrc;
    const uint16_t* t_ml = mlist;
    // descriptor: bits 0..10 target class, 11..13 tile group, 14..29 chunk of 16 targets, 30 = last batch.
    // A short last batch is padded with the all-empty table row (skipping the padded k-steps was measured
    // and loses: the branch costs more than the DMMAs it saves, profiles/r03h_one_gpu.log)
    auto t_open = [&]() {
        if (w >= total) {
            t_desc = -1;
            return;
        }
        while (w >= istart[wcls + 1])
            ++wcls;
        const int idx = w - istart[wcls], nc = nch[wcls];
        const int tg = idx / nc, j = idx - tg * nc;
        t_nm = mcnt[wcls];
        t_nkb = (t_nm + 4 * KB - 1) / (4 * KB);
        t_kb = 0;
        t_gp = p.gsrc + p.gpos[wcls] + 16 * j + (g >> 1) + 4 * (g & 1);
        t_ml = mlist + wcls * MP;
        t_desc = wcls | (tg << 11) | (j << 14);
    };
    // lane (g, t) feeds B[k = t][n = g] of both n8 tiles; n index gg of tile nt is the target address
    // 16 j + 8 nt + (gg >> 1) + 4 (gg & 1): the four lanes of an accumulator row hold consecutive targets
    auto t_load = [&](uint2(&q)[KB], int(&m)[KB], int& desc) {
        desc = t_desc;
        if (t_desc < 0) {
#pragma unroll
            for (int u = 0; u < KB; ++u) {
                q[u] = make_uint2(0x7fffffffu, 0x7fffffffu);
                m[u] = MP;
            }
            return;
        }
#pragma unroll
        for (int u = 0; u < KB; ++u) {
            const int k = 4 * (t_kb * KB + u) + t;
            m[u] = k < t_nm ? static_cast<int>(t_ml[k]) : MP; // MP: the all-empty table row, zero column of A'
            const uint32_t* e = t_gp + static_cast<size_t>(m[u]) * p.gld;
            q[u] = make_uint2(__ldg(e), __ldg(e + 8));
        }
        if (++t_kb == t_nkb) {
            desc |= 1 << 30;
            w += NW;
            t_open();
        }
    };
    auto gather = [&](uint32_t e) {
        const uint32_t body = e & 0x7fffffffu;
        double v = 0.0;
        if (body != 0x7fffffffu) {
            const uint32_t base = srcbase[body >> 20];
            if (base != kNoRow)
                v = __ldg(p.C + static_cast<size_t>(base) + (body & 0xfffffu));
        }
        return __hiloint2double(__double2hiint(v) ^ static_cast<int>(e & 0x80000000u), __double2loint(v));
    };

    uint2 tq[KB];
    int tm[KB], tdesc;
    double gb0[KB], gb1[KB];
    int gm[KB], gdesc;
    t_open();
    t_load(tq, tm, tdesc);
#pragma unroll
    for (int u = 0; u < KB; ++u) {
        gb0[u] = gather(tq[u].x);
        gb1[u] = gather(tq[u].y);
        gm[u] = tm[u];
    }
    gdesc = tdesc;
    t_load(tq, tm, tdesc);

    double acc[TG][2][4];
#pragma unroll
    for (int i = 0; i < TG; ++i)
#pragma unroll
        for (int nt = 0; nt < 2; ++nt)
            acc[i][nt][0] = acc[i][nt][1] = acc[i][nt][2] = acc[i][nt][3] = 0.0;
    bool fresh = true;
    int tiles = 0;
    const double* arow[TG][2];
    const uint8_t* rl = rlist;
    while (gdesc >= 0) {
        double b0[KB], b1[KB];
        int mm[KB];
        const int desc = gdesc;
#pragma unroll
        for (int u = 0; u < KB; ++u) {
            b0[u] = gb0[u];
            b1[u] = gb1[u];
            mm[u] = gm[u];
        }
        // operands of the next batch, table entries of the one after
#pragma unroll
        for (int u = 0; u < KB; ++u) {
            gb0[u] = gather(tq[u].x);
            gb1[u] = gather(tq[u].y);
            gm[u] = tm[u];
        }
        gdesc = tdesc;
        t_load(tq, tm, tdesc);

        const int cls = desc & 0x7ff;
        if (fresh) {
            const int tg = (desc >> 11) & 7;
            rl = rlist + cls * ROWS + tg * TG * 16;
            tiles = ((rcnt[cls] + 15) >> 4) - tg * TG;
#pragma unroll
            for (int i = 0; i < TG; ++i) {
                const int r0 = i < tiles ? rl[16 * i + g] : ROWS, r1 = i < tiles ? rl[16 * i + g + 8] : ROWS;
                arow[i][0] = Asm + r0 * S;
                arow[i][1] = Asm + r1 * S;
            }
            fresh = false;
        }
#pragma unroll
        for (int u = 0; u < KB; ++u) {
#pragma unroll
            for (int i = 0; i < TG; ++i) {
                if (i == 0 || i < tiles) {
                    const double a0 = arow[i][0][mm[u]], a1 = arow[i][1][mm[u]];
                    dmma_16x8x4(acc[i][0], a0, a1, b0[u]);
                    dmma_16x8x4(acc[i][1], a0, a1, b1[u]);
                }
            }
        }
        if (desc & (1 << 30)) {
            // accumulator (row, tile nt, n index 2t + c) is target address 16 j + 8 nt + 4 c + t of the class
            const int c0 = 16 * ((desc >> 14) & 0xffff) + t, ncls = nbc[cls];
#pragma unroll
            for (int i = 0; i < TG; ++i) {
                if (i == 0 || i < tiles) {
#pragma unroll
                    for (int h = 0; h < 2; ++h) {
                        const int r = rl[16 * i + g + 8 * h];
                        if (r < ROWS) {
                            double* dst = p.sigma + static_cast<size_t>(rowoff[r * NC + cls]) + c0;
#pragma unroll
                            for (int nt = 0; nt < 2; ++nt)
#pragma unroll
                                for (int c = 0; c < 2; ++c)
                                    if (c0 + 8 * nt + 4 * c < ncls)
                                        atomicAdd(dst + 8 * nt + 4 * c, acc[i][nt][2 * h + c]);
                        }
                    }
                }
#pragma unroll
                for (int nt = 0; nt < 2; ++nt)
                    acc[i][nt][0] = acc[i][nt][1] = acc[i][nt][2] = acc[i][nt][3] = 0.0;
            }
            fresh = true;
        }
    }
}

struct SSParams {
    const double* X;
    double* Y;
    const uint32_t* klist; // (N-2)-electron strings of this launch (blockIdx.x)
    const uint32_t* h2_ptr;
    const uint64_t* h2_ent;
    const uint32_t *s_cls, *s_loc;
    const int64_t* xoff; // [ncls_row * ngroups] offset of (row class, column group) in X, -1 if absent
    const int64_t* yoff; // same for Y
    const uint32_t* g_ncols;
    const uint32_t* g_ldx;
    const uint32_t* g_ldy;
    const uint32_t* tiles; // 2 per blockIdx.y: column group, first column of the SSG_TW-wide panel
    const double* W;       // [P*P] antisymmetrised (and one-electron folded) integrals
    int P;
    int ngroups;
    int nk; // number of strings in klist
};
// spaces with several determinant blocks: per column group the strings that have a parent in a block of
// the group (list[ptr[g] .. ptr[g + 1])), instead of SSParams::klist; ptr == NULL: klist serves every group
struct SSGroupLists {
    const uint32_t* ptr;
    const uint32_t* list;
};

// (the F2_SSG_* macros are tuning hooks for scripts/tune_ssg.sh; the defaults are the measured best)
#ifndef F2_SSG_TW
#define F2_SSG_TW 512
#endif
#ifndef F2_SSG_KG
#define F2_SSG_KG 8
#endif
#ifndef F2_SSG_PFB
#define F2_SSG_PFB 2
#endif
#ifndef F2_SSG_MINB
#define F2_SSG_MINB 2
#endif
constexpr int SSG_TW = F2_SSG_TW; // columns per CTA: a panel of all rows x SSG_TW columns of X and Y stays in L2
constexpr int SSG_KG = F2_SSG_KG; // (N-2)-electron strings per CTA, software-pipelined
__host__ __device__ constexpr int ssg_astride(int rows) { // smallest stride >= rows, == 4 (mod 16)
    int s = rows;
    while (s % 16 != 4)
        ++s;
    return s;
}
constexpr size_t ssg_smem_bytes(int rows) {
    return 2 * size_t(rows) * ssg_astride(rows) * 8 + 3 * size_t(rows) * (8 + 8 + 4) + 16;
}

__device__ __forceinline__ double flip_sign(double v, unsigned signbit31) {
    return __hiloint2double(__double2hiint(v) ^ static_cast<int>(signbit31 & 0x80000000u), __double2loint(v));
}

// Same-spin term (one- and two-electron) through the (N-2)-electron strings K, reduction form.
// One CTA per (group of SSG_KG strings K, panel of SSG_TW columns).  Per K the m x m block
// W[pair_a, pair_b] of antisymmetrised integrals is copied into shared memory with cp.async while the
// previous K multiplies (the phases s_a s_b are applied to the B operand and to the result instead of
// the block, so the copy needs no registers).  Every warp takes 16-column chunks: B fragments are the
// parent rows X[I_b, chunk] loaded straight from global memory into registers (full 32-byte sectors,
// no staging), A fragments come from the shared block, and the m x 16 result is added to Y[I_a, chunk]
// with red.global.add.f64.  No colouring, no read of Y: one launch per spin and operand; K varies
// fastest over the grid so that a column panel of X and Y is reused out of L2 by all K.
// Reference counterpart: CISigmaBuilder::H2_hz_same_spin and H1_hz
// (ci_sigma_builder_harrison_zarrabian.cc:14-195).
template <int NM16, bool HAS8, bool AL16, bool GK>
__global__ void __launch_bounds__(256, F2_SSG_MINB) ssg_kernel(const SSParams p, const SSGroupLists gk) {
    constexpr int ROWS = NM16 * 16 + (HAS8 ? 8 : 0);
    constexpr int S = ssg_astride(ROWS);
    extern __shared__ __align__(16) unsigned char smem_raw[];
    double* Abuf = reinterpret_cast<double*>(smem_raw); // [2][ROWS][S]
    long long* msrc3 = reinterpret_cast<long long*>(Abuf + 2 * ROWS * S); // [3][ROWS]
    long long* mdst3 = msrc3 + 3 * ROWS;
    unsigned* minfo3 = reinterpret_cast<unsigned*>(mdst3 + 3 * ROWS);
    __shared__ int s_m[3];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int g = lane >> 2, t = lane & 3;
    const int grp = p.tiles[2 * blockIdx.y];
    const int j0 = p.tiles[2 * blockIdx.y + 1];
    const int ncols = p.g_ncols[grp];
    const long long ldx = p.g_ldx[grp], ldy = p.g_ldy[grp];
    // GK: per-group string lists (a separate instance: the common-list kernel keeps its register allocation)
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

    auto decode = [&](int j) { // warp 0: parent list of the j-th string of this CTA into slot j % 3
        const int slot = j % 3;
        long long* msrc = msrc3 + slot * ROWS;
        long long* mdst = mdst3 + slot * ROWS;
        unsigned* minfo = minfo3 + slot * ROWS;
        const uint32_t K = GK ? gklist[k_first + j] : p.klist[k_first + j];
        const int e0 = p.h2_ptr[K], e1 = p.h2_ptr[K + 1];
        int cnt = 0;
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
                    info = static_cast<unsigned>((ent >> 32) & 0xffff) |
                           (static_cast<unsigned>((ent >> 48) & 1) << 31);
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
        for (int r = cnt + lane; r < ROWS; r += 32) {
            msrc[r] = -1;
            mdst[r] = -1;
            minfo[r] = 0;
        }
        if (lane == 0)
            s_m[slot] = cnt;
    };
    auto copy_block = [&](int j) { // all threads: W[pair_a, pair_b] of string j into buffer j % 2 (async)
        const unsigned* minfo = minfo3 + (j % 3) * ROWS;
        const int m = s_m[j % 3];
        double* A = Abuf + (j & 1) * ROWS * S;
        // only the 16-row tiles that hold parents are multiplied (nothing at all for a string without parents
        // in this column group, the usual case in GAS and symmetry-blocked spaces), so only they are copied
        const int rows_used = ((m + 15) >> 4) << 4;
        for (int idx = tid; idx < (rows_used < ROWS ? rows_used : ROWS) * S; idx += 256) {
            const int a = idx / S, b = idx - a * S;
            const bool in = a < m && b < m;
            const double* src = in ? p.W + static_cast<size_t>(minfo[a] & 0xffff) * p.P + (minfo[b] & 0xffff) : p.W;
            cp_async8(A + idx, src, in ? 8 : 0);
        }
        cp_async_commit();
    };

    if (warp == 0) {
        decode(0);
        if (k_cnt > 1)
            decode(1);
    }
    __syncthreads();
    copy_block(0);
    cp_async_wait<0>();
    __syncthreads();

    const int jend = j0 + SSG_TW < ncols ? j0 + SSG_TW : ncols;
    const int nchunks = (jend - j0 + 15) >> 4;
    constexpr int PFB = F2_SSG_PFB; // k-steps per register batch; the next batch loads while this one multiplies
    double cur[PFB][2], nxt[PFB][2];
    unsigned scur[PFB], snxt[PFB];
    // parent row 4ks + t of the string in slot `slot`, columns cb and cb + 8 (zeros past the end)
    auto loadb = [&](int slot, int ks, int cb, double (&v)[2], unsigned& sg) {
        v[0] = 0.0;
        v[1] = 0.0;
        sg = 0;
        const int b = 4 * ks + t;
        if (b < s_m[slot]) {
            const double* src = p.X + msrc3[slot * ROWS + b] + cb;
            sg = minfo3[slot * ROWS + b];
            if (cb < jend)
                v[0] = __ldg(src);
            if (cb + 8 < jend)
                v[1] = __ldg(src + 8);
        }
    };
    // B operand: n8 tile nt, n index gg holds column n0 + 8 nt + (gg >> 1) + 4 (gg & 1), so that the
    // four lanes of an accumulator row hold four consecutive columns (one full sector per reduction)
    const int cboff = (g >> 1) + 4 * (g & 1);
    bool primed = false; // cur already holds the first batch of the next (string, chunk) of this warp
    for (int j = 0; j < k_cnt; ++j) {
        if (j + 1 < k_cnt)
            copy_block(j + 1);
        if (warp == 0 && j + 2 < k_cnt)
            decode(j + 2);
        const int slot = j % 3;
        const int m = s_m[slot];
        const long long* mdst = mdst3 + slot * ROWS;
        const unsigned* minfo = minfo3 + slot * ROWS;
        const int nks = (m + 3) >> 2;
        const double* ap = Abuf + (j & 1) * ROWS * S + g * S + t;
        for (int chunk = warp; chunk < nchunks && m > 0; chunk += 8) {
            const int n0 = j0 + (chunk << 4);
            const int cb = n0 + cboff;
            double acc[NM16 > 0 ? NM16 : 1][2][4];
            double acc8[2][2];
#pragma unroll
            for (int i = 0; i < NM16; ++i)
#pragma unroll
                for (int nt = 0; nt < 2; ++nt)
                    acc[i][nt][0] = acc[i][nt][1] = acc[i][nt][2] = acc[i][nt][3] = 0.0;
            acc8[0][0] = acc8[0][1] = acc8[1][0] = acc8[1][1] = 0.0;
            if (!primed) {
#pragma unroll
                for (int i = 0; i < PFB; ++i)
                    loadb(slot, i, cb, cur[i], scur[i]);
            }
            // the work item after this one: next chunk of the same string, else the first chunk of the
            // next string (its parent list was decoded two strings ahead and is visible since the last
            // barrier); its first batch is loaded during the last multiplications of this one, so the
            // L2 latency of a new chunk hides behind the reductions, the barrier and the tensor work
            const bool same = chunk + 8 < nchunks;
            const bool has_next = same || (j + 1 < k_cnt && s_m[(j + 1) % 3] > 0);
            const int nslot = same ? slot : (j + 1) % 3;
            const int ncb = j0 + ((same ? chunk + 8 : warp) << 4) + cboff;
            for (int kb = 0; kb < nks; kb += PFB) {
                if (kb + PFB < nks) {
#pragma unroll
                    for (int i = 0; i < PFB; ++i)
                        loadb(slot, kb + PFB + i, cb, nxt[i], snxt[i]);
                } else if (has_next) {
#pragma unroll
                    for (int i = 0; i < PFB; ++i)
                        loadb(nslot, i, ncb, nxt[i], snxt[i]);
                }
#pragma unroll
                for (int i = 0; i < PFB; ++i) {
                    const int ks = kb + i;
                    if (ks < nks) {
                        const double b0 = flip_sign(cur[i][0], scur[i]), b1 = flip_sign(cur[i][1], scur[i]);
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
#pragma unroll
                for (int i = 0; i < PFB; ++i) {
                    cur[i][0] = nxt[i][0];
                    cur[i][1] = nxt[i][1];
                    scur[i] = snxt[i];
                }
            }
            primed = has_next;
            // accumulator (row, tile nt, n index 2t + c) is column n0 + 8 nt + 4 c + t
            const int c0 = n0 + t;
#pragma unroll
            for (int mt = 0; mt < NM16; ++mt)
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    const int a = 16 * mt + g + 8 * h;
                    if (a < m) {
                        double* dst = p.Y + mdst[a] + c0;
                        const unsigned sg = minfo[a];
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
                    const unsigned sg = minfo[a];
#pragma unroll
                    for (int nt = 0; nt < 2; ++nt)
#pragma unroll
                        for (int c = 0; c < 2; ++c)
                            if (c0 + 8 * nt + 4 * c < jend)
                                atomicAdd(dst + 8 * nt + 4 * c, flip_sign(acc8[nt][c], sg));
                }
            }
        }
        cp_async_wait<0>();
        __syncthreads();
    }
}

// ---------------------------------------------------------------------------------------------
// Shape-generic kernels (FP64 FMA pipe, no tensor tiles): the sigma terms for spaces whose parent lists or
// single-replacement lists are longer than the tensor-core kernels hold in shared memory and registers (more than 112
// parents per (N-2)-electron string, e.g. 4 electrons of one spin in 20 orbitals; more than 128 single replacements per
// alpha string, e.g. 8 alpha electrons in 24 orbitals).  Same operands, lists, plans and launch grids as ssg_kernel /
// t3_kernel; the panel of columns a CTA works on is cut into strips of `tw` columns (a power of two, at most the
// number of threads) so that the staged rows fit in shared memory whatever the list length.
constexpr int GEN_THREADS = 256;

constexpr size_t gen_smem_bytes(int rows, int tw, int extra8 = 0) {
    return size_t(rows) * tw * 8 + size_t(rows) * (8 + 8 + 4) + size_t(extra8) * 8 + 64;
}

// same-spin term: one CTA per ((N-2)-electron string K, column panel); Y[I_a, col] += s_a sum_b W[pair_a, pair_b] s_b X[I_b, col]
template <bool GK>
__global__ void __launch_bounds__(GEN_THREADS) ss_generic_kernel(const SSParams p, const SSGroupLists gk, const int rows,
                                                                 const int tw) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    double* Xs = reinterpret_cast<double*>(smem_raw);               // [rows][tw], phase applied
    long long* msrc = reinterpret_cast<long long*>(Xs + rows * tw); // [rows]
    long long* mdst = msrc + rows;
    unsigned* minfo = reinterpret_cast<unsigned*>(mdst + rows);
    __shared__ int s_m;
    const int tid = threadIdx.x, lane = tid & 31;
    const int grp = p.tiles[2 * blockIdx.y];
    const int j0 = p.tiles[2 * blockIdx.y + 1];
    const int ncols = p.g_ncols[grp];
    const long long ldx = p.g_ldx[grp], ldy = p.g_ldy[grp];
    uint32_t K;
    if constexpr (GK) {
        const int gnk = static_cast<int>(gk.ptr[grp + 1] - gk.ptr[grp]);
        if (static_cast<int>(blockIdx.x) >= gnk)
            return;
        K = gk.list[gk.ptr[grp] + blockIdx.x];
    } else {
        K = p.klist[blockIdx.x];
    }
    if (tid < 32) { // parents of K that have a block in this column group, compacted in list order
        const int e0 = p.h2_ptr[K], e1 = p.h2_ptr[K + 1];
        int cnt = 0;
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
                if (pos < rows) {
                    msrc[pos] = src;
                    mdst[pos] = dst;
                    minfo[pos] = info;
                }
            }
            cnt += __popc(mask);
        }
        if (lane == 0)
            s_m = cnt < rows ? cnt : rows;
    }
    __syncthreads();
    const int m = s_m;
    if (m == 0)
        return;
    const int jend = j0 + SSG_TW < ncols ? j0 + SSG_TW : ncols;
    const int c = tid & (tw - 1), a_first = tid / tw, a_step = GEN_THREADS / tw;
    for (int s0 = j0; s0 < jend; s0 += tw) {
        for (int idx = tid; idx < m * tw; idx += GEN_THREADS) {
            const int r = idx / tw, cc = idx - r * tw;
            double v = 0.0;
            if (s0 + cc < jend)
                v = flip_sign(__ldg(p.X + msrc[r] + s0 + cc), minfo[r]);
            Xs[idx] = v;
        }
        __syncthreads();
        if (s0 + c < jend) {
            for (int a = a_first; a < m; a += a_step) {
                const double* w = p.W + static_cast<size_t>(minfo[a] & 0xffff) * p.P;
                double acc = 0.0;
                for (int b = 0; b < m; ++b)
                    acc += __ldg(w + (minfo[b] & 0xffff)) * Xs[b * tw + c];
                atomicAdd(p.Y + mdst[a] + s0 + c, flip_sign(acc, minfo[a]));
            }
        }
        __syncthreads();
    }
}

// opposite-spin term, scatter form: one CTA per (output alpha string Ia, beta class of the intermediate column);
// G[m, Jb] = sum_r s_r Vx[m, kl_r] C[Ja_r, Jb], then sigma[Ia, Ib] += s' G[m, Jb] through the beta replacement table
__global__ void __launch_bounds__(GEN_THREADS) t3_generic_kernel(const T3Params p, const int rows, const int tw) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    double* Cs = reinterpret_cast<double*>(smem_raw);               // [rows][tw], phase applied
    long long* ksrc = reinterpret_cast<long long*>(Cs + rows * tw); // [rows]
    long long* unused = ksrc + rows;                                // (keeps the layout of gen_smem_bytes)
    unsigned* kinfo = reinterpret_cast<unsigned*>(unused + rows);   // [rows]
    long long* rowbase = reinterpret_cast<long long*>(kinfo + rows + (rows & 1)); // [nclsB]
    __shared__ int s_Kn;
    const int tid = threadIdx.x, lane = tid & 31;
    const uint32_t Ia = p.items[2 * blockIdx.x];
    const int cBp = p.items[2 * blockIdx.x + 1] & 0xffff;
    const int cA = p.a_cls[Ia];
    const uint32_t ia = p.a_loc[Ia];
    const int nBp = p.nB_cls[cBp];
    if (tid < 32) {
        const int e0 = p.as_ptr[Ia], e1 = p.as_ptr[Ia + 1];
        int cnt = 0;
        for (int base = e0; base < e1; base += 32) {
            const int e = base + lane;
            bool ok = false;
            long long src = -1;
            unsigned info = 0;
            if (e < e1) {
                const uint64_t ent = p.as_ent[e];
                const uint32_t Ja = static_cast<uint32_t>(ent);
                const long long off = p.blk_off[p.a_cls[Ja] * p.nclsB + cBp];
                if (off >= 0) {
                    ok = true;
                    src = off + static_cast<long long>(p.a_loc[Ja]) * nBp;
                    info = static_cast<unsigned>((ent >> 32) & 0xffff) | (static_cast<unsigned>((ent >> 48) & 1) << 31);
                }
            }
            const unsigned mask = __ballot_sync(0xffffffffu, ok);
            if (ok) {
                const int pos = cnt + __popc(mask & ((1u << lane) - 1u));
                if (pos < rows) {
                    ksrc[pos] = src;
                    kinfo[pos] = info;
                }
            }
            cnt += __popc(mask);
        }
        if (lane == 0)
            s_Kn = cnt < rows ? cnt : rows;
    }
    for (int cc = tid; cc < p.nclsB; cc += GEN_THREADS) {
        const long long off = p.sblk_off[cA * p.nclsB + cc];
        rowbase[cc] = off >= 0 ? off + static_cast<long long>(ia - p.srow0[cA]) * p.nB_cls[cc] : -1;
    }
    __syncthreads();
    const int Kn = s_Kn;
    if (Kn == 0)
        return;
    const int ldv = p.norb * p.norb;
    const int c = tid & (tw - 1), m_first = tid / tw, m_step = GEN_THREADS / tw;
    const uint32_t* tgt0 = p.tgt + p.b_off[cBp];
    for (int s0 = 0; s0 < nBp; s0 += tw) {
        for (int idx = tid; idx < Kn * tw; idx += GEN_THREADS) {
            const int r = idx / tw, cc = idx - r * tw;
            double v = 0.0;
            if (s0 + cc < nBp)
                v = flip_sign(__ldg(p.C + ksrc[r] + s0 + cc), kinfo[r]);
            Cs[idx] = v;
        }
        __syncthreads();
        if (s0 + c < nBp) {
            for (int m = m_first; m < p.M; m += m_step) {
                const uint32_t x = __ldg(tgt0 + static_cast<size_t>(m) * p.nbs + s0 + c);
                if (x == kNoTarget)
                    continue;
                const long long base = rowbase[(x >> 20) & 0x7ff];
                if (base < 0)
                    continue;
                const double* v = p.Vx + static_cast<size_t>(m) * ldv;
                double acc = 0.0;
                for (int r = 0; r < Kn; ++r)
                    acc += __ldg(v + (kinfo[r] & 0xffff)) * Cs[r * tw + c];
                atomicAdd(p.sigma + base + (x & 0xfffffu), (x >> 31) ? -acc : acc);
            }
        }
        __syncthreads();
    }
}

// ---------------------------------------------------------------------------------------------
// The beta replacement tables of the opposite-spin kernels, made on the device from the beta single-replacement list
// (ELL by source string; 3.4e7 entries at 3 GAS / 20 orbitals, 0.7 s of host loops before).
__global__ void fill_u32_kernel(uint32_t* __restrict__ p, uint32_t v, int64_t n) {
    const int64_t stride = static_cast<int64_t>(gridDim.x) * blockDim.x;
    for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += stride)
        p[i] = v;
}

// tgt[m][Jb] = sign << 31 | class(Ib) << 20 | address(Ib) for <Ib| E_m |Jb> = sign (scatter form); pre-filled with all ones
__global__ void tgt_table_kernel(const uint64_t* __restrict__ bs_ent, int64_t nbs, int width,
                                 const uint16_t* __restrict__ rowmap, uint32_t* __restrict__ tgt) {
    const int64_t idx = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (idx >= nbs * width)
        return;
    const uint64_t ent = bs_ent[idx];
    if (ent == ~0ull)
        return;
    const int64_t gJ = idx % nbs;
    const uint32_t m = rowmap[(ent >> 40) & 0xffff];
    tgt[static_cast<int64_t>(m) * nbs + gJ] = static_cast<uint32_t>(ent & 0xfffff) |
                                              (static_cast<uint32_t>((ent >> 28) & 0x7ff) << 20) |
                                              (static_cast<uint32_t>((ent >> 56) & 1) << 31);
}

// its inverse for one beta class (gather form): gsrc[m][Ib] = Jb | sign; pre-filled with nbs (the zero slot)
__global__ void gsrc_table_kernel(const uint32_t* __restrict__ tgt, int M, int64_t nbs, int64_t gld, uint32_t* __restrict__ gsrc) {
    const int64_t idx = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (idx >= static_cast<int64_t>(M) * nbs)
        return;
    const uint32_t x = tgt[idx];
    if (x == kNoTarget)
        return;
    const int64_t m = idx / nbs, gJ = idx - m * nbs;
    gsrc[m * gld + (x & 0xfffffu)] = static_cast<uint32_t>(gJ) | (x & 0x80000000u);
}

// the inverse for several beta classes (t3gp_kernel): table columns are the targets class by class (gpos), an entry
// names its source as class << 20 | address; pre-filled with 0x7fffffff.  srccls[target class][m][2]: the (at most two)
// source classes of integral row m, -2 in slot 0 = more than two (pre-filled with -1).
__global__ void gm_table_kernel(const uint32_t* __restrict__ tgt, int M, int Mpad, int64_t nbs, int64_t gld,
                                const uint32_t* __restrict__ gpos, const uint32_t* __restrict__ b_cls,
                                const uint32_t* __restrict__ b_loc, uint32_t* __restrict__ gsrc, int* __restrict__ srccls) {
    const int64_t idx = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (idx >= static_cast<int64_t>(M) * nbs)
        return;
    const uint32_t x = tgt[idx];
    if (x == kNoTarget)
        return;
    const int64_t m = idx / nbs, gJ = idx - m * nbs;
    const uint32_t ct = (x >> 20) & 0x7ff;
    const int cj = static_cast<int>(b_cls[gJ]);
    gsrc[m * gld + gpos[ct] + (x & 0xfffffu)] = (static_cast<uint32_t>(cj) << 20) | b_loc[gJ] | (x & 0x80000000u);
    int* sc = srccls + (static_cast<int64_t>(ct) * Mpad + m) * 2;
    if (sc[0] == cj || sc[0] == -2 || sc[1] == cj) // (plain reads: a stale value only costs the atomics below)
        return;
    const int o0 = atomicCAS(&sc[0], -1, cj);
    if (o0 == -1 || o0 == cj || o0 == -2)
        return;
    const int o1 = atomicCAS(&sc[1], -1, cj);
    if (o1 == -1 || o1 == cj)
        return;
    atomicExch(&sc[0], -2);
}

// sigma = E * C  (ACC: sigma += E * C, when the opposite-spin term has already been streamed in)
template <bool ACC>
__global__ void scale_copy_kernel(const double* __restrict__ x, double* __restrict__ y, double s,
                                  int64_t n) {
    const int64_t stride = static_cast<int64_t>(gridDim.x) * blockDim.x;
    for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += stride)
        y[i] = ACC ? y[i] + s * x[i] : s * x[i];
}

struct TrBlock {
    int64_t src_off, dst_off; // row-major [nA][ldA] at src_off  <->  row-major [nB][nA] at dst_off
    int64_t sig_off;          // where the same rows sit in sigma (== src_off unless sigma is a local slab)
    int32_t nA, nB;
    int64_t ldA;
    uint32_t tile0; // first tile id of this block
    uint32_t tiles_j;
};

// forward: T[dst_off + j*nA + i] = S[src_off + i*ldA + j];  back: Sigma[sig_off + i*ldA + j] += T[...]
// local_src (forward only): `in` is the local row-slab vector, rows sit at sig_off instead of src_off
template <bool BACK>
__global__ void transpose_kernel(const double* __restrict__ in, double* __restrict__ out,
                                 const TrBlock* __restrict__ blocks, int nblocks, bool local_src) {
    __shared__ double tile[32][33];
    // locate the block of this tile (blocks are few; binary search over tile0)
    int lo = 0, hi = nblocks - 1;
    while (lo < hi) {
        const int mid = (lo + hi + 1) >> 1;
        if (blocks[mid].tile0 <= blockIdx.x)
            lo = mid;
        else
            hi = mid - 1;
    }
    const TrBlock b = blocks[lo];
    const uint32_t tl = blockIdx.x - b.tile0;
    const int ti = tl / b.tiles_j, tj = tl % b.tiles_j;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5; // 32 x 8
    if (!BACK) {
        for (int r = ty; r < 32; r += 8) {
            const int i = ti * 32 + r, j = tj * 32 + tx;
            if (i < b.nA && j < b.nB)
                tile[r][tx] = in[(local_src ? b.sig_off : b.src_off) + static_cast<int64_t>(i) * b.ldA + j];
        }
        __syncthreads();
        for (int r = ty; r < 32; r += 8) {
            const int j = tj * 32 + r, i = ti * 32 + tx;
            if (i < b.nA && j < b.nB)
                out[b.dst_off + static_cast<int64_t>(j) * b.nA + i] = tile[tx][r];
        }
    } else {
        for (int r = ty; r < 32; r += 8) {
            const int j = tj * 32 + r, i = ti * 32 + tx;
            if (i < b.nA && j < b.nB)
                tile[tx][r] = in[b.dst_off + static_cast<int64_t>(j) * b.nA + i];
        }
        __syncthreads();
        for (int r = ty; r < 32; r += 8) {
            const int i = ti * 32 + r, j = tj * 32 + tx;
            if (i < b.nA && j < b.nB)
                out[b.sig_off + static_cast<int64_t>(i) * b.ldA + j] += tile[r][tx];
        }
    }
}

struct CopyPiece {
    int64_t src_off, dst_off, n;
    uint32_t cta0; // first CTA of this piece (1024 elements per CTA)
};

// dst[dst_off + i] (=|+=) s * src[src_off + i] for a list of contiguous pieces
template <bool ACC>
__global__ void piece_copy_kernel(const double* __restrict__ src, double* __restrict__ dst,
                                  const CopyPiece* __restrict__ pieces, int npieces, double s) {
    int lo = 0, hi = npieces - 1;
    while (lo < hi) {
        const int mid = (lo + hi + 1) >> 1;
        if (pieces[mid].cta0 <= blockIdx.x)
            lo = mid;
        else
            hi = mid - 1;
    }
    const CopyPiece pc = pieces[lo];
    const int64_t base = static_cast<int64_t>(blockIdx.x - pc.cta0) * 1024;
    for (int k = threadIdx.x; k < 1024; k += blockDim.x) {
        const int64_t i = base + k;
        if (i < pc.n) {
            if (ACC)
                dst[pc.dst_off + i] += s * src[pc.src_off + i];
            else
                dst[pc.dst_off + i] = s * src[pc.src_off + i];
        }
    }
}

struct HdiagBlock {
    int64_t off;
    uint32_t a0, b0; // first global alpha / beta string index of the block's classes
    int32_t nA, nB;
    uint32_t tile0;
};

// Spin-symmetric diagonal for determinant-basis solves without a CSF adapter: the determinant energies averaged
// over the determinants of the same orbital configuration (same doubly / singly occupied orbitals, same Ms).  Only
// the exchange between singly occupied orbitals depends on how the spins are distributed:
//   E_avg = E_det + sum_{p<q open, same spin in this determinant} K_pq - P_same(N_open) sum_{p<q open} K_pq
// with K_pq = <pq|qp> and P_same the fraction of spin distributions that give p and q the same spin.  A
// preconditioner built from it commutes with S^2, so spin-pure trial vectors stay spin-pure.
struct SpinAvg {
    const double* kx;    // [norb * norb] exchange integrals
    const double* psame; // [65] P_same by number of open shells
};

__device__ __forceinline__ double open_shell_exchange(uint64_t w, const double* __restrict__ kx, int norb) {
    double s = 0.0;
    uint64_t u = w;
    while (u) {
        const int p = __ffsll(static_cast<long long>(u)) - 1;
        u &= u - 1;
        uint64_t v = u;
        while (v) {
            const int q = __ffsll(static_cast<long long>(v)) - 1;
            v &= v - 1;
            s += kx[p * norb + q];
        }
    }
    return s;
}

// E_det = E + ea[Ia] + eb[Ib] + sum_{q in Ib} ja[Ia][q]   (SlaterRules::energy)
__global__ void hdiag_kernel(double* __restrict__ out, const HdiagBlock* __restrict__ blocks,
                             int nblocks, const double* __restrict__ ea,
                             const double* __restrict__ eb, const double* __restrict__ ja,
                             const uint64_t* __restrict__ bstr, int norb, double E,
                             const uint64_t* __restrict__ astr = nullptr, SpinAvg avg = SpinAvg{nullptr, nullptr}) {
    int lo = 0, hi = nblocks - 1;
    while (lo < hi) {
        const int mid = (lo + hi + 1) >> 1;
        if (blocks[mid].tile0 <= blockIdx.x)
            lo = mid;
        else
            hi = mid - 1;
    }
    const HdiagBlock b = blocks[lo];
    const int i = blockIdx.x - b.tile0; // one alpha row per CTA
    const double* jrow = ja + static_cast<size_t>(b.a0 + i) * norb;
    const double e0 = E + ea[b.a0 + i];
    const uint64_t wa = avg.kx ? astr[b.a0 + i] : 0ull;
    for (int j = threadIdx.x; j < b.nB; j += blockDim.x) {
        const uint64_t wb = bstr[b.b0 + j];
        uint64_t w = wb;
        double e = e0 + eb[b.b0 + j];
        while (w) {
            const int q = __ffsll(static_cast<long long>(w)) - 1;
            e += jrow[q];
            w &= w - 1;
        }
        if (avg.kx) {
            const uint64_t open = wa ^ wb;
            e += open_shell_exchange(wa & open, avg.kx, norb) + open_shell_exchange(wb & open, avg.kx, norb) -
                 avg.psame[__popcll(open)] * open_shell_exchange(open, avg.kx, norb);
        }
        out[b.off + static_cast<int64_t>(i) * b.nB + j] = e;
    }
}

} // namespace f2
