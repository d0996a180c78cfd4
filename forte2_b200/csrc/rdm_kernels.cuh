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
    // point-group symmetry: Gamma[(A B), (A' B')] vanishes unless the two operator tuples have the same irrep h, and a
    // column (Ka, Kb) has entries in the rows of irrep h only when irrep(Ka) x irrep(Kb) = h x symmetry of the state.
    // The block of one h is then built from the hole strings of matching irreps only: ka0 / kb0 index these lists
    // (NULL: all hole strings), rowmap sends an operator row to its row inside the block of h (-1: another irrep).
    const uint32_t *kaList, *kbList;
    const int32_t* rowmap;
};

constexpr int RDM_GATHER_THREADS = 256;
constexpr int RDM_MAXPA = 512; // parents of one Ka staged per pass

__global__ void __launch_bounds__(RDM_GATHER_THREADS) rdm_gather_kernel(const RdmGatherParams p) {
    __shared__ uint32_t s_cls[RDM_MAXPA], s_loc[RDM_MAXPA], s_row[RDM_MAXPA]; // row = opA * Pb | sign << 31
    const uint32_t Ka = p.kaList ? p.kaList[p.ka0 + blockIdx.x] : p.ka0 + blockIdx.x;
    const uint32_t kb = blockIdx.y * RDM_GATHER_THREADS + threadIdx.x; // within the chunk
    const uint32_t a0 = p.ptrA[Ka], a1 = p.ptrA[Ka + 1];
    const int64_t col = static_cast<int64_t>(blockIdx.x) * p.nkb + kb;
    uint32_t b0 = 0, b1 = 0;
    if (kb < p.nkb) {
        const uint32_t Kb = p.kbList ? p.kbList[p.kb0 + kb] : p.kb0 + kb;
        b0 = p.ptrB[Kb];
        b1 = p.ptrB[Kb + 1];
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
                int64_t row = (s_row[i] & 0x7fffffffu) + opB;
                if (p.rowmap) {
                    row = p.rowmap[row];
                    if (row < 0)
                        continue;
                }
                p.D[row * p.ldd + col] = ((s_row[i] >> 31) ^ negB) ? -v : v;
            }
        }
    }
}

// ---------------------------------------------------------------------------------------------
// C[M x M] += A[M x K] B[M x K]^T, both row-major with the K index contiguous (leading dimension ld, even).
// BM = tile edge: 128 for the full operator sets, 64 / 32 for the small row sets of one irrep (256 / 8 = 32 rows of the
// opposite-spin 2-RDM at 16 orbitals in D2h: a 128-row tile would multiply 16 x more than there is).
constexpr int RG_BM = 128, RG_BK = 16, RG_STAGES = 3, RG_THREADS = 256;
constexpr int RG_LDS = RG_BK + 4; // 20 doubles: rows g = 0..7 of a fragment load fall on disjoint banks
constexpr size_t rdm_gemm_smem_bytes(int bm = RG_BM) { return size_t(RG_STAGES) * (2 * bm) * RG_LDS * 8; }

struct RdmGemmParams {
    const double* A;
    const double* B;
    double* C;
    int M;          // rows of A and of B, C is M x M (leading dimension M)
    int64_t K, ld;  // columns of this chunk, leading dimension of A and B
    int ksplit;     // CTAs along K per tile
    const int32_t* unmap; // NULL, or row of the block -> row of the full result (leading dimension ldc)
    int ldc;
};

template <int BM>
__global__ void __launch_bounds__(RG_THREADS) rdm_gemm_kernel(const RdmGemmParams p) {
    constexpr int BN = BM, NI = BM / 32, NJ = BM / 32; // per warp: NI 16-row tiles x NJ 8-column tiles
    extern __shared__ __align__(16) unsigned char rg_smem[];
    double* As = reinterpret_cast<double*>(rg_smem);                 // [STAGES][BM][LDS]
    double* Bs = As + RG_STAGES * BM * RG_LDS;                       // [STAGES][BN][LDS]
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int g = lane >> 2, t = lane & 3;
    const int tiles_n = (p.M + BN - 1) / BN;
    const int tile = blockIdx.x / p.ksplit, ks = blockIdx.x - tile * p.ksplit;
    const int m0 = (tile / tiles_n) * BM, n0 = (tile % tiles_n) * BN;
    // this CTA's share of the K range, in steps of BK
    const int64_t nsteps_all = (p.K + RG_BK - 1) / RG_BK;
    const int64_t s0 = nsteps_all * ks / p.ksplit, s1 = nsteps_all * (ks + 1) / p.ksplit;
    const int nsteps = static_cast<int>(s1 - s0);
    if (nsteps <= 0)
        return;
    const int wm = (warp >> 2) * (BM / 2), wn = (warp & 3) * (BN / 4); // warp tile (BM / 2) x (BN / 4)

    auto load_stage = [&](int step, int stage) {
        // 2 BM rows (BM of A, BN of B) x 16 doubles = 8 pieces of 16 bytes per row
        const int64_t k0 = (s0 + step) * RG_BK;
#pragma unroll
        for (int it = 0; it < (2 * BM * 8) / RG_THREADS; ++it) {
            const int idx = tid + it * RG_THREADS;
            const int row = idx >> 3, piece = idx & 7;
            const bool isB = row >= BM;
            const int r = isB ? row - BM : row;
            const int grow = (isB ? n0 : m0) + r;
            const int64_t k = k0 + 2 * piece;
            int bytes = 0;
            if (grow < p.M && k < p.K)
                bytes = (p.K - k >= 2) ? 16 : 8;
            const double* src = (isB ? p.B : p.A) + (bytes ? static_cast<int64_t>(grow) * p.ld + k : 0);
            double* dst = (isB ? Bs : As) + (stage * BM + r) * RG_LDS + 2 * piece;
            cp_async16(dst, src, bytes);
        }
        cp_async_commit();
    };

    double acc[NI][NJ][4];
#pragma unroll
    for (int i = 0; i < NI; ++i)
#pragma unroll
        for (int j = 0; j < NJ; ++j)
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
        const double* Ast = As + (step % RG_STAGES) * BM * RG_LDS;
        const double* Bst = Bs + (step % RG_STAGES) * BN * RG_LDS;
#pragma unroll
        for (int kk = 0; kk < RG_BK; kk += 4) {
            double a0[NI], a1[NI], b[NJ];
#pragma unroll
            for (int i = 0; i < NI; ++i) {
                a0[i] = Ast[(wm + 16 * i + g) * RG_LDS + kk + t];
                a1[i] = Ast[(wm + 16 * i + g + 8) * RG_LDS + kk + t];
            }
#pragma unroll
            for (int j = 0; j < NJ; ++j)
                b[j] = Bst[(wn + 8 * j + g) * RG_LDS + kk + t];
#pragma unroll
            for (int i = 0; i < NI; ++i)
#pragma unroll
                for (int j = 0; j < NJ; ++j)
                    dmma_16x8x4(acc[i][j], a0[i], a1[i], b[j]);
        }
    }
    cp_async_wait<0>();
    // accumulator (tile i, j): rows wm + 16 i + g (+8), columns wn + 8 j + 2 t (+1)
#pragma unroll
    for (int i = 0; i < NI; ++i)
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            const int row = m0 + wm + 16 * i + g + 8 * h;
            if (row >= p.M)
                continue;
            const int64_t crow = static_cast<int64_t>(p.unmap ? p.unmap[row] : row) * p.ldc;
#pragma unroll
            for (int j = 0; j < NJ; ++j)
#pragma unroll
                for (int c = 0; c < 2; ++c) {
                    const int colc = n0 + wn + 8 * j + 2 * t + c;
                    if (colc < p.M)
                        atomicAdd(p.C + crow + (p.unmap ? p.unmap[colc] : colc), acc[i][j][2 * h + c]);
                }
        }
}

// ---------------------------------------------------------------------------------------------
// One-spin blocks (1-RDMs, same-spin 2-RDMs, same-spin 3-RDMs with short parent lists) without the D matrix: for an
// (N - level)-electron string K of the acting spin with parents I_a = K + A (a = 0 .. m-1),
//   Gamma[A, A'] += s_A s_A' sum over the determinants of the other spin  L[I_A, .] R[I_A', .]
// is the m x m Gram matrix of the parent rows -- the transpose of the same-spin sigma contraction (ssg_kernel: rows in,
// rows out; here rows in, a small matrix out).  One CTA per K runs over the whole row length: the parent rows stream
// through shared memory in chunks of 64 columns (cp.async, two stages), warp (i, cs) multiplies row tile i against all
// m parents for the k-steps cs, cs + CS, ... of the chunk (DMMA m16n8k4), and the CTA adds its m x m result into Gamma
// once.  Operands are the CI vectors themselves for alpha (a row = one alpha string) and their transposed row slabs
// for beta (the same slabs and plans as the beta same-spin sigma term: SSPlan, BetaPass of sigma.cu).
struct RdmGramParams {
    const double* XL;
    const double* XR;
    double* out; // [Pops][Pops]
    const uint32_t* hptr;
    const uint64_t* hent;
    const uint32_t *s_cls, *s_loc;
    const int64_t* xoff; // [classes of the acting spin][column groups] offset in the operand, -1 if absent
    const uint32_t* g_ncols;
    const uint32_t* g_ldx;
    int ngroups, Pops;
    uint32_t kfirst; // first hole string of this launch (a slice of them: f2_sigma_rdm_part)
};

constexpr int GRAM_THREADS = 384, GRAM_CHUNK = 64, GRAM_TS = 68;
constexpr size_t rdm_gram_smem_bytes(int mt, bool same) {
    return size_t(2) * (same ? 1 : 2) * (16 * mt) * GRAM_TS * 8 + size_t(16 * mt) * (8 + 4 + 4) + 64;
}

template <int MT, bool AL16, bool SAME>
__global__ void __launch_bounds__(GRAM_THREADS, 1) rdm_gram_kernel(const RdmGramParams p) {
    constexpr int ROWS = 16 * MT, CS = (GRAM_THREADS / 32) / MT, NOP = SAME ? 1 : 2;
    constexpr int TILE = ROWS * GRAM_TS;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    double* tiles = reinterpret_cast<double*>(smem_raw);                   // [2][NOP][ROWS][TS]
    long long* src = reinterpret_cast<long long*>(tiles + 2 * NOP * TILE); // [ROWS] offset of the parent row in the group
    uint32_t* pstr = reinterpret_cast<uint32_t*>(src + ROWS);              // [ROWS] parent string
    uint32_t* pinfo = pstr + ROWS;                                         // [ROWS] operator index | sign << 31
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int g = lane >> 2, t = lane & 3;
    const uint32_t K = p.kfirst + blockIdx.x;
    const uint32_t e0 = p.hptr[K], e1 = p.hptr[K + 1];
    const int m = static_cast<int>(e1 - e0) < ROWS ? static_cast<int>(e1 - e0) : ROWS;
    if (m == 0)
        return;
    for (int r = tid; r < ROWS; r += GRAM_THREADS) {
        uint32_t I = 0, info = 0;
        if (r < m) {
            const uint64_t ent = p.hent[e0 + r];
            I = static_cast<uint32_t>(ent);
            info = static_cast<uint32_t>((ent >> 32) & 0xffff) | (static_cast<uint32_t>((ent >> 48) & 1) << 31);
        }
        pstr[r] = I;
        pinfo[r] = info;
    }
    const int wi = warp / CS, cs = warp - wi * CS; // row tile, k-step phase
    double acc[2 * MT][4];
#pragma unroll
    for (int j = 0; j < 2 * MT; ++j)
        acc[j][0] = acc[j][1] = acc[j][2] = acc[j][3] = 0.0;

    for (int grp = 0; grp < p.ngroups; ++grp) {
        const int ncols = static_cast<int>(p.g_ncols[grp]);
        const long long ldx = p.g_ldx[grp];
        __syncthreads(); // the previous group's tiles and offsets are no longer read; pstr is visible
        int any = 0;
        for (int r = tid; r < ROWS; r += GRAM_THREADS) {
            long long o = -1;
            if (r < m) {
                const uint32_t I = pstr[r];
                const long long xo = p.xoff[static_cast<size_t>(p.s_cls[I]) * p.ngroups + grp];
                if (xo >= 0) {
                    o = xo + static_cast<long long>(p.s_loc[I]) * ldx;
                    any = 1;
                }
            }
            src[r] = o;
        }
        if (!__syncthreads_or(any) || ncols == 0)
            continue;
        const int nchunks = (ncols + GRAM_CHUNK - 1) / GRAM_CHUNK;
        auto stage = [&](int chunk, int buf) {
            if (chunk < nchunks) {
                const int c0 = chunk * GRAM_CHUNK;
                constexpr int PIECE = AL16 ? 2 : 1, PPR = GRAM_CHUNK / PIECE; // doubles per piece, pieces per row
                for (int idx = tid; idx < NOP * ROWS * PPR; idx += GRAM_THREADS) {
                    const int op = idx / (ROWS * PPR), rem = idx - op * ROWS * PPR;
                    const int r = rem / PPR, pc = rem - r * PPR;
                    const int col = c0 + pc * PIECE;
                    const long long o = src[r];
                    int bytes = 0;
                    if (o >= 0 && col < ncols)
                        bytes = (AL16 && col + 1 < ncols) ? 16 : 8;
                    const double* gsrc = (op == 0 ? p.XL : p.XR) + (bytes ? o + col : 0);
                    double* dst = tiles + (buf * NOP + op) * TILE + r * GRAM_TS + pc * PIECE;
                    if (AL16)
                        cp_async16(dst, gsrc, bytes);
                    else
                        cp_async8(dst, gsrc, bytes);
                }
            }
            cp_async_commit();
        };
        stage(0, 0);
        for (int chunk = 0; chunk < nchunks; ++chunk) {
            cp_async_wait<0>();
            __syncthreads(); // chunk landed for everyone; the other buffer is no longer read
            stage(chunk + 1, (chunk + 1) & 1);
            const double* TL = tiles + ((chunk & 1) * NOP) * TILE;
            const double* TR = SAME ? TL : TL + TILE;
            const double* ap = TL + (16 * wi + g) * GRAM_TS + t;
            const double* bp = TR + g * GRAM_TS + t;
#pragma unroll 2
            for (int ks = cs; ks < GRAM_CHUNK / 4; ks += CS) {
                const double a0 = ap[4 * ks], a1 = ap[8 * GRAM_TS + 4 * ks];
#pragma unroll
                for (int j = 0; j < 2 * MT; ++j)
                    dmma_16x8x4(acc[j], a0, a1, bp[8 * j * GRAM_TS + 4 * ks]);
            }
        }
        cp_async_wait<0>();
    }
    // accumulator (tile j, element 2h + c): parent a = 16 wi + g + 8 h (of L), parent b = 8 j + 2 t + c (of R)
#pragma unroll
    for (int h = 0; h < 2; ++h) {
        const int a = 16 * wi + g + 8 * h;
        if (a >= m)
            continue;
        const uint32_t ia = pinfo[a];
        double* orow = p.out + static_cast<size_t>(ia & 0xffff) * p.Pops;
#pragma unroll
        for (int j = 0; j < 2 * MT; ++j)
#pragma unroll
            for (int c = 0; c < 2; ++c) {
                const int b = 8 * j + 2 * t + c;
                if (b < m) {
                    const uint32_t ib = pinfo[b];
                    const double v = acc[j][2 * h + c];
                    if (v != 0.0)
                        atomicAdd(orow + (ib & 0xffff), ((ia ^ ib) >> 31) ? -v : v);
                }
            }
    }
}

// ---------------------------------------------------------------------------------------------
// Opposite-spin 2-RDM for spaces with one determinant block, in the gather form of the opposite-spin sigma term
// (t3g_kernel read backwards): one CTA per alpha string Ja of the right vector,
//   D[(v y), Ib]            = <Ib| E^b_vy |Jb> R[Ja, Jb]          gathered from the row R[Ja, :] in shared memory
//   G[(v y), r]             = sum_Ib D[(v y), Ib] L[Ia_r, Ib]     Ia_r = s_r E^a_{u_r x_r} Ja, the single replacements of Ja
//   Gamma[u_r, v, x_r, y]  += s_r G[(v y), r]
// The contraction index is the beta string Ib: 2 n^2 K_a nbs flops per CTA on the FP64 tensor pipe instead of the
// 2 n^4 nbs of the D-matrix form, and nothing of size n^2 x ndet is ever written.  Warp w owns MW 16-row tiles of (v y);
// the rows L[Ia_r, chunk] are staged with cp.async (two stages), the table entries (source column and phase of every
// ((v y), Ib)) are read 16 bytes per lane.  Within a group of 16 beta strings lane t supplies the strings 4 t + s to
// k-step s (the sum over Ib does not care), which is what makes the 16-byte table reads line up with the fragments.
struct RdmAbParams {
    const double* L;
    const double* R;
    double* out; // [n][n][n][n]
    const uint32_t* as_ptr;
    const uint64_t* as_ent;
    const uint32_t* gtab; // [mrows][gld]: source column | phase << 31; nbs = no source (the zero slot)
    int64_t nbs, gld;
    int norb, n2;
    uint32_t first; // first alpha string of this launch
    int nsm;        // columns of the row of R staged in shared memory (LONG: the rest is gathered from L2)
};

constexpr int AB_THREADS = 256, AB_CHUNK = 32, AB_TS = 33;
constexpr size_t rdm_ab_smem_bytes(int64_t nbs, int ntp) {
    return size_t(nbs + 2) * 8 + size_t(2) * (8 * ntp) * AB_TS * 8 + 64;
}

// gtab from the beta single-replacement list (ELL by source string): <Ib| E_ij |Jb> = sign  ->  gtab[i n + j][Ib] = Jb | neg << 31
__global__ void rdm_abtab_kernel(const uint64_t* __restrict__ bs_ent, int64_t nbs, int width, int64_t gld, uint32_t* gtab) {
    const int64_t idx = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (idx >= nbs * width)
        return;
    const uint64_t ent = bs_ent[idx];
    if (ent == ~0ull)
        return;
    const int64_t Jb = idx % nbs;
    const uint32_t Ib = static_cast<uint32_t>(ent & 0xfffffff);
    const uint32_t ij = static_cast<uint32_t>((ent >> 40) & 0xffff);
    gtab[static_cast<int64_t>(ij) * gld + Ib] = static_cast<uint32_t>(Jb) | (static_cast<uint32_t>((ent >> 56) & 1) << 31);
}

// LONG: the row of R is longer than shared memory (CAS(18,18): 389 KB): nsm columns are staged, the others are gathered
// from global memory, where the row stays L2-resident while its CTA runs (as t3g_kernel does)
template <int MW, int NTP, bool LONG>
__global__ void __launch_bounds__(AB_THREADS, 1) rdm_ab_kernel(const RdmAbParams p) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int g = lane >> 2, t = lane & 3;
    const int64_t nbs = p.nbs;
    const int nsm = LONG ? p.nsm : static_cast<int>(nbs);
    double* row = reinterpret_cast<double*>(smem_raw); // [nsm + 2], the last two are zero
    double* Ls = row + nsm + 2;                        // [2][8 NTP][AB_TS]
    __shared__ long long s_lrow[128];                  // offset of row Ia_r in L, -1 past the list
    __shared__ int s_ux[128];                          // (u n + x) | phase << 30, -1 past the list
    const uint32_t Ja = p.first + blockIdx.x;
    const double* rsrc = p.R + static_cast<int64_t>(Ja) * nbs;
    for (int i = tid; i < nsm + 2; i += AB_THREADS)
        row[i] = i < nsm ? __ldg(rsrc + i) : 0.0;
    const int e0 = p.as_ptr[Ja], e1 = p.as_ptr[Ja + 1];
    const int Kn = e1 - e0 < 128 ? e1 - e0 : 128;
    for (int r = tid; r < 128; r += AB_THREADS) {
        if (r < Kn) {
            // <Ja| E_kl |Ia_r> = s  <=>  Ia_r = s E_lk Ja: creator u = l, annihilator x = k
            const uint64_t ent = p.as_ent[e0 + r];
            const int kl = static_cast<int>((ent >> 32) & 0xffff);
            const int k = kl / p.norb, l = kl - k * p.norb;
            s_ux[r] = (l * p.norb + k) | (static_cast<int>((ent >> 48) & 1) << 30);
            s_lrow[r] = static_cast<long long>(static_cast<uint32_t>(ent)) * nbs;
        } else {
            s_ux[r] = -1;
            s_lrow[r] = -1;
        }
    }
    __syncthreads();
    const int nchunks = static_cast<int>((nbs + AB_CHUNK - 1) / AB_CHUNK);
    const int mbase = warp * MW * 16;
    for (int r0 = 0; r0 < Kn; r0 += 8 * NTP) { // pass over the output rows
        double acc[MW][NTP][4];
#pragma unroll
        for (int i = 0; i < MW; ++i)
#pragma unroll
            for (int j = 0; j < NTP; ++j)
                acc[i][j][0] = acc[i][j][1] = acc[i][j][2] = acc[i][j][3] = 0.0;
        auto stage = [&](int chunk, int buf) {
            if (chunk < nchunks) {
                const int64_t c0 = static_cast<int64_t>(chunk) * AB_CHUNK;
                for (int idx = tid; idx < 8 * NTP * AB_CHUNK; idx += AB_THREADS) {
                    const int r = idx / AB_CHUNK, c = idx - r * AB_CHUNK;
                    const long long o = r0 + r < 128 ? s_lrow[r0 + r] : -1;
                    const bool in = o >= 0 && c0 + c < nbs;
                    cp_async8(Ls + (buf * 8 * NTP + r) * AB_TS + c, p.L + (in ? o + c0 + c : 0), in ? 8 : 0);
                }
            }
            cp_async_commit();
        };
        __syncthreads(); // the previous pass no longer reads the stages
        stage(0, 0);
        for (int chunk = 0; chunk < nchunks; ++chunk) {
            cp_async_wait<0>();
            __syncthreads();
            stage(chunk + 1, (chunk + 1) & 1);
            const double* Lb = Ls + ((chunk & 1) * 8 * NTP) * AB_TS;
#pragma unroll
            for (int half = 0; half < AB_CHUNK / 16; ++half) {
                const int64_t ib0 = static_cast<int64_t>(chunk) * AB_CHUNK + 16 * half;
                // table entries of this lane's four beta strings ib0 + 4 t + {0..3}, rows g and g + 8 of each tile
                uint4 q[MW][2];
#pragma unroll
                for (int i = 0; i < MW; ++i)
#pragma unroll
                    for (int h = 0; h < 2; ++h)
                        q[i][h] = __ldg(reinterpret_cast<const uint4*>(p.gtab + static_cast<int64_t>(mbase + 16 * i + g + 8 * h) * p.gld +
                                                                       ib0 + 4 * t));
                auto gather = [&](uint32_t e) {
                    const uint32_t idx = e & 0x7fffffffu; // nbs = no source
                    double v;
                    if constexpr (LONG) {
                        v = row[idx < static_cast<uint32_t>(nsm) ? idx : static_cast<uint32_t>(nsm)];
                        if (idx >= static_cast<uint32_t>(nsm) && idx < static_cast<uint32_t>(nbs))
                            v = __ldg(rsrc + idx);
                    } else {
                        v = row[idx];
                    }
                    return __hiloint2double(__double2hiint(v) ^ static_cast<int>(e & 0x80000000u), __double2loint(v));
                };
#pragma unroll
                for (int s4 = 0; s4 < 4; ++s4) {
                    double b[NTP];
#pragma unroll
                    for (int j = 0; j < NTP; ++j)
                        b[j] = Lb[(8 * j + g) * AB_TS + 16 * half + 4 * t + s4];
#pragma unroll
                    for (int i = 0; i < MW; ++i) {
                        const uint32_t ea = s4 == 0 ? q[i][0].x : s4 == 1 ? q[i][0].y : s4 == 2 ? q[i][0].z : q[i][0].w;
                        const uint32_t eb = s4 == 0 ? q[i][1].x : s4 == 1 ? q[i][1].y : s4 == 2 ? q[i][1].z : q[i][1].w;
                        const double a0 = gather(ea), a1 = gather(eb);
#pragma unroll
                        for (int j = 0; j < NTP; ++j)
                            dmma_16x8x4(acc[i][j], a0, a1, b[j]);
                    }
                }
            }
        }
        cp_async_wait<0>();
        // accumulator (tile i, j, element 2h + c): (v y) = mbase + 16 i + g + 8 h, output row r = r0 + 8 j + 2 t + c
        const int n = p.norb;
#pragma unroll
        for (int i = 0; i < MW; ++i)
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const int mrow = mbase + 16 * i + g + 8 * h;
                if (mrow >= p.n2)
                    continue;
                const int v = mrow / n, y = mrow - v * n;
#pragma unroll
                for (int j = 0; j < NTP; ++j)
#pragma unroll
                    for (int c = 0; c < 2; ++c) {
                        const int r = r0 + 8 * j + 2 * t + c;
                        const int ux = r < 128 ? s_ux[r] : -1;
                        if (ux < 0)
                            continue;
                        const int u = (ux & 0xffff) / n, x = (ux & 0xffff) - u * n;
                        const double val = acc[i][j][2 * h + c];
                        if (val != 0.0)
                            atomicAdd(p.out + ((static_cast<int64_t>(u) * n + v) * n + x) * n + y, (ux & (1 << 30)) ? -val : val);
                    }
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
