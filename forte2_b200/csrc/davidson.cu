// Device kernels of the Davidson-Liu eigensolver that drives the sigma builder, and their C ABI
// (include/forte2_b200.h, f2_dl_*).  Reference counterpart: forte2/helpers/davidsonliu.py:184-484
// (numpy, vectors as strided columns of (size, max_subspace) arrays).  Here every vector is a
// contiguous row of a (max_subspace, ld) device array and every step of the iteration is one
// streaming pass over the vectors involved:
//
//   dl_dots_kernel      out[i, j] = <A_i | X_j>                       subspace matrix, overlaps
//   dl_combine_kernel   Y_j = beta Y_j + sum_i coef[i, j] A_i         collapse, Ritz vectors,
//                                                                     Gram-Schmidt projection (+ norm)
//   dl_residual_kernel  R_j = (S - lam_j B) alpha_j, |R_j|^2, then the Davidson preconditioner
//                       R_j / (lam_j - Hdiag) with the small-denominator mask (davidsonliu.py:266-283)
//   dl_scale_kernel     v *= s
//
// All of them are HBM-bound (a few flops per 8 bytes); reductions are two-stage (per-CTA partials
// in a caller-provided workspace, then one CTA sums them in a fixed order) so results do not depend
// on scheduling.
#include <cstdint>
#include <cstring>
#include <stdexcept>
#include <string>

#include <cuda_runtime.h>

#include "handles.hpp"
#include "nvtx_ranges.hpp"

namespace f2 {

constexpr int DL_MAXV = 16;      // vectors per side of one call
constexpr int DL_THREADS = 256;
constexpr int DL_CTAS = 148 * 4; // grid of every streaming pass
constexpr int DL_WS = DL_CTAS * DL_MAXV * DL_MAXV; // doubles of workspace a caller must provide

struct DlCoef {
    double c[DL_MAXV * DL_MAXV];
};
struct DlLam {
    double c[DL_MAXV];
};

// sum `nvals` per-thread values over the CTA and leave them in part[blockIdx.x * nvals + v]
template <int NV>
__device__ __forceinline__ void cta_partials(double (&acc)[NV], int nvals, double* __restrict__ part) {
    __shared__ double red[DL_THREADS / 32][NV];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int v = 0; v < NV; ++v) {
        double x = acc[v];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1)
            x += __shfl_xor_sync(0xffffffffu, x, o);
        if (lane == 0)
            red[warp][v] = x;
    }
    __syncthreads();
    if (threadIdx.x < nvals) {
        double x = 0.0;
#pragma unroll
        for (int w = 0; w < DL_THREADS / 32; ++w)
            x += red[w][threadIdx.x % NV];
        part[static_cast<size_t>(blockIdx.x) * nvals + threadIdx.x] = x;
    }
}

// out[v] = sum over CTAs of part[cta * nvals + v], fixed order
__global__ void dl_finish_kernel(const double* __restrict__ part, int nctas, int nvals,
                                 double* __restrict__ out) {
    const int v = blockIdx.x * blockDim.x + threadIdx.x;
    if (v >= nvals)
        return;
    double x = 0.0;
    for (int c = 0; c < nctas; ++c)
        x += part[static_cast<size_t>(c) * nvals + v];
    out[v] = x;
}

constexpr int DL_TILE = 4096; // elements a CTA reads contiguously from every row before moving on

template <int PT, int QT>
__global__ void __launch_bounds__(DL_THREADS)
    dl_dots_kernel(const double* __restrict__ A, int64_t lda, int p, const double* __restrict__ X,
                   int64_t ldx, int q, int64_t n, double* __restrict__ part) {
    double acc[PT * QT];
#pragma unroll
    for (int i = 0; i < PT * QT; ++i)
        acc[i] = 0.0;
    const int64_t ntiles = (n + DL_TILE - 1) / DL_TILE;
    for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const int64_t t0 = tile * DL_TILE;
        const int64_t t1 = t0 + DL_TILE < n ? t0 + DL_TILE : n;
        for (int64_t e = t0 + threadIdx.x; e < t1; e += DL_THREADS) {
            double x[QT], a[PT];
#pragma unroll
            for (int j = 0; j < QT; ++j)
                x[j] = j < q ? __ldg(X + j * ldx + e) : 0.0;
#pragma unroll
            for (int i = 0; i < PT; ++i)
                a[i] = i < p ? __ldg(A + i * lda + e) : 0.0;
#pragma unroll
            for (int i = 0; i < PT; ++i)
#pragma unroll
                for (int j = 0; j < QT; ++j)
                    acc[i * QT + j] = fma(a[i], x[j], acc[i * QT + j]);
        }
    }
    __shared__ double red[DL_THREADS / 32][PT * QT];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int v = 0; v < PT * QT; ++v) {
        double y = acc[v];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1)
            y += __shfl_xor_sync(0xffffffffu, y, o);
        if (lane == 0)
            red[warp][v] = y;
    }
    __syncthreads();
    for (int v = threadIdx.x; v < p * q; v += DL_THREADS) {
        const int i = v / q, j = v - i * q;
        double y = 0.0;
#pragma unroll
        for (int w = 0; w < DL_THREADS / 32; ++w)
            y += red[w][i * QT + j];
        part[static_cast<size_t>(blockIdx.x) * (p * q) + v] = y;
    }
}

// Y_j = beta * Y_j + sum_i coef[i, j] * A_i for j < k.  Safe when rows of Y alias rows of A:
// a thread reads every A_i[e] before it writes any Y_j[e].  WANT_SS: per-CTA partials of |Y_j|^2.
// coef.c[i * DL_MAXV + j]: compile-time strides, so that the unrolled FMAs take their coefficient straight from
// the constant bank; PT = p rounded up to a multiple of 4 (zero coefficients, no loads, for the padding rows).
template <int KT, int PT, bool WANT_SS>
__global__ void __launch_bounds__(DL_THREADS)
    dl_combine_kernel(const double* A, int64_t lda, int p, const DlCoef coef, double* Y, int64_t ldy,
                      int k, double beta, int64_t n, double* __restrict__ part) {
    double ss[KT];
#pragma unroll
    for (int j = 0; j < KT; ++j)
        ss[j] = 0.0;
    const int64_t ntiles = (n + DL_TILE - 1) / DL_TILE;
    for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const int64_t t0 = tile * DL_TILE;
        const int64_t t1 = t0 + DL_TILE < n ? t0 + DL_TILE : n;
        for (int64_t e = t0 + threadIdx.x; e < t1; e += DL_THREADS) {
            double y[KT], a[PT > 0 ? PT : 1];
#pragma unroll
            for (int i = 0; i < PT; ++i)
                a[i] = i < p ? A[i * lda + e] : 0.0;
#pragma unroll
            for (int j = 0; j < KT; ++j)
                y[j] = (beta != 0.0 && j < k) ? beta * Y[j * ldy + e] : 0.0;
#pragma unroll
            for (int i = 0; i < PT; ++i)
#pragma unroll
                for (int j = 0; j < KT; ++j)
                    y[j] = fma(coef.c[i * DL_MAXV + j], a[i], y[j]);
#pragma unroll
            for (int j = 0; j < KT; ++j)
                if (j < k) {
                    Y[j * ldy + e] = y[j];
                    if (WANT_SS)
                        ss[j] = fma(y[j], y[j], ss[j]);
                }
        }
    }
    if (WANT_SS)
        cta_partials<KT>(ss, k, part);
}

// R_j = sum_i alpha[i*k+j] (S_i - lam_j B_i); partials of |R_j|^2; then R_j <- R_j / (lam_j - h) where
// |lam_j - h| > thresh, 0 elsewhere (davidsonliu.py:266-283).  hdiag == nullptr: no preconditioning.
template <int KT>
__global__ void __launch_bounds__(DL_THREADS)
    dl_residual_kernel(const double* __restrict__ B, const double* __restrict__ S, int64_t ld, int m,
                       const DlCoef alpha, const DlLam lam, const double* __restrict__ hdiag,
                       double* __restrict__ R, int64_t ldr, int k, double thresh, int64_t n,
                       double* __restrict__ part) {
    double ss[KT];
#pragma unroll
    for (int j = 0; j < KT; ++j)
        ss[j] = 0.0;
    const int64_t stride = static_cast<int64_t>(gridDim.x) * DL_THREADS;
    for (int64_t e = static_cast<int64_t>(blockIdx.x) * DL_THREADS + threadIdx.x; e < n; e += stride) {
        double sb[KT], sv[KT];
#pragma unroll
        for (int j = 0; j < KT; ++j)
            sb[j] = sv[j] = 0.0;
        for (int i = 0; i < m; ++i) {
            const double b = __ldg(B + i * ld + e), s = __ldg(S + i * ld + e);
#pragma unroll
            for (int j = 0; j < KT; ++j) {
                const double a = alpha.c[i * k + (j < k ? j : 0)];
                sb[j] = fma(a, b, sb[j]);
                sv[j] = fma(a, s, sv[j]);
            }
        }
        const double h = hdiag ? __ldg(hdiag + e) : 0.0;
#pragma unroll
        for (int j = 0; j < KT; ++j)
            if (j < k) {
                double r = sv[j] - lam.c[j] * sb[j];
                ss[j] = fma(r, r, ss[j]);
                if (hdiag) {
                    const double d = lam.c[j] - h;
                    r = fabs(d) > thresh ? r / d : 0.0;
                }
                R[j * ldr + e] = r;
            }
    }
    cta_partials<KT>(ss, k, part);
}

__global__ void __launch_bounds__(DL_THREADS) dl_scale_kernel(double* __restrict__ v, double s, int64_t n) {
    const int64_t stride = static_cast<int64_t>(gridDim.x) * DL_THREADS;
    for (int64_t e = static_cast<int64_t>(blockIdx.x) * DL_THREADS + threadIdx.x; e < n; e += stride)
        v[e] *= s;
}

// ---------------------------------------------------------------------------------------------
// Fused passes of one Davidson-Liu iteration (at most 4 roots, 16 basis vectors).  The candidate block is
// orthonormalised as a block instead of vector by vector (davidsonliu.py:415-484 does modified Gram-Schmidt
// with one pass over the basis per candidate and cycle): every pass below reads the basis once, writes the k
// candidates once and returns, from the same registers, the overlaps the NEXT pass needs.
//
//   RESID:  z_j = precondition( sum_i alpha[i,j] (S_i - lam_j B_i) )      residual kernel
//   else :  z_j = sum_l T[l,j] ( Y_l - sum_i c[i,l] B_i )                 projection + in-block transform
//   both :  out = [ |r_j|^2 before preconditioning (RESID only) ; d[i,j] = <B_i|z_j> ; g[j,j'] = <z_j|z_j'> ]
//
// HBM-bound: 8 (2m + k) or 8 (m + 2k) bytes per element for a few hundred flops.
constexpr int DLF_MAXK = 4;
constexpr int DLF_THREADS = 128;
constexpr int DLF_CTAS = 148 * 4;

struct DlfCoef {
    double c[DL_MAXV * DLF_MAXK]; // [i * k + j]
};
struct DlfSmall {
    double t[DLF_MAXK * DLF_MAXK];
};

// Coefficients sit in the kernel parameters with COMPILE-TIME strides (coef.c[i * DLF_MAXK + j], T.t[l * DLF_MAXK + j],
// zero outside the live m x k part): after unrolling every coefficient is a constant-bank operand of its DFMA.  (The
// first version indexed them with the runtime k: an address computation and a load per FMA, 666 instructions per
// element and issue-bound at 2.2 TB/s, profiles/r2c_dl_fused.md.)  A CTA walks contiguous tiles of DLF_TILE
// elements, so that each of the up to 28 row streams it reads stays inside one DRAM page for a while.
constexpr int DLF_TILE = 2048;

template <int KT, int MT, bool RESID>
__global__ void __launch_bounds__(DLF_THREADS)
    dl_fused_kernel(const double* __restrict__ B, const double* __restrict__ S, int64_t ld, int m, const double* Yin,
                    int64_t ldin, int kin, const DlfCoef coef, const DlfSmall T, const double* __restrict__ hdiag,
                    double thresh, double* Zout, int64_t ldout, int kout, int64_t n, double* __restrict__ part) {
    constexpr int NV = (RESID ? KT : 0) + MT * KT + KT * KT;
    constexpr int D0 = RESID ? KT : 0;
    double acc[NV];
#pragma unroll
    for (int v = 0; v < NV; ++v)
        acc[v] = 0.0;
    const int64_t ntiles = (n + DLF_TILE - 1) / DLF_TILE;
    for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const int64_t t0 = tile * DLF_TILE;
        const int64_t t1 = t0 + DLF_TILE < n ? t0 + DLF_TILE : n;
        for (int64_t e = t0 + threadIdx.x; e < t1; e += DLF_THREADS) {
            double b[MT], z[KT];
#pragma unroll
            for (int i = 0; i < MT; ++i)
                b[i] = i < m ? __ldg(B + i * ld + e) : 0.0;
            if constexpr (RESID) {
                double sb[KT], sv[KT];
#pragma unroll
                for (int j = 0; j < KT; ++j)
                    sb[j] = sv[j] = 0.0;
#pragma unroll
                for (int i = 0; i < MT; ++i) {
                    const double s = i < m ? __ldg(S + i * ld + e) : 0.0;
#pragma unroll
                    for (int j = 0; j < KT; ++j) {
                        sb[j] = fma(coef.c[i * DLF_MAXK + j], b[i], sb[j]);
                        sv[j] = fma(coef.c[i * DLF_MAXK + j], s, sv[j]);
                    }
                }
                const double h = hdiag ? __ldg(hdiag + e) : 0.0;
#pragma unroll
                for (int j = 0; j < KT; ++j) {
                    double r = sv[j] - T.t[j] * sb[j]; // T.t[j] = lam_j; zero coefficients give r = 0 for j >= k
                    acc[j] = fma(r, r, acc[j]);
                    if (hdiag) {
                        const double d = T.t[j] - h;
                        r = fabs(d) > thresh ? r / d : 0.0;
                    }
                    z[j] = r;
                }
            } else {
                double y[KT];
#pragma unroll
                for (int l = 0; l < KT; ++l)
                    y[l] = l < kin ? Yin[l * ldin + e] : 0.0;
#pragma unroll
                for (int i = 0; i < MT; ++i)
#pragma unroll
                    for (int l = 0; l < KT; ++l)
                        y[l] = fma(-coef.c[i * DLF_MAXK + l], b[i], y[l]);
#pragma unroll
                for (int j = 0; j < KT; ++j) {
                    double x = 0.0;
#pragma unroll
                    for (int l = 0; l < KT; ++l)
                        x = fma(T.t[l * DLF_MAXK + j], y[l], x); // zero outside kin x kout
                    z[j] = x;
                }
            }
#pragma unroll
            for (int j = 0; j < KT; ++j)
                if (j < kout)
                    Zout[j * ldout + e] = z[j];
#pragma unroll
            for (int i = 0; i < MT; ++i)
#pragma unroll
                for (int j = 0; j < KT; ++j)
                    acc[D0 + i * KT + j] = fma(b[i], z[j], acc[D0 + i * KT + j]);
#pragma unroll
            for (int j = 0; j < KT; ++j)
#pragma unroll
                for (int jj = j; jj < KT; ++jj) // upper triangle; the finish kernel mirrors it
                    acc[D0 + MT * KT + j * KT + jj] = fma(z[j], z[jj], acc[D0 + MT * KT + j * KT + jj]);
        }
    }
    __shared__ double red[DLF_THREADS / 32][NV];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int v = 0; v < NV; ++v) {
        double x = acc[v];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1)
            x += __shfl_xor_sync(0xffffffffu, x, o);
        if (lane == 0)
            red[warp][v] = x;
    }
    __syncthreads();
    for (int v = threadIdx.x; v < NV; v += DLF_THREADS) {
        double x = 0.0;
#pragma unroll
        for (int w = 0; w < DLF_THREADS / 32; ++w)
            x += red[w][v];
        part[static_cast<size_t>(blockIdx.x) * NV + v] = x;
    }
}

// out[(rn2: k) ; d: m x kout ; g: kout x kout] from the padded per-CTA partials [ (KT) ; MT x KT ; KT x KT ]
__global__ void dl_fused_finish_kernel(const double* __restrict__ part, int nctas, int nv, int KT, int MT, bool resid,
                                       int m, int kout, double* __restrict__ out) {
    const int o = blockIdx.x * blockDim.x + threadIdx.x;
    const int n_r = resid ? kout : 0;
    const int total = n_r + m * kout + kout * kout;
    if (o >= total)
        return;
    int src;
    if (o < n_r)
        src = o;
    else if (o < n_r + m * kout) {
        const int q = o - n_r;
        src = (resid ? KT : 0) + (q / kout) * KT + q % kout;
    } else {
        const int q = o - n_r - m * kout;
        const int a = q / kout, b = q % kout;
        src = (resid ? KT : 0) + MT * KT + (a < b ? a : b) * KT + (a < b ? b : a);
    }
    double x = 0.0;
    for (int c = 0; c < nctas; ++c)
        x += part[static_cast<size_t>(c) * nv + src];
    out[o] = x;
}

namespace {

void check_launch() {
    const cudaError_t err = cudaGetLastError();
    if (err != cudaSuccess)
        throw CudaError(std::string("CUDA error: ") + cudaGetErrorString(err));
}

int grid_for(int64_t n) {
    const int64_t need = (n + DL_THREADS - 1) / DL_THREADS;
    return static_cast<int>(need < 1 ? 1 : (need < DL_CTAS ? need : DL_CTAS));
}

template <int PT>
void dots_q(int q, int grid, cudaStream_t st, const double* A, int64_t lda, int p, const double* X,
            int64_t ldx, int64_t n, double* part) {
    switch (q) {
    case 1: dl_dots_kernel<PT, 1><<<grid, DL_THREADS, 0, st>>>(A, lda, p, X, ldx, q, n, part); break;
    case 2: dl_dots_kernel<PT, 2><<<grid, DL_THREADS, 0, st>>>(A, lda, p, X, ldx, q, n, part); break;
    case 3: dl_dots_kernel<PT, 3><<<grid, DL_THREADS, 0, st>>>(A, lda, p, X, ldx, q, n, part); break;
    default: dl_dots_kernel<PT, 4><<<grid, DL_THREADS, 0, st>>>(A, lda, p, X, ldx, q, n, part); break;
    }
}

} // namespace
} // namespace f2

using namespace f2;

extern "C" {

int64_t f2_dl_workspace_doubles(void) { return DL_WS; }

int f2_dl_dots(const double* A, int64_t lda, int p, const double* X, int64_t ldx, int q, int64_t n,
               double* out_dev, double* work_dev, void* stream) {
    try {
        f2::NvtxRange nvtx("davidson: subspace dots");
        if (p < 1 || p > DL_MAXV || q < 1 || q > DL_MAXV || n < 0)
            throw std::invalid_argument("f2_dl_dots: need 1 <= p, q <= 16 and n >= 0");
        cudaStream_t st = static_cast<cudaStream_t>(stream);
        const int64_t tiles = (n + DL_TILE - 1) / DL_TILE;
        const int grid = static_cast<int>(tiles < 1 ? 1 : (tiles < DL_CTAS ? tiles : DL_CTAS));
        for (int j0 = 0; j0 < q; j0 += 4) {
            const int qc = q - j0 < 4 ? q - j0 : 4;
            // partials of this column chunk: [grid][p*qc]; finished into a p x qc scratch, then placed
            double* part = work_dev;
            if (p <= 4)
                dots_q<4>(qc, grid, st, A, lda, p, X + j0 * ldx, ldx, n, part);
            else if (p <= 8)
                dots_q<8>(qc, grid, st, A, lda, p, X + j0 * ldx, ldx, n, part);
            else if (p <= 12)
                dots_q<12>(qc, grid, st, A, lda, p, X + j0 * ldx, ldx, n, part);
            else
                dots_q<16>(qc, grid, st, A, lda, p, X + j0 * ldx, ldx, n, part);
            check_launch();
            double* fin = work_dev + static_cast<size_t>(grid) * p * qc;
            dl_finish_kernel<<<1, 64, 0, st>>>(part, grid, p * qc, fin);
            check_launch();
            // out[i*q + j0 + j] = fin[i*qc + j]
            if (cudaMemcpy2DAsync(out_dev + j0, static_cast<size_t>(q) * sizeof(double), fin,
                                  static_cast<size_t>(qc) * sizeof(double), static_cast<size_t>(qc) * sizeof(double),
                                  p, cudaMemcpyDeviceToDevice, st) != cudaSuccess)
                throw CudaError("f2_dl_dots: cudaMemcpy2DAsync failed");
        }
        return F2_OK;
    } catch (const std::exception& e) {
        return status_of(e);
    }
}

int f2_dl_combine(const double* A, int64_t lda, int p, const double* coef_host, double* Y, int64_t ldy,
                  int k, double beta, int64_t n, double* sumsq_dev, double* work_dev, void* stream) {
    try {
        f2::NvtxRange nvtx("davidson: combine");
        if (p < 0 || p > DL_MAXV || k < 1 || k > DL_MAXV || n < 0)
            throw std::invalid_argument("f2_dl_combine: need 0 <= p <= 16, 1 <= k <= 16 and n >= 0");
        cudaStream_t st = static_cast<cudaStream_t>(stream);
        DlCoef c;
        std::memset(&c, 0, sizeof(c));
        for (int i = 0; i < p; ++i)
            for (int jj = 0; jj < k; ++jj)
                c.c[i * DL_MAXV + jj] = coef_host[i * k + jj];
        const int64_t tiles = (n + DL_TILE - 1) / DL_TILE;
        const int grid = static_cast<int>(tiles < 1 ? 1 : (tiles < DL_CTAS ? tiles : DL_CTAS));
        const bool ss = sumsq_dev != nullptr;
        const int pt = (p + 3) / 4 * 4;
#define F2_COMBINE_P(KT, PT)                                                                                \
    do {                                                                                                    \
        if (ss)                                                                                             \
            dl_combine_kernel<KT, PT, true><<<grid, DL_THREADS, 0, st>>>(A, lda, p, c, Y, ldy, k, beta, n, work_dev); \
        else                                                                                                \
            dl_combine_kernel<KT, PT, false><<<grid, DL_THREADS, 0, st>>>(A, lda, p, c, Y, ldy, k, beta, n, work_dev); \
    } while (0)
#define F2_COMBINE(KT)                                                                                      \
    do {                                                                                                    \
        if (pt == 0)                                                                                        \
            F2_COMBINE_P(KT, 0);                                                                            \
        else if (pt == 4)                                                                                   \
            F2_COMBINE_P(KT, 4);                                                                            \
        else if (pt == 8)                                                                                   \
            F2_COMBINE_P(KT, 8);                                                                            \
        else if (pt == 12)                                                                                  \
            F2_COMBINE_P(KT, 12);                                                                           \
        else                                                                                                \
            F2_COMBINE_P(KT, 16);                                                                           \
    } while (0)
        if (k <= 1)
            F2_COMBINE(1);
        else if (k <= 2)
            F2_COMBINE(2);
        else if (k <= 4)
            F2_COMBINE(4);
        else if (k <= 8)
            F2_COMBINE(8);
        else
            F2_COMBINE(16);
#undef F2_COMBINE
#undef F2_COMBINE_P
        check_launch();
        if (ss) {
            dl_finish_kernel<<<1, 64, 0, st>>>(work_dev, grid, k, sumsq_dev);
            check_launch();
        }
        return F2_OK;
    } catch (const std::exception& e) {
        return status_of(e);
    }
}

int f2_dl_residual(const double* B, const double* S, int64_t ld, int m, const double* alpha_host,
                   const double* lam_host, const double* hdiag, double* R, int64_t ldr, int k,
                   double denom_threshold, int64_t n, double* rnorm2_dev, double* work_dev, void* stream) {
    try {
        f2::NvtxRange nvtx("davidson: residual");
        if (m < 1 || m > DL_MAXV || k < 1 || k > DL_MAXV || n < 0)
            throw std::invalid_argument("f2_dl_residual: need 1 <= m, k <= 16 and n >= 0");
        cudaStream_t st = static_cast<cudaStream_t>(stream);
        DlCoef a;
        DlLam l;
        std::memset(&a, 0, sizeof(a));
        std::memset(&l, 0, sizeof(l));
        std::memcpy(a.c, alpha_host, sizeof(double) * m * k);
        std::memcpy(l.c, lam_host, sizeof(double) * k);
        const int grid = grid_for(n);
#define F2_RESID(KT)                                                                                       \
    dl_residual_kernel<KT><<<grid, DL_THREADS, 0, st>>>(B, S, ld, m, a, l, hdiag, R, ldr, k, denom_threshold, n, work_dev)
        if (k <= 1)
            F2_RESID(1);
        else if (k <= 2)
            F2_RESID(2);
        else if (k <= 4)
            F2_RESID(4);
        else if (k <= 8)
            F2_RESID(8);
        else
            F2_RESID(16);
#undef F2_RESID
        check_launch();
        dl_finish_kernel<<<1, 64, 0, st>>>(work_dev, grid, k, rnorm2_dev);
        check_launch();
        return F2_OK;
    } catch (const std::exception& e) {
        return status_of(e);
    }
}

int f2_dl_scale(double* v, double s, int64_t n, void* stream) {
    try {
        dl_scale_kernel<<<grid_for(n), DL_THREADS, 0, static_cast<cudaStream_t>(stream)>>>(v, s, n);
        check_launch();
        return F2_OK;
    } catch (const std::exception& e) {
        return status_of(e);
    }
}

} // extern "C"

namespace {
template <int KT, int MT, bool RESID>
void launch_fused(int grid, cudaStream_t st, const double* B, const double* S, int64_t ld, int m, const double* Yin,
                  int64_t ldin, int kin, const DlfCoef& coef, const DlfSmall& T, const double* hdiag, double thresh,
                  double* Zout, int64_t ldout, int kout, int64_t n, double* work, double* out) {
    constexpr int NV = (RESID ? KT : 0) + MT * KT + KT * KT;
    dl_fused_kernel<KT, MT, RESID><<<grid, DLF_THREADS, 0, st>>>(B, S, ld, m, Yin, ldin, kin, coef, T, hdiag, thresh,
                                                                 Zout, ldout, kout, n, work);
    check_launch();
    const int total = (RESID ? kout : 0) + m * kout + kout * kout;
    dl_fused_finish_kernel<<<(total + 63) / 64, 64, 0, st>>>(work, grid, NV, KT, MT, RESID, m, kout, out);
    check_launch();
}

template <bool RESID>
void dispatch_fused(int grid, cudaStream_t st, const double* B, const double* S, int64_t ld, int m, const double* Yin,
                    int64_t ldin, int kin, const DlfCoef& coef, const DlfSmall& T, const double* hdiag, double thresh,
                    double* Zout, int64_t ldout, int kout, int64_t n, double* work, double* out) {
    const int kt = kin > kout ? kin : kout;
    const int mt = m <= 4 ? 4 : (m <= 8 ? 8 : (m <= 12 ? 12 : 16));
#define F2_FUSED(KT, MT)                                                                                            \
    launch_fused<KT, MT, RESID>(grid, st, B, S, ld, m, Yin, ldin, kin, coef, T, hdiag, thresh, Zout, ldout, kout, n, \
                                work, out)
#define F2_FUSED_M(KT)                                                                                              \
    do {                                                                                                            \
        if (mt == 4)                                                                                                \
            F2_FUSED(KT, 4);                                                                                        \
        else if (mt == 8)                                                                                           \
            F2_FUSED(KT, 8);                                                                                        \
        else if (mt == 12)                                                                                          \
            F2_FUSED(KT, 12);                                                                                       \
        else                                                                                                        \
            F2_FUSED(KT, 16);                                                                                       \
    } while (0)
    if (kt == 1)
        F2_FUSED_M(1);
    else if (kt == 2)
        F2_FUSED_M(2);
    else if (kt == 3)
        F2_FUSED_M(3);
    else
        F2_FUSED_M(4);
#undef F2_FUSED_M
#undef F2_FUSED
}

int fused_grid(int64_t n) {
    const int64_t need = (n + DLF_TILE - 1) / DLF_TILE;
    return static_cast<int>(need < 1 ? 1 : (need < DLF_CTAS ? need : DLF_CTAS));
}
} // namespace

extern "C" {

int f2_dl_residual_fused(const double* B, const double* S, int64_t ld, int m, const double* alpha_host,
                         const double* lam_host, const double* hdiag, double* R, int64_t ldr, int k,
                         double denom_threshold, int64_t n, double* out_dev, double* work_dev, void* stream) {
    try {
        f2::NvtxRange nvtx("davidson: residual + preconditioner + projection");
        if (m < 1 || m > DL_MAXV || k < 1 || k > DLF_MAXK || n < 0)
            throw std::invalid_argument("f2_dl_residual_fused: need 1 <= m <= 16, 1 <= k <= 4 and n >= 0");
        DlfCoef a;
        DlfSmall l;
        std::memset(&a, 0, sizeof(a));
        std::memset(&l, 0, sizeof(l));
        for (int i = 0; i < m; ++i)
            for (int j = 0; j < k; ++j)
                a.c[i * DLF_MAXK + j] = alpha_host[i * k + j];
        std::memcpy(l.t, lam_host, sizeof(double) * k);
        dispatch_fused<true>(fused_grid(n), static_cast<cudaStream_t>(stream), B, S, ld, m, nullptr, 0, k, a, l, hdiag,
                             denom_threshold, R, ldr, k, n, work_dev, out_dev);
        return F2_OK;
    } catch (const std::exception& e) {
        return status_of(e);
    }
}

int f2_dl_project(const double* B, int64_t ld, int m, const double* Yin, int64_t ldin, int kin, const double* c_host,
                  const double* T_host, double* Zout, int64_t ldout, int kout, int64_t n, double* out_dev,
                  double* work_dev, void* stream) {
    try {
        f2::NvtxRange nvtx("davidson: block Gram-Schmidt pass");
        if (m < 0 || m > DL_MAXV || kin < 1 || kin > DLF_MAXK || kout < 1 || kout > DLF_MAXK || n < 0)
            throw std::invalid_argument("f2_dl_project: need 0 <= m <= 16, 1 <= kin, kout <= 4 and n >= 0");
        DlfCoef c;
        DlfSmall T;
        std::memset(&c, 0, sizeof(c));
        std::memset(&T, 0, sizeof(T));
        for (int i = 0; i < m; ++i)
            for (int l = 0; l < kin; ++l)
                c.c[i * DLF_MAXK + l] = c_host[i * kin + l];
        for (int l = 0; l < kin; ++l)
            for (int j = 0; j < kout; ++j)
                T.t[l * DLF_MAXK + j] = T_host[l * kout + j];
        dispatch_fused<false>(fused_grid(n), static_cast<cudaStream_t>(stream), B, nullptr, ld, m, Yin, ldin, kin, c, T,
                              nullptr, 0.0, Zout, ldout, kout, n, work_dev, out_dev);
        return F2_OK;
    } catch (const std::exception& e) {
        return status_of(e);
    }
}

} // extern "C"
