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

template <int PT, int QT>
__global__ void __launch_bounds__(DL_THREADS)
    dl_dots_kernel(const double* __restrict__ A, int64_t lda, int p, const double* __restrict__ X,
                   int64_t ldx, int q, int64_t n, double* __restrict__ part) {
    double acc[PT * QT];
#pragma unroll
    for (int i = 0; i < PT * QT; ++i)
        acc[i] = 0.0;
    const int64_t stride = static_cast<int64_t>(gridDim.x) * DL_THREADS;
    for (int64_t e = static_cast<int64_t>(blockIdx.x) * DL_THREADS + threadIdx.x; e < n; e += stride) {
        double x[QT];
#pragma unroll
        for (int j = 0; j < QT; ++j)
            x[j] = j < q ? __ldg(X + j * ldx + e) : 0.0;
#pragma unroll
        for (int i = 0; i < PT; ++i) {
            if (i < p) {
                const double a = __ldg(A + i * lda + e);
#pragma unroll
                for (int j = 0; j < QT; ++j)
                    acc[i * QT + j] = fma(a, x[j], acc[i * QT + j]);
            }
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

// Y_j = beta * Y_j + sum_i coef[i * k + j] * A_i for j < k.  Safe when rows of Y alias rows of A:
// a thread reads every A_i[e] before it writes any Y_j[e].  WANT_SS: per-CTA partials of |Y_j|^2.
template <int KT, bool WANT_SS>
__global__ void __launch_bounds__(DL_THREADS)
    dl_combine_kernel(const double* A, int64_t lda, int p, const DlCoef coef, double* Y, int64_t ldy,
                      int k, double beta, int64_t n, double* __restrict__ part) {
    double ss[KT];
#pragma unroll
    for (int j = 0; j < KT; ++j)
        ss[j] = 0.0;
    const int64_t stride = static_cast<int64_t>(gridDim.x) * DL_THREADS;
    for (int64_t e = static_cast<int64_t>(blockIdx.x) * DL_THREADS + threadIdx.x; e < n; e += stride) {
        double y[KT];
#pragma unroll
        for (int j = 0; j < KT; ++j)
            y[j] = (beta != 0.0 && j < k) ? beta * Y[j * ldy + e] : 0.0;
        for (int i = 0; i < p; ++i) {
            const double a = A[i * lda + e];
#pragma unroll
            for (int j = 0; j < KT; ++j)
                y[j] = fma(coef.c[i * k + (j < k ? j : 0)], a, y[j]);
        }
#pragma unroll
        for (int j = 0; j < KT; ++j)
            if (j < k) {
                Y[j * ldy + e] = y[j];
                if (WANT_SS)
                    ss[j] = fma(y[j], y[j], ss[j]);
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
        if (p < 1 || p > DL_MAXV || q < 1 || q > DL_MAXV || n < 0)
            throw std::invalid_argument("f2_dl_dots: need 1 <= p, q <= 16 and n >= 0");
        cudaStream_t st = static_cast<cudaStream_t>(stream);
        const int grid = grid_for(n);
        for (int j0 = 0; j0 < q; j0 += 4) {
            const int qc = q - j0 < 4 ? q - j0 : 4;
            // partials of this column chunk: [grid][p*qc]; finished into a p x qc scratch, then placed
            double* part = work_dev;
            if (p <= 4)
                dots_q<4>(qc, grid, st, A, lda, p, X + j0 * ldx, ldx, n, part);
            else if (p <= 8)
                dots_q<8>(qc, grid, st, A, lda, p, X + j0 * ldx, ldx, n, part);
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
        if (p < 0 || p > DL_MAXV || k < 1 || k > DL_MAXV || n < 0)
            throw std::invalid_argument("f2_dl_combine: need 0 <= p <= 16, 1 <= k <= 16 and n >= 0");
        cudaStream_t st = static_cast<cudaStream_t>(stream);
        DlCoef c;
        std::memset(&c, 0, sizeof(c));
        if (p > 0)
            std::memcpy(c.c, coef_host, sizeof(double) * p * k);
        const int grid = grid_for(n);
        const bool ss = sumsq_dev != nullptr;
#define F2_COMBINE(KT)                                                                                     \
    do {                                                                                                   \
        if (ss)                                                                                            \
            dl_combine_kernel<KT, true><<<grid, DL_THREADS, 0, st>>>(A, lda, p, c, Y, ldy, k, beta, n, work_dev); \
        else                                                                                               \
            dl_combine_kernel<KT, false><<<grid, DL_THREADS, 0, st>>>(A, lda, p, c, Y, ldy, k, beta, n, work_dev); \
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
