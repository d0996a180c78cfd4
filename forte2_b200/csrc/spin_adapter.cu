// CSF <-> determinant transform of the CI vector (spin adaptation), GPU-resident.
//
// Behavioural contract: forte2::CISpinAdapter (forte2/determinant/ci_spin_adapter.cc:69-97 transforms,
// :99-200 prepare_couplings, :202-288 CSF construction) and the default Davidson preconditioner of
// CISigmaBuilder::form_Hdiag_csf (forte2/ci/ci_sigma_builder.cc:317-370).  Same CSF numbering
// (configurations in first-seen order over the determinants in CI-vector order, spin couplings of a
// configuration in lexicographic up/down order), same determinant phases, same coefficients.
//
// Not the reference's data structures: no hash of every determinant (addresses come from the string
// tables), and the coupling table is kept twice, CSR by CSF and CSR by determinant, so that both
// directions of the transform are gather-type sparse matrix-vector products on the GPU (one warp per
// row, no atomics, deterministic), bound by the HBM read of the table (12 B per coupling).
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <memory>
#include <numeric>
#include <unordered_map>
#include <vector>

#include <cuda_runtime.h>

#include "handles.hpp"
#include "nvtx_ranges.hpp"

namespace f2 {
namespace {

#define F2A_CUDA(call)                                                                             \
    do {                                                                                           \
        cudaError_t err__ = (call);                                                                \
        if (err__ != cudaSuccess)                                                                  \
            throw CudaError(std::string("CUDA error: ") + cudaGetErrorString(err__) + " at " +     \
                            __FILE__ + ":" + std::to_string(__LINE__));                            \
    } while (0)

// <spin coupling | spin occupation> for N unpaired electrons: product of the Clebsch-Gordan
// coefficients <S_{i-1} M_{i-1}; 1/2 m_i | S_i M_i> along the genealogical path.
// path[i] = 1: S decreases at step i; occ[i] = 1: electron i is beta (m = -1/2).
double path_overlap(int N, const std::vector<char>& path, const std::vector<char>& occ) {
    double o = 1.0;
    int S2 = 0, M2 = 0; // twice S and M after step i
    for (int i = 0; i < N; ++i) {
        const int ds = path[i] ? -1 : 1, dm = occ[i] ? -1 : 1;
        S2 += ds;
        M2 += dm;
        if (std::abs(M2) > S2)
            return 0.0;
        if (ds == 1)
            o *= std::sqrt(0.5 * (S2 + dm * M2) / S2);
        else
            o *= -dm * std::sqrt(0.5 * (S2 + 2.0 - dm * M2) / (S2 + 2.0));
    }
    return o;
}

// all sequences of length N with `ones` ones, in lexicographic order (0 < 1), optionally keeping the
// first element fixed at 0 and only the sequences whose running sum of (+1 for 0, -1 for 1) stays >= 0
std::vector<std::vector<char>> sequences(int N, int ones, bool genealogical) {
    std::vector<std::vector<char>> out;
    if (N == 0) {
        out.emplace_back();
        return out;
    }
    if (ones < 0 || ones > N)
        return out;
    std::vector<char> s(N, 0);
    std::fill(s.begin() + (N - ones), s.end(), 1);
    do {
        if (genealogical) {
            bool ok = true;
            for (int i = 0, p = 0; i < N && ok; ++i) {
                p += s[i] ? -1 : 1;
                ok = p >= 0;
            }
            if (!ok)
                continue;
        }
        out.push_back(s);
    } while (std::next_permutation(s.begin() + (genealogical ? 1 : 0), s.end()));
    return out;
}

struct OpenShellTable { // everything that depends only on the number of unpaired electrons N
    bool built = false;
    std::vector<std::vector<char>> occs;                  // spin occupations with 2 Ms
    std::vector<uint32_t> csf_ptr;                        // per coupling: range in (occ_id, coef)
    std::vector<uint32_t> occ_id;
    std::vector<double> coef;
    size_t ncsf() const { return csf_ptr.empty() ? 0 : csf_ptr.size() - 1; }
};

struct ConfKey {
    uint64_t docc, socc;
    bool operator==(const ConfKey& o) const { return docc == o.docc && socc == o.socc; }
};
struct ConfHash {
    size_t operator()(const ConfKey& k) const {
        uint64_t x = k.docc * 0x9e3779b97f4a7c15ull ^ (k.socc + 0x7f4a7c15ull + (k.docc << 6) + (k.docc >> 2));
        x ^= x >> 29;
        return static_cast<size_t>(x * 0xbf58476d1ce4e5b9ull);
    }
};

__global__ void csr_rows_kernel(const int64_t* __restrict__ ptr, const uint32_t* __restrict__ col,
                                const double* __restrict__ val, const double* __restrict__ x,
                                double* __restrict__ y, int64_t nrows, bool square) {
    // one warp per row: y[row] = sum_e val[e]^(1 or 2) x[col[e]]
    const int lane = threadIdx.x & 31;
    const int64_t warp = (static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
    const int64_t nwarps = (static_cast<int64_t>(gridDim.x) * blockDim.x) >> 5;
    for (int64_t row = warp; row < nrows; row += nwarps) {
        const int64_t e0 = ptr[row], e1 = ptr[row + 1];
        double acc = 0.0;
        for (int64_t e = e0 + lane; e < e1; e += 32) {
            const double v = val[e];
            acc += (square ? v * v : v) * x[col[e]];
        }
#pragma unroll
        for (int off = 16; off > 0; off >>= 1)
            acc += __shfl_xor_sync(0xffffffffu, acc, off);
        if (lane == 0)
            y[row] = acc;
    }
}

template <typename T> struct Dev {
    T* p = nullptr;
    size_t n = 0;
    ~Dev() {
        if (p)
            cudaFree(p);
    }
    void upload(const std::vector<T>& v) {
        if (p)
            cudaFree(p);
        p = nullptr;
        n = v.size();
        F2A_CUDA(cudaMalloc(&p, std::max<size_t>(n, 1) * sizeof(T)));
        if (n)
            F2A_CUDA(cudaMemcpy(p, v.data(), n * sizeof(T), cudaMemcpyHostToDevice));
    }
    void alloc(size_t count) {
        if (count <= n && p)
            return;
        if (p)
            cudaFree(p);
        p = nullptr;
        F2A_CUDA(cudaMalloc(&p, std::max<size_t>(count, 1) * sizeof(T)));
        n = count;
    }
};

} // namespace
} // namespace f2

using namespace f2;

struct f2_adapter {
    int twoS = 0, twoMs = 0, norb = 0;
    int64_t ndet = 0, ncsf = 0, nconf = 0, ncoupling = 0;
    // CSR by CSF (reference order) and CSR by determinant
    std::vector<int64_t> csf_ptr, det_ptr;
    std::vector<uint32_t> csf_det, det_csf;
    std::vector<double> csf_coef, det_coef;
    std::vector<int64_t> conf_csf; // nconf + 1: first CSF of every configuration
    // device copies
    int device = -1;
    Dev<int64_t> d_csf_ptr, d_det_ptr;
    Dev<uint32_t> d_csf_det, d_det_csf;
    Dev<double> d_csf_coef, d_det_coef, d_stage_a, d_stage_b, d_hdiag;
};

namespace {

void build_table(f2_adapter& A, const CIStringsHost& S) {
    A.norb = S.norb;
    A.ndet = S.ndet;
    if (A.twoS < 0)
        throw std::invalid_argument("CISpinAdapter: twoS must be non-negative");
    if (S.na - S.nb != A.twoMs)
        throw std::invalid_argument("CISpinAdapter: twoMs (" + std::to_string(A.twoMs) +
                                    ") does not match na - nb of the string space (" +
                                    std::to_string(S.na - S.nb) + ")");
    if (S.ndet >= (int64_t(1) << 32))
        throw std::overflow_error("CISpinAdapter: more than 2^32 determinants");
    const auto& SA = S.space(0, K_STR_N);
    const auto& SB = S.space(1, K_STR_N);

    // configurations in first-seen order over the determinants in CI-vector order, keeping those with
    // at least twoS unpaired electrons (ci_spin_adapter.cc:148-161)
    std::unordered_map<ConfKey, uint32_t, ConfHash> seen;
    std::vector<ConfKey> confs;
    for (const auto& blk : S.blocks)
        for (int64_t ia = 0; ia < blk.nA; ++ia) {
            const uint64_t a = SA.cls[blk.cA][ia];
            for (int64_t ib = 0; ib < blk.nB; ++ib) {
                const uint64_t b = SB.cls[blk.cB][ib];
                const ConfKey k{a & b, a ^ b};
                if (__builtin_popcountll(k.socc) < A.twoS)
                    continue;
                if (seen.emplace(k, static_cast<uint32_t>(confs.size())).second)
                    confs.push_back(k);
            }
        }
    A.nconf = static_cast<int64_t>(confs.size());

    std::vector<OpenShellTable> tables(S.norb + 1);
    auto table_of = [&](int N) -> const OpenShellTable& {
        OpenShellTable& T = tables[N];
        if (!T.built) {
            T.built = true;
            T.csf_ptr.assign(1, 0);
            if ((N + A.twoS) % 2 == 0 && (N + A.twoMs) % 2 == 0 && std::abs(A.twoMs) <= N) {
                const auto paths = sequences(N, (N - A.twoS) / 2, true);
                T.occs = sequences(N, (N - A.twoMs) / 2, false);
                for (const auto& path : paths) {
                    for (size_t j = 0; j < T.occs.size(); ++j) {
                        const double o = path_overlap(N, path, T.occs[j]);
                        if (std::fabs(o) > 0.0) {
                            T.occ_id.push_back(static_cast<uint32_t>(j));
                            T.coef.push_back(o);
                        }
                    }
                    T.csf_ptr.push_back(static_cast<uint32_t>(T.occ_id.size()));
                }
            }
        }
        return T;
    };

    // sizes first, then the fill
    int64_t ncsf = 0, ncoupling = 0;
    for (const auto& c : confs) {
        const auto& T = table_of(__builtin_popcountll(c.socc));
        ncsf += static_cast<int64_t>(T.ncsf());
        ncoupling += static_cast<int64_t>(T.occ_id.size());
    }
    if (ncsf >= (int64_t(1) << 32))
        throw std::overflow_error("CISpinAdapter: more than 2^32 CSFs");
    A.ncsf = ncsf;
    A.ncoupling = ncoupling;
    A.csf_ptr.assign(static_cast<size_t>(ncsf) + 1, 0);
    A.csf_det.resize(static_cast<size_t>(ncoupling));
    A.csf_coef.resize(static_cast<size_t>(ncoupling));
    A.conf_csf.assign(static_cast<size_t>(A.nconf) + 1, 0);

    int64_t csf = 0, e = 0;
    std::vector<int> orb;
    std::vector<int64_t> det_of_occ;
    std::vector<double> sign_of_occ;
    for (size_t ci = 0; ci < confs.size(); ++ci) {
        const ConfKey& c = confs[ci];
        A.conf_csf[ci] = csf;
        const int N = __builtin_popcountll(c.socc);
        const auto& T = table_of(N);
        orb.clear();
        for (uint64_t w = c.socc; w; w &= w - 1)
            orb.push_back(__builtin_ctzll(w));
        // determinant and phase of every spin occupation of this configuration: the unpaired electrons
        // are created from the last orbital to the first on top of the doubly occupied core, an alpha
        // creator passing the alpha electrons below it, a beta creator passing every alpha electron and
        // the beta electrons below it (ci_spin_adapter.cc:216-231, determinant.hpp:340-360)
        det_of_occ.assign(T.occs.size(), -1);
        sign_of_occ.assign(T.occs.size(), 1.0);
        for (size_t j = 0; j < T.occs.size(); ++j) {
            uint64_t a = c.docc, b = c.docc;
            int parity = 0;
            for (int k = N - 1; k >= 0; --k) {
                const int p = orb[k];
                if (T.occs[j][k]) {
                    parity += popcount_below(b, p) + __builtin_popcountll(a);
                    b |= 1ull << p;
                } else {
                    parity += popcount_below(a, p);
                    a |= 1ull << p;
                }
            }
            det_of_occ[j] = S.determinant_index(a, b);
            sign_of_occ[j] = (parity & 1) ? -1.0 : 1.0;
        }
        for (size_t q = 0; q < T.ncsf(); ++q) {
            for (uint32_t t = T.csf_ptr[q]; t < T.csf_ptr[q + 1]; ++t) {
                const uint32_t j = T.occ_id[t];
                if (det_of_occ[j] < 0)
                    throw std::runtime_error("CISpinAdapter: a determinant of a configuration is missing "
                                             "from the CI space (the space is not closed under spin flips)");
                A.csf_det[e] = static_cast<uint32_t>(det_of_occ[j]);
                A.csf_coef[e] = sign_of_occ[j] * T.coef[t];
                ++e;
            }
            A.csf_ptr[++csf] = e;
        }
    }
    A.conf_csf[confs.size()] = csf;

    // transpose: CSR by determinant (entries of a determinant in increasing CSF order)
    A.det_ptr.assign(static_cast<size_t>(A.ndet) + 1, 0);
    for (int64_t k = 0; k < ncoupling; ++k)
        A.det_ptr[A.csf_det[k] + 1] += 1;
    for (int64_t d = 0; d < A.ndet; ++d)
        A.det_ptr[d + 1] += A.det_ptr[d];
    A.det_csf.resize(static_cast<size_t>(ncoupling));
    A.det_coef.resize(static_cast<size_t>(ncoupling));
    std::vector<int64_t> fill(A.det_ptr.begin(), A.det_ptr.end() - 1);
    for (int64_t q = 0; q < ncsf; ++q)
        for (int64_t k = A.csf_ptr[q]; k < A.csf_ptr[q + 1]; ++k) {
            const int64_t pos = fill[A.csf_det[k]]++;
            A.det_csf[pos] = static_cast<uint32_t>(q);
            A.det_coef[pos] = A.csf_coef[k];
        }
}

void to_device(f2_adapter& A, int device) {
    int count = 0;
    if (cudaGetDeviceCount(&count) != cudaSuccess || count == 0)
        throw CudaError("forte2_b200: no CUDA device available (the spin adapter has no CPU fallback)");
    if (A.device == device)
        return;
    F2A_CUDA(cudaSetDevice(device));
    A.d_csf_ptr.upload(A.csf_ptr);
    A.d_csf_det.upload(A.csf_det);
    A.d_csf_coef.upload(A.csf_coef);
    A.d_det_ptr.upload(A.det_ptr);
    A.d_det_csf.upload(A.det_csf);
    A.d_det_coef.upload(A.det_coef);
    // cudaMemcpy from pageable memory returns once the data is staged: the DMA tail must have landed before a
    // kernel on the caller's (possibly non-blocking) stream reads the tables
    F2A_CUDA(cudaDeviceSynchronize());
    A.device = device;
}

void launch_rows(const int64_t* ptr, const uint32_t* col, const double* val, const double* x, double* y,
                 int64_t nrows, bool square, cudaStream_t st) {
    if (nrows == 0)
        return;
    const int64_t want = (nrows + 7) / 8; // 8 warps per CTA
    const int grid = static_cast<int>(std::min<int64_t>(want, 148 * 16));
    csr_rows_kernel<<<grid, 256, 0, st>>>(ptr, col, val, x, y, nrows, square);
    F2A_CUDA(cudaGetLastError());
}

int current_device_or(int fallback) {
    int d = fallback;
    if (cudaGetDevice(&d) != cudaSuccess)
        d = fallback;
    return d;
}

} // namespace

extern "C" {

f2_adapter* f2_adapter_create(const f2_strings* strings, int twoS, int twoMs) {
    try {
        if (!strings)
            throw std::invalid_argument("null strings handle");
        auto A = std::make_unique<f2_adapter>();
        A->twoS = twoS;
        A->twoMs = twoMs;
        build_table(*A, *strings->host);
        return A.release();
    } catch (const std::exception& e) {
        status_of(e);
        return nullptr;
    }
}

void f2_adapter_destroy(f2_adapter* h) { delete h; }

int f2_adapter_info(const f2_adapter* h, int64_t* out4) {
    out4[0] = h->nconf;
    out4[1] = h->ncsf;
    out4[2] = h->ndet;
    out4[3] = h->ncoupling;
    return F2_OK;
}

int64_t f2_adapter_table(const f2_adapter* h, int64_t* bounds, int64_t* det, double* coeff, int64_t cap) {
    if (bounds)
        std::copy(h->csf_ptr.begin(), h->csf_ptr.end(), bounds);
    if (det && coeff && cap >= h->ncoupling)
        for (int64_t k = 0; k < h->ncoupling; ++k) {
            det[k] = h->csf_det[k];
            coeff[k] = h->csf_coef[k];
        }
    return h->ncoupling;
}

int f2_adapter_conf_ranges(const f2_adapter* h, int64_t* out) {
    std::copy(h->conf_csf.begin(), h->conf_csf.end(), out);
    return F2_OK;
}

int f2_adapter_csf_to_det(f2_adapter* h, const double* csf_dev, double* det_dev, int device, void* stream) {
    try {
        f2::NvtxRange nvtx("CSF -> determinant transform");
        to_device(*h, device);
        F2A_CUDA(cudaSetDevice(device));
        launch_rows(h->d_det_ptr.p, h->d_det_csf.p, h->d_det_coef.p, csf_dev, det_dev, h->ndet, false,
                    static_cast<cudaStream_t>(stream));
        return F2_OK;
    } catch (const std::exception& e) {
        return status_of(e);
    }
}

int f2_adapter_det_to_csf(f2_adapter* h, const double* det_dev, double* csf_dev, int device, void* stream) {
    try {
        f2::NvtxRange nvtx("determinant -> CSF transform");
        to_device(*h, device);
        F2A_CUDA(cudaSetDevice(device));
        launch_rows(h->d_csf_ptr.p, h->d_csf_det.p, h->d_csf_coef.p, det_dev, csf_dev, h->ncsf, false,
                    static_cast<cudaStream_t>(stream));
        return F2_OK;
    } catch (const std::exception& e) {
        return status_of(e);
    }
}

// numpy-facing forms (what the reference's methods receive): staged through the device, synchronous
int f2_adapter_csf_to_det_host(f2_adapter* h, const double* csf_host, int64_t ncsf, double* det_host, int64_t ndet) {
    try {
        if (ncsf != h->ncsf || ndet != h->ndet)
            throw std::invalid_argument("CISpinAdapter.csf_C_to_det_C: expected " + std::to_string(h->ncsf) +
                                        " CSF and " + std::to_string(h->ndet) + " determinant coefficients");
        const int dev = h->device >= 0 ? h->device : current_device_or(0);
        to_device(*h, dev);
        F2A_CUDA(cudaSetDevice(dev));
        h->d_stage_a.alloc(static_cast<size_t>(ncsf));
        h->d_stage_b.alloc(static_cast<size_t>(ndet));
        F2A_CUDA(cudaMemcpy(h->d_stage_a.p, csf_host, static_cast<size_t>(ncsf) * 8, cudaMemcpyHostToDevice));
        launch_rows(h->d_det_ptr.p, h->d_det_csf.p, h->d_det_coef.p, h->d_stage_a.p, h->d_stage_b.p, h->ndet,
                    false, nullptr);
        F2A_CUDA(cudaMemcpy(det_host, h->d_stage_b.p, static_cast<size_t>(ndet) * 8, cudaMemcpyDeviceToHost));
        return F2_OK;
    } catch (const std::exception& e) {
        return status_of(e);
    }
}

int f2_adapter_det_to_csf_host(f2_adapter* h, const double* det_host, int64_t ndet, double* csf_host, int64_t ncsf) {
    try {
        if (ncsf != h->ncsf || ndet != h->ndet)
            throw std::invalid_argument("CISpinAdapter.det_C_to_csf_C: expected " + std::to_string(h->ndet) +
                                        " determinant and " + std::to_string(h->ncsf) + " CSF coefficients");
        const int dev = h->device >= 0 ? h->device : current_device_or(0);
        to_device(*h, dev);
        F2A_CUDA(cudaSetDevice(dev));
        h->d_stage_a.alloc(static_cast<size_t>(ncsf));
        h->d_stage_b.alloc(static_cast<size_t>(ndet));
        F2A_CUDA(cudaMemcpy(h->d_stage_b.p, det_host, static_cast<size_t>(ndet) * 8, cudaMemcpyHostToDevice));
        launch_rows(h->d_csf_ptr.p, h->d_csf_det.p, h->d_csf_coef.p, h->d_stage_b.p, h->d_stage_a.p, h->ncsf,
                    false, nullptr);
        F2A_CUDA(cudaMemcpy(csf_host, h->d_stage_a.p, static_cast<size_t>(ncsf) * 8, cudaMemcpyDeviceToHost));
        return F2_OK;
    } catch (const std::exception& e) {
        return status_of(e);
    }
}

// H_II ~ sum_d c_dI^2 E_d from the determinant energies (device vector of ndet doubles)
int f2_adapter_hdiag_csf(f2_adapter* h, const double* det_energies_dev, double* out_csf_dev, int device,
                         void* stream) {
    try {
        f2::NvtxRange nvtx("CSF preconditioner");
        to_device(*h, device);
        F2A_CUDA(cudaSetDevice(device));
        launch_rows(h->d_csf_ptr.p, h->d_csf_det.p, h->d_csf_coef.p, det_energies_dev, out_csf_dev, h->ncsf, true,
                    static_cast<cudaStream_t>(stream));
        return F2_OK;
    } catch (const std::exception& e) {
        return status_of(e);
    }
}

} // extern "C"
