// Replacement lists built on the GPU (reference counterpart: the serial, hash-map based
// builders of forte2/ci/ci_strings_makers.cc:75-322).  The strings themselves are enumerated on the host (milliseconds,
// strings.cc); every list entry needs the address of a string that differs from a known one by one or two orbitals, and
// that look-up -- 10^7 to 10^8 of them for the 3-GAS / 20-orbital space -- is what the time goes into.  Here:
//
//   as_*   alpha single replacements, CSR by output string     one thread per string: count, scan, fill
//   bs_*   beta single replacements, ELL by source string      one thread per string: count, max, fill
//   h2_*   two-hole members, CSR by (N-2)-electron string      one thread per string emits (K << 32 | I, entry) for each
//                                                              pair of its electrons; radix sort; row pointer by search
//
// A string is found by binary search in the sorted occupation words of its space (the table the host look-up uses,
// StringSpace::key), which yields its global index.  Entry formats and entry order are those of the host builder
// (strings.cc build_device_lists), bit for bit: tests/test_lists_gpu.py compares every array.
#include <algorithm>
#include <cstdint>
#include <stdexcept>
#include <string>
#include <vector>

#include <cub/device/device_radix_sort.cuh>
#include <cub/device/device_reduce.cuh>
#include <cub/device/device_scan.cuh>
#include <cuda_runtime.h>

#include "lists_device.hpp"
#include "nvtx_ranges.hpp"

namespace f2 {

namespace {

#define F2L_CUDA(call)                                                                                        \
    do {                                                                                                      \
        cudaError_t err__ = (call);                                                                           \
        if (err__ != cudaSuccess)                                                                             \
            throw std::runtime_error(std::string("CUDA error in the list builder: ") + cudaGetErrorString(err__)); \
    } while (0)

template <class T> struct Buf {
    T* p = nullptr;
    size_t n = 0;
    Buf() = default;
    explicit Buf(size_t count) { alloc(count); }
    Buf(const Buf&) = delete;
    Buf& operator=(const Buf&) = delete;
    ~Buf() {
        if (p)
            cudaFree(p);
    }
    void alloc(size_t count) {
        if (p)
            cudaFree(p);
        p = nullptr;
        n = count;
        if (count)
            F2L_CUDA(cudaMalloc(&p, count * sizeof(T)));
    }
    void upload(const std::vector<T>& v) {
        alloc(v.size());
        if (!v.empty())
            F2L_CUDA(cudaMemcpy(p, v.data(), v.size() * sizeof(T), cudaMemcpyHostToDevice));
    }
    void download(std::vector<T>& v, size_t count) const {
        v.resize(count);
        if (count)
            F2L_CUDA(cudaMemcpy(v.data(), p, count * sizeof(T), cudaMemcpyDeviceToHost));
    }
};

// sorted words of a string space and the global index of each
struct Table {
    const uint64_t* key;
    const uint32_t* gidx;
    uint32_t n;
};

__device__ __forceinline__ bool find(const Table& t, uint64_t w, uint32_t& g) {
    uint32_t lo = 0, hi = t.n;
    while (lo < hi) {
        const uint32_t mid = (lo + hi) >> 1;
        if (__ldg(t.key + mid) < w)
            lo = mid + 1;
        else
            hi = mid;
    }
    if (lo >= t.n || __ldg(t.key + lo) != w)
        return false;
    g = __ldg(t.gidx + lo);
    return true;
}

__device__ __forceinline__ uint64_t below(uint64_t w, int n) { return w & ((1ull << n) - 1ull); }
__device__ __forceinline__ bool odd_below(uint64_t w, int n) { return __popcll(below(w, n)) & 1; }

// single replacements of string g in the order (annihilated orbital ascending, created orbital ascending):
// FILL = false counts them, FILL = true writes them through emit(t, target global index, annihilated, created, negative)
template <bool FILL, class Emit>
__device__ __forceinline__ uint32_t singles(const Table& t, uint64_t s, int norb, Emit emit) {
    uint32_t cnt = 0;
    for (int k = 0; k < norb; ++k) {
        if (!((s >> k) & 1))
            continue;
        const bool nk = odd_below(s, k);
        const uint64_t s1 = s & ~(1ull << k);
        for (int l = 0; l < norb; ++l) {
            if ((s1 >> l) & 1)
                continue;
            uint32_t J;
            if (!find(t, s1 | (1ull << l), J))
                continue;
            if (FILL)
                emit(cnt, J, k, l, nk != odd_below(s1, l));
            ++cnt;
        }
    }
    return cnt;
}

__global__ void singles_count_kernel(Table t, const uint64_t* __restrict__ str, uint32_t n, int norb, uint32_t* cnt) {
    const uint32_t g = blockIdx.x * blockDim.x + threadIdx.x;
    if (g < n)
        cnt[g] = singles<false>(t, str[g], norb, [](uint32_t, uint32_t, int, int, bool) {});
}

// alpha: <Ia| E_kl |Ja> = sign seen from the output row Ia (Ja = sign a+_l a_k Ia): Ja | (k norb + l) << 32 | neg << 48
__global__ void alpha_fill_kernel(Table t, const uint64_t* __restrict__ str, uint32_t n, int norb,
                                  const uint32_t* __restrict__ ptr, uint64_t* ent) {
    const uint32_t g = blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= n)
        return;
    uint64_t* out = ent + ptr[g];
    singles<true>(t, str[g], norb, [&](uint32_t i, uint32_t J, int k, int l, bool neg) {
        out[i] = uint64_t(J) | (uint64_t(k * norb + l) << 32) | (uint64_t(neg) << 48);
    });
}

// beta: <Ib| E_ij |Jb> = sign seen from the source column Jb (Ib = sign a+_i a_j Jb), entry t at [t nbs + Jb]:
// address(Ib) | class(Ib) << 28 | (i norb + j) << 40 | neg << 56
__global__ void beta_fill_kernel(Table t, const uint64_t* __restrict__ str, uint32_t n, int norb,
                                 const uint32_t* __restrict__ cls, const uint32_t* __restrict__ loc, uint64_t* ent) {
    const uint32_t g = blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= n)
        return;
    singles<true>(t, str[g], norb, [&](uint32_t i, uint32_t I, int j, int c, bool neg) {
        ent[static_cast<size_t>(i) * n + g] = uint64_t(loc[I]) | (uint64_t(cls[I]) << 28) |
                                              (uint64_t(c * norb + j) << 40) | (uint64_t(neg) << 56);
    });
}

// two-hole members: string g = K + {p > q}, a_p a_q |g> = sign |K>; one record per pair of electrons, in slot
// g * npairs + (running pair count); pairs whose K is outside the 2h space get the all-ones key (sorted to the end)
__global__ void holes2_emit_kernel(Table t2, const uint64_t* __restrict__ str, uint32_t n, int norb, int npairs,
                                   uint64_t* keys, uint64_t* vals) {
    const uint32_t g = blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= n)
        return;
    const uint64_t s = str[g];
    size_t slot = static_cast<size_t>(g) * npairs;
    for (int q = 0; q < norb; ++q) {
        if (!((s >> q) & 1))
            continue;
        const uint64_t s1 = s & ~(1ull << q);
        const bool nq = odd_below(s1, q);
        for (int p = q + 1; p < norb; ++p) {
            if (!((s >> p) & 1))
                continue;
            const uint64_t s2 = s1 & ~(1ull << p);
            uint32_t K;
            if (find(t2, s2, K)) {
                keys[slot] = (uint64_t(K) << 32) | g;
                vals[slot] = uint64_t(g) | (uint64_t(p * (p - 1) / 2 + q) << 32) | (uint64_t(nq != odd_below(s2, p)) << 48);
            } else {
                keys[slot] = ~0ull;
                vals[slot] = ~0ull;
            }
            ++slot;
        }
    }
}

// ptr[K] = first sorted record whose key is at least K << 32 (K = 0 .. nK; records with the all-ones key stay behind
// every real K because K < 2^32 - 1)
__global__ void holes2_ptr_kernel(const uint64_t* __restrict__ keys, uint64_t nrec, uint32_t nK, uint32_t* ptr) {
    const uint32_t K = blockIdx.x * blockDim.x + threadIdx.x;
    if (K > nK)
        return;
    const uint64_t want = uint64_t(K) << 32;
    uint64_t lo = 0, hi = nrec;
    while (lo < hi) {
        const uint64_t mid = (lo + hi) >> 1;
        if (__ldg(keys + mid) < want)
            lo = mid + 1;
        else
            hi = mid;
    }
    ptr[K] = static_cast<uint32_t>(lo);
}

struct SpaceOnDevice {
    Buf<uint64_t> key, str;
    Buf<uint32_t> gidx;
    Table table() const { return Table{key.p, gidx.p, static_cast<uint32_t>(key.n)}; }
};

void put_space(const StringSpace& S, SpaceOnDevice& d, bool with_strings) {
    std::vector<uint32_t> gidx(S.key.size());
    for (size_t i = 0; i < S.key.size(); ++i)
        gidx[i] = S.cls_off[S.key_cls[i]] + S.key_add[i];
    d.key.upload(S.key);
    d.gidx.upload(gidx);
    if (with_strings) {
        std::vector<uint64_t> flat;
        flat.reserve(S.nstr());
        for (const auto& c : S.cls)
            flat.insert(flat.end(), c.begin(), c.end());
        d.str.upload(flat);
    }
}

uint32_t grid_for(size_t n) { return static_cast<uint32_t>((n + 127) / 128); }

} // namespace

void build_lists_on_device(const CIStringsHost& h) {
    NvtxRange nvtx("replacement lists on the device");
    DeviceLists& d = h.lists_storage();
    h.lists_prologue();
    const auto& A = h.space(0, K_STR_N);
    const auto& B = h.space(1, K_STR_N);
    const int norb = h.norb;
    if (A.nstr() >= (size_t(1) << 32) - 1 || B.nstr() >= (size_t(1) << 28))
        throw std::overflow_error("CIStrings: too many strings for the packed list format");
    SpaceOnDevice dA, dB;
    put_space(A, dA, true);
    put_space(B, dB, true);
    Buf<unsigned char> tmp;
    auto scratch = [&](size_t bytes) {
        if (tmp.n < bytes)
            tmp.alloc(bytes);
        return static_cast<void*>(tmp.p);
    };

    // ---- alpha singles, CSR
    {
        const uint32_t n = static_cast<uint32_t>(A.nstr());
        d.as_ptr.assign(size_t(n) + 1, 0);
        d.as_ent.clear();
        d.as_max = 0;
        if (n > 0) {
            Buf<uint32_t> cnt(size_t(n) + 1), ptr(size_t(n) + 1);
            F2L_CUDA(cudaMemset(cnt.p, 0, (size_t(n) + 1) * sizeof(uint32_t)));
            singles_count_kernel<<<grid_for(n), 128>>>(dA.table(), dA.str.p, n, norb, cnt.p);
            size_t bytes = 0;
            F2L_CUDA(cub::DeviceScan::ExclusiveSum(nullptr, bytes, cnt.p, ptr.p, n + 1));
            F2L_CUDA(cub::DeviceScan::ExclusiveSum(scratch(bytes), bytes, cnt.p, ptr.p, n + 1));
            Buf<uint32_t> mx(1);
            F2L_CUDA(cub::DeviceReduce::Max(nullptr, bytes, cnt.p, mx.p, n));
            F2L_CUDA(cub::DeviceReduce::Max(scratch(bytes), bytes, cnt.p, mx.p, n));
            ptr.download(d.as_ptr, size_t(n) + 1);
            std::vector<uint32_t> m;
            mx.download(m, 1);
            d.as_max = static_cast<int>(m[0]);
            const size_t total = d.as_ptr[n];
            // (a 32-bit row pointer: the host builder has the same limit)
            Buf<uint64_t> ent(total);
            if (total > 0) {
                alpha_fill_kernel<<<grid_for(n), 128>>>(dA.table(), dA.str.p, n, norb, ptr.p, ent.p);
                F2L_CUDA(cudaGetLastError());
                ent.download(d.as_ent, total);
            }
        }
    }
    // ---- beta singles, ELL
    {
        const uint32_t n = static_cast<uint32_t>(B.nstr());
        d.bs_width = 0;
        d.bs_ent.clear();
        if (n > 0) {
            Buf<uint32_t> cnt(n), mx(1), cls, loc;
            singles_count_kernel<<<grid_for(n), 128>>>(dB.table(), dB.str.p, n, norb, cnt.p);
            size_t bytes = 0;
            F2L_CUDA(cub::DeviceReduce::Max(nullptr, bytes, cnt.p, mx.p, n));
            F2L_CUDA(cub::DeviceReduce::Max(scratch(bytes), bytes, cnt.p, mx.p, n));
            std::vector<uint32_t> m;
            mx.download(m, 1);
            d.bs_width = static_cast<int>(m[0]);
            const size_t total = size_t(m[0]) * n;
            if (total > 0) {
                cls.upload(d.b_cls);
                loc.upload(d.b_loc);
                Buf<uint64_t> ent(total);
                F2L_CUDA(cudaMemset(ent.p, 0xff, total * sizeof(uint64_t))); // kNullEntry
                beta_fill_kernel<<<grid_for(n), 128>>>(dB.table(), dB.str.p, n, norb, cls.p, loc.p, ent.p);
                F2L_CUDA(cudaGetLastError());
                ent.download(d.bs_ent, total);
            }
        }
    }
    // ---- two-hole members, CSR by (N-2)-electron string
    for (int spin = 0; spin < 2; ++spin) {
        const int ne = static_cast<int>(spin == 0 ? h.na : h.nb);
        const SpaceOnDevice& dS = spin == 0 ? dA : dB;
        auto& ptr = d.h2_ptr[spin];
        auto& ent = d.h2_ent[spin];
        ptr.assign(1, 0);
        ent.clear();
        if (ne == 1)
            h.lists_one_electron(spin);
        if (ne >= 2) {
            const auto& H = h.space(spin, K_STR_2H);
            const uint32_t nK = static_cast<uint32_t>(H.nstr());
            const uint32_t n = static_cast<uint32_t>(dS.str.n);
            const int npairs = ne * (ne - 1) / 2;
            const size_t nrec = size_t(n) * npairs;
            if (nrec >= (size_t(1) << 32))
                throw std::overflow_error("CIStrings: more than 2^32 two-hole list entries");
            ptr.assign(size_t(nK) + 1, 0);
            if (nrec > 0 && nK > 0) {
                SpaceOnDevice dH;
                put_space(H, dH, false);
                Buf<uint64_t> k0(nrec), v0(nrec), k1(nrec), v1(nrec);
                holes2_emit_kernel<<<grid_for(n), 128>>>(dH.table(), dS.str.p, n, norb, npairs, k0.p, v0.p);
                F2L_CUDA(cudaGetLastError());
                size_t bytes = 0;
                F2L_CUDA(cub::DeviceRadixSort::SortPairs(nullptr, bytes, k0.p, k1.p, v0.p, v1.p, nrec));
                F2L_CUDA(cub::DeviceRadixSort::SortPairs(scratch(bytes), bytes, k0.p, k1.p, v0.p, v1.p, nrec));
                Buf<uint32_t> dptr(size_t(nK) + 1);
                holes2_ptr_kernel<<<grid_for(size_t(nK) + 1), 128>>>(k1.p, nrec, nK, dptr.p);
                F2L_CUDA(cudaGetLastError());
                dptr.download(ptr, size_t(nK) + 1);
                v1.download(ent, ptr[nK]);
            }
        }
        h.lists_epilogue(spin);
    }
    F2L_CUDA(cudaDeviceSynchronize());
    h.lists_adopt();
}

} // namespace f2
