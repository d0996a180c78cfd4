// C ABI over CIStringsHost (include/forte2_b200.h); no CUDA calls in this file, so the strings
// half of the library works on a machine without a GPU.
#include <cstring>
#include <memory>
#include <string>

#include "handles.hpp"

namespace f2 {
thread_local std::string g_last_error;
thread_local int g_last_status = F2_OK;

int status_of(const std::exception& e) {
    g_last_error = e.what();
    int st = F2_ERUNTIME;
    if (dynamic_cast<const std::invalid_argument*>(&e))
        st = F2_EINVAL;
    else if (dynamic_cast<const std::overflow_error*>(&e))
        st = F2_EOVERFLOW;
    else if (dynamic_cast<const std::out_of_range*>(&e))
        st = F2_EINDEX;
    else if (dynamic_cast<const CudaError*>(&e))
        st = F2_ECUDA;
    g_last_status = st;
    return st;
}
} // namespace f2

using namespace f2;

extern "C" {

const char* f2_last_error(void) { return g_last_error.c_str(); }
int f2_last_status(void) { return g_last_status; }
int f2_abi_version(void) { return 1; }

f2_strings* f2_strings_create(int64_t na, int64_t nb, int symmetry, int ngas, const int* gas_sizes,
                              const int* orbsym_flat, int n_min, const int* gas_min, int n_max,
                              const int* gas_max) {
    try {
        std::vector<std::vector<int>> orbsym;
        size_t k = 0;
        for (int g = 0; g < ngas; ++g) {
            orbsym.emplace_back(orbsym_flat + k, orbsym_flat + k + gas_sizes[g]);
            k += gas_sizes[g];
        }
        std::vector<int> gmin(gas_min, gas_min + n_min), gmax(gas_max, gas_max + n_max);
        auto h = std::make_unique<f2_strings>();
        h->host = std::make_unique<CIStringsHost>(na, nb, symmetry, orbsym, gmin, gmax);
        return h.release();
    } catch (const std::exception& e) {
        status_of(e);
        return nullptr;
    }
}

void f2_strings_destroy(f2_strings* h) { delete h; }

int f2_strings_info(const f2_strings* h, int64_t* out) {
    const auto& s = *h->host;
    const int64_t v[12] = {s.na,
                           s.nb,
                           s.nas,
                           s.nbs,
                           s.ndet,
                           s.norb,
                           s.nirrep,
                           s.ngas,
                           static_cast<int64_t>(s.blocks.size()),
                           s.symmetry,
                           static_cast<int64_t>(s.nclsA()),
                           static_cast<int64_t>(s.nclsB())};
    std::memcpy(out, v, sizeof(v));
    return F2_OK;
}

int f2_strings_blocks(const f2_strings* h, int64_t* out) {
    for (const auto& b : h->host->blocks) {
        *out++ = b.cA;
        *out++ = b.cB;
        *out++ = b.offset;
        *out++ = b.nA;
        *out++ = b.nB;
    }
    return F2_OK;
}

int64_t f2_strings_occupations(const f2_strings* h, int which, int64_t* out, int64_t cap) {
    const auto& s = *h->host;
    std::vector<int64_t> v;
    if (which < 2) {
        for (const auto& o : (which == 0 ? s.aocc : s.bocc))
            for (int x : o)
                v.push_back(x);
    } else {
        for (const auto& [a, b] : s.occ_pairs) {
            v.push_back(static_cast<int64_t>(a));
            v.push_back(static_cast<int64_t>(b));
        }
    }
    if (out && cap >= static_cast<int64_t>(v.size()))
        std::memcpy(out, v.data(), v.size() * sizeof(int64_t));
    return static_cast<int64_t>(v.size());
}

int64_t f2_strings_export(const f2_strings* h, int kind, int spin, int64_t* out, int64_t cap) {
    try {
        // cache the last export so that the sizing call and the copying call do the work once
        auto& c = h->export_cache;
        if (!(c.valid && c.kind == kind && c.spin == spin)) {
            c.data = h->host->export_kind(kind, spin);
            c.kind = kind;
            c.spin = spin;
            c.valid = true;
        }
        const int64_t n = static_cast<int64_t>(c.data.size());
        if (out && cap >= n) {
            std::memcpy(out, c.data.data(), n * sizeof(int64_t));
            c.valid = false;
            c.data = {};
        }
        return n;
    } catch (const std::exception& e) {
        status_of(e);
        return -1;
    }
}

int f2_strings_determinants(const f2_strings* h, uint64_t* out) {
    h->host->determinants(out);
    return F2_OK;
}

int f2_strings_determinant(const f2_strings* h, int64_t address, uint64_t* out2) {
    try {
        h->host->determinant(address, out2[0], out2[1]);
        return F2_OK;
    } catch (const std::exception& e) {
        return status_of(e);
    }
}

int64_t f2_strings_determinant_index(const f2_strings* h, uint64_t a, uint64_t b) {
    try {
        return h->host->determinant_index(a, b);
    } catch (const std::exception& e) {
        status_of(e);
        return -1;
    }
}

} // extern "C"
