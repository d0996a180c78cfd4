// TEST INFRASTRUCTURE — NOT PART OF THE PRODUCT.
//
// C-ABI glue over the *unmodified* reference classes forte2::CIStrings, forte2::CISigmaBuilder,
// forte2::SlaterRules and forte2::CISpinAdapter, compiled from /root/reference in place by
// oracle/Makefile into oracle/_ref/libforte2_ref.so.  Only tests/, __graft_entry__.smoke() and
// bench.py's cpu_baseline / --impl reference legs load that library (as the checker or the timed
// CPU baseline, never as the product).
//
// Every exported list is flattened to int64 rows in the reference's own container order
// (std::map key order, then push_back order; forte2/ci/ci_strings_makers.cc:75-322), which is the
// canonical order the product's f2_strings_export() reproduces bit-exactly.
#include <array>
#include <cstdint>
#include <cstring>
#include <memory>
#include <string>
#include <vector>

#include "ci/ci_occupation.h"
#include "ci/ci_sigma_builder.h"
#include "ci/ci_strings.h"
#include "ci/ci_strings_makers.h"
#include "determinant/ci_spin_adapter.h"
#include "determinant/slater_rules.h"
#include "helpers/logger.h"

using namespace forte2;

namespace {

thread_local std::string g_err;

struct RefStrings {
    std::unique_ptr<CIStrings> s;
    std::vector<std::vector<int>> orbsym;
    std::vector<int> gas_size;
    std::vector<std::vector<size_t>> gas_mos;
};

struct RefSigma {
    RefStrings* strings;
    std::vector<double> Hbuf, Vbuf;
    np_matrix H;
    np_tensor4 V;
    std::unique_ptr<CISigmaBuilder> b;
};

struct RefAdapter {
    std::unique_ptr<CISpinAdapter> a;
    std::vector<Determinant> dets;
};

np_vector wrap_vec(double* p, size_t n) {
    size_t shape[1] = {n};
    return np_vector(p, 1, shape);
}

// kinds understood by ref_strings_export (shared with include/forte2_b200.h)
enum Kind {
    K_STR_N = 0,
    K_STR_1H = 1,
    K_STR_2H = 2,
    K_STR_3H = 3,
    K_STR_1H1P = 4,
    K_VO = 10,
    K_VO2 = 11,
    K_1H = 12,
    K_1H2 = 13,
    K_2H = 14,
    K_3H = 15,
};

using occ_list = std::vector<std::array<int, 6>>;

occ_list holes(const occ_list& occ, int n) {
    occ_list o = occ;
    for (int i = 0; i < n; ++i)
        o = generate_1h_occupations(o);
    return o;
}

StringList strings_of_kind(const RefStrings& r, int kind, int spin) {
    const auto& L = *r.s;
    if (kind == K_STR_N)
        return spin == 0 ? L.alpha_strings() : L.beta_strings();
    const auto& base = spin == 0 ? L.gas_alpha_occupations() : L.gas_beta_occupations();
    occ_list occ;
    if (kind == K_STR_1H1P)
        occ = spin == 0 ? L.gas_alpha_1h1p_occupations() : L.gas_beta_1h1p_occupations();
    else
        occ = holes(base, kind - K_STR_N);
    auto sc = L.string_class();
    return make_strings_with_occupation(L.ngas_spaces(), L.nirrep(), r.gas_size, r.gas_mos, occ,
                                        sc);
}

template <typename Map, typename Row> void flatten_hole(const Map& m, std::vector<int64_t>& out, Row row) {
    for (const auto& [key, vec] : m) {
        const auto& [cK, aK, cI] = key;
        for (const auto& e : vec) {
            out.push_back(cK);
            out.push_back(static_cast<int64_t>(aK));
            out.push_back(cI);
            row(e, out);
        }
    }
}

} // namespace

extern "C" {

const char* ref_last_error() { return g_err.c_str(); }

void ref_set_log_level(int level) {
    Logger::getInstance().setLevel(static_cast<Logger::Level>(level));
}

void* ref_strings_create(int na, int nb, int symmetry, int ngas, const int* gas_sizes,
                         const int* orbsym_flat, int n_min, const int* gas_min, int n_max,
                         const int* gas_max) {
    try {
        auto r = std::make_unique<RefStrings>();
        size_t k = 0;
        for (int g = 0; g < ngas; ++g) {
            std::vector<int> sym(orbsym_flat + k, orbsym_flat + k + gas_sizes[g]);
            std::vector<size_t> mos(gas_sizes[g]);
            for (int i = 0; i < gas_sizes[g]; ++i)
                mos[i] = k + i;
            k += gas_sizes[g];
            r->orbsym.push_back(sym);
            r->gas_size.push_back(gas_sizes[g]);
            r->gas_mos.push_back(mos);
        }
        std::vector<int> gmin(gas_min, gas_min + n_min), gmax(gas_max, gas_max + n_max);
        r->s = std::make_unique<CIStrings>(na, nb, symmetry, r->orbsym, gmin, gmax);
        return r.release();
    } catch (const std::exception& e) {
        g_err = e.what();
        return nullptr;
    }
}

void ref_strings_destroy(void* h) { delete static_cast<RefStrings*>(h); }

// out[0..9] = na nb nas nbs ndet norb nirrep ngas nblocks symmetry
void ref_strings_info(void* h, int64_t* out) {
    const auto& L = *static_cast<RefStrings*>(h)->s;
    out[0] = L.na();
    out[1] = L.nb();
    out[2] = L.nas();
    out[3] = L.nbs();
    out[4] = L.ndet();
    out[5] = L.norb();
    out[6] = L.nirrep();
    out[7] = L.ngas_spaces();
    out[8] = L.determinant_classes().size();
    out[9] = L.symmetry();
}

// rows of (class_Ia, class_Ib, offset, nIa, nIb)
void ref_strings_blocks(void* h, int64_t* out) {
    const auto& L = *static_cast<RefStrings*>(h)->s;
    for (const auto& [n, cIa, cIb] : L.determinant_classes()) {
        *out++ = cIa;
        *out++ = cIb;
        *out++ = L.block_offset(n);
        *out++ = L.alpha_address()->strpcls(cIa);
        *out++ = L.beta_address()->strpcls(cIb);
    }
}

// which: 0 = alpha occupations (rows of 6), 1 = beta occupations (rows of 6), 2 = pairs (rows of 2)
int64_t ref_strings_occupations(void* h, int which, int64_t* out, int64_t cap) {
    const auto& L = *static_cast<RefStrings*>(h)->s;
    std::vector<int64_t> v;
    if (which < 2) {
        for (const auto& o : (which == 0 ? L.gas_alpha_occupations() : L.gas_beta_occupations()))
            for (int x : o)
                v.push_back(x);
    } else {
        for (const auto& [a, b] : L.gas_occupations()) {
            v.push_back(a);
            v.push_back(b);
        }
    }
    if (out && cap >= static_cast<int64_t>(v.size()))
        std::memcpy(out, v.data(), v.size() * sizeof(int64_t));
    return v.size();
}

// Flatten strings / lists to int64.  Returns the number of int64 produced; copies only when the
// caller's buffer is large enough (call once with out = NULL to size it).
//   strings: [nclasses, strpcls[0..nclasses), then the uint64 words class by class]
//   VO   rows: class_I class_J p q sign I J
//   VO2  rows: class_K class_I K sign(±1, ±2) pq J
//   1H   rows: class_K add_K class_I sign p I
//   1H2  rows: class_K class_I K sign p I
//   2H   rows: class_K add_K class_I sign p q I
//   3H   rows: class_K add_K class_I sign p q r I
int64_t ref_strings_export(void* h, int kind, int spin, int64_t* out, int64_t cap) {
    auto* r = static_cast<RefStrings*>(h);
    const auto& L = *r->s;
    std::vector<int64_t> v;
    try {
        if (kind <= K_STR_1H1P) {
            auto sl = strings_of_kind(*r, kind, spin);
            v.push_back(sl.size());
            for (const auto& c : sl)
                v.push_back(c.size());
            for (const auto& c : sl)
                for (const auto& s : c)
                    v.push_back(static_cast<int64_t>(s.get_word(0)));
        } else if (kind == K_VO) {
            const auto& addr = spin == 0 ? L.alpha_address() : L.beta_address();
            for (int cI = 0; cI < addr->nclasses(); ++cI)
                for (int cJ = 0; cJ < addr->nclasses(); ++cJ) {
                    const auto& el =
                        spin == 0 ? L.get_alpha_vo_list(cI, cJ) : L.get_beta_vo_list(cI, cJ);
                    for (const auto& [pq, vec] : el)
                        for (const auto& e : vec) {
                            v.insert(v.end(), {cI, cJ, std::get<0>(pq), std::get<1>(pq),
                                               static_cast<int64_t>(e.sign), e.I, e.J});
                        }
                }
        } else if (kind == K_VO2) {
            const auto& addrK = spin == 0 ? L.alpha_address_1h1p() : L.beta_address_1h1p();
            const auto& addrI = spin == 0 ? L.alpha_address() : L.beta_address();
            for (int cK = 0; cK < addrK->nclasses(); ++cK)
                for (int cI = 0; cI < addrI->nclasses(); ++cI) {
                    const auto& el =
                        spin == 0 ? L.get_alpha_vo_list2(cK, cI) : L.get_beta_vo_list2(cK, cI);
                    for (size_t K = 0; K < el.size(); ++K)
                        for (const auto& e : el[K])
                            v.insert(v.end(), {cK, cI, static_cast<int64_t>(K),
                                               static_cast<int64_t>(e.sign),
                                               static_cast<int64_t>(e.pq),
                                               static_cast<int64_t>(e.J)});
                }
        } else if (kind == K_1H2) {
            const auto& addrK = spin == 0 ? L.alpha_address_1h() : L.beta_address_1h();
            const auto& addrI = spin == 0 ? L.alpha_address() : L.beta_address();
            for (int cK = 0; cK < addrK->nclasses(); ++cK)
                for (int cI = 0; cI < addrI->nclasses(); ++cI) {
                    const auto& el =
                        spin == 0 ? L.get_alpha_1h_list2(cK, cI) : L.get_beta_1h_list2(cK, cI);
                    for (size_t K = 0; K < el.size(); ++K)
                        for (const auto& e : el[K])
                            v.insert(v.end(), {cK, cI, static_cast<int64_t>(K), e.sign, e.p, e.J});
                }
        } else if (kind == K_1H || kind == K_2H || kind == K_3H) {
            // hole lists are keyed (class_K, add_K, class_I); the accessor API only allows point
            // look-ups, so walk every possible key in key order (identical to std::map order)
            const int nh = kind == K_1H ? 1 : (kind == K_2H ? 2 : 3);
            const auto addrK = nh == 1 ? (spin == 0 ? L.alpha_address_1h() : L.beta_address_1h())
                               : nh == 2
                                   ? (spin == 0 ? L.alpha_address_2h() : L.beta_address_2h())
                                   : (spin == 0 ? L.alpha_address_3h() : L.beta_address_3h());
            const auto& addrI = spin == 0 ? L.alpha_address() : L.beta_address();
            for (int cK = 0; cK < addrK->nclasses(); ++cK)
                for (size_t aK = 0; aK < addrK->strpcls(cK); ++aK)
                    for (int cI = 0; cI < addrI->nclasses(); ++cI) {
                        if (nh == 1) {
                            const auto& vec = spin == 0 ? L.get_alpha_1h_list(cK, aK, cI)
                                                        : L.get_beta_1h_list(cK, aK, cI);
                            for (const auto& e : vec)
                                v.insert(v.end(),
                                         {cK, static_cast<int64_t>(aK), cI, e.sign, e.p, e.J});
                        } else if (nh == 2) {
                            const auto& vec = spin == 0 ? L.get_alpha_2h_list(cK, aK, cI)
                                                        : L.get_beta_2h_list(cK, aK, cI);
                            for (const auto& e : vec)
                                v.insert(v.end(), {cK, static_cast<int64_t>(aK), cI, e.sign, e.p,
                                                   e.q, static_cast<int64_t>(e.J)});
                        } else {
                            const auto& vec = spin == 0 ? L.get_alpha_3h_list(cK, aK, cI)
                                                        : L.get_beta_3h_list(cK, aK, cI);
                            for (const auto& e : vec)
                                v.insert(v.end(), {cK, static_cast<int64_t>(aK), cI, e.sign, e.p,
                                                   e.q, e.r, static_cast<int64_t>(e.J)});
                        }
                    }
        } else {
            g_err = "unknown export kind";
            return -1;
        }
    } catch (const std::exception& e) {
        g_err = e.what();
        return -1;
    }
    if (out && cap >= static_cast<int64_t>(v.size()))
        std::memcpy(out, v.data(), v.size() * sizeof(int64_t));
    return v.size();
}

// alpha/beta occupation words of every determinant in CI-vector order (2 x ndet uint64)
void ref_strings_determinants(void* h, uint64_t* out) {
    const auto& L = *static_cast<RefStrings*>(h)->s;
    auto dets = L.make_determinants();
    for (size_t i = 0; i < dets.size(); ++i) {
        out[2 * i] = dets[i].a_string().get_word(0);
        out[2 * i + 1] = dets[i].b_string().get_word(0);
    }
}

void* ref_sigma_create(void* hs, double E, const double* H, const double* V, int log_level) {
    try {
        auto* rs = static_cast<RefStrings*>(hs);
        const size_t n = rs->s->norb();
        auto r = std::make_unique<RefSigma>();
        r->strings = rs;
        r->Hbuf.assign(H, H + n * n);
        r->Vbuf.assign(V, V + n * n * n * n);
        size_t s2[2] = {n, n}, s4[4] = {n, n, n, n};
        r->H = np_matrix(r->Hbuf.data(), 2, s2);
        r->V = np_tensor4(r->Vbuf.data(), 4, s4);
        r->b = std::make_unique<CISigmaBuilder>(*rs->s, E, r->H, r->V, log_level);
        return r.release();
    } catch (const std::exception& e) {
        g_err = e.what();
        return nullptr;
    }
}

void ref_sigma_destroy(void* h) { delete static_cast<RefSigma*>(h); }

// status: 0 ok, 1 invalid_argument (ValueError), 2 runtime_error, 3 overflow_error
static int status_of(const std::exception& e) {
    g_err = e.what();
    if (dynamic_cast<const std::invalid_argument*>(&e))
        return 1;
    if (dynamic_cast<const std::overflow_error*>(&e))
        return 3;
    return 2;
}

int ref_sigma_set_algorithm(void* h, const char* name) {
    try {
        static_cast<RefSigma*>(h)->b->set_algorithm(name);
        return 0;
    } catch (const std::exception& e) {
        return status_of(e);
    }
}

const char* ref_sigma_get_algorithm(void* h) {
    g_err = static_cast<RefSigma*>(h)->b->get_algorithm();
    return g_err.c_str();
}

int ref_sigma_set_memory(void* h, int mb) {
    try {
        static_cast<RefSigma*>(h)->b->set_memory(mb);
        return 0;
    } catch (const std::exception& e) {
        return status_of(e);
    }
}

int ref_sigma_apply(void* h, double* C, double* sigma, int64_t n) {
    try {
        static_cast<RefSigma*>(h)->b->Hamiltonian(wrap_vec(C, n), wrap_vec(sigma, n));
        return 0;
    } catch (const std::exception& e) {
        return status_of(e);
    }
}

void ref_sigma_times(void* h, double* out4) {
    auto t = static_cast<RefSigma*>(h)->b->avg_build_time();
    for (int i = 0; i < 4; ++i)
        out4[i] = t[i];
}

// SlaterRules::energy of every determinant (CI-vector order)
void ref_det_energies(void* hsig, double* out) {
    auto* r = static_cast<RefSigma*>(hsig);
    const size_t n = r->strings->s->norb();
    SlaterRules sr(n, 0.0, r->H, r->V);
    (void)sr;
    // use a fresh SlaterRules with the builder's scalar: rebuild from stored integrals
    auto dets = r->strings->s->make_determinants();
    // scalar energy is not retrievable from the builder; the caller adds E itself
    for (size_t i = 0; i < dets.size(); ++i)
        out[i] = sr.energy(dets[i]);
}

// dense <I|H|J> over the CI space from SlaterRules (independent of the sigma algorithms); E added
// on the diagonal by SlaterRules when scalar != 0
void ref_dense_hamiltonian(void* hsig, double E, double* out) {
    auto* r = static_cast<RefSigma*>(hsig);
    const size_t n = r->strings->s->norb();
    SlaterRules sr(n, E, r->H, r->V);
    auto dets = r->strings->s->make_determinants();
    const size_t nd = dets.size();
    for (size_t i = 0; i < nd; ++i)
        for (size_t j = 0; j < nd; ++j)
            out[i * nd + j] = (i == j) ? sr.energy(dets[i]) : sr.slater_rules(dets[i], dets[j]);
}

// ---- CSF spin adapter (forte2/determinant/ci_spin_adapter.cc:69-97) ----
void* ref_adapter_create(void* hs, int twoS, int twoMs) {
    try {
        auto* rs = static_cast<RefStrings*>(hs);
        auto r = std::make_unique<RefAdapter>();
        r->a = std::make_unique<CISpinAdapter>(twoS, twoMs, rs->s->norb());
        r->a->set_log_level(0);
        r->dets = rs->s->make_determinants();
        r->a->prepare_couplings(r->dets);
        return r.release();
    } catch (const std::exception& e) {
        g_err = e.what();
        return nullptr;
    }
}
void ref_adapter_destroy(void* h) { delete static_cast<RefAdapter*>(h); }
int64_t ref_adapter_ncsf(void* h) { return static_cast<RefAdapter*>(h)->a->ncsf(); }
int64_t ref_adapter_ndet(void* h) { return static_cast<RefAdapter*>(h)->a->ndet(); }
void ref_adapter_csf_to_det(void* h, double* csf, int64_t ncsf, double* det, int64_t ndet) {
    static_cast<RefAdapter*>(h)->a->csf_C_to_det_C(wrap_vec(csf, ncsf), wrap_vec(det, ndet));
}
void ref_adapter_det_to_csf(void* h, double* det, int64_t ndet, double* csf, int64_t ncsf) {
    static_cast<RefAdapter*>(h)->a->det_C_to_csf_C(wrap_vec(det, ndet), wrap_vec(csf, ncsf));
}
// CSF table as CSR: bounds[ncsf+1], then (det index, coefficient) pairs
int64_t ref_adapter_table(void* h, int64_t* bounds, int64_t* det_idx, double* coeff, int64_t cap) {
    auto* r = static_cast<RefAdapter*>(h);
    const size_t ncsf = r->a->ncsf();
    int64_t k = 0;
    for (size_t i = 0; i < ncsf; ++i) {
        if (bounds)
            bounds[i] = k;
        for (const auto& [d, c] : r->a->csf(i)) {
            if (det_idx && k < cap) {
                det_idx[k] = d;
                coeff[k] = c;
            }
            ++k;
        }
    }
    if (bounds)
        bounds[ncsf] = k;
    return k;
}
// approximate CSF-basis diagonal, forte2/ci/ci_sigma_builder.cc:317-370
void ref_hdiag_csf(void* hsig, void* hadapter, int full, double* out) {
    auto* r = static_cast<RefSigma*>(hsig);
    auto* a = static_cast<RefAdapter*>(hadapter);
    auto v = r->b->form_Hdiag_csf(a->dets, *a->a, full != 0);
    std::memcpy(out, v.data(), v.shape(0) * sizeof(double));
}
// dense CSF-basis Hamiltonian and one element of it, forte2/ci/ci_sigma_builder.cc:278-297,372-392
// returns 0, or 1 when the reference throws (its memory check); the message through ref_last_error
int ref_form_H_csf(void* hsig, void* hadapter, double* out) {
    auto* r = static_cast<RefSigma*>(hsig);
    auto* a = static_cast<RefAdapter*>(hadapter);
    try {
        auto m = r->b->form_H_csf(a->dets, *a->a);
        std::memcpy(out, m.data(), m.shape(0) * m.shape(1) * sizeof(double));
        return 0;
    } catch (const std::exception& e) {
        g_err = e.what();
        return 1;
    }
}
double ref_slater_rules_csf(void* hsig, void* hadapter, int64_t I, int64_t J) {
    auto* r = static_cast<RefSigma*>(hsig);
    auto* a = static_cast<RefAdapter*>(hadapter);
    return r->b->slater_rules_csf(a->dets, *a->a, static_cast<size_t>(I), static_cast<size_t>(J));
}
// Reduced density matrices (forte2/ci/ci_sigma_builder_1rdm.cc:9-84, ci_sigma_builder_2rdm.cc:45-282): the
// checker of the next widening step.  kind: 0 a_1rdm, 1 b_1rdm, 2 sf_1rdm (n x n);
// 3 aa_2rdm, 4 bb_2rdm (npairs x npairs, p > q); 5 ab_2rdm, 6 sf_2rdm, 7 sf_2cumulant (n^4).  Returns the number
// of doubles written (or needed when out is NULL), -1 when the reference throws.
int64_t ref_rdm(void* hsig, int kind, double* Cl, double* Cr, int64_t n, double* out, int64_t cap) {
    auto* r = static_cast<RefSigma*>(hsig);
    try {
        auto L = wrap_vec(Cl, n), R = wrap_vec(Cr, n);
        auto put = [&](const double* src, size_t count) -> int64_t {
            if (out && static_cast<int64_t>(count) <= cap)
                std::memcpy(out, src, count * sizeof(double));
            return static_cast<int64_t>(count);
        };
        switch (kind) {
        case 0: { auto m = r->b->compute_a_1rdm(L, R); return put(m.data(), m.shape(0) * m.shape(1)); }
        case 1: { auto m = r->b->compute_b_1rdm(L, R); return put(m.data(), m.shape(0) * m.shape(1)); }
        case 2: { auto m = r->b->compute_sf_1rdm(L, R); return put(m.data(), m.shape(0) * m.shape(1)); }
        case 3: { auto m = r->b->compute_aa_2rdm(L, R); return put(m.data(), m.shape(0) * m.shape(1)); }
        case 4: { auto m = r->b->compute_bb_2rdm(L, R); return put(m.data(), m.shape(0) * m.shape(1)); }
        case 5: { auto t = r->b->compute_ab_2rdm(L, R); return put(t.data(), t.shape(0) * t.shape(1) * t.shape(2) * t.shape(3)); }
        case 6: { auto t = r->b->compute_sf_2rdm(L, R); return put(t.data(), t.shape(0) * t.shape(1) * t.shape(2) * t.shape(3)); }
        case 7: { auto t = r->b->compute_sf_2cumulant(L, R); return put(t.data(), t.shape(0) * t.shape(1) * t.shape(2) * t.shape(3)); }
        case 8: { auto m = r->b->compute_aaa_3rdm(L, R); return put(m.data(), m.shape(0) * m.shape(1)); }
        case 9: { auto t = r->b->compute_aab_3rdm(L, R); return put(t.data(), t.shape(0) * t.shape(1) * t.shape(2) * t.shape(3)); }
        case 10: { auto t = r->b->compute_abb_3rdm(L, R); return put(t.data(), t.shape(0) * t.shape(1) * t.shape(2) * t.shape(3)); }
        case 11: { auto m = r->b->compute_bbb_3rdm(L, R); return put(m.data(), m.shape(0) * m.shape(1)); }
        case 12: { auto t = r->b->compute_sf_3rdm(L, R);
                   size_t c = 1; for (int d = 0; d < 6; ++d) c *= t.shape(d); return put(t.data(), c); }
        case 13: { auto t = r->b->compute_sf_3cumulant(L, R);
                   size_t c = 1; for (int d = 0; d < 6; ++d) c *= t.shape(d); return put(t.data(), c); }
        default: g_err = "ref_rdm: unknown kind"; return -1;
        }
    } catch (const std::exception& e) {
        g_err = e.what();
        return -1;
    }
}

// transition 1-RDMs between two CI spaces with the same orbitals and electron numbers (CISigmaBuilder::compute_*_1trdm,
// ci_sigma_builder_1trdm.cc:11-133): kind 0 alpha, 1 beta, 2 spin-free; left vector of the left builder's space
int64_t ref_trdm(void* hleft, void* hright, int kind, double* Cl, int64_t nl, double* Cr, int64_t nr, double* out,
                 int64_t cap) {
    auto* l = static_cast<RefSigma*>(hleft);
    auto* r = static_cast<RefSigma*>(hright);
    try {
        auto L = wrap_vec(Cl, nl), R = wrap_vec(Cr, nr);
        auto m = kind == 0 ? l->b->compute_a_1trdm(*r->b, L, R)
                           : (kind == 1 ? l->b->compute_b_1trdm(*r->b, L, R) : l->b->compute_sf_1trdm(*r->b, L, R));
        const int64_t count = static_cast<int64_t>(m.shape(0) * m.shape(1));
        if (out && count <= cap)
            std::memcpy(out, m.data(), count * sizeof(double));
        return count;
    } catch (const std::exception& e) {
        g_err = e.what();
        return -1;
    }
}

} // extern "C"
