// Slater-Condon matrix elements <I|H|J> between determinants of the CI space, on the host.
// Used for the small guess Hamiltonian of the Davidson-Liu solver (forte2/ci/ci.py:411-452 calls
// SlaterRules::slater_rules, forte2/determinant/slater_rules.cc:123-245, for the lowest-energy
// determinants); O(n_guess^2) elements, never on the hot path.  No CUDA calls in this file.
//
// Conventions: V[p,q,r,s] = <pq|rs> physicist order, a determinant is
// an alpha word and a beta word, the phase of a mixed-spin replacement is the product of the two
// same-spin phases (no cross-spin phase).
#include <cstdint>
#include <vector>

#include "handles.hpp"

namespace f2 {
namespace {

struct Ints {
    int n;
    double E;
    const double *H, *V;
    double h(int p, int q) const { return H[p * n + q]; }
    double v(int p, int q, int r, int s) const {
        return V[((static_cast<size_t>(p) * n + q) * n + r) * n + s];
    }
};

// phase of a+_a a_i acting on word w (i occupied, a empty): (-1)^(occupied orbitals strictly between)
int single_phase(uint64_t w, int i, int a) {
    const int lo = i < a ? i : a, hi = i < a ? a : i;
    const uint64_t between = (hi - lo <= 1) ? 0ull : (((1ull << hi) - 1ull) & ~((1ull << (lo + 1)) - 1ull));
    return (__builtin_popcountll(w & between) & 1) ? -1 : 1;
}

double diagonal(const Ints& I, uint64_t a, uint64_t b) {
    double e = I.E;
    for (int p = 0; p < I.n; ++p) {
        const bool pa = (a >> p) & 1, pb = (b >> p) & 1;
        if (pa)
            e += I.h(p, p);
        if (pb)
            e += I.h(p, p);
        for (int q = 0; q < I.n; ++q) {
            const bool qa = (a >> q) & 1, qb = (b >> q) & 1;
            if (q < p) {
                if (pa && qa)
                    e += I.v(p, q, p, q) - I.v(p, q, q, p);
                if (pb && qb)
                    e += I.v(p, q, p, q) - I.v(p, q, q, p);
            }
            if (pa && qb)
                e += I.v(p, q, p, q);
        }
    }
    return e;
}

// one spin differs by i -> x (i in I, x in J); `same` is that spin's word of I, `other` the other spin's
double single(const Ints& I, uint64_t same, uint64_t other, int i, int x) {
    double e = I.h(i, x);
    for (int p = 0; p < I.n; ++p) {
        if (((same >> p) & 1) && p != i)
            e += I.v(i, p, x, p) - I.v(i, p, p, x);
        if ((other >> p) & 1)
            e += I.v(i, p, x, p);
    }
    return single_phase(same, i, x) * e;
}

double element(const Ints& I, uint64_t Ia, uint64_t Ib, uint64_t Ja, uint64_t Jb) {
    const int da = __builtin_popcountll(Ia ^ Ja) / 2, db = __builtin_popcountll(Ib ^ Jb) / 2;
    if (da + db > 2)
        return 0.0;
    if (da + db == 0)
        return diagonal(I, Ia, Ib);
    auto first = [](uint64_t w) { return __builtin_ctzll(w); };
    if (da == 1 && db == 0)
        return single(I, Ia, Ib, first(Ia & ~Ja), first(Ja & ~Ia));
    if (da == 0 && db == 1)
        return single(I, Ib, Ia, first(Ib & ~Jb), first(Jb & ~Ib));
    if (da == 1 && db == 1) {
        const int i = first(Ia & ~Ja), x = first(Ja & ~Ia), j = first(Ib & ~Jb), y = first(Jb & ~Ib);
        return single_phase(Ia, i, x) * single_phase(Ib, j, y) * I.v(i, j, x, y);
    }
    // two replacements in one spin: i < j leave, x < y enter
    const bool alpha = da == 2;
    const uint64_t w = alpha ? Ia : Ib, w2 = alpha ? Ja : Jb;
    uint64_t out = w & ~w2, in = w2 & ~w;
    const int i = first(out);
    out &= out - 1;
    const int j = first(out);
    const int x = first(in);
    in &= in - 1;
    const int y = first(in);
    // i -> x on w, then j -> y on the result
    const uint64_t mid = (w & ~(1ull << i)) | (1ull << x);
    const int ph = single_phase(w, i, x) * single_phase(mid, j, y);
    return ph * (I.v(i, j, x, y) - I.v(i, j, y, x));
}

} // namespace
} // namespace f2

using namespace f2;

extern "C" int f2_slater_matrix(const f2_strings* h, double E, const double* H, const double* V,
                                const int64_t* addresses, int n, double* out) {
    try {
        const CIStringsHost& S = *h->host;
        const Ints I{S.norb, E, H, V};
        std::vector<uint64_t> a(n), b(n);
        for (int k = 0; k < n; ++k) {
            if (addresses[k] < 0 || addresses[k] >= S.ndet)
                throw std::out_of_range("f2_slater_matrix: determinant address out of range");
            S.determinant(addresses[k], a[k], b[k]);
        }
        for (int r = 0; r < n; ++r)
            for (int c = 0; c <= r; ++c) {
                const double x = element(I, a[r], b[r], a[c], b[c]);
                out[static_cast<size_t>(r) * n + c] = x;
                out[static_cast<size_t>(c) * n + r] = x;
            }
        return F2_OK;
    } catch (const std::exception& e) {
        return status_of(e);
    }
}

// <I|S^2|J> between determinants of the CI space, on the host: the spin control of determinant-basis solves
// (forte2/sci/sci.py:405-455 builds the same matrix over its guess determinants with det.spin2,
// forte2/determinant/determinant_helpers.cc:154-203).  S^2 = S- S+ + Sz (Sz + 1):
//   diagonal      Ms (Ms + 1) + (number of singly occupied orbitals that carry a beta electron)
//   off-diagonal  -phase when J is I with the spins of two singly occupied orbitals exchanged, 0 otherwise;
//                 phase = product of the four same-spin string phases (no cross-spin phase, as everywhere here)
extern "C" int f2_spin2_matrix(const f2_strings* h, const int64_t* addresses, int n, double* out) {
    try {
        const CIStringsHost& S = *h->host;
        std::vector<uint64_t> a(n), b(n);
        for (int k = 0; k < n; ++k) {
            if (addresses[k] < 0 || addresses[k] >= S.ndet)
                throw std::out_of_range("f2_spin2_matrix: determinant address out of range");
            S.determinant(addresses[k], a[k], b[k]);
        }
        const double ms = 0.5 * (S.na - S.nb);
        auto below = [](uint64_t w, int p) { return __builtin_popcountll(w & ((1ull << p) - 1ull)); };
        for (int r = 0; r < n; ++r)
            for (int c = 0; c <= r; ++c) {
                double x = 0.0;
                const uint64_t da = a[r] ^ a[c], db = b[r] ^ b[c];
                if (da == 0 && db == 0) {
                    x = ms * (ms + 1.0) + __builtin_popcountll(b[r] & ~a[r]);
                } else if (da == db && __builtin_popcountll(da) == 2 && (a[r] & da) == (b[c] & db) &&
                           (a[c] & da) == (b[r] & db) && __builtin_popcountll(a[r] & da) == 1) {
                    const int j = __builtin_ctzll(a[r] & da); // alpha in r, beta in c
                    const int i = __builtin_ctzll(a[c] & da); // alpha in c, beta in r
                    const int par = below(a[c], i) + below(b[c], j) + below(a[r], j) + below(b[r], i);
                    x = (par & 1) ? 1.0 : -1.0;
                }
                out[static_cast<size_t>(r) * n + c] = x;
                out[static_cast<size_t>(c) * n + r] = x;
            }
        return F2_OK;
    } catch (const std::exception& e) {
        return status_of(e);
    }
}
