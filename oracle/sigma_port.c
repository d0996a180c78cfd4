/* TEST INFRASTRUCTURE — NOT PART OF THE PRODUCT.  Only tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline leg may load this; nothing under forte2_b200/ does.
 *
 * Plain-C restatement of the reference's determinant-basis Hamiltonian action, sigma = H C, by
 * Slater rules over an explicit determinant list.  It restates
 *   forte2/determinant/slater_rules.cc:86-121   SlaterRules::energy            -> port_energy
 *   forte2/determinant/slater_rules.cc:127-143  singles_coupling_a/_b          -> single_coupling
 *   forte2/determinant/slater_rules.cc:154-200  SlaterRules::slater_rules      -> the four loops below
 *   forte2/determinant/bitarray.hpp:575-583     slater_sign                    -> sign_below
 * which is what the reference itself uses to cross-check its sigma algorithms
 * (tests/ci/test_ci_exact_diagonalization.py:51-79: "hz" == exact diagonalisation).
 * Determinant = (alpha word, beta word), bit p = orbital p, no cross-spin phase
 * (slater_rules.cc:168-173).  V is physicist <pq|rs>, C-contiguous.
 * Parity pinned: checked against oracle/_ref (the compiled reference) in tests/test_oracle.py and
 * against tests/golden/sigma_*.npz.
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

typedef struct {
    uint64_t a, b;
    int64_t idx;
} det_key;

static int cmp_key(const void* x, const void* y) {
    const det_key* p = (const det_key*)x;
    const det_key* q = (const det_key*)y;
    if (p->a != q->a)
        return p->a < q->a ? -1 : 1;
    if (p->b != q->b)
        return p->b < q->b ? -1 : 1;
    return 0;
}

static int64_t find_det(const det_key* keys, int64_t n, uint64_t a, uint64_t b) {
    int64_t lo = 0, hi = n - 1;
    while (lo <= hi) {
        const int64_t mid = (lo + hi) / 2;
        if (keys[mid].a == a && keys[mid].b == b)
            return keys[mid].idx;
        if (keys[mid].a < a || (keys[mid].a == a && keys[mid].b < b))
            lo = mid + 1;
        else
            hi = mid - 1;
    }
    return -1;
}

static double sign_below(uint64_t w, int n) {
    const uint64_t mask = n >= 64 ? ~0ull : ((1ull << n) - 1ull);
    return (__builtin_popcountll(w & mask) & 1) ? -1.0 : 1.0;
}

#define V4(p, q, r, s) V[(((size_t)(p) * n + (q)) * n + (r)) * n + (s)]

/* term selection for the multi-process tests: the Hamiltonian is the sum of these four parts */
enum { PART_SCALAR = 1, PART_ALPHA = 2, PART_BETA = 4, PART_ALPHA_BETA = 8, PART_ALL = 15 };

/* <I|H|I>, slater_rules.cc:86-121, split into scalar / alpha same-spin / beta same-spin / alpha-beta */
static double port_energy_parts(int n, double E, const double* H, const double* V, uint64_t a, uint64_t b,
                                int parts) {
    double e = (parts & PART_SCALAR) ? E : 0.0;
    for (int p = 0; p < n; ++p) {
        if ((a >> p) & 1) {
            if (parts & PART_ALPHA) {
                e += H[p * n + p];
                for (int q = 0; q < p; ++q)
                    if ((a >> q) & 1)
                        e += V4(q, p, q, p) - V4(q, p, p, q);
            }
            if (parts & PART_ALPHA_BETA)
                for (int q = 0; q < n; ++q)
                    if ((b >> q) & 1)
                        e += V4(p, q, p, q);
        }
        if (((b >> p) & 1) && (parts & PART_BETA)) {
            e += H[p * n + p];
            for (int q = 0; q < p; ++q)
                if ((b >> q) & 1)
                    e += V4(q, p, q, p) - V4(q, p, p, q);
        }
    }
    return e;
}

static double port_energy(int n, double E, const double* H, const double* V, uint64_t a, uint64_t b) {
    return port_energy_parts(n, E, H, V, a, b, PART_ALL);
}

/* h(i,a) + sum_j same-spin (<ij|aj> - <ij|ja>) + sum_j other-spin <ij|aj>, slater_rules.cc:127-143 */
static double single_coupling(int n, const double* H, const double* V, int i, int a, uint64_t same,
                              uint64_t other, int want_same, int want_other) {
    double c = want_same ? H[i * n + a] : 0.0;
    for (int j = 0; j < n; ++j) {
        if (want_same && ((same >> j) & 1))
            c += V4(i, j, a, j) - V4(i, j, j, a);
        if (want_other && ((other >> j) & 1))
            c += V4(i, j, a, j);
    }
    return c;
}

/* sigma[I] = sum_J <I|H|J> C[J] over the ndet determinants dets[2*I] (alpha), dets[2*I+1] (beta) */
int port_sigma_parts(int64_t ndet, const uint64_t* dets, int n, double E, const double* H, const double* V,
                     const double* C, double* sigma, int parts) {
    det_key* keys = (det_key*)malloc((size_t)(ndet > 0 ? ndet : 1) * sizeof(det_key));
    if (!keys)
        return 1;
    for (int64_t I = 0; I < ndet; ++I) {
        keys[I].a = dets[2 * I];
        keys[I].b = dets[2 * I + 1];
        keys[I].idx = I;
    }
    qsort(keys, (size_t)ndet, sizeof(det_key), cmp_key);
    for (int64_t I = 0; I < ndet; ++I) {
        const uint64_t a = dets[2 * I], b = dets[2 * I + 1];
        double acc = port_energy_parts(n, E, H, V, a, b, parts) * C[I];
        for (int spin = 0; spin < 2; ++spin) {
            const uint64_t s = spin == 0 ? a : b, o = spin == 0 ? b : a;
            const int want_same = parts & (spin == 0 ? PART_ALPHA : PART_BETA);
            const int want_ab = parts & PART_ALPHA_BETA;
            /* single replacements i -> x of this spin (slater_rules.cc:186-198) */
            for (int i = 0; i < n; ++i) {
                if (!((s >> i) & 1))
                    continue;
                const uint64_t s1 = s & ~(1ull << i);
                const double si = sign_below(s, i);
                for (int x = 0; x < n; ++x) {
                    if ((s >> x) & 1)
                        continue;
                    const uint64_t s2 = s1 | (1ull << x);
                    const double sg = si * sign_below(s1, x);
                    const int64_t J = spin == 0 ? find_det(keys, ndet, s2, o) : find_det(keys, ndet, o, s2);
                    if (J >= 0 && (want_same || want_ab))
                        acc += sg * single_coupling(n, H, V, i, x, s, o, want_same, want_ab) * C[J];
                    /* same-spin double replacements (i,j) -> (x,y), i<j, x<y (slater_rules.cc:174-184) */
                    for (int j = i + 1; j < n && want_same; ++j) {
                        if (!((s2 >> j) & 1) || j == x)
                            continue;
                        const uint64_t s3 = s2 & ~(1ull << j);
                        const double sj = sign_below(s2, j);
                        for (int y = x + 1; y < n; ++y) {
                            if ((s2 >> y) & 1 || y == i)
                                continue;
                            if ((s >> y) & 1)
                                continue;
                            const uint64_t s4 = s3 | (1ull << y);
                            const double sg2 = sg * sj * sign_below(s3, y);
                            const int64_t J2 =
                                spin == 0 ? find_det(keys, ndet, s4, o) : find_det(keys, ndet, o, s4);
                            if (J2 >= 0)
                                acc += sg2 * (V4(i, j, x, y) - V4(i, j, y, x)) * C[J2];
                        }
                    }
                    /* opposite-spin double replacement, counted once from the alpha side
                     * (slater_rules.cc:166-172): <i j|x y> with i,x alpha and j,y beta */
                    if (spin == 0 && want_ab) {
                        for (int j = 0; j < n; ++j) {
                            if (!((b >> j) & 1))
                                continue;
                            const uint64_t b1 = b & ~(1ull << j);
                            const double sbj = sign_below(b, j);
                            for (int y = 0; y < n; ++y) {
                                if ((b >> y) & 1)
                                    continue;
                                const double sgb = sbj * sign_below(b1, y);
                                const int64_t J3 = find_det(keys, ndet, s2, b1 | (1ull << y));
                                if (J3 >= 0)
                                    acc += sg * sgb * V4(i, j, x, y) * C[J3];
                            }
                        }
                    }
                }
            }
        }
        sigma[I] = acc;
    }
    free(keys);
    return 0;
}

int port_sigma(int64_t ndet, const uint64_t* dets, int n, double E, const double* H, const double* V,
               const double* C, double* sigma) {
    return port_sigma_parts(ndet, dets, n, E, H, V, C, sigma, PART_ALL);
}

/* determinant energies for every determinant (SlaterRules::energies, slater_rules.cc:145-152) */
void port_energies(int64_t ndet, const uint64_t* dets, int n, double E, const double* H, const double* V,
                   double* out) {
    for (int64_t I = 0; I < ndet; ++I)
        out[I] = port_energy(n, E, H, V, dets[2 * I], dets[2 * I + 1]);
}

/* ---------------------------------------------------------------------------------------------------
 * Sampled checker (sizes no full oracle sigma can reach: CAS(16,16), CAS(18,18), 3 GAS / 20 orbitals).
 * For nsamp determinants I = (a[s], b[s]) list every determinant J of the FULL Fock sector (same na, nb)
 * that H connects to I -- I itself, the single replacements of either spin, the same-spin and the
 * opposite-spin double replacements -- together with <I|H|J> (SlaterRules::slater_rules,
 * forte2/determinant/slater_rules.cc:154-245; phases: bitarray.hpp:575-583).  The caller (oracle/sampled.py)
 * maps (alpha word, beta word) to CI-vector addresses, drops the J outside the CI space (P H P, what
 * the reference's sigma computes) and sums <I|H|J> C[J].
 * Every sample produces the same count: 1 + Sa + Sb + Daa + Dbb + Sa Sb with S = N (n - N),
 * D = C(N,2) C(n-N,2).  Returns that count (out_* may be NULL to size the buffers). */
static int64_t connections_one(int n, double E, const double* H, const double* V, uint64_t wa, uint64_t wb,
                               uint64_t* oa, uint64_t* ob, double* oh) {
    int64_t k = 0;
#define EMIT(xa, xb, val)                                                                                    \
    do {                                                                                                     \
        if (oa) {                                                                                            \
            oa[k] = (xa);                                                                                    \
            ob[k] = (xb);                                                                                    \
            oh[k] = (val);                                                                                   \
        }                                                                                                    \
        ++k;                                                                                                 \
    } while (0)
    EMIT(wa, wb, port_energy(n, E, H, V, wa, wb));
    for (int spin = 0; spin < 2; ++spin) {
        const uint64_t w = spin == 0 ? wa : wb, o = spin == 0 ? wb : wa;
        for (int i = 0; i < n; ++i) {
            if (!((w >> i) & 1))
                continue;
            const uint64_t w1 = w & ~(1ull << i);
            const double si = sign_below(w, i);
            for (int x = 0; x < n; ++x) {
                if ((w >> x) & 1)
                    continue;
                const uint64_t w2 = w1 | (1ull << x);
                const double sg = si * sign_below(w1, x);
                const double h1 = sg * single_coupling(n, H, V, i, x, w, o, 1, 1);
                if (spin == 0)
                    EMIT(w2, wb, h1);
                else
                    EMIT(wa, w2, h1);
                /* same-spin doubles (i < j) -> (x < y) */
                for (int j = i + 1; j < n; ++j) {
                    if (!((w >> j) & 1))
                        continue;
                    const uint64_t w3 = w2 & ~(1ull << j);
                    const double sj = sign_below(w2, j);
                    for (int y = x + 1; y < n; ++y) {
                        if ((w >> y) & 1)
                            continue;
                        const uint64_t w4 = w3 | (1ull << y);
                        const double h2 = sg * sj * sign_below(w3, y) * (V4(i, j, x, y) - V4(i, j, y, x));
                        if (spin == 0)
                            EMIT(w4, wb, h2);
                        else
                            EMIT(wa, w4, h2);
                    }
                }
                /* opposite-spin doubles, once, from the alpha side */
                if (spin == 0)
                    for (int j = 0; j < n; ++j) {
                        if (!((wb >> j) & 1))
                            continue;
                        const uint64_t b1 = wb & ~(1ull << j);
                        const double sbj = sign_below(wb, j);
                        for (int y = 0; y < n; ++y) {
                            if ((wb >> y) & 1)
                                continue;
                            EMIT(w2, b1 | (1ull << y), sg * sbj * sign_below(b1, y) * V4(i, j, x, y));
                        }
                    }
            }
        }
    }
#undef EMIT
    return k;
}

int64_t port_connections(int n, double E, const double* H, const double* V, int64_t nsamp, const uint64_t* a,
                         const uint64_t* b, uint64_t* out_a, uint64_t* out_b, double* out_h) {
    if (nsamp <= 0)
        return 0;
    const int64_t per = connections_one(n, E, H, V, a[0], b[0], NULL, NULL, NULL);
    if (!out_a)
        return per;
    for (int64_t s = 0; s < nsamp; ++s)
        if (connections_one(n, E, H, V, a[s], b[s], out_a + s * per, out_b + s * per, out_h + s * per) != per)
            return -1; /* samples with different electron counts */
    return per;
}
