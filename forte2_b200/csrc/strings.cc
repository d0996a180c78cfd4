// Host-side CI string model; see strings.hpp for the contract and the reference citations.
#include "strings.hpp"

#include <algorithm>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstdlib>
#include <cstring>
#include <numeric>

namespace f2 {

// ------------------------------------------------------------------------------------------
// occupation patterns (forte2/ci/ci_occupation.cc:35-142)
// ------------------------------------------------------------------------------------------
namespace {

struct OccBuilder {
    occ_t lo{}, hi{}, size{};
    int nspaces = 0;
    std::vector<occ_t>*a, *b;
    std::vector<std::pair<size_t, size_t>>* pairs;

    static size_t intern(std::vector<occ_t>& v, const occ_t& o) {
        auto it = std::find(v.begin(), v.end(), o);
        if (it != v.end())
            return static_cast<size_t>(it - v.begin());
        v.push_back(o);
        return v.size() - 1;
    }

    // depth-first over the spaces; every space but the last enumerates (n_alpha, n_beta) with
    // n_alpha ascending (outer) and n_beta ascending (inner); the last space takes what is left
    void walk(int space, int ra, int rb, occ_t oa, occ_t ob) {
        if (space == nspaces - 1) {
            const int tot = ra + rb;
            if (tot <= hi[space] && tot >= lo[space] && std::max(ra, rb) <= size[space] &&
                std::min(ra, rb) >= 0) {
                oa[space] = ra;
                ob[space] = rb;
                const size_t ia = intern(*a, oa);
                const size_t ib = intern(*b, ob);
                pairs->emplace_back(ia, ib);
            }
            return;
        }
        const int a_lo = std::max(lo[space] - size[space], 0);
        const int a_hi = std::min(hi[space], size[space]);
        for (int xa = a_lo; xa <= a_hi; ++xa) {
            const int b_lo = std::max(lo[space] - xa, 0);
            const int b_hi = std::min(hi[space] - xa, size[space]);
            for (int xb = b_lo; xb <= b_hi; ++xb) {
                if (xa + xb > hi[space] || xa + xb < lo[space])
                    continue;
                oa[space] = xa;
                ob[space] = xb;
                walk(space + 1, ra - xa, rb - xb, oa, ob);
            }
        }
    }
};

} // namespace

static void occupation_patterns(int na, int nb, const std::vector<int>& gmin,
                                const std::vector<int>& gmax, const std::vector<int>& gsize,
                                int& nspaces, std::vector<occ_t>& a, std::vector<occ_t>& b,
                                std::vector<std::pair<size_t, size_t>>& pairs) {
    OccBuilder ob;
    ob.a = &a;
    ob.b = &b;
    ob.pairs = &pairs;
    for (int n = 0; n < kMaxGas; ++n) {
        const int sz = n < static_cast<int>(gsize.size()) ? gsize[n] : 0;
        if (!sz)
            continue;
        int emax = std::min(2 * sz, na + nb);
        if (static_cast<int>(gmax.size()) > n)
            emax = std::min(gmax[n], emax);
        int emin = 0;
        if (static_cast<int>(gmin.size()) > n)
            emin = gmin[n];
        ob.hi[n] = emax;
        ob.lo[n] = emin;
        ob.size[n] = sz;
        ob.nspaces += 1;
    }
    if (ob.nspaces == 0) { // no orbitals at all: one empty pattern (ci_occupation.cc:136-140)
        nspaces = 1;
        a.assign(1, occ_t{});
        b.assign(1, occ_t{});
        pairs.assign(1, {0, 0});
        return;
    }
    nspaces = ob.nspaces;
    ob.walk(0, na, nb, occ_t{}, occ_t{});
}

std::vector<occ_t> one_hole_occupations(const std::vector<occ_t>& occ) {
    std::vector<occ_t> out;
    for (const auto& o : occ)
        for (int n = 0; n < kMaxGas; ++n)
            if (o[n] >= 1) {
                occ_t x = o;
                x[n] -= 1;
                if (std::find(out.begin(), out.end(), x) == out.end())
                    out.push_back(x);
            }
    return out;
}

std::vector<occ_t> one_hole_one_particle_occupations(const std::vector<occ_t>& occ) {
    int used = 0;
    for (const auto& o : occ)
        for (int n = 0; n < kMaxGas; ++n)
            if (o[n] > 0)
                used = std::max(used, n + 1);
    std::vector<occ_t> out;
    for (const auto& o : occ)
        for (int n = 0; n < used; ++n)
            if (o[n] >= 1)
                for (int s = 0; s < used; ++s) {
                    occ_t x = o;
                    x[n] -= 1;
                    x[s] += 1;
                    if (std::find(out.begin(), out.end(), x) == out.end())
                        out.push_back(x);
                }
    return out;
}

// ------------------------------------------------------------------------------------------
// string spaces
// ------------------------------------------------------------------------------------------
void StringSpace::index() {
    cls_off.assign(cls.size() + 1, 0);
    for (size_t c = 0; c < cls.size(); ++c)
        cls_off[c + 1] = cls_off[c] + static_cast<uint32_t>(cls[c].size());
    const size_t n = cls_off.back();
    std::vector<uint32_t> order(n);
    std::vector<uint64_t> flat(n);
    std::vector<uint32_t> fcls(n), fadd(n);
    size_t k = 0;
    for (size_t c = 0; c < cls.size(); ++c)
        for (size_t i = 0; i < cls[c].size(); ++i, ++k) {
            flat[k] = cls[c][i];
            fcls[k] = static_cast<uint32_t>(c);
            fadd[k] = static_cast<uint32_t>(i);
        }
    std::iota(order.begin(), order.end(), 0u);
    std::sort(order.begin(), order.end(), [&](uint32_t x, uint32_t y) { return flat[x] < flat[y]; });
    key.resize(n);
    key_add.resize(n);
    key_cls.resize(n);
    for (size_t i = 0; i < n; ++i) {
        key[i] = flat[order[i]];
        key_add[i] = fadd[order[i]];
        key_cls[i] = fcls[order[i]];
    }
    built = true;
}

bool StringSpace::find(uint64_t s, uint32_t& add, uint32_t& c) const {
    auto it = std::lower_bound(key.begin(), key.end(), s);
    if (it == key.end() || *it != s)
        return false;
    const size_t i = static_cast<size_t>(it - key.begin());
    add = key_add[i];
    c = key_cls[i];
    return true;
}

int CIStringsHost::symmetry_of(uint64_t s) const {
    int h = 0;
    while (s) {
        const int p = __builtin_ctzll(s);
        h ^= mo_sym[p];
        s &= s - 1;
    }
    return h;
}

std::vector<occ_t> CIStringsHost::occs_of(int spin, int kind) const {
    const auto& base = spin == 0 ? aocc : bocc;
    if (kind == K_STR_N)
        return base;
    if (kind == K_STR_1H1P)
        return one_hole_one_particle_occupations(base);
    std::vector<occ_t> o = base;
    for (int i = 0; i < kind; ++i)
        o = one_hole_occupations(o);
    return o;
}

// Strings of one occupation pattern: per GAS space every way of placing that space's electrons,
// in lexicographic order of the bit sequence (orbital 0 most significant, empty < occupied),
// which starts from "highest orbitals occupied"; spaces combined with the last space fastest;
// then bucketed by irrep keeping that order (ci_strings_makers.cc:13-73).
void CIStringsHost::build_space(int spin, int kind) const {
    StringSpace& sp = sp_[spin][kind];
    if (sp.built)
        return;
    sp.nirrep = nirrep;
    sp.occs = occs_of(spin, kind);
    sp.cls.assign(sp.occs.size() * nirrep, {});
    for (size_t io = 0; io < sp.occs.size(); ++io) {
        const occ_t& occ = sp.occs[io];
        std::vector<std::vector<uint64_t>> per_space(ngas);
        bool empty = false;
        for (int g = 0; g < ngas; ++g) {
            const int m = gas_size[g], k = occ[g];
            if (k < 0 || k > m) {
                empty = true;
                continue;
            }
            std::vector<char> bits(m, 0);
            for (int i = m - k; i < m; ++i)
                bits[i] = 1;
            do {
                uint64_t w = 0;
                for (int i = 0; i < m; ++i)
                    if (bits[i])
                        w |= 1ull << (gas_start[g] + i);
                per_space[g].push_back(w);
            } while (std::next_permutation(bits.begin(), bits.end()));
        }
        if (empty)
            continue;
        // mixed-radix counter, last space fastest
        std::vector<size_t> idx(ngas, 0);
        while (true) {
            uint64_t w = 0;
            for (int g = 0; g < ngas; ++g)
                w |= per_space[g][idx[g]];
            sp.cls[io * nirrep + symmetry_of(w)].push_back(w);
            int g = ngas - 1;
            while (g >= 0 && ++idx[g] == per_space[g].size()) {
                idx[g] = 0;
                --g;
            }
            if (g < 0)
                break;
        }
    }
    sp.index();
}

const StringSpace& CIStringsHost::space(int spin, int kind) const {
    if (!sp_[spin][kind].built)
        build_space(spin, kind);
    return sp_[spin][kind];
}

CIStringsHost::CIStringsHost(int64_t na_, int64_t nb_, int symmetry_,
                             const std::vector<std::vector<int>>& orbital_symmetry,
                             const std::vector<int>& gas_min, const std::vector<int>& gas_max)
    : na(static_cast<int>(na_)), nb(static_cast<int>(nb_)), symmetry(symmetry_) {
    if (na_ < 0)
        throw std::invalid_argument("Number of alpha electrons must be non-negative.");
    if (nb_ < 0)
        throw std::invalid_argument("Number of beta electrons must be non-negative.");
    if (gas_min.size() > orbital_symmetry.size() || gas_max.size() > orbital_symmetry.size())
        throw std::invalid_argument(
            "CIStrings: The number of GAS spaces specified by gas_min (" +
            std::to_string(gas_min.size()) + ") and gas_max (" + std::to_string(gas_max.size()) +
            ") must be less than or equal to the number of orbital symmetries (" +
            std::to_string(orbital_symmetry.size()) + ")");
    if (orbital_symmetry.size() > static_cast<size_t>(kMaxGas))
        throw std::invalid_argument("CIStrings: at most 6 GAS spaces are supported");
    for (const auto& space : orbital_symmetry) {
        gas_start.push_back(norb);
        gas_size.push_back(static_cast<int>(space.size()));
        for (int s : space) {
            mo_sym.push_back(s);
            nirrep = std::max(nirrep, std::abs(s + 1));
        }
        norb += static_cast<int>(space.size());
    }
    if (norb > 64)
        throw std::invalid_argument("CIStrings: at most 64 orbitals are supported");
    for (int s : mo_sym)
        if (s < 0 || s >= nirrep)
            throw std::invalid_argument("CIStrings: negative orbital irrep label");
    // ngas = number of non-empty spaces (1 when there is none), as the reference reports it
    occupation_patterns(na, nb, gas_min, gas_max, gas_size, ngas, aocc, bocc, occ_pairs);
    while (static_cast<int>(gas_size.size()) < ngas) {
        gas_size.push_back(0);
        gas_start.push_back(norb);
    }

    build_space(0, K_STR_N);
    build_space(1, K_STR_N);
    const auto& A = sp_[0][K_STR_N];
    const auto& B = sp_[1][K_STR_N];
    nas = static_cast<int64_t>(A.nstr());
    nbs = static_cast<int64_t>(B.nstr());

    // determinant blocks (ci_string_class.cc:57-65): for each allowed (alpha pattern, beta
    // pattern) and each alpha irrep h: (class(apat,h), class(bpat, h ^ symmetry))
    block_of.assign(A.nclasses() * B.nclasses(), -1);
    for (const auto& [ia, ib] : occ_pairs)
        for (int h = 0; h < nirrep; ++h) {
            const int hb = h ^ symmetry;
            if (hb < 0 || hb >= nirrep)
                throw std::out_of_range("CIStrings: target symmetry is outside the irrep range");
            DetBlock blk;
            blk.cA = static_cast<int>(ia) * nirrep + h;
            blk.cB = static_cast<int>(ib) * nirrep + hb;
            blk.nA = static_cast<int64_t>(A.strpcls(blk.cA));
            blk.nB = static_cast<int64_t>(B.strpcls(blk.cB));
            blk.offset = ndet;
            ndet += blk.nA * blk.nB;
            block_of[blk.cA * B.nclasses() + blk.cB] = static_cast<int64_t>(blocks.size());
            blocks.push_back(blk);
        }
}

void CIStringsHost::determinants(uint64_t* out) const {
    const auto& A = space(0, K_STR_N);
    const auto& B = space(1, K_STR_N);
    for (const auto& blk : blocks) {
        uint64_t* o = out + 2 * blk.offset;
        for (int64_t i = 0; i < blk.nA; ++i)
            for (int64_t j = 0; j < blk.nB; ++j) {
                *o++ = A.cls[blk.cA][i];
                *o++ = B.cls[blk.cB][j];
            }
    }
}

void CIStringsHost::determinant(int64_t address, uint64_t& a, uint64_t& b) const {
    // same search as the reference (ci_strings.cc:209-230): first block whose end exceeds address
    size_t n = 0;
    int64_t start = 0;
    for (; n < blocks.size(); ++n) {
        if (start + blocks[n].nA * blocks[n].nB > address)
            break;
        start += blocks[n].nA * blocks[n].nB;
    }
    if (n == blocks.size())
        throw std::out_of_range("CIStrings::determinant: address out of range");
    const auto& blk = blocks[n];
    const int64_t shift = address - start;
    a = space(0, K_STR_N).cls[blk.cA][shift / blk.nB];
    b = space(1, K_STR_N).cls[blk.cB][shift % blk.nB];
}

int64_t CIStringsHost::determinant_index(uint64_t a, uint64_t b) const {
    uint32_t ia, ca, ib, cb;
    if (!space(0, K_STR_N).find(a, ia, ca) || !space(1, K_STR_N).find(b, ib, cb))
        throw std::out_of_range("CIStrings::determinant_index: string not in the CI space");
    const int64_t blk = block_of[ca * nclsB() + cb];
    if (blk < 0)
        throw std::runtime_error("Block index not found");
    return blocks[blk].offset + static_cast<int64_t>(ia) * blocks[blk].nB + ib;
}

// ------------------------------------------------------------------------------------------
// reference-format exports; row layouts documented in include/forte2_b200.h
// ------------------------------------------------------------------------------------------
namespace {
template <int W> void stable_sort_rows(std::vector<int64_t>& v, int nkey) {
    const size_t n = v.size() / W;
    std::vector<uint32_t> order(n);
    std::iota(order.begin(), order.end(), 0u);
    std::stable_sort(order.begin(), order.end(), [&](uint32_t x, uint32_t y) {
        for (int k = 0; k < nkey; ++k) {
            const int64_t a = v[size_t(x) * W + k], b = v[size_t(y) * W + k];
            if (a != b)
                return a < b;
        }
        return false;
    });
    std::vector<int64_t> out(v.size());
    for (size_t i = 0; i < n; ++i)
        std::memcpy(&out[i * W], &v[size_t(order[i]) * W], W * sizeof(int64_t));
    v.swap(out);
}
} // namespace

std::vector<int64_t> CIStringsHost::export_kind(int kind, int spin) const {
    std::vector<int64_t> v;
    if (kind >= K_STR_N && kind <= K_STR_1H1P) {
        const auto& sp = space(spin, kind);
        v.push_back(static_cast<int64_t>(sp.nclasses()));
        for (const auto& c : sp.cls)
            v.push_back(static_cast<int64_t>(c.size()));
        for (const auto& c : sp.cls)
            for (uint64_t s : c)
                v.push_back(static_cast<int64_t>(s));
        return v;
    }
    if (kind >= 20 && kind <= 28) {
        // device-format tables exactly as the kernels read them (DeviceLists, strings.hpp), for the host models
        // of the kernels in tests/: 20 as_ptr, 21 as_ent, 22 [width, bs_ent...], 23 h2_ptr[spin], 24 h2_ent[spin],
        // 25 class of every string of `spin`, 26 its address in the class, 27 block offsets [nclsA][nclsB]
        // (-1 = no block), 28 strings per class of `spin`
        const DeviceLists& d = device_lists();
        auto put = [&](const auto& src) {
            for (auto x : src)
                v.push_back(static_cast<int64_t>(x));
        };
        switch (kind) {
        case 20: put(d.as_ptr); break;
        case 21: put(d.as_ent); break;
        case 22: v.push_back(d.bs_width); put(d.bs_ent); break;
        case 23: put(d.h2_ptr[spin]); break;
        case 24: put(d.h2_ent[spin]); break;
        case 25: put(spin == 0 ? d.a_cls : d.b_cls); break;
        case 26: put(spin == 0 ? d.a_loc : d.b_loc); break;
        case 27: {
            std::vector<int64_t> off(nclsA() * nclsB(), -1);
            for (const auto& b : blocks)
                off[b.cA * nclsB() + b.cB] = b.offset;
            put(off);
            break;
        }
        case 28: {
            const auto& sp = space(spin, K_STR_N);
            for (size_t c = 0; c < sp.nclasses(); ++c)
                v.push_back(static_cast<int64_t>(sp.strpcls(c)));
            break;
        }
        }
        return v;
    }
    if (kind >= 30 && kind <= 37) {
        // hole lists of the reduced density matrices (HoleLists, strings.hpp): 30 + level = ptr, 34 + level = entries
        const HoleLists& h = hole_lists(spin, (kind - 30) & 3);
        if (kind < 34)
            for (auto x : h.ptr)
                v.push_back(static_cast<int64_t>(x));
        else
            for (auto x : h.ent)
                v.push_back(static_cast<int64_t>(x));
        return v;
    }
    const auto& N = space(spin, K_STR_N);
    if (kind == K_VO || kind == K_VO2) {
        // J = sign a+_p a_q I, (p,q) outer loops, strings inner (ci_strings_makers.cc:75-161)
        const auto& src = kind == K_VO ? N : space(spin, K_STR_1H1P);
        for (int p = 0; p < norb; ++p)
            for (int q = 0; q < norb; ++q)
                for (size_t c = 0; c < src.nclasses(); ++c)
                    for (size_t i = 0; i < src.cls[c].size(); ++i) {
                        uint64_t s = src.cls[c][i];
                        if (!((s >> q) & 1))
                            continue;
                        int sign = slater_sign(s, q);
                        s &= ~(1ull << q);
                        if ((s >> p) & 1)
                            continue;
                        sign *= slater_sign(s, p);
                        s |= 1ull << p;
                        uint32_t aj, cj;
                        if (!N.find(s, aj, cj))
                            continue;
                        if (kind == K_VO) {
                            v.insert(v.end(), {int64_t(c), int64_t(cj), p, q, sign, int64_t(i),
                                               int64_t(aj)});
                        } else {
                            const int hi = std::max(p, q), lo = std::min(p, q);
                            v.insert(v.end(), {int64_t(c), int64_t(cj), int64_t(i),
                                               sign * (p == q ? 2 : 1), int64_t(hi) * (hi + 1) / 2 + lo,
                                               int64_t(aj)});
                        }
                    }
        if (kind == K_VO)
            stable_sort_rows<7>(v, 4);
        else
            stable_sort_rows<6>(v, 3);
        return v;
    }
    if (kind == K_1H || kind == K_1H2) {
        if (norb == 0)
            return v;
        const auto& H = space(spin, K_STR_1H);
        for (size_t c = 0; c < N.nclasses(); ++c)
            for (size_t i = 0; i < N.cls[c].size(); ++i) {
                const uint64_t s = N.cls[c][i];
                for (int p = 0; p < norb; ++p) {
                    if (!((s >> p) & 1))
                        continue;
                    uint32_t ak, ck;
                    if (!H.find(s & ~(1ull << p), ak, ck))
                        continue;
                    const int sign = slater_sign(s, p);
                    if (kind == K_1H)
                        v.insert(v.end(), {int64_t(ck), int64_t(ak), int64_t(c), sign, p, int64_t(i)});
                    else
                        v.insert(v.end(), {int64_t(ck), int64_t(c), int64_t(ak), sign, p, int64_t(i)});
                }
            }
        stable_sort_rows<6>(v, 3);
        return v;
    }
    if (kind == K_2H) {
        const auto& H = space(spin, K_STR_2H);
        for (size_t c = 0; c < N.nclasses(); ++c)
            for (size_t i = 0; i < N.cls[c].size(); ++i) {
                const uint64_t s = N.cls[c][i];
                for (int q = 0; q < norb; ++q)
                    for (int p = q + 1; p < norb; ++p) {
                        if (!((s >> p) & 1) || !((s >> q) & 1))
                            continue;
                        const uint64_t s1 = s & ~(1ull << q);
                        const uint64_t s2 = s1 & ~(1ull << p);
                        uint32_t ak, ck;
                        if (!H.find(s2, ak, ck))
                            continue;
                        const int sign = slater_sign(s1, q) * slater_sign(s2, p);
                        v.insert(v.end(),
                                 {int64_t(ck), int64_t(ak), int64_t(c), sign, p, q, int64_t(i)});
                    }
            }
        stable_sort_rows<7>(v, 3);
        return v;
    }
    if (kind == K_3H) {
        const auto& H = space(spin, K_STR_3H);
        for (size_t c = 0; c < N.nclasses(); ++c)
            for (size_t i = 0; i < N.cls[c].size(); ++i) {
                const uint64_t s = N.cls[c][i];
                for (int r = 0; r < norb; ++r)
                    for (int q = r + 1; q < norb; ++q)
                        for (int p = q + 1; p < norb; ++p) {
                            if (!((s >> p) & 1) || !((s >> q) & 1) || !((s >> r) & 1))
                                continue;
                            const uint64_t s1 = s & ~(1ull << r);
                            const uint64_t s2 = s1 & ~(1ull << q);
                            const uint64_t s3 = s2 & ~(1ull << p);
                            uint32_t ak, ck;
                            if (!H.find(s3, ak, ck))
                                continue;
                            const int sign =
                                slater_sign(s1, r) * slater_sign(s2, q) * slater_sign(s3, p);
                            v.insert(v.end(), {int64_t(ck), int64_t(ak), int64_t(c), sign, p, q, r,
                                               int64_t(i)});
                        }
            }
        stable_sort_rows<8>(v, 3);
        return v;
    }
    throw std::invalid_argument("unknown export kind");
}

// ------------------------------------------------------------------------------------------
// device-format lists
// ------------------------------------------------------------------------------------------
const DeviceLists& CIStringsHost::device_lists() const {
    if (!dl_built_)
        build_device_lists();
    return dl_;
}

const HoleLists& CIStringsHost::hole_lists(int spin, int level) const {
    if (spin < 0 || spin > 1 || level < 0 || level > 3)
        throw std::invalid_argument("hole_lists: spin in {0, 1}, level in 0..3");
    HoleLists& h = holes_[spin][level];
    if (h.built)
        return h;
    const StringSpace& S = space(spin, K_STR_N);
    const int ne = spin == 0 ? na : nb;
    h.ptr.assign(1, 0);
    h.ent.clear();
    h.kword.clear();
    h.max_parents = 0;
    if (level == 0) {
        for (size_t g = 0; g < S.nstr(); ++g) {
            h.ent.push_back(static_cast<uint64_t>(g));
            h.ptr.push_back(static_cast<uint32_t>(h.ent.size()));
        }
        for (const auto& c : S.cls)
            h.kword.insert(h.kword.end(), c.begin(), c.end());
        h.max_parents = S.nstr() ? 1 : 0;
        h.built = true;
        return h;
    }
    if (ne >= level) {
        struct Item {
            uint64_t k, ent;
        };
        std::vector<Item> items;
        for (size_t c = 0; c < S.nclasses(); ++c)
            for (size_t x = 0; x < S.cls[c].size(); ++x) {
                const uint64_t w = S.cls[c][x];
                const uint64_t g = S.cls_off[c] + x;
                int occ[64], nocc = 0;
                for (int p = 0; p < norb; ++p)
                    if ((w >> p) & 1)
                        occ[nocc++] = p;
                // every increasing tuple o1 < o2 < o3 of occupied orbitals, removed in that order
                for (int i1 = 0; i1 < nocc; ++i1) {
                    const int o1 = occ[i1];
                    const int s1 = slater_sign(w, o1);
                    const uint64_t w1 = w & ~(1ull << o1);
                    if (level == 1) {
                        items.push_back({w1, g | (uint64_t(o1) << 32) | (uint64_t(s1 < 0) << 48)});
                        continue;
                    }
                    for (int i2 = i1 + 1; i2 < nocc; ++i2) {
                        const int o2 = occ[i2];
                        const int s2 = s1 * slater_sign(w1, o2);
                        const uint64_t w2 = w1 & ~(1ull << o2);
                        if (level == 2) {
                            const uint64_t op = uint64_t(o2) * (o2 - 1) / 2 + o1;
                            items.push_back({w2, g | (op << 32) | (uint64_t(s2 < 0) << 48)});
                            continue;
                        }
                        for (int i3 = i2 + 1; i3 < nocc; ++i3) {
                            const int o3 = occ[i3];
                            const int s3 = s2 * slater_sign(w2, o3);
                            const uint64_t w3 = w2 & ~(1ull << o3);
                            const uint64_t op = uint64_t(o3) * (o3 - 1) * (o3 - 2) / 6 + uint64_t(o2) * (o2 - 1) / 2 + o1;
                            if (op > 0xffff)
                                throw std::overflow_error("hole_lists: orbital triple index exceeds 16 bits");
                            items.push_back({w3, g | (op << 32) | (uint64_t(s3 < 0) << 48)});
                        }
                    }
                }
            }
        std::stable_sort(items.begin(), items.end(), [](const Item& a, const Item& b) { return a.k < b.k; });
        if (items.size() >= (size_t(1) << 32))
            throw std::overflow_error("hole_lists: more than 2^32 entries");
        for (size_t i = 0; i < items.size(); ++i) {
            if (i > 0 && items[i].k != items[i - 1].k)
                h.ptr.push_back(static_cast<uint32_t>(h.ent.size()));
            if (i == 0 || items[i].k != items[i - 1].k)
                h.kword.push_back(items[i].k);
            h.ent.push_back(items[i].ent);
        }
        if (!items.empty())
            h.ptr.push_back(static_cast<uint32_t>(h.ent.size()));
        for (size_t k = 0; k + 1 < h.ptr.size(); ++k)
            h.max_parents = std::max<int>(h.max_parents, h.ptr[k + 1] - h.ptr[k]);
    }
    h.built = true;
    return h;
}

void CIStringsHost::lists_prologue() const {
    DeviceLists& d = dl_;
    const auto& A = space(0, K_STR_N);
    const auto& B = space(1, K_STR_N);
    auto fill_cls = [](const StringSpace& S, std::vector<uint32_t>& c, std::vector<uint32_t>& l) {
        c.resize(S.nstr());
        l.resize(S.nstr());
        for (size_t k = 0; k < S.nclasses(); ++k)
            for (size_t i = 0; i < S.cls[k].size(); ++i) {
                c[S.cls_off[k] + i] = static_cast<uint32_t>(k);
                l[S.cls_off[k] + i] = static_cast<uint32_t>(i);
            }
    };
    fill_cls(A, d.a_cls, d.a_loc);
    fill_cls(B, d.b_cls, d.b_loc);
    if (B.nclasses() > 4096 || norb * norb > 65535)
        throw std::overflow_error("CIStrings: class or orbital-pair index exceeds the packed list format");
}

// one electron of this spin: the only same-spin operator is h_pq a+_p a_q; treat the vacuum as the single "hole"
// string whose parents are all one-electron strings and let the pair index carry the occupied orbital
void CIStringsHost::lists_one_electron(int spin) const {
    DeviceLists& d = dl_;
    const auto& S = space(spin, K_STR_N);
    auto& ptr = d.h2_ptr[spin];
    auto& ent = d.h2_ent[spin];
    ptr.assign(2, 0);
    ent.clear();
    for (size_t c = 0; c < S.nclasses(); ++c)
        for (size_t x = 0; x < S.cls[c].size(); ++x)
            ent.push_back(uint64_t(S.cls_off[c] + x) | (uint64_t(__builtin_ctzll(S.cls[c][x])) << 32));
    ptr[1] = static_cast<uint32_t>(ent.size());
}

// after h2_ptr / h2_ent of `spin` are in place: the longest parent list and the launch order of the (N-2)-electron
// strings that have parents.  Neighbouring strings share most of their parents, so strings that run at the same time
// (consecutive entries of the list) are taken far apart: a stride permutation (golden-ratio stride made coprime with
// the count) -- the reductions of concurrent CTAs then hit different rows of sigma.
void CIStringsHost::lists_epilogue(int spin) const {
    DeviceLists& d = dl_;
    const auto& ptr = d.h2_ptr[spin];
    auto& order = d.h2_color_list[spin];
    auto& cptr = d.h2_color_ptr[spin];
    order.clear();
    cptr.assign(1, 0);
    d.h2_max[spin] = 0;
    std::vector<uint32_t> live;
    for (size_t K = 0; K + 1 < ptr.size(); ++K) {
        const int m = static_cast<int>(ptr[K + 1] - ptr[K]);
        if (m > 0)
            live.push_back(static_cast<uint32_t>(K));
        d.h2_max[spin] = std::max(d.h2_max[spin], m);
    }
    const uint64_t n = live.size();
    if (n == 0)
        return;
    uint64_t stride = static_cast<uint64_t>(0.6180339887498949 * static_cast<double>(n)) | 1ull;
    while (std::gcd(stride, n) != 1)
        stride += 2;
    order.resize(n);
    for (uint64_t i = 0; i < n; ++i)
        order[i] = live[(i * stride) % n];
    cptr.push_back(static_cast<uint32_t>(n));
}

void CIStringsHost::build_device_lists() const {
    DeviceLists& d = dl_;
    lists_prologue();
    const auto& A = space(0, K_STR_N);
    const auto& B = space(1, K_STR_N);

    // alpha singles from the output row: Ja = sign a+_l a_k Ia  <=>  <Ia|E_kl|Ja> = sign
    d.as_ptr.assign(A.nstr() + 1, 0);
    d.as_ent.clear();
    d.as_max = 0;
    for (size_t c = 0; c < A.nclasses(); ++c)
        for (size_t i = 0; i < A.cls[c].size(); ++i) {
            const uint64_t s = A.cls[c][i];
            const size_t g = A.cls_off[c] + i;
            for (int k = 0; k < norb; ++k) {
                if (!((s >> k) & 1))
                    continue;
                const int sk = slater_sign(s, k);
                const uint64_t s1 = s & ~(1ull << k);
                for (int l = 0; l < norb; ++l) {
                    if ((s1 >> l) & 1)
                        continue;
                    const uint64_t s2 = s1 | (1ull << l);
                    uint32_t aj, cj;
                    if (!A.find(s2, aj, cj))
                        continue;
                    const int sign = sk * slater_sign(s1, l);
                    const uint64_t J = A.cls_off[cj] + aj;
                    d.as_ent.push_back(J | (uint64_t(k * norb + l) << 32) |
                                       (uint64_t(sign < 0) << 48));
                }
            }
            d.as_ptr[g + 1] = static_cast<uint32_t>(d.as_ent.size());
            d.as_max = std::max<int>(d.as_max, d.as_ptr[g + 1] - d.as_ptr[g]);
        }

    // beta scatter from the source column: Ib = sign a+_i a_j Jb  <=>  <Ib|E_ij|Jb> = sign
    {
        std::vector<std::vector<uint64_t>> per(B.nstr());
        size_t width = 0;
        for (size_t c = 0; c < B.nclasses(); ++c)
            for (size_t x = 0; x < B.cls[c].size(); ++x) {
                const uint64_t s = B.cls[c][x];
                auto& out = per[B.cls_off[c] + x];
                for (int j = 0; j < norb; ++j) {
                    if (!((s >> j) & 1))
                        continue;
                    const int sj = slater_sign(s, j);
                    const uint64_t s1 = s & ~(1ull << j);
                    for (int i = 0; i < norb; ++i) {
                        if ((s1 >> i) & 1)
                            continue;
                        uint32_t ai, ci;
                        if (!B.find(s1 | (1ull << i), ai, ci))
                            continue;
                        const int sign = sj * slater_sign(s1, i);
                        out.push_back(uint64_t(ai) | (uint64_t(ci) << 28) |
                                      (uint64_t(i * norb + j) << 40) | (uint64_t(sign < 0) << 56));
                    }
                }
                width = std::max(width, out.size());
            }
        d.bs_width = static_cast<int>(width);
        d.bs_ent.assign(width * B.nstr(), kNullEntry);
        for (size_t g = 0; g < B.nstr(); ++g)
            for (size_t t = 0; t < per[g].size(); ++t)
                d.bs_ent[t * B.nstr() + g] = per[g][t];
    }

    // two-hole members
    for (int spin = 0; spin < 2; ++spin) {
        const auto& S = spin == 0 ? A : B;
        const int ne = spin == 0 ? na : nb;
        auto& ptr = d.h2_ptr[spin];
        auto& ent = d.h2_ent[spin];
        ptr.assign(1, 0);
        ent.clear();
        if (ne == 1)
            lists_one_electron(spin);
        if (ne >= 2) {
            const auto& H = space(spin, K_STR_2H);
            const size_t nK = H.nstr();
            std::vector<std::vector<uint64_t>> per(nK);
            for (size_t c = 0; c < S.nclasses(); ++c)
                for (size_t x = 0; x < S.cls[c].size(); ++x) {
                    const uint64_t s = S.cls[c][x];
                    const size_t g = S.cls_off[c] + x;
                    for (int q = 0; q < norb; ++q)
                        for (int p = q + 1; p < norb; ++p) {
                            if (!((s >> p) & 1) || !((s >> q) & 1))
                                continue;
                            const uint64_t s1 = s & ~(1ull << q);
                            const uint64_t s2 = s1 & ~(1ull << p);
                            uint32_t ak, ck;
                            if (!H.find(s2, ak, ck))
                                continue;
                            const int sign = slater_sign(s1, q) * slater_sign(s2, p);
                            const uint32_t K = H.cls_off[ck] + ak;
                            per[K].push_back(uint64_t(g) | (uint64_t(p * (p - 1) / 2 + q) << 32) |
                                             (uint64_t(sign < 0) << 48));
                        }
                }
            ptr.assign(nK + 1, 0);
            for (size_t K = 0; K < nK; ++K) {
                ent.insert(ent.end(), per[K].begin(), per[K].end());
                ptr[K + 1] = static_cast<uint32_t>(ent.size());
            }
        }
        lists_epilogue(spin);
    }
    dl_built_ = true;
}

} // namespace f2
