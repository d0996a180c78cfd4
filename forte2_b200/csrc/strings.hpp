// Host-side model of the CI string spaces: GAS/ORMAS occupation patterns, string classes
// (occupation pattern x irrep), string enumeration in the reference's address order, string
// addressing, determinant blocks, and the sign-carrying replacement lists.
//
// Behavioural contract (bit-exact): forte2/ci/ci_occupation.cc:35-142,361-417,
// forte2/ci/ci_strings.cc:34-184, forte2/ci/ci_strings_makers.cc:13-322,
// forte2/ci/ci_string_class.cc:7-102, forte2/ci/ci_string_address.cc:6-23.
// The data structures are not the reference's (no hash maps, no std::map of vectors): strings
// live in flat per-class tables with a sorted index for look-ups, and every list is a flat
// CSR/ELL array that is uploaded to the GPU as is.
#pragma once
#include <array>
#include <cstdint>
#include <stdexcept>
#include <string>
#include <utility>
#include <vector>

namespace f2 {

constexpr int kMaxGas = 6; // forte2/ci/ci_occupation.h:10
using occ_t = std::array<int, kMaxGas>;

enum Kind : int {
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

inline int popcount_below(uint64_t w, int n) {
    return __builtin_popcountll(w & ((n >= 64) ? ~0ull : ((1ull << n) - 1ull)));
}
/// (-1)^(number of occupied orbitals below n); forte2/determinant/bitarray.hpp:575-583
inline int slater_sign(uint64_t w, int n) { return (popcount_below(w, n) & 1) ? -1 : 1; }

/// One family of strings (N-electron, 1h, 2h, 3h or 1h1p) of one spin.
struct StringSpace {
    int nirrep = 1;
    std::vector<occ_t> occs;                // occupation patterns, reference order
    std::vector<std::vector<uint64_t>> cls; // class -> strings in address order
    std::vector<uint32_t> cls_off;          // class -> first global index (size nclasses+1)
    // sorted look-up table: string -> (address in class, class)
    std::vector<uint64_t> key;
    std::vector<uint32_t> key_add, key_cls;
    bool built = false;

    size_t nclasses() const { return cls.size(); }
    size_t nstr() const { return cls_off.empty() ? 0 : cls_off.back(); }
    size_t strpcls(size_t c) const { return cls[c].size(); }
    bool find(uint64_t s, uint32_t& add, uint32_t& c) const;
    void index();
};

struct DetBlock {
    int cA, cB;
    int64_t offset, nA, nB;
};

/// Packed device-format list entries (see DESIGN.md "data layout in HBM").
/// alpha single replacement seen from the OUTPUT row Ia:  <Ia| E^a_kl |Ja> = sign
///   bits  0..31 Ja (global alpha index), 32..47 k*norb+l, bit 48 sign (1 = negative)
/// beta scatter entry seen from the SOURCE column Jb:   <Ib| E^b_ij |Jb> = sign
///   bits  0..27 Ib address inside its class, 28..39 class of Ib, 40..55 row m of the integral
///   matrix, bit 56 sign, all-ones = padding
/// two-hole member: string I = K + {p,q}, a_p a_q |I> = sign |K>
///   bits 0..31 I (global index), 32..47 pair index p(p-1)/2+q, bit 48 sign
constexpr uint64_t kNullEntry = ~0ull;

struct DeviceLists {
    // alpha singles, CSR over global alpha string index
    std::vector<uint32_t> as_ptr;
    std::vector<uint64_t> as_ent;
    int as_max = 0;
    // beta scatter, ELL: entry t of string Jb at [t * nbs + Jb]
    std::vector<uint64_t> bs_ent;
    int bs_width = 0;
    // two-hole members per spin, CSR over global 2h string index, plus the launch order of the 2h strings that
    // have parents (a stride permutation: strings that run concurrently share few parents)
    std::vector<uint32_t> h2_ptr[2];
    std::vector<uint64_t> h2_ent[2];
    std::vector<uint32_t> h2_color_ptr[2];  // {0, count}
    std::vector<uint32_t> h2_color_list[2]; // 2h string ids in launch order
    int h2_max[2] = {0, 0};
    // per global string: class and address in class
    std::vector<uint32_t> a_cls, a_loc, b_cls, b_loc;
};

/// Hole lists for the reduced density matrices: every (N - level)-electron string K that is a sub-string of some
/// N-electron string of one spin, CSR over K (K indexed by increasing occupation word), entry = one parent:
///   bits 0..31 I (global index of the N-electron string), 32..47 index of the decreasing tuple of the removed
///   orbitals (p; p (p-1)/2 + q; p (p-1) (p-2)/6 + q (q-1)/2 + r), bit 48 sign of removing them smallest first.
/// Level 0 is the identity (K = I).  Reference counterparts: the 1h / 2h / 3h lists of
/// forte2/ci/ci_strings_makers.cc:215-322 as the RDM code uses them (ci_sigma_builder_{1,2,3}rdm.cc).
struct HoleLists {
    std::vector<uint32_t> ptr;
    std::vector<uint64_t> ent;
    std::vector<uint64_t> kword; // occupation word of every K
    int max_parents = 0;
    bool built = false;
    size_t nK() const { return ptr.empty() ? 0 : ptr.size() - 1; }
};

class CIStringsHost {
  public:
    CIStringsHost(int64_t na, int64_t nb, int symmetry,
                  const std::vector<std::vector<int>>& orbital_symmetry,
                  const std::vector<int>& gas_min, const std::vector<int>& gas_max);

    int na, nb, symmetry, norb = 0, nirrep = 1, ngas = 0;
    std::vector<int> gas_size, gas_start, mo_sym;
    std::vector<occ_t> aocc, bocc;
    std::vector<std::pair<size_t, size_t>> occ_pairs;
    std::vector<DetBlock> blocks;
    std::vector<int64_t> block_of; // [cA * nclsB + cB] -> block id or -1
    int64_t ndet = 0, nas = 0, nbs = 0;

    /// kind in {N,1H,2H,3H,1H1P}; built on first use except N, 1H, 2H
    const StringSpace& space(int spin, int kind) const;
    size_t nclsA() const { return space(0, K_STR_N).nclasses(); }
    size_t nclsB() const { return space(1, K_STR_N).nclasses(); }

    /// Reference-format flat export (row layouts in include/forte2_b200.h).
    std::vector<int64_t> export_kind(int kind, int spin) const;
    /// (alpha word, beta word) of every determinant in CI-vector order
    void determinants(uint64_t* out) const;
    void determinant(int64_t address, uint64_t& a, uint64_t& b) const;
    int64_t determinant_index(uint64_t a, uint64_t b) const;

    /// Device-format lists; rowmap[i*norb+j] = row of the integral matrix used for E_ij.
    const DeviceLists& device_lists() const;

    int symmetry_of(uint64_t s) const;

    /// level in 0..3
    const HoleLists& hole_lists(int spin, int level) const;

    /// pieces of the list construction, for the device builder (csrc/lists_device.cu): it runs the prologue (class and
    /// address of every string), fills as_* / bs_* / h2_ptr / h2_ent of lists_storage() from the GPU, then the epilogue
    /// per spin, and marks the lists as built
    bool lists_built() const { return dl_built_; }
    DeviceLists& lists_storage() const { return dl_; }
    void lists_prologue() const;
    void lists_one_electron(int spin) const;
    void lists_epilogue(int spin) const;
    void lists_adopt() const { dl_built_ = true; }

  private:
    mutable StringSpace sp_[2][5];
    mutable DeviceLists dl_;
    mutable HoleLists holes_[2][4];
    mutable bool dl_built_ = false;
    void build_space(int spin, int kind) const;
    std::vector<occ_t> occs_of(int spin, int kind) const;
    void build_device_lists() const;
};

std::vector<occ_t> one_hole_occupations(const std::vector<occ_t>& occ);
std::vector<occ_t> one_hole_one_particle_occupations(const std::vector<occ_t>& occ);

} // namespace f2
