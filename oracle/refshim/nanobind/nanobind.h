// Stand-in for <nanobind/nanobind.h>, used ONLY to compile the reference's CI sources
// out-of-tree into oracle/_ref/ (test infrastructure; never linked into the product).
//
// The reference's compute code (forte2/ci/*.cc, forte2/determinant/*.cc) touches nanobind only
// through the ndarray handle type declared in forte2/helpers/ndarray.h:18-35,87-160:
//   nb::ndarray<nb::numpy, T, nb::ndim<N>[, order]>::{data, shape, stride, ndim, view, cast}
//   nb::capsule(ptr, deleter), nb::object
// This header and ndarray.h next to it provide exactly that surface with plain C++ objects.
#pragma once

#include <cstddef>
#include <cstdint>
#include <memory>

namespace nanobind {

struct numpy {};
struct c_contig {};
struct f_contig {};
template <std::size_t N> struct ndim {
    static constexpr std::size_t value = N;
};

/// Owner handle: keeps an allocation alive for as long as any array refers to it.
struct object {
    std::shared_ptr<void> keep;
};

struct capsule : object {
    capsule() = default;
    capsule(void* p, void (*deleter)(void*) noexcept) { keep = std::shared_ptr<void>(p, deleter); }
};

} // namespace nanobind
