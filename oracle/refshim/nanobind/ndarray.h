// Stand-in for <nanobind/ndarray.h> (see nanobind.h in this directory for why it exists).
// A strided, non-owning-or-capsule-owned N-d array handle with element strides.
#pragma once

#include <array>
#include <cstddef>
#include <cstdint>

#include "nanobind.h"

namespace nanobind {

namespace detail {
template <typename T, std::size_t N> struct view_t {
    T* ptr;
    std::array<std::int64_t, N> st;
    template <typename... I> T& operator()(I... idx) const {
        static_assert(sizeof...(I) == N, "wrong number of indices");
        const std::int64_t ii[N] = {static_cast<std::int64_t>(idx)...};
        std::int64_t off = 0;
        for (std::size_t d = 0; d < N; ++d)
            off += ii[d] * st[d];
        return ptr[off];
    }
};
} // namespace detail

template <typename Framework, typename T, typename Dim, typename Order = void> class ndarray {
  public:
    static constexpr std::size_t N = Dim::value;

    ndarray() = default;

    ndarray(T* data, std::size_t nd, const std::size_t* shape, object owner = {},
            const std::int64_t* strides = nullptr)
        : data_(data), owner_(owner) {
        (void)nd;
        for (std::size_t d = 0; d < N; ++d)
            shape_[d] = shape[d];
        if (strides) {
            for (std::size_t d = 0; d < N; ++d)
                stride_[d] = strides[d];
        } else {
            std::int64_t s = 1;
            for (std::size_t d = N; d-- > 0;) {
                stride_[d] = s;
                s *= static_cast<std::int64_t>(shape_[d]);
            }
        }
    }

    T* data() const { return data_; }
    std::size_t ndim() const { return N; }
    std::size_t shape(std::size_t d) const { return shape_[d]; }
    std::int64_t stride(std::size_t d) const { return stride_[d]; }
    std::size_t size() const {
        std::size_t s = 1;
        for (std::size_t d = 0; d < N; ++d)
            s *= shape_[d];
        return s;
    }
    detail::view_t<T, N> view() const { return {data_, stride_}; }
    object cast() const { return owner_; }

  private:
    T* data_ = nullptr;
    std::array<std::size_t, N> shape_{};
    std::array<std::int64_t, N> stride_{};
    object owner_;
};

} // namespace nanobind
