// Replacement lists built on the GPU (csrc/lists_device.cu).
#pragma once
#include "strings.hpp"

namespace f2 {

/// Fills the list cache of `h` (what CIStringsHost::device_lists() returns) from kernels on the current CUDA device;
/// the result is identical, array by array, to the host builder's.
void build_lists_on_device(const CIStringsHost& h);

} // namespace f2
