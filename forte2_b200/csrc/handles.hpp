// Opaque handle types behind include/forte2_b200.h.
#pragma once
#include <map>
#include <memory>
#include <mutex>
#include <stdexcept>
#include <string>
#include <vector>

#include "../../include/forte2_b200.h"
#include "strings.hpp"

namespace f2 {

struct CudaError : std::runtime_error {
    using std::runtime_error::runtime_error;
};

extern thread_local std::string g_last_error;
int status_of(const std::exception& e);

struct DevStrings; // GPU-resident copy of the lists (sigma.cu)
void destroy_dev_strings(DevStrings* d);

} // namespace f2

struct f2_strings {
    std::unique_ptr<f2::CIStringsHost> host;
    struct {
        bool valid = false;
        int kind = 0, spin = 0;
        std::vector<int64_t> data;
    } mutable export_cache;
    // the replacement lists live on the GPU with the strings object, not with the builder, because
    // forte2 rebuilds the builder on every CI run but keeps CIStrings (forte2/ci/ci.py:171-192)
    // One copy per device, created on first use and kept until the strings object dies: builders on different
    // devices share one CIStrings and hold raw pointers into this map, so an entry is never replaced or freed early.
    std::map<int, f2::DevStrings*> dev;
    std::mutex dev_mutex;
    ~f2_strings() {
        for (auto& kv : dev)
            f2::destroy_dev_strings(kv.second);
    }
};
