// NVTX ranges around the phases of the sigma build, the RDM step, the Davidson passes and the list construction
// (the reference has only its wall-clock timers, ci_sigma_builder.cc:187-218; these are the hooks a
// timeline tool sees).  Header-only NVTX 3: without a tool attached a push / pop is one load and a branch.
#pragma once
#include <nvtx3/nvToolsExt.h>

namespace f2 {

struct NvtxRange {
    explicit NvtxRange(const char* name) { nvtxRangePushA(name); }
    ~NvtxRange() { nvtxRangePop(); }
    NvtxRange(const NvtxRange&) = delete;
    NvtxRange& operator=(const NvtxRange&) = delete;
};

// a range that follows the phases of one function: next("name") closes the previous phase and opens the next one
struct NvtxPhases {
    bool open = false;
    void next(const char* name) {
        if (open)
            nvtxRangePop();
        nvtxRangePushA(name);
        open = true;
    }
    void close() {
        if (open)
            nvtxRangePop();
        open = false;
    }
    ~NvtxPhases() { close(); }
};

} // namespace f2
