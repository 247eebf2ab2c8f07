// nvtx_ranges.h — NVTX ranges around the stages of the search path (the reference has Go's pprof/trace hooks around the same
// stages).  Header-only NVTX v3: without a profiler attached a range is a couple of predictable branches; with ncu / nsys the timeline
// shows reader → H2D → query preparation → probe → hit sort → D2H → post-filter → writer per batch and per part.
#pragma once
#include <nvtx3/nvToolsExt.h>

namespace kmcpg {

struct NvtxRange {
    explicit NvtxRange(const char *name) { nvtxRangePushA(name); }
    ~NvtxRange() { nvtxRangePop(); }
    NvtxRange(const NvtxRange &) = delete;
    NvtxRange &operator=(const NvtxRange &) = delete;
};

}  // namespace kmcpg
