// pgi_nvtx.h — NVTX ranges around the phases of the path (the reference times its phases with
// RunningStatistics, pose_graph_builder.h:398-399 ff.; the counters are exported as pgi_stats / pgb_counters, the ranges
// make the same phases visible on a profiler timeline).  Header-only NVTX v3: without an attached tool a range is a
// load and a branch.
#pragma once
#if defined(__has_include)
#if __has_include(<nvtx3/nvToolsExt.h>)
#include <nvtx3/nvToolsExt.h>
#define PGI_HAVE_NVTX 1
#endif
#endif

struct PgiNvtxRange {
#ifdef PGI_HAVE_NVTX
    explicit PgiNvtxRange(const char *name) { nvtxRangePushA(name); }
    ~PgiNvtxRange() { nvtxRangePop(); }
#else
    explicit PgiNvtxRange(const char *) {}
#endif
    PgiNvtxRange(const PgiNvtxRange &) = delete;
    PgiNvtxRange &operator=(const PgiNvtxRange &) = delete;
};
