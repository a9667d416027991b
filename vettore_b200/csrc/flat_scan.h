// flat_scan.h — host interface of the K1/K4 scan kernels.
#pragma once
#include "common.cuh"

namespace vb {

struct ScanParams;
typedef void (*ScanKernel)(const ScanParams);

struct ScanPlan {
    ScanKernel kernel = nullptr;
    int nv = 0, r = 0;
    uint32_t grid_x = 0;     // CTAs per query
    uint32_t cap = 0;        // collector capacity (entries)
    size_t smem = 0;         // dynamic shared memory bytes
};

// Picks the kernel variant, grid and collector capacity for `n` rows of `dims` scored
// elements and k results (k <= kMaxFusedK unless dump). Env knobs for tuning runs:
// VB_SCAN_R (rows per warp step), VB_SCAN_CTAS_PER_SM.
Status plan_flat_scan(int metric, uint32_t dims, uint32_t n, uint32_t k, bool dump, ScanPlan* plan);

// Launches plan.kernel over grid (plan.grid_x, nq). ScanParams.k/.cap are taken from the plan.
Status run_flat_scan(const ScanPlan& plan, ScanParams params, uint32_t nq, cudaStream_t stream);

int device_sm_count();

}  // namespace vb
