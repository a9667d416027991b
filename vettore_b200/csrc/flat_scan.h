// flat_scan.h — host interface of the K1/K4 scan kernels.
#pragma once
#include <cuda.h>

#include "common.cuh"

namespace vb {

struct ScanParams;
struct StreamGeom;
typedef void (*ScanKernel)(const ScanParams);
typedef void (*StreamKernel)(const ScanParams, const StreamGeom, const CUtensorMap);

struct ScanPlan {
    ScanKernel kernel = nullptr;          // kernel A (register-staged loads)
    StreamKernel stream_kernel = nullptr; // kernel B (TMA-staged ring), when eligible
    int nv = 0, r = 0;
    uint32_t grid_x = 0;     // CTAs per query
    uint32_t cap = 0;        // collector capacity (entries)
    size_t smem = 0;         // dynamic shared memory bytes
    uint32_t stages = 0, tile_bytes = 0;  // kernel B ring geometry
    uint32_t stream_threads = 0;          // kernel B block size (consumer warps + producer warp)
    // kernel B geometry in shared memory; use_tmap: prefix scan fed by a 2D tensor map (built at launch)
    uint32_t row_floats = 0, tail_rem = 4, tile_rows = 0;
    bool use_tmap = false;
    // kernel C (prefix_lane.cu): one row per lane, 128-byte swizzled boxes, tensor map `lane_cols` columns wide
    bool lane_rows = false;
    uint32_t lane_cols = 0;
};

// Picks the kernel variant, grid and collector capacity for `n` rows of `dims` scored
// elements (row stride `row_stride` floats) and k results (k <= kMaxFusedK unless dump).
// `contiguous` = no row list and no data beyond `dims` in a row: such scans take the
// TMA-staged kernel B.
// Env knobs for tuning runs: VB_SCAN_R, VB_SCAN_CTAS_PER_SM, VB_SCAN_NO_STREAM,
// VB_STREAM_STAGES, VB_STREAM_RPW, VB_STREAM_WARPS.
// `layout`: kScanRowList (row list or anything else), kScanWholeRows (every row, nothing beyond `dims` in a
// row), kScanPrefixAllRows (every row, only the first `dims` columns are scored).
enum { kScanRowList = 0, kScanWholeRows = 1, kScanPrefixAllRows = 2 };
Status plan_flat_scan(int metric, uint32_t dims, size_t row_stride, int layout, uint32_t n, uint32_t k,
                      bool dump, ScanPlan* plan);

// Launches plan.kernel over grid (plan.grid_x, nq). ScanParams.k/.cap are taken from the plan.
Status run_flat_scan(const ScanPlan& plan, ScanParams params, uint32_t nq, cudaStream_t stream);

int device_sm_count();

}  // namespace vb
