// flat_scan.cu — variant selection and launch of the K1/K4 scan kernels.
#include "flat_scan.cuh"
#include "flat_scan.h"
#include "runtime.h"

#include <cstdlib>
#include <cstring>
#include <map>
#include <mutex>

namespace vb {

#define VB_DECL(M)                                         \
    ScanKernel flat_scan_kernel_metric_##M(int nv, int r); \
    StreamKernel flat_stream_kernel_metric_##M(int nv, int rpw, int warps);
VB_DECL(0) VB_DECL(1) VB_DECL(2) VB_DECL(3) VB_DECL(4) VB_DECL(5) VB_DECL(6) VB_DECL(7) VB_DECL(8) VB_DECL(9)
#undef VB_DECL

static ScanKernel lookup(int metric, int nv, int r) {
    switch (metric) {
        case 0: return flat_scan_kernel_metric_0(nv, r);
        case 1: return flat_scan_kernel_metric_1(nv, r);
        case 2: return flat_scan_kernel_metric_2(nv, r);
        case 3: return flat_scan_kernel_metric_3(nv, r);
        case 4: return flat_scan_kernel_metric_4(nv, r);
        case 5: return flat_scan_kernel_metric_5(nv, r);
        case 6: return flat_scan_kernel_metric_6(nv, r);
        case 7: return flat_scan_kernel_metric_7(nv, r);
        case 8: return flat_scan_kernel_metric_8(nv, r);
        case 9: return flat_scan_kernel_metric_9(nv, r);
    }
    return nullptr;
}

static StreamKernel lookup_stream(int metric, int nv, int rpw, int warps) {
    switch (metric) {
        case 0: return flat_stream_kernel_metric_0(nv, rpw, warps);
        case 1: return flat_stream_kernel_metric_1(nv, rpw, warps);
        case 2: return flat_stream_kernel_metric_2(nv, rpw, warps);
        case 3: return flat_stream_kernel_metric_3(nv, rpw, warps);
        case 4: return flat_stream_kernel_metric_4(nv, rpw, warps);
        case 5: return flat_stream_kernel_metric_5(nv, rpw, warps);
        case 6: return flat_stream_kernel_metric_6(nv, rpw, warps);
        case 7: return flat_stream_kernel_metric_7(nv, rpw, warps);
        case 8: return flat_stream_kernel_metric_8(nv, rpw, warps);
        case 9: return flat_stream_kernel_metric_9(nv, rpw, warps);
    }
    return nullptr;
}

StreamKernel prefix_lane_kernel_entry();
int prefix_lane_max_warps();
int prefix_lane_max_stages();
int prefix_lane_sync_rounds();

int device_sm_count() {
    static thread_local int cached_dev = -1, cached = 0;
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return 0;
    if (dev != cached_dev) {
        cudaDeviceProp prop;
        if (cudaGetDeviceProperties(&prop, dev) != cudaSuccess) return 0;
        cached = prop.multiProcessorCount;
        cached_dev = dev;
    }
    return cached;
}

static int env_int(const char* name, int dflt) {
    const char* v = std::getenv(name);
    return v && *v ? std::atoi(v) : dflt;
}

static uint32_t next_pow2(uint32_t v) {
    uint32_t p = 32;
    while (p < v) p <<= 1;
    return p;
}

// Collector capacity for k results with up to `slack` pushes between two checkpoints: room for k kept entries plus
// the pushes between two compactions; beyond k = 256 the launch-wide bound (pivot ladder) keeps pushes rare, so half
// of k again is plenty and the ring keeps its stages. (More head-room — 4k, 8k — measured no faster at any k.)
static uint32_t collector_cap(uint32_t k, uint32_t slack) {
    return next_pow2(k <= 256 ? 2 * k + slack : k + k / 2 + slack);
}

// Kernel B: whole contiguous rows streamed through a shared-memory ring by TMA bulk copies.
static Status plan_stream(int metric, int nv, size_t row_stride, uint32_t row_floats, uint32_t tail_rem, bool use_tmap,
                          uint32_t n, uint32_t k, ScanPlan* plan, bool* taken) {
    (void)row_stride;
    *taken = false;
    int rpw = nv <= 2 ? 4 : (nv <= 4 ? 2 : 1);
    int warps = nv >= 12 ? 8 : 16;
    const int rpw_env = env_int("VB_STREAM_RPW", 0), warps_env = env_int("VB_STREAM_WARPS", 0);
    if ((rpw_env > 0 || warps_env > 0) &&
        lookup_stream(metric, nv, rpw_env > 0 ? rpw_env : rpw, warps_env > 0 ? warps_env : warps)) {
        if (rpw_env > 0) rpw = rpw_env;
        if (warps_env > 0) warps = warps_env;
    }
    StreamKernel kernel = lookup_stream(metric, nv, rpw, warps);
    if (!kernel) return Status::Ok();
    const uint32_t tile_rows = warps * rpw;
    const uint32_t tile_bytes = (uint32_t)(((size_t)tile_rows * row_floats * 4 + 127) & ~(size_t)127);
    const uint32_t groups_per_sync = std::max(1, kStreamSyncEvery / (kGroupRows / rpw));
    const uint32_t slack = groups_per_sync * kGroupRows * warps;
    // room for k kept entries plus the pushes between two compactions; beyond k = 256 the launch-wide bound
    // (pivot ladder) keeps pushes rare, so half of k again is plenty and the ring keeps its stages
    const uint32_t cap = collector_cap(k, slack);
    const size_t budget = 200 * 1024;
    if ((size_t)cap * 16 + 2 * (size_t)tile_bytes > budget) return Status::Ok();
    uint32_t stages = (uint32_t)std::min<size_t>(kStreamMaxStages, (budget - (size_t)cap * 16) / tile_bytes);
    const int stages_env = env_int("VB_STREAM_STAGES", 0);
    if (stages_env >= 2 && (uint32_t)stages_env < stages) stages = stages_env;
    const size_t smem = (size_t)stages * tile_bytes + (size_t)cap * 16;
    VB_TRY(ensure_dynamic_smem_for(kernel, budget));   // cached per (device, kernel)
    const int sms = device_sm_count();
    if (sms <= 0) return Status::Cuda("no CUDA device");
    const uint32_t tiles = (n + tile_rows - 1) / tile_rows;
    plan->stream_kernel = kernel;
    plan->nv = nv;
    plan->r = rpw;
    plan->stream_threads = warps * 32 + 32;
    plan->grid_x = std::min<uint32_t>(tiles, (uint32_t)sms);
    plan->cap = cap;
    plan->smem = smem;
    plan->stages = stages;
    plan->tile_bytes = tile_bytes;
    plan->row_floats = row_floats;
    plan->tail_rem = tail_rem;
    plan->tile_rows = tile_rows;
    plan->use_tmap = use_tmap;
    *taken = true;
    return Status::Ok();
}

// Kernel C (prefix_lane.cu): true-cosine prefix scoring of every row, one row per lane, tiles of 32 rows as
// ceil(dims / 32) swizzled TMA boxes. Ring slots are a multiple of the consumer warps (a slot belongs to one warp).
static Status plan_lane(uint32_t dims, uint32_t n, uint32_t k, ScanPlan* plan, bool* taken) {
    *taken = false;
    const uint32_t nb = (dims + 31) / 32;
    const uint32_t tile_bytes = nb * 32 * 128;
    const size_t budget = 214 * 1024;
    // consumer warps: as many as the ring can give one tile each (the per-row dependency chain is long — conversions,
    // f64 FMAs — and only other warps hide it), at most kLaneMaxWarps; what is left deepens the ring
    uint32_t warps = (uint32_t)prefix_lane_max_warps();
    const int warps_env = env_int("VB_LANE_WARPS", 0);
    if (warps_env > 0 && (uint32_t)warps_env < warps) warps = warps_env;
    const uint32_t sync_rounds = (uint32_t)std::min(std::max(env_int("VB_LANE_SYNC_ROUNDS", prefix_lane_sync_rounds()), 1), 8);
    uint32_t cap = 0;
    for (;; --warps) {
        if (warps < 2) return Status::Ok();
        const uint32_t slack = sync_rounds * warps * 32;
        cap = collector_cap(k, slack);
        if ((size_t)cap * 16 + (size_t)warps * tile_bytes <= budget) break;
    }
    uint32_t depth = (uint32_t)std::min<size_t>((budget - (size_t)cap * 16) / ((size_t)warps * tile_bytes),
                                                (size_t)prefix_lane_max_stages() / (warps * nb));
    const int depth_env = env_int("VB_LANE_DEPTH", 0);
    if (depth_env >= 1 && (uint32_t)depth_env < depth) depth = depth_env;
    if (depth < 1) return Status::Ok();
    const uint32_t stages = depth;   // StreamGeom.stages carries the ring depth in tiles per consumer warp
    StreamKernel kernel = prefix_lane_kernel_entry();
    VB_TRY(ensure_dynamic_smem_for(kernel, budget));
    const int sms = device_sm_count();
    if (sms <= 0) return Status::Cuda("no CUDA device");
    plan->stream_kernel = kernel;
    plan->nv = 1;
    plan->r = 32;
    plan->stream_threads = warps * 32 + 32;
    plan->grid_x = std::min<uint32_t>((n + 31) / 32, (uint32_t)sms);
    plan->cap = cap;
    plan->smem = (size_t)warps * stages * tile_bytes + (size_t)cap * 16;
    plan->stages = stages;
    plan->tile_bytes = tile_bytes;
    plan->row_floats = nb * 32;
    plan->tail_rem = sync_rounds;   // kernel C reads StreamGeom.tail_rem as its checkpoint cadence (it masks no tail)
    plan->tile_rows = 32;
    plan->use_tmap = true;
    plan->lane_rows = true;
    plan->lane_cols = dims;
    *taken = true;
    return Status::Ok();
}

Status plan_flat_scan(int metric, uint32_t dims, size_t row_stride, int layout, uint32_t n, uint32_t k,
                      bool dump, ScanPlan* plan) {
    if (metric < 0 || metric > 9) return Status::Ref("unknown metric");
    if (dims == 0 || n == 0) return Status::Cuda("empty scan");
    const uint32_t nvec = (dims + 3) / 4;
    const uint32_t need = (nvec + 31) / 32;
    int nv, r;
    if (need <= 1) { nv = 1; r = 8; }
    else if (need <= 2) { nv = 2; r = 4; }
    else if (need <= 3) { nv = 3; r = 4; }
    else if (need <= 4) { nv = 4; r = 4; }
    else if (need <= 6) { nv = 6; r = 2; }
    else if (need <= 8) { nv = 8; r = 2; }
    else if (need <= 12) { nv = 12; r = 1; }
    else { nv = 0; r = 2; }
    const int r_env = env_int("VB_SCAN_R", 0);
    if (r_env > 0 && lookup(metric, nv, r_env)) r = r_env;
    ScanKernel kernel = lookup(metric, nv, r);
    if (!kernel) return Status::Cuda("no scan kernel variant");

    if (!dump && k > (uint32_t)kMaxFusedK) return Status::Cuda("k beyond fused collector");
    *plan = ScanPlan{};
    if (!dump && nv > 0 && !env_int("VB_SCAN_NO_STREAM", 0)) {
        bool taken = false;
        const bool whole_rows = layout == kScanWholeRows && (size_t)nvec * 4 == row_stride;
        if (layout != kScanRowList && metric == kCosineTrue && dims <= 128 && (row_stride & 3) == 0 && n >= 4096 &&
            !env_int("VB_SCAN_NO_LANE", 0)) {
            // true-cosine prefix (or a narrow whole row, e.g. the dense prefix mirror) of every row: one row per lane
            VB_TRY(plan_lane(dims, n, k, plan, &taken));
        }
        if (taken) return Status::Ok();
        if (whole_rows) {
            VB_TRY(plan_stream(metric, nv, row_stride, /*row_floats=*/(uint32_t)row_stride, 4, false, n, k, plan, &taken));
        } else if (layout != kScanRowList && nvec * 4 <= 256 && (row_stride & 3) == 0 && !env_int("VB_SCAN_NO_PREFIX_STREAM", 0)) {
            // every row, only its first columns: a 2D tensor map moves just those columns (box <= 256 floats wide)
            VB_TRY(plan_stream(metric, nv, row_stride, nvec * 4, dims - 4 * (nvec - 1), true, n, k, plan, &taken));
        }
        if (taken) return Status::Ok();
    }
    const uint32_t slack = kSyncEvery * kScanWarps * r;
    const uint32_t kk = dump ? 1 : k;
    const uint32_t cap = collector_cap(kk, slack);
    const size_t smem = (size_t)cap * 16;

    VB_TRY(ensure_dynamic_smem_for(kernel, smem));      // cached per (device, kernel); occupancy is device independent here
    static std::mutex mu;
    static std::map<std::pair<const void*, size_t>, int> occ_cache;
    int per_sm = 0;
    {
        std::lock_guard<std::mutex> g(mu);
        auto key = std::make_pair((const void*)kernel, smem);
        auto it = occ_cache.find(key);
        if (it == occ_cache.end()) {
            VB_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, kScanThreads, smem));
            if (per_sm < 1) return Status::Cuda("scan kernel does not fit on an SM");
            occ_cache[key] = per_sm;
        } else {
            per_sm = it->second;
        }
    }
    const int per_sm_env = env_int("VB_SCAN_CTAS_PER_SM", 0);
    if (per_sm_env > 0 && per_sm_env < per_sm) per_sm = per_sm_env;
    const int sms = device_sm_count();
    if (sms <= 0) return Status::Cuda("no CUDA device");
    const uint32_t tile_rows = kScanWarps * r;
    const uint32_t tiles = (n + tile_rows - 1) / tile_rows;
    plan->kernel = kernel;
    plan->nv = nv;
    plan->r = r;
    plan->grid_x = std::min<uint32_t>(tiles, (uint32_t)(sms * per_sm));
    plan->cap = cap;
    plan->smem = smem;
    return Status::Ok();
}

// Which kernel the calling thread's last scan launched (tests, tuning): 0 kernel A, 1 kernel B over whole rows,
// 2 kernel B over a prefix box, 3 kernel C (one row per lane).
static thread_local int t_last_scan = 0;
extern "C" int vb_debug_scan_path() { return t_last_scan; }

Status run_flat_scan(const ScanPlan& plan, ScanParams params, uint32_t nq, cudaStream_t stream) {
    params.cap = plan.cap;
    t_last_scan = plan.lane_rows ? 3 : plan.stream_kernel ? (plan.use_tmap ? 2 : 1) : 0;
    dim3 grid(plan.grid_x, nq);
    if (plan.stream_kernel) {
        StreamGeom geom{plan.stages, plan.tile_bytes, plan.row_floats, plan.tail_rem, plan.use_tmap ? 1u : 0u};
        CUtensorMap tmap;
        std::memset(&tmap, 0, sizeof(tmap));
        if (plan.lane_rows)
            VB_TRY(make_tmap_rows_sw128_cols(params.rows, params.n, params.row_stride, plan.lane_cols, plan.tile_rows, &tmap));
        else if (plan.use_tmap)
            VB_TRY(make_tmap_rows_prefix(params.rows, params.n, params.row_stride, plan.row_floats, plan.tile_rows, &tmap));
        plan.stream_kernel<<<grid, plan.stream_threads, plan.smem, stream>>>(params, geom, tmap);
    } else {
        plan.kernel<<<grid, kScanThreads, plan.smem, stream>>>(params);
    }
    VB_CUDA(cudaGetLastError());
    return Status::Ok();
}

}  // namespace vb
