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
    const uint32_t cap = next_pow2(k <= 256 ? 2 * k + slack : k + k / 2 + slack);
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
        if (layout == kScanWholeRows && (size_t)nvec * 4 == row_stride) {
            VB_TRY(plan_stream(metric, nv, row_stride, /*row_floats=*/(uint32_t)row_stride, 4, false, n, k, plan, &taken));
        } else if (layout != kScanRowList && nvec * 4 <= 256 && (row_stride & 3) == 0 && !env_int("VB_SCAN_NO_PREFIX_STREAM", 0)) {
            // every row, only its first columns: a 2D tensor map moves just those columns (box <= 256 floats wide)
            VB_TRY(plan_stream(metric, nv, row_stride, nvec * 4, dims - 4 * (nvec - 1), true, n, k, plan, &taken));
        }
        if (taken) return Status::Ok();
    }
    const uint32_t slack = kSyncEvery * kScanWarps * r;
    const uint32_t kk = dump ? 1 : k;
    const uint32_t cap = next_pow2(kk <= 256 ? 2 * kk + slack : kk + kk / 2 + slack);   // as in plan_stream
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

Status run_flat_scan(const ScanPlan& plan, ScanParams params, uint32_t nq, cudaStream_t stream) {
    params.cap = plan.cap;
    dim3 grid(plan.grid_x, nq);
    if (plan.stream_kernel) {
        StreamGeom geom{plan.stages, plan.tile_bytes, plan.row_floats, plan.tail_rem, plan.use_tmap ? 1u : 0u};
        CUtensorMap tmap;
        std::memset(&tmap, 0, sizeof(tmap));
        if (plan.use_tmap)
            VB_TRY(make_tmap_rows_prefix(params.rows, params.n, params.row_stride, plan.row_floats, plan.tile_rows, &tmap));
        plan.stream_kernel<<<grid, plan.stream_threads, plan.smem, stream>>>(params, geom, tmap);
    } else {
        plan.kernel<<<grid, kScanThreads, plan.smem, stream>>>(params);
    }
    VB_CUDA(cudaGetLastError());
    return Status::Ok();
}

}  // namespace vb
