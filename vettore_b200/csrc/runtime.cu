// runtime.cu — buffers, search contexts and the context pool (see runtime.h).
#include "runtime.h"

#include <map>
#include <utility>

namespace vb {

Status DeviceBuf::reserve(size_t bytes) {
    if (bytes <= cap) return Status::Ok();
    if (p) { cudaFree(p); p = nullptr; cap = 0; }
    size_t want = bytes + bytes / 4 + 256;
    VB_CUDA(cudaMalloc(&p, want));
    cap = want;
    return Status::Ok();
}
void DeviceBuf::release() {
    if (p) cudaFree(p);
    p = nullptr;
    cap = 0;
}

Status PinnedBuf::reserve(size_t bytes) {
    if (bytes <= cap) return Status::Ok();
    if (p) { cudaFreeHost(p); p = nullptr; cap = 0; }
    size_t want = bytes + bytes / 4 + 256;
    VB_CUDA(cudaMallocHost(&p, want));
    cap = want;
    return Status::Ok();
}
void PinnedBuf::release() {
    if (p) cudaFreeHost(p);
    p = nullptr;
    cap = 0;
}

__global__ void arm_ctrl_kernel(u64* g_thresh, uint32_t* done, uint32_t* err_row, uint32_t nq) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < nq) { g_thresh[i] = kKeyMax; done[i] = 0u; err_row[i] = kNoError; }
}

Status SearchCtx::arm_ctrl(uint32_t nq) {
    if (nq <= ctrl_queries) return Status::Ok();
    uint32_t slots = nq < 16 ? 16 : nq;
    VB_TRY(ctrl.reserve((size_t)slots * 16));
    ctrl_queries = slots;
    arm_ctrl_kernel<<<(slots + 255) / 256, 256, 0, stream>>>(g_thresh(), done(), err_row(), slots);
    VB_CUDA(cudaGetLastError());
    return Status::Ok();
}

void SearchCtx::destroy() {
    cudaSetDevice(device);
    DeviceBuf* dbufs[] = {&queries, &q_norms, &cand_keys, &cand_pays, &cand_counts, &ctrl, &out_keys, &result,
                          &row_sel, &row_sel2, &staging, &staging_rank, &dump_keys, &dump_pays, &dump_keys2, &dump_pays2,
                          &sort_tmp, &hist, &misc};
    for (DeviceBuf* b : dbufs) b->release();
    h_queries.release();
    h_result.release();
    h_misc.release();
    if (stream) cudaStreamDestroy(stream);
    stream = nullptr;
}

CtxPool::~CtxPool() {
    // Process teardown: the CUDA context may already be gone; leak rather than crash.
}

Status CtxPool::acquire(SearchCtx** out) {
    int dev = 0;
    VB_CUDA(cudaGetDevice(&dev));
    {
        std::lock_guard<std::mutex> g(mu_);
        for (size_t i = 0; i < free_.size(); ++i) {
            if (free_[i]->device == dev) {
                *out = free_[i];
                free_.erase(free_.begin() + i);
                return Status::Ok();
            }
        }
    }
    SearchCtx* ctx = new SearchCtx();
    ctx->device = dev;
    cudaError_t e = cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking);
    if (e != cudaSuccess) {
        delete ctx;
        return Status::Cuda(cudaGetErrorString(e));
    }
    *out = ctx;
    return Status::Ok();
}

void CtxPool::release(SearchCtx* ctx) {
    std::lock_guard<std::mutex> g(mu_);
    free_.push_back(ctx);
}

CtxPool& ctx_pool() {
    static CtxPool* pool = new CtxPool();  // intentionally leaked (see ~CtxPool)
    return *pool;
}

Status ensure_dynamic_smem(const void* kernel, size_t bytes) {
    static std::mutex mu;
    static std::map<std::pair<int, const void*>, size_t> granted;
    // The 48 KB a kernel gets without opting in covers its STATIC shared memory too (the general MaxSim kernel holds
    // 26 KB of tiles statically: 32 KB of collector on top of that already needs the attribute), so only requests
    // that fit next to any kernel's static part are waved through.
    if (bytes <= 8 * 1024) return Status::Ok();
    int dev = 0;
    VB_CUDA(cudaGetDevice(&dev));
    std::lock_guard<std::mutex> g(mu);
    size_t& have = granted[{dev, kernel}];
    if (have >= bytes) return Status::Ok();
    VB_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
    have = bytes;
    return Status::Ok();
}

}  // namespace vb
