// common.cuh — shared device/host definitions for the Vettore B200 scan path.
#pragma once
#include <cuda_runtime.h>

#include <cstdint>
#include <cstring>
#include <string>

namespace vb {

// Metric codes, reference distances.rs:25-38. kCosineTrue is the f64 renormalising
// cosine that search.rs:56-60 / multi_vector.rs:74-75 use instead of the flat dot.
enum Metric : int {
    kL2 = 0, kL2Squared = 1, kCosine = 2, kInnerProduct = 3, kNegativeInnerProduct = 4,
    kManhattan = 5, kChebyshev = 6, kHamming = 7, kJaccard = 8, kCosineTrue = 9
};

typedef unsigned long long u64;

constexpr u64 kKeyMax = ~0ull;
constexpr uint32_t kNoError = 0xFFFFFFFFu;
constexpr int kMaxFusedK = 1024;   // largest k the fused collector handles

// f32::total_cmp as an ascending unsigned key (flat.rs:36-38 orders ranks this way).
__host__ __device__ __forceinline__ uint32_t order_key(float f) {
#ifdef __CUDA_ARCH__
    uint32_t b = __float_as_uint(f);
#else
    uint32_t b;
    std::memcpy(&b, &f, 4);
#endif
    return (b & 0x80000000u) ? ~b : (b ^ 0x80000000u);
}

// distances.rs:113-119. The subtraction must be one IEEE f32 operation.
__device__ __forceinline__ float rank_value(int metric, float raw) {
    if (metric == kCosine || metric == kCosineTrue) return __fsub_rn(1.0f, raw);
    if (metric == kInnerProduct) return -raw;
    return raw;
}

// distances.rs:122-128.
__device__ __forceinline__ float similarity_value(int metric, float raw) {
    if (metric == kCosine || metric == kCosineTrue || metric == kInnerProduct) return raw;
    if (metric == kNegativeInnerProduct) return -raw;
    return __fdiv_rn(1.0f, __fadd_rn(1.0f, raw));
}

// Streaming 128-bit load: read-only path, no L1 allocation (rows are touched once per query).
__device__ __forceinline__ float4 ldg_stream(const float4* p) {
    float4 v;
    asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0, %1, %2, %3}, [%4];"
                 : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p));
    return v;
}
__device__ __forceinline__ uint4 ldg_stream(const uint4* p) {
    uint4 v;
    asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0, %1, %2, %3}, [%4];"
                 : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p));
    return v;
}

__device__ __forceinline__ u64 ld_volatile_u64(const u64* p) {
    u64 v;
    asm volatile("ld.volatile.global.u64 %0, [%1];" : "=l"(v) : "l"(p));
    return v;
}
__device__ __forceinline__ uint32_t ld_volatile_u32(const uint32_t* p) {
    uint32_t v;
    asm volatile("ld.volatile.global.u32 %0, [%1];" : "=r"(v) : "l"(p));
    return v;
}

template <typename T>
__device__ __forceinline__ T warp_sum(T v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}

// ---- host-side error plumbing -----------------------------------------------------
struct Status {
    int code = 0;
    std::string msg;
    bool ok() const { return code == 0; }
    static Status Ok() { return {}; }
    static Status Ref(const char* m) { return Status{1, m}; }          // reference-visible string
    static Status Cuda(const std::string& m) { return Status{2, "cuda: " + m}; }
};

#define VB_CUDA(expr)                                                                   \
    do {                                                                                \
        cudaError_t e__ = (expr);                                                       \
        if (e__ != cudaSuccess)                                                         \
            return ::vb::Status::Cuda(std::string(cudaGetErrorString(e__)) + " (" #expr ")"); \
    } while (0)

#define VB_TRY(expr)                      \
    do {                                  \
        ::vb::Status s__ = (expr);        \
        if (!s__.ok()) return s__;        \
    } while (0)

}  // namespace vb
