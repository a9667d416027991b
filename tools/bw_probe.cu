// bw_probe.cu — read-bandwidth ceilings on this GPU for the access patterns the scan kernels use.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o build/bw_probe tools/bw_probe.cu
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstdlib>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); exit(1);} } while (0)

__device__ __forceinline__ float4 ldg_stream(const float4* p) {
    float4 v;
    asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p));
    return v;
}

template <int U>
__global__ void __launch_bounds__(256) read_ldg(const float4* __restrict__ src, size_t n4, float* out) {
    float acc = 0.f;
    size_t stride = (size_t)gridDim.x * blockDim.x;
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    for (; i + (U - 1) * stride < n4; i += U * stride) {
        float4 v[U];
#pragma unroll
        for (int u = 0; u < U; ++u) v[u] = ldg_stream(src + i + u * stride);
#pragma unroll
        for (int u = 0; u < U; ++u) acc += v[u].x + v[u].y + v[u].z + v[u].w;
    }
    for (; i < n4; i += stride) { float4 v = ldg_stream(src + i); acc += v.x + v.y + v.z + v.w; }
    if (acc == 123.456f) out[0] = acc;
}

// contiguous chunk per CTA per iteration (like the scan kernels' tiles)
template <int U>
__global__ void __launch_bounds__(256) read_ldg_tiled(const float4* __restrict__ src, size_t n4, float* out) {
    float acc = 0.f;
    const size_t tile4 = (size_t)256 * U;  // float4 per tile
    const size_t tiles = n4 / tile4;
    for (size_t t = blockIdx.x; t < tiles; t += gridDim.x) {
        const float4* p = src + t * tile4 + threadIdx.x;
        float4 v[U];
#pragma unroll
        for (int u = 0; u < U; ++u) v[u] = ldg_stream(p + u * 256);
#pragma unroll
        for (int u = 0; u < U; ++u) acc += v[u].x + v[u].y + v[u].z + v[u].w;
    }
    if (acc == 123.456f) out[0] = acc;
}

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    return ok != 0;
}

// TMA bulk ring, consumers only touch one word per 128 B then release (pure streaming ceiling)
__global__ void __launch_bounds__(288, 1) read_tma(const unsigned char* __restrict__ src, size_t bytes, uint32_t tile_bytes,
                                                   uint32_t stages, int touch, float* out) {
    extern __shared__ __align__(128) unsigned char smem[];
    __shared__ __align__(8) uint64_t full_bar[16], empty_bar[16];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) {
        for (uint32_t s = 0; s < stages; ++s) {
            asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(&full_bar[s])), "r"(1));
            asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(&empty_bar[s])), "r"(8));
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    const size_t tiles = bytes / tile_bytes;
    if (warp == 8) {
        if (lane == 0) {
            uint32_t it = 0;
            for (size_t t = blockIdx.x; t < tiles; t += gridDim.x, ++it) {
                uint32_t s = it % stages, ph = (it / stages) & 1;
                while (!mbar_try_wait(&empty_bar[s], ph ^ 1)) {}
                asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&full_bar[s])), "r"(tile_bytes) : "memory");
                asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                             ::"r"(smem_u32(smem + (size_t)s * tile_bytes)), "l"(src + t * tile_bytes), "r"(tile_bytes), "r"(smem_u32(&full_bar[s])) : "memory");
            }
        }
        return;
    }
    float acc = 0.f;
    uint32_t it = 0;
    for (size_t t = blockIdx.x; t < tiles; t += gridDim.x, ++it) {
        uint32_t s = it % stages, ph = (it / stages) & 1;
        while (!mbar_try_wait(&full_bar[s], ph)) {}
        if (touch) {
            const float4* p = reinterpret_cast<const float4*>(smem + (size_t)s * tile_bytes);
            for (uint32_t i = threadIdx.x; i < tile_bytes / 16; i += 256) { float4 v = p[i]; acc += v.x + v.y + v.z + v.w; }
        }
        __syncwarp();
        if (lane == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(&empty_bar[s])) : "memory");
    }
    if (acc == 123.456f) out[0] = acc;
}

template <typename F>
float time_it(F f, int iters = 20) {
    cudaEvent_t a, b;
    cudaEventCreate(&a); cudaEventCreate(&b);
    for (int i = 0; i < 3; ++i) f();
    cudaEventRecord(a);
    for (int i = 0; i < iters; ++i) f();
    cudaEventRecord(b);
    CK(cudaEventSynchronize(b));
    float ms; cudaEventElapsedTime(&ms, a, b);
    return ms / iters;
}

int main(int argc, char** argv) {
    // default: the 1M x 768 fp32 corpus; argv[1] = size in MB (e.g. 12800 for the 100M x 1024-bit codes)
    const size_t bytes = (argc > 1 ? (size_t)atoll(argv[1]) : (size_t)3072) * 1000 * 1000;
    unsigned char* d; float* out;
    CK(cudaMalloc(&d, bytes)); CK(cudaMalloc(&out, 4));
    CK(cudaMemset(d, 1, bytes));
    const size_t n4 = bytes / 16;
    auto report = [&](const char* name, float ms) { printf("%-40s %8.4f ms  %8.1f GB/s\n", name, ms, bytes / ms / 1e6); };
    for (int mult : {2, 4, 8}) {
        char nm[64];
        snprintf(nm, 64, "ldg strided U=8 grid=148x%d", mult);
        report(nm, time_it([&] { read_ldg<8><<<148 * mult, 256>>>((const float4*)d, n4, out); }));
        snprintf(nm, 64, "ldg tiled   U=8 grid=148x%d", mult);
        report(nm, time_it([&] { read_ldg_tiled<8><<<148 * mult, 256>>>((const float4*)d, n4, out); }));
        snprintf(nm, 64, "ldg tiled   U=16 grid=148x%d", mult);
        report(nm, time_it([&] { read_ldg_tiled<16><<<148 * mult, 256>>>((const float4*)d, n4, out); }));
    }
    CK(cudaFuncSetAttribute(read_tma, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    for (uint32_t tile : {16384u, 32768u, 49152u}) {
        for (uint32_t stages : {2u, 4u}) {
            for (int touch : {0, 1}) {
                if ((size_t)tile * stages > 200 * 1024) continue;
                char nm[64];
                snprintf(nm, 64, "tma ring tile=%uK stages=%u touch=%d", tile / 1024, stages, touch);
                report(nm, time_it([&] { read_tma<<<148, 288, tile * stages>>>(d, bytes, tile, stages, touch, out); }));
            }
        }
    }
    // copy for reference (read+write, counts both directions like MEASURED_PEAKS)
    unsigned char* d2; CK(cudaMalloc(&d2, bytes));
    float ms = time_it([&] { cudaMemcpyAsync(d2, d, bytes, cudaMemcpyDeviceToDevice); });
    printf("%-40s %8.4f ms  %8.1f GB/s (read+write)\n", "cudaMemcpy D2D", ms, 2.0 * bytes / ms / 1e6);
    CK(cudaDeviceSynchronize());
    return 0;
}
