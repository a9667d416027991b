// tc_probe.cu — one-tile self-test of the tcgen05 building blocks in tc.cuh: TMA 128B-swizzled
// tile load, 3xTF32 split, A operand written to TMEM (tcgen05.st), B operand in the UMMA
// K-major SWIZZLE_128B shared-memory layout, tcgen05.mma kind::tf32 (TS), tcgen05.ld.
// D[128][32] = A[128][K] . B[32][K]^T. Exercised by tests/test_tc_gpu.py.
#include <cuda_runtime.h>

#include <mutex>

#include "tc.cuh"

namespace vb {

typedef CUresult (*TmapEncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                 const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                 CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static TmapEncodeFn tmap_encode_fn() {
    static TmapEncodeFn encode = nullptr;
    static std::once_flag once;
    std::call_once(once, [] {
        void* fn = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres) == cudaSuccess &&
            qres == cudaDriverEntryPointSuccess)
            encode = reinterpret_cast<TmapEncodeFn>(fn);
    });
    return encode;
}

static Status make_tmap_rows(const float* base, uint64_t rows, uint64_t row_stride_floats, uint32_t box_cols,
                             uint32_t box_rows, CUtensorMapSwizzle swizzle, CUtensorMap* out) {
    TmapEncodeFn encode = tmap_encode_fn();
    if (!encode) return Status::Cuda("cuTensorMapEncodeTiled unavailable");
    const cuuint64_t gdim[2] = {row_stride_floats, rows};
    const cuuint64_t gstride[1] = {row_stride_floats * sizeof(float)};
    const cuuint32_t box[2] = {box_cols, box_rows};
    const cuuint32_t estride[2] = {1, 1};
    CUresult r = encode(out, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(base), gdim, gstride, box, estride,
                        CU_TENSOR_MAP_INTERLEAVE_NONE, swizzle, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return Status::Cuda("cuTensorMapEncodeTiled failed (" + std::to_string((int)r) + ")");
    return Status::Ok();
}

Status make_tmap_rows_sw128(const float* base, uint64_t rows, uint64_t row_stride_floats, uint32_t box_rows,
                            CUtensorMap* out) {
    return make_tmap_rows(base, rows, row_stride_floats, 32, box_rows, CU_TENSOR_MAP_SWIZZLE_128B, out);
}

Status make_tmap_rows_sw128_cols(const float* base, uint64_t rows, uint64_t row_stride_floats, uint32_t cols,
                                 uint32_t box_rows, CUtensorMap* out) {
    TmapEncodeFn encode = tmap_encode_fn();
    if (!encode) return Status::Cuda("cuTensorMapEncodeTiled unavailable");
    if (cols == 0 || cols > row_stride_floats || box_rows == 0 || box_rows > 256) return Status::Cuda("tensor map: box out of range");
    const cuuint64_t gdim[2] = {cols, rows};
    const cuuint64_t gstride[1] = {row_stride_floats * sizeof(float)};
    const cuuint32_t box[2] = {32, box_rows};
    const cuuint32_t estride[2] = {1, 1};
    CUresult r = encode(out, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(base), gdim, gstride, box, estride,
                        CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return Status::Cuda("cuTensorMapEncodeTiled failed (" + std::to_string((int)r) + ")");
    return Status::Ok();
}

Status make_tmap_rows_prefix(const float* base, uint64_t rows, uint64_t row_stride_floats, uint32_t box_cols,
                             uint32_t box_rows, CUtensorMap* out) {
    if (box_cols == 0 || box_cols > 256 || (box_cols & 3) || box_rows == 0 || box_rows > 256)
        return Status::Cuda("prefix tensor map: box out of range");
    return make_tmap_rows(base, rows, row_stride_floats, box_cols, box_rows, CU_TENSOR_MAP_SWIZZLE_NONE, out);
}

__global__ void __launch_bounds__(128) tc_probe_kernel(const __grid_constant__ CUtensorMap tmap_a, const float* B,
                                                       int K, int mode, float* D) {
    extern __shared__ __align__(1024) unsigned char smem[];
    __shared__ __align__(8) uint64_t bars[2];
    __shared__ uint32_t tmem_slot;
    const int t = threadIdx.x, warp = t >> 5;
    const int KB = K / 32;
    unsigned char* a_tile = smem;                       // KB blocks x [128 rows x 128 B]
    unsigned char* b_hi = a_tile + (size_t)KB * 16384;  // KB blocks x [32 rows x 128 B]
    unsigned char* b_lo = b_hi + (size_t)KB * 4096;

    if (t == 0) {
        tc::mbar_init(&bars[0], 1);
        tc::mbar_init(&bars[1], 1);
        tc::mbar_fence_init();
    }
    if (warp == 0) tc::tmem_alloc(&tmem_slot, 512);
    for (int idx = t; idx < 32 * K; idx += 128) {
        const int n = idx / K, k = idx % K;
        const float x = B[idx];
        const float hi = mode == 11 ? x : __uint_as_float(__float_as_uint(x) & 0xFFFFE000u);   // 11: raw fp32 words, the hardware cuts them to TF32
        const uint32_t off = (uint32_t)(k / 32) * 4096u + tc::sw128_offset(n, k % 32);
        *reinterpret_cast<float*>(b_hi + off) = hi;
        *reinterpret_cast<float*>(b_lo + off) = x - hi;
    }
    tc::fence_proxy_async_smem();
    tc::fence_before_sync();
    __syncthreads();
    tc::fence_after_sync();
    const uint32_t tbase = tmem_slot;

    if (t == 0) {
        tc::mbar_arrive_expect_tx(&bars[0], (uint32_t)KB * 16384u);
        for (int kb = 0; kb < KB; ++kb) tc::tma_load_2d(a_tile + (size_t)kb * 16384, &tmap_a, kb * 32, 0, &bars[0]);
    }
    tc::mbar_wait(&bars[0], 0);

    const uint32_t lane_addr = tbase + ((uint32_t)(warp * 32) << 16);
    for (int kb = 0; kb < KB; ++kb) {
        uint32_t hi[32], lo[32];
#pragma unroll
        for (int c = 0; c < 8; ++c) {
            const float4 v = *reinterpret_cast<const float4*>(a_tile + (size_t)kb * 16384 + t * 128 + ((c ^ (t & 7)) << 4));
            const float xs[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                const uint32_t h = __float_as_uint(xs[e]) & 0xFFFFE000u;
                hi[c * 4 + e] = h;
                lo[c * 4 + e] = __float_as_uint(xs[e] - __uint_as_float(h));
            }
        }
        tc::tmem_st32(lane_addr + kb * 32, hi);
        tc::tmem_st32(lane_addr + 128 + kb * 32, lo);
    }
    tc::tmem_st_wait();
    tc::fence_before_sync();
    __syncthreads();

    if (t == 0) {
        tc::fence_after_sync();
        const uint32_t idesc = tc::umma_idesc_tf32(128, 32);
        const uint32_t d_tmem = tbase + 256;
        uint32_t acc = 0;
        for (int kb = 0; kb < KB; ++kb) {
            for (int ks = 0; ks < 4; ++ks) {
                const uint64_t bh = tc::umma_smem_desc_sw128(tc::smem_addr(b_hi + kb * 4096 + ks * 32));
                const uint64_t bl = tc::umma_smem_desc_sw128(tc::smem_addr(b_lo + kb * 4096 + ks * 32));
                const uint32_t ah = tbase + kb * 32 + ks * 8, al = tbase + 128 + kb * 32 + ks * 8;
                if (mode == 11) {   // SS: A straight from the TMA-swizzled shared-memory tile, unmasked
                    tc::umma_tf32_ss(d_tmem, tc::umma_smem_desc_sw128(tc::smem_addr(a_tile + (size_t)kb * 16384 + ks * 32)), bh, idesc, acc);
                    acc = 1;
                    continue;
                }
                tc::umma_tf32_ts(d_tmem, ah, bh, idesc, acc);
                acc = 1;
                if (mode == 3) {
                    tc::umma_tf32_ts(d_tmem, ah, bl, idesc, 1);
                    tc::umma_tf32_ts(d_tmem, al, bh, idesc, 1);
                }
            }
        }
        tc::umma_commit(&bars[1]);
    }
    tc::mbar_wait(&bars[1], 0);
    tc::fence_after_sync();
    uint32_t r[32];
    tc::tmem_ld32(lane_addr + 256, r);
    tc::tmem_ld_wait();
#pragma unroll
    for (int j = 0; j < 32; ++j) D[t * 32 + j] = __uint_as_float(r[j]);
    tc::fence_before_sync();
    __syncthreads();
    if (warp == 0) tc::tmem_dealloc(tbase, 512);
}

}  // namespace vb

// Self-test entry (not part of the drop-in ABI): host A[128][K], B[32][K] -> host D[128][32].
// mode 1 = single TF32 pass (A from TMEM), 3 = 3xTF32, 11 = single pass with both operands from shared memory (SS). Returns 0 or a cudaError_t / -1.
extern "C" int vb_debug_tc_probe(const float* hA, const float* hB, int K, int mode, float* hD) {
    if (K <= 0 || K > 128 || K % 32) return -1;
    float *dA = nullptr, *dB = nullptr, *dD = nullptr;
    cudaError_t e;
    if ((e = cudaMalloc(&dA, 128 * K * sizeof(float))) != cudaSuccess) return (int)e;
    cudaMalloc(&dB, 32 * K * sizeof(float));
    cudaMalloc(&dD, 128 * 32 * sizeof(float));
    cudaMemcpy(dA, hA, 128 * K * sizeof(float), cudaMemcpyHostToDevice);
    cudaMemcpy(dB, hB, 32 * K * sizeof(float), cudaMemcpyHostToDevice);
    CUtensorMap tmap;
    vb::Status s = vb::make_tmap_rows_sw128(dA, 128, K, 128, &tmap);
    int rc = 0;
    if (!s.ok()) rc = -2;
    if (rc == 0) {
        const size_t smem = (size_t)(K / 32) * (16384 + 8192) + 1024;
        cudaFuncSetAttribute(vb::tc_probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        vb::tc_probe_kernel<<<1, 128, smem>>>(tmap, dB, K, mode, dD);
        e = cudaDeviceSynchronize();
        if (e != cudaSuccess) rc = (int)e;
        else cudaMemcpy(hD, dD, 128 * 32 * sizeof(float), cudaMemcpyDeviceToHost);
    }
    cudaFree(dA);
    cudaFree(dB);
    cudaFree(dD);
    return rc;
}


// ---------------------------------------------------------------------------------------------------------
// Measured dense TF32 tensor-pipe peak of THIS device ("measure TF32 directly on the box"): one
// persistent CTA per SM issues back-to-back tcgen05.mma kind::tf32 (M = 128, N = 256, K = 8, both operands from
// shared memory, two TMEM accumulators alternating) with no loads, no epilogue and no waits in between. Not part
// of the drop-in ABI; bench.py uses it as the denominator of the tensor-bound rooflines.
namespace vb {
__global__ void __launch_bounds__(128) tf32_peak_kernel(int iters) {
    extern __shared__ __align__(1024) unsigned char smem[];
    __shared__ __align__(8) uint64_t bar;
    __shared__ uint32_t tmem_slot;
    const int t = threadIdx.x, warp = t >> 5;
    for (int i = t; i < (16384 + 32768) / 4; i += 128) reinterpret_cast<float*>(smem)[i] = 1.0f;
    if (t == 0) { tc::mbar_init(&bar, 1); tc::mbar_fence_init(); }
    if (warp == 0) tc::tmem_alloc(&tmem_slot, 512);
    tc::fence_proxy_async_smem();
    tc::fence_before_sync();
    __syncthreads();
    tc::fence_after_sync();
    const uint32_t tbase = tmem_slot;
    if (t == 0) {
        const uint32_t idesc = tc::umma_idesc_tf32(128, 256);
        const uint64_t a0 = tc::umma_smem_desc_sw128(tc::smem_addr(smem));
        const uint64_t b0 = tc::umma_smem_desc_sw128(tc::smem_addr(smem + 16384));
        for (int i = 0; i < iters; ++i)
#pragma unroll
            for (uint32_t ks = 0; ks < 4; ++ks)
                tc::umma_tf32_ss(tbase + (uint32_t)(i & 1) * 256u, a0 + (uint64_t)(ks * 2u), b0 + (uint64_t)(ks * 2u), idesc, 1u);
        tc::umma_commit(&bar);
    }
    tc::mbar_wait(&bar, 0);
    tc::fence_before_sync();
    __syncthreads();
    if (warp == 0) tc::tmem_dealloc(tbase, 512);
}
}  // namespace vb

extern "C" int vb_debug_tf32_peak(int iters, float* tflops) {
    int dev = 0, sms = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return -1;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const size_t smem = 16384 + 32768 + 1024;
    if (cudaFuncSetAttribute(vb::tf32_peak_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) return -2;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    vb::tf32_peak_kernel<<<sms, 128, smem>>>(iters / 8 + 1);            // warm-up
    float best = 0.0f;
    for (int rep = 0; rep < 5; ++rep) {
        cudaEventRecord(e0);
        vb::tf32_peak_kernel<<<sms, 128, smem>>>(iters);
        cudaEventRecord(e1);
        if (cudaEventSynchronize(e1) != cudaSuccess) return -3;
        float ms = 0.0f;
        cudaEventElapsedTime(&ms, e0, e1);
        const double flops = (double)sms * iters * 4.0 * 2.0 * 128.0 * 256.0 * 8.0;
        best = fmaxf(best, (float)(flops / (ms * 1e-3) / 1e12));
    }
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    *tflops = best;
    return cudaGetLastError() == cudaSuccess ? 0 : -4;
}
