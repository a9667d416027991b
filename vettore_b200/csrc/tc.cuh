// tc.cuh — tcgen05 / TMEM / TMA building blocks for the tensor-core kernels (sm_100a).
//
// The dense contractions of this path (MaxSim, batched flat scan) must keep fp32 accuracy
// (1e-5), so they run as 3xTF32: x = hi + lo with hi = the 19 bits the TF32 datapath reads
// and lo = x - hi (exact in fp32); x.y ~= hi.hi + hi.lo + lo.hi accumulated in fp32 in TMEM.
#pragma once
#include <cuda.h>

#include "common.cuh"

namespace vb {
namespace tc {

__device__ __forceinline__ uint32_t smem_addr(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// One lane of a fully converged warp (elect.sync). Keeping the surrounding code warp-uniform lets
// the compiler hold descriptors in uniform registers instead of broadcasting them per issue.
__device__ __forceinline__ bool elect_one() {
    uint32_t pred;
    asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred));
    return pred != 0;
}

// ---- mbarrier -------------------------------------------------------------------------
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_addr(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_fence_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok) : "r"(smem_addr(bar)), "r"(parity) : "memory");
    return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    while (!mbar_try_wait(bar, parity)) {}
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_addr(bar)) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_addr(bar)), "r"(bytes) : "memory");
}

// ---- TMA: 2D tiled tensor load, 128-byte swizzle -----------------------------------------
__device__ __forceinline__ void tma_load_2d(void* dst_smem, const CUtensorMap* tmap, uint32_t x, uint32_t y,
                                            uint64_t* bar) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
        ::"r"(smem_addr(dst_smem)), "l"(tmap), "r"(x), "r"(y), "r"(smem_addr(bar)) : "memory");
}
__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap* tmap) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(tmap) : "memory");
}

// Byte offset of element (row, col) of a [rows x 32 floats] block stored with the 128-byte
// swizzle (TMA CU_TENSOR_MAP_SWIZZLE_128B / UMMA SWIZZLE_128B, K-major): 16-byte chunk c of
// row r lives at chunk position c ^ (r & 7) of that row's 128 bytes. Block base 1024-aligned.
__device__ __forceinline__ uint32_t sw128_offset(uint32_t row, uint32_t col) {
    const uint32_t chunk = col >> 2;
    return row * 128u + (((chunk ^ (row & 7u)) << 4) | ((col & 3u) << 2));
}

// ---- TMEM allocation ------------------------------------------------------------------
// One full warp. Writes the TMEM base address (lane 0, column c) to *slot_smem.
__device__ __forceinline__ void tmem_alloc(uint32_t* slot_smem, uint32_t ncols) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_addr(slot_smem)), "r"(ncols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void fence_before_sync() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void fence_after_sync() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// ---- TMEM <-> registers: 32 lanes x 32-bit, 32 consecutive columns -----------------------
// The calling warp reaches lanes [32 * (warp % 4), +32); taddr = base | lane << 16 | column.
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&r)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
          "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
          "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
          "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr) : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&r)[16]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
          "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr) : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tmem_st32(uint32_t taddr, const uint32_t (&r)[32]) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
        "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
        "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};"
        ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]),
          "r"(r[8]), "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]),
          "r"(r[16]), "r"(r[17]), "r"(r[18]), "r"(r[19]), "r"(r[20]), "r"(r[21]), "r"(r[22]), "r"(r[23]),
          "r"(r[24]), "r"(r[25]), "r"(r[26]), "r"(r[27]), "r"(r[28]), "r"(r[29]), "r"(r[30]), "r"(r[31])
        : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// ---- UMMA descriptors -------------------------------------------------------------------
// Shared-memory operand, K-major, 128-byte swizzle: rows of 128 bytes (32 tf32), 8-row
// groups 1024 bytes apart (SBO); version 1 (Blackwell); layout type 2 = SWIZZLE_128B.
__device__ __forceinline__ uint64_t umma_smem_desc_sw128(uint32_t smem_byte_addr) {
    uint64_t d = 0;
    d |= (uint64_t)((smem_byte_addr >> 4) & 0x3FFFu);        // start address, bits [0,14)
    d |= (uint64_t)0 << 16;                                    // leading byte offset (unused: one atom along K)
    d |= (uint64_t)((1024u >> 4) & 0x3FFFu) << 32;             // stride byte offset, bits [32,46)
    d |= (uint64_t)1 << 46;                                    // descriptor version
    d |= (uint64_t)2 << 61;                                    // SWIZZLE_128B
    return d;
}
// Instruction descriptor, kind::tf32, fp32 accumulate, A and B K-major: M x N x 8 per issue.
__host__ __device__ constexpr uint32_t umma_idesc_tf32(uint32_t M, uint32_t N) {
    return (1u << 4)            // D format: F32
           | (2u << 7)          // A format: TF32
           | (2u << 10)         // B format: TF32
           | ((N >> 3) << 17)   // N / 8
           | ((M >> 4) << 24);  // M / 16
}
// D[tmem] (+)= A[tmem] * B[smem]^T, issued by ONE thread.
__device__ __forceinline__ void umma_tf32_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc,
                                             uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}"
        ::"r"(d_tmem), "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(accumulate) : "memory");
}
// D[tmem] (+)= A[smem] * B[smem]^T
__device__ __forceinline__ void umma_tf32_ss(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc,
                                             uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate) : "memory");
}
// Arrives on the mbarrier once every previously issued MMA of this thread has completed.
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_addr(bar))
                 : "memory");
}

// ---- CTA pair (cta_group::2): two CTAs of one cluster share an MMA of M = 256 -----------------------
// Both CTAs of the pair execute alloc / dealloc (one warp each); only the leader (cluster rank 0) issues MMAs and
// commits; the commit arrives on the barrier at the same shared-memory offset in every CTA of `cta_mask`.
__device__ __forceinline__ uint32_t cluster_ctarank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// Shared-memory address of `p` (own CTA) as seen in CTA `rank` of the cluster.
__device__ __forceinline__ uint32_t mapa_shared(const void* p, uint32_t rank) {
    uint32_t r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(smem_addr(p)), "r"(rank));
    return r;
}
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t cluster_addr) {
    asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
// (Waits on barriers the other CTA arrives on use the plain mbar_wait: an acquire at CLUSTER scope makes ptxas put a
// CCTL.IVALL — a full L1 invalidate — behind every successful wait, which made the paired K2 kernel 2.3x slower on
// its load path alone. No generic-proxy data crosses the CTAs here: the barriers order async-proxy TMA writes, tensor
// core reads and tcgen05.ld completions, which carry their own fences.)
__device__ __forceinline__ void tmem_alloc2(uint32_t* slot_smem, uint32_t ncols) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_addr(slot_smem)), "r"(ncols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc2(uint32_t taddr, uint32_t ncols) {
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
// D[tmem, both CTAs] (+)= A[smem of each CTA: its 128 rows] * B[smem of each CTA: its half of N]^T, M = 256.
__device__ __forceinline__ void umma_tf32_ss2(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc,
                                              uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void umma_commit2(uint64_t* bar, uint16_t cta_mask) {
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
                 ::"r"(smem_addr(bar)), "h"(cta_mask) : "memory");
}

// 3xTF32 split: the tensor core reads the top 19 bits of an fp32 word (truncation), so
// hi is x itself and lo = x - trunc19(x), exact in fp32.
__device__ __forceinline__ float tf32_lo(float x) {
    return x - __uint_as_float(__float_as_uint(x) & 0xFFFFE000u);
}

}  // namespace tc

// Host: 2D tensor map over a row-major fp32 matrix [rows, row_stride], box = 32 floats x
// box_rows, 128-byte swizzle, zero fill out of bounds. Uses the driver entry point through the
// runtime so the library does not link libcuda directly.
Status make_tmap_rows_sw128(const float* base, uint64_t rows, uint64_t row_stride_floats, uint32_t box_rows,
                            CUtensorMap* out);
// The first `cols` columns of the same matrix as a tensor `cols` wide (columns beyond are out of bounds: zero filled),
// box = 32 floats x box_rows, 128-byte swizzle. Feeds the lane-per-row prefix kernel (prefix_lane.cu).
Status make_tmap_rows_sw128_cols(const float* base, uint64_t rows, uint64_t row_stride_floats, uint32_t cols,
                                 uint32_t box_rows, CUtensorMap* out);
// Same matrix, no swizzle: box = the first `box_cols` floats (<= 256, multiple of 4) x box_rows (<= 256) rows,
// landing densely packed in shared memory. Feeds the prefix scans (only the scored columns leave HBM).
Status make_tmap_rows_prefix(const float* base, uint64_t rows, uint64_t row_stride_floats, uint32_t box_cols,
                             uint32_t box_rows, CUtensorMap* out);

}  // namespace vb
