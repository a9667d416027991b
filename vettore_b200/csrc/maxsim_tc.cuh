// maxsim_tc.cuh — what the two tensor-core MaxSim kernels share (maxsim_tc.cu: uniform documents,
// maxsim_tcr.cu: ragged documents): warp roles, ring / TMEM geometry and the transposing max butterflies.
#pragma once
#include "tc.cuh"

namespace vb {

constexpr int kTcEpiWarps = 8, kTcSplitWarps = 8;   // epilogue: two groups of 4 warps, alternating tiles
constexpr int kTcEpiGroupWarps = 4;
constexpr int kTcProducerWarp = kTcEpiWarps + kTcSplitWarps, kTcMmaWarp = kTcProducerWarp + 1;
constexpr int kTcThreads = (kTcMmaWarp + 1) * 32;
constexpr int kTcStages = 4;     // ring stages of one chunk (<= 2 K blocks = 32 KB) each
constexpr int kTcAccBufs = 2;     // accumulator buffers (MMA <-> epilogue double buffering)
constexpr int kTcChains = 4;      // independent accumulators per buffer: consecutive MMAs never depend on each other
constexpr int kTcTile = 128;     // tokens per tile (UMMA M)
constexpr int kTcN = 32;         // query tokens (UMMA N), zero padded
constexpr uint32_t kTcChunkBytes = 2 * 16384;   // 2 K blocks of [128 rows x 128 B]
// TMEM columns: A operand double-buffered per chunk, buffer u at [128 u, +128): hi [0,64) lo [64,128);
// accumulator buffer b at 256 + 128 b: kTcChains partial accumulators of 32 columns (summed by the epilogue).
constexpr uint32_t kTcAccCol = 256;


// Max over the 32 lanes of a warp for 32 per-lane values at once: after the butterfly lane q
// holds max over lanes of v[q]. 16 + 8 + 4 + 2 + 1 = 31 shuffles instead of 32 x 5.
__device__ __forceinline__ float warp_transpose_max(float (&v)[32], int lane) {
#pragma unroll
    for (int half = 16; half >= 1; half >>= 1) {
        const bool hi = (lane & half) != 0;
#pragma unroll
        for (int i = 0; i < half; ++i) {
            const float send = hi ? v[i] : v[i + half];
            const float keep = hi ? v[i + half] : v[i];
            v[i] = fmaxf(keep, __shfl_xor_sync(0xffffffffu, send, half));
        }
    }
    return v[0];
}

// Same for 16 per-lane values: 8 + 4 + 2 + 1 exchanges, then one plain step; lanes 2c and 2c + 1 hold column c.
__device__ __forceinline__ float warp_transpose_max16(float (&v)[16], int lane) {
#pragma unroll
    for (int half = 8; half >= 1; half >>= 1) {
        const bool hi = (lane & (half * 2)) != 0;
#pragma unroll
        for (int i = 0; i < half; ++i) {
            const float send = hi ? v[i] : v[i + half];
            const float keep = hi ? v[i + half] : v[i];
            v[i] = fmaxf(keep, __shfl_xor_sync(0xffffffffu, send, half * 2));
        }
    }
    return fmaxf(v[0], __shfl_xor_sync(0xffffffffu, v[0], 1));
}

}  // namespace vb
