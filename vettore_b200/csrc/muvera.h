// muvera.h — MUVERA fixed-dimensional encoding, batched on the device (see muvera.cu).
#pragma once
#include "runtime.h"

namespace vb {

struct MuveraConfig {       // muvera.rs:13-21
    size_t dimension = 0, num_repetitions = 0, num_simhash_projections = 0;
    uint64_t seed = 0;
    size_t projection_dimension = 0;
    bool has_final = false;               // final_projection_dimension is Some(...)
    size_t final_projection_dimension = 0;
};

// Validates the configuration (muvera.rs:77-108, :29-44) and returns the encoding sizes before / after the
// optional count sketch.
Status muvera_output_dimension(const MuveraConfig& c, size_t* full_size, size_t* final_size);

// Encodes `ndocs` multi-vectors: document i owns vectors [doc_vec[i], doc_vec[i+1]) of the ragged list
// (vals, vec_off). mode 0 = query (sum per partition), 1 = document (running average). out: host,
// [ndocs][final_size]. Bit-identical to the reference arithmetic (order-preserving f64 accumulation).
Status muvera_encode_batch(SearchCtx& ctx, const MuveraConfig& c, size_t ndocs, const float* vals, const uint64_t* vec_off,
                           const uint64_t* doc_vec, int mode, float* out, size_t out_capacity);

}  // namespace vb
