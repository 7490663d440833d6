// Launch parameters of the class-balance noise augmentation kernel (augment.cu).  Internal header.
#pragma once
#include <cuda_runtime.h>
#include <cstdint>

namespace sept {

struct AddNoiseParams {
    float* data;               // (n_rows, row_elems) windows, updated in place
    const int64_t* job_row;    // [n_jobs] row every job updates (distinct rows)
    const int32_t* job_ptr;    // [n_jobs + 1] CSR into draw_id
    const int64_t* draw_id;    // [total draws] global draw index of every noise sample added to the row, in order
    int n_jobs;
    int row_elems;             // multiple of 4
    uint64_t seed;
    float std;                 // 0.05 in the reference
    const float* noise;        // null: Philox(seed, draw); else (n_draws_total, row_elems) supplied by the caller
};

cudaError_t launch_add_noise(const AddNoiseParams& p, cudaStream_t stream);

}  // namespace sept
