// Launch parameters of the cloak / gradient-reversal kernels (cloak.cu).  Internal header.
#pragma once
#include <cuda_runtime.h>
#include <cstddef>
#include <cstdint>

namespace sept {

constexpr int kCloakSlices = 16;     // batch slices (grid.y); also the row count of the backward's partial buffer

struct CloakFwdParams {
    const float* x;          // (B, wf)
    const float* locs;       // (wf)
    const float* rhos;       // (wf)
    const float* mask;       // (wf) or null
    const float* eps;        // (wf) or null -> Philox(seed, offset)
    uint64_t seed, offset;
    const uint64_t* draw_dev; // null, or device counter: the Philox offset becomes offset + *draw_dev * ceil(wf / 4)
    int per_sample;          // 0: one (wf) eps for the whole batch (the reference's forward); 1: batch element b has its own
                             // eps -- eps / eps_out are (B, wf), Philox draw index = *draw_dev + b (a batched stand-in for
                             // the reference's one-forward-per-window evaluation loop, adversary_cloak_evaluation.py:73-83)
    float eps_std;           // 0.1 in the reference (cloak_models.py:37)
    float min_scale, max_scale;
    int batch, wf;
    float* out;              // (B, wf)
    float* eps_out;          // (wf) or null: the eps that was used (unmasked)
    float* noise_out;        // (wf) or null: locs + sigma * eps * mask  (sample_noise())
};

struct CloakBwdParams {
    const float* g_a;        // (B, wf) gradient wrt the noisy output
    const float* g_b;        // (B, wf) or null: second upstream gradient, reversed with lambda
    float lambda;
    const float* eps;        // (wf) as returned by the forward
    const float* rhos;       // (wf)
    const float* mask;       // (wf) or null
    float min_scale, max_scale;
    float reg_coef;          // added to d loss / d sigma_i (0 for none)
    int batch, wf;
    float* partial;          // workspace (kCloakSlices, wf)
    unsigned* counters;      // workspace, ceil(wf / 512) zero-initialised words; left zeroed
    float* dlocs;            // (wf)
    float* drhos;            // (wf) or null
    float* dx;               // (B, wf) or null
};

int cloak_slices(int batch);
cudaError_t launch_cloak_fwd(const CloakFwdParams& p, cudaStream_t stream);
cudaError_t launch_cloak_bwd(const CloakBwdParams& p, cudaStream_t stream);
cudaError_t launch_counter_add(uint64_t* counter, uint64_t inc, cudaStream_t stream);
cudaError_t launch_grl_bwd(const float* g, float lambda, size_t n, float* dx, cudaStream_t stream);

}  // namespace sept
