// Class-balance noise augmentation on the device (sm_100a).
//
// The reference tops every minority class up to the size of the largest one by re-drawing windows of that class and adding
// N(0, 0.05) noise (preprocess_data/preprocess_adversary_data.py:392-421).  In the reference the new key ALIASES the dict of
// the window it was drawn from and the noisy array is written through that alias, so the source window and all of its
// copies end up sharing one array: the original plus every noise sample drawn for it, added in draw order.  The host plan
// (augmentation.py) reproduces the draws; this kernel applies them: row r <- ((row r + n_1) + n_2) + ... for the draws
// of every source row.  HBM-bound: one read and one write of each touched row, float4 accesses.
#include <cuda_runtime.h>
#include <cstdint>

#include "augment.h"
#include "philox.cuh"

namespace sept {

constexpr int kAugThreads = 256;

__global__ void __launch_bounds__(kAugThreads) add_noise_kernel(const AddNoiseParams p) {
    const int job = blockIdx.y;
    const int quads = p.row_elems / 4;
    const int first = p.job_ptr[job], last = p.job_ptr[job + 1];
    float* row = p.data + p.job_row[job] * (int64_t)p.row_elems;
    for (int i4 = blockIdx.x * kAugThreads + threadIdx.x; i4 < quads; i4 += gridDim.x * kAugThreads) {
        float4 v = *reinterpret_cast<const float4*>(row + 4 * i4);
        for (int d = first; d < last; ++d) {
            const int64_t draw = p.draw_id[d];
            float4 n;
            if (p.noise) n = *reinterpret_cast<const float4*>(p.noise + draw * (int64_t)p.row_elems + 4 * i4);
            else n = normal4(p.seed, (uint64_t)draw * (uint64_t)quads, (uint32_t)i4, p.std);
            v.x += n.x; v.y += n.y; v.z += n.z; v.w += n.w;
        }
        *reinterpret_cast<float4*>(row + 4 * i4) = v;
    }
}

cudaError_t launch_add_noise(const AddNoiseParams& p, cudaStream_t stream) {
    if (p.n_jobs == 0) return cudaSuccess;
    const int quads = p.row_elems / 4;
    int gx = (quads + kAugThreads - 1) / kAugThreads;
    if (gx > 32) gx = 32;
    add_noise_kernel<<<dim3((unsigned)gx, (unsigned)p.n_jobs), kAugThreads, 0, stream>>>(p);
    return cudaGetLastError();
}

}  // namespace sept
