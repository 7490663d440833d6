// Band-limited sinc resampling of a ragged batch (sm_100a), e.g. 44.1 kHz -> 16 kHz ahead of framing.
//
// Replaces torchaudio.transforms.Resample(sample_rate, 16000) as called for the MSP-Improv corpus
// (feature_extraction/audio_feature_extraction.py:139-141 -> torchaudio/functional/functional.py
// _get_sinc_resample_kernel / _apply_sinc_resample_kernel): Hann-windowed sinc, lowpass_filter_width 6, rolloff 0.99,
// evaluated as `up` polyphase FIR rows applied with stride `orig`.  torchaudio convolves every row over its full
// 2*width + orig taps; outside |t| < lowpass_filter_width the window is cos^2(pi/2) ~ 1e-33, so each row is stored and
// applied on its ~2*width + 2 tap support only (the dropped taps are below fp32 resolution of the sum by > 25 orders).
//
// One thread per output sample: y[m*up + j] = sum_k w[j][k] * x[m*orig + k - width].  Neighbouring outputs read
// overlapping input windows, which L1 serves; the 160 x 36 weight table stays L1 resident.  HBM-bound by a wide margin
// is not reached: 34 FMAs per 6.8 bytes of traffic put it at the L1/FMA balance point.
#include <cuda_runtime.h>
#include <cstdint>

#include "resample.h"

namespace sept {

constexpr int kResampleThreads = 256;

__global__ void __launch_bounds__(kResampleThreads) resample_kernel(const ResampleParams p) {
    __shared__ int u_first;
    const long long g0 = (long long)blockIdx.x * kResampleThreads;
    if (threadIdx.x == 0) {
        int lo = 0, hi = p.n_utts - 1;                         // utterance of the block's first output sample
        while (lo < hi) {
            const int mid = (lo + hi) >> 1;
            if (p.out_off[mid + 1] > g0) hi = mid; else lo = mid + 1;
        }
        u_first = lo;
    }
    __syncthreads();
    const long long g = g0 + threadIdx.x;
    if (g >= p.total_out) return;
    int u = u_first;
    while (g >= p.out_off[u + 1]) ++u;                         // a block rarely spans more than two utterances
    const long long o = g - p.out_off[u];
    const long long m = o / p.up;
    const int j = (int)(o - m * p.up);
    const long long base = p.in_off[u];
    const long long n_in = p.in_off[u + 1] - base;
    const int k0 = p.k_lo[j];
    const float* w = p.w + (long long)j * p.taps;
    const long long i0 = m * p.orig + k0 - p.width;            // input index of the first stored tap
    const float* x = p.in + base;
    float acc = 0.f;
    if (i0 >= 0 && i0 + p.taps <= n_in) {
#pragma unroll 4
        for (int k = 0; k < p.taps; ++k) acc = fmaf(__ldg(w + k), __ldg(x + i0 + k), acc);
    } else {
        for (int k = 0; k < p.taps; ++k) {
            const long long i = i0 + k;
            if (i >= 0 && i < n_in) acc = fmaf(__ldg(w + k), __ldg(x + i), acc);   // zero padding outside the utterance
        }
    }
    p.out[g] = acc;
}

// 16-bit PCM -> float in [-1, 1): x / 32768, what torchaudio.load(normalize=True) does on the host before the reference's
// callables see the audio (audio_feature_extraction.py:182).  Lets a bulk job ship half the bytes over PCIe.
__global__ void __launch_bounds__(256) pcm16_to_f32_kernel(const short* __restrict__ in, long long n, float* __restrict__ out) {
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) out[i] = (float)in[i] * (1.0f / 32768.0f);
}

cudaError_t launch_pcm16_to_f32(const short* in, long long n, float* out, cudaStream_t stream) {
    if (n <= 0) return cudaSuccess;
    long long blocks = (n + 255) / 256;
    if (blocks > 148 * 16) blocks = 148 * 16;
    pcm16_to_f32_kernel<<<(unsigned)blocks, 256, 0, stream>>>(in, n, out);
    return cudaGetLastError();
}

cudaError_t launch_resample(const ResampleParams& p, cudaStream_t stream) {
    if (p.total_out <= 0) return cudaSuccess;
    const long long blocks = (p.total_out + kResampleThreads - 1) / kResampleThreads;
    resample_kernel<<<(unsigned)blocks, kResampleThreads, 0, stream>>>(p);
    return cudaGetLastError();
}

}  // namespace sept
