// Per-speaker feature statistics and normalisation (sm_100a).
//
// The reference appends every frame of every training window to a per-speaker Python list and reduces it with
// numpy (preprocess_data/preprocess_adversary_data.py:26-27, 41-48, 357-364): a frame that lies in k overlapping
// windows is counted k times.  Here the multiplicity is a closed-form weight and the statistics are weighted
// Welford accumulations: one CTA per utterance (thread = feature, coalesced rows), then one CTA per speaker merges
// its utterances' partials with Chan's formula in list order (deterministic, no atomics).
//   znorm   (x - mean) / (std + 1e-5)            :378
//   min_max (x - min) / (max - min) * 2 - 1      :380
#include <cuda_runtime.h>
#include <cfloat>
#include <cstdint>

#include "norm.h"

namespace sept {

// number of training windows (start i*shift, length win) that contain frame t of a T-frame utterance (:44-48)
__device__ __forceinline__ int frame_weight(int t, int T, int win, int shift, bool whole) {
    if (whole || T < win) return 1;
    const int n_win = (T - win) / shift + 1;
    int i_hi = t / shift;
    if (i_hi > n_win - 1) i_hi = n_win - 1;
    const int num = t - win + 1;
    const int i_lo = num <= 0 ? 0 : (num + shift - 1) / shift;
    return i_hi >= i_lo ? i_hi - i_lo + 1 : 0;
}

__global__ void __launch_bounds__(128) utt_partial_kernel(const SpeakerStatsParams p) {
    const int u = blockIdx.x;
    const long long f0 = p.frame_off[u];
    const int T = (int)(p.frame_off[u + 1] - f0);
    const bool whole = p.whole ? p.whole[u] != 0 : false;
    const int F = p.n_feat;
    for (int f = threadIdx.x; f < F; f += blockDim.x) {
        float n = 0.f, mean = 0.f, m2 = 0.f, mn = FLT_MAX, mx = -FLT_MAX;
        const float* col = p.feat + f0 * F + f;
        for (int t = 0; t < T; ++t) {
            const int w = frame_weight(t, T, p.win_len, p.shift_len, whole);
            if (w == 0) continue;
            const float x = col[(long long)t * F];
            n += (float)w;
            const float d = x - mean;
            mean += d * ((float)w / n);
            m2 += (float)w * d * (x - mean);
            mn = fminf(mn, x);
            mx = fmaxf(mx, x);
        }
        float* o = p.utt_partial + ((long long)u * kStatRows) * F + f;
        o[0] = n; o[F] = mean; o[2 * F] = m2; o[3 * F] = mn; o[4 * F] = mx;
    }
}

__global__ void __launch_bounds__(128) speaker_merge_kernel(const SpeakerStatsParams p) {
    const int s = blockIdx.x;
    const int F = p.n_feat;
    for (int f = threadIdx.x; f < F; f += blockDim.x) {
        float n = 0.f, mean = 0.f, m2 = 0.f, mn = FLT_MAX, mx = -FLT_MAX;
        for (int j = p.spk_ptr[s]; j < p.spk_ptr[s + 1]; ++j) {
            const float* o = p.utt_partial + ((long long)p.spk_utts[j] * kStatRows) * F + f;
            const float nb = o[0];
            if (nb == 0.f) continue;
            const float tot = n + nb, d = o[F] - mean;
            mean += d * (nb / tot);
            m2 += o[2 * F] + d * d * (n * nb / tot);
            n = tot;
            mn = fminf(mn, o[3 * F]);
            mx = fmaxf(mx, o[4 * F]);
        }
        float* r = p.stats + ((long long)s * kStatRows) * F + f;
        r[0] = n; r[F] = mean; r[2 * F] = n > 0.f ? sqrtf(m2 / n) : 0.f; r[3 * F] = mn; r[4 * F] = mx;
    }
}

__device__ __forceinline__ float normalize_one(float x, float mean, float sd, float mn, float mx, int mode) {
    if (mode == 0) return (x - mean) / (sd + 1e-5f);
    return (x - mn) / (mx - mn) * 2.0f - 1.0f;
}

__global__ void __launch_bounds__(256) normalize_kernel(const NormalizeParams p) {
    const int F = p.n_feat;
    int u, t_begin, rows, T;
    float* out;
    if (p.win_utt) {
        u = p.win_utt[blockIdx.x];
        t_begin = p.win_t0[blockIdx.x];
        rows = p.win_len;
        out = p.out + (long long)blockIdx.x * p.win_len * F;
    } else {
        u = blockIdx.x;
        t_begin = 0;
        rows = (int)(p.frame_off[u + 1] - p.frame_off[u]);
        out = p.out + p.frame_off[u] * F;
    }
    const long long f0 = p.frame_off[u];
    T = (int)(p.frame_off[u + 1] - f0);
    const float* st = p.stats + ((long long)p.spk_of_utt[u] * kStatRows) * F;
    const int total = rows * F;
    for (int i = threadIdx.x; i < total; i += blockDim.x) {
        const int r = i / F, f = i - r * F;
        const int t = t_begin + r;
        const float x = t < T ? p.feat[(f0 + t) * F + f] : 0.f;
        out[i] = normalize_one(x, st[F + f], st[2 * F + f], st[3 * F + f], st[4 * F + f], p.mode);
    }
}

cudaError_t launch_speaker_stats(const SpeakerStatsParams& p, cudaStream_t stream) {
    if (p.n_utts > 0) utt_partial_kernel<<<p.n_utts, 128, 0, stream>>>(p);
    if (p.n_spk > 0) speaker_merge_kernel<<<p.n_spk, 128, 0, stream>>>(p);
    return cudaGetLastError();
}

cudaError_t launch_normalize(const NormalizeParams& p, cudaStream_t stream) {
    const int blocks = p.win_utt ? p.n_windows : p.n_utts;
    if (blocks > 0) normalize_kernel<<<blocks, 256, 0, stream>>>(p);
    return cudaGetLastError();
}

}  // namespace sept
