// Launch parameters of the per-speaker normalisation kernels (norm.cu).  Internal header.
#pragma once
#include <cuda_runtime.h>
#include <cstddef>
#include <cstdint>

namespace sept {

constexpr int kStatRows = 5;   // per (utterance | speaker) and feature: n, mean, M2 | std, min, max

struct SpeakerStatsParams {
    const float* feat;           // (total_frames, F) frame major
    const int64_t* frame_off;    // [n_utts + 1]
    const uint8_t* whole;        // [n_utts] or null: 1 = utterance is appended whole (test split), 0 = per window
    int n_utts, n_feat;
    int win_len, shift_len;      // preprocess_adversary_data.py:131 (200, 50)
    float* utt_partial;          // workspace (n_utts, 5, F): n, mean, M2, min, max
    const int32_t* spk_ptr;      // [n_spk + 1] CSR over spk_utts
    const int32_t* spk_utts;     // utterance ids grouped by speaker
    int n_spk;
    float* stats;                // (n_spk, 5, F): count, mean, std (ddof 0), min, max
};

struct NormalizeParams {
    const float* feat;           // (total_frames, F)
    const int64_t* frame_off;    // [n_utts + 1]
    const int32_t* spk_of_utt;   // [n_utts]
    const float* stats;          // (n_spk, 5, F)
    int n_feat;
    int mode;                    // 0 znorm, 1 min_max
    // frame mode: one CTA per utterance, out (total_frames, F)
    int n_utts;
    // window mode: one CTA per window, out (n_windows, win_len, F); rows past the utterance end are the
    // normalised zero padding of preprocess_adversary_data.py:29-35
    const int32_t* win_utt;      // [n_windows] or null (frame mode)
    const int32_t* win_t0;       // [n_windows]
    int n_windows, win_len;
    float* out;
};

cudaError_t launch_speaker_stats(const SpeakerStatsParams& p, cudaStream_t stream);
cudaError_t launch_normalize(const NormalizeParams& p, cudaStream_t stream);

}  // namespace sept
