// Host-side builders of the immutable constants the extraction kernels keep in shared memory:
// periodic Hann window, HTK mel filterbank in tap-list (CSR) form, real-FFT split twiddles, ortho DCT-II.
// Plain C++ (no CUDA), shared by the library (plan creation) and by tests/hostsim.
//
// Reference semantics restated (third-party torchaudio 2.11.0 / torch):
//   window  torch.hann_window(n_fft), periodic                         (audio_feature_extraction.py:34,40)
//   fbank   torchaudio.functional.melscale_fbanks(n_freqs, 0, 8000, n_mels, 16000, None, "htk")
//                                                                      (functional.py:518-587, 492-515)
//   dct     torchaudio.functional.create_dct(40, 128, "ortho")         (functional.py:636-667)
#pragma once
#include <cmath>
#include <cstdint>
#include <vector>

namespace sept {

struct MelTap { int32_t pos; float w; };  // pos: pk2 slot in the frame pair's power tile; w: fb[k][m] * 0.25

inline std::vector<float> make_hann_periodic(int n_fft) {
    // torch evaluates the angle in float32 (arange * (2*pi/N)); mirror that, take the cosine exactly
    std::vector<float> w(n_fft);
    const float step = (float)(2.0 * M_PI / (double)n_fft);
    for (int n = 0; n < n_fft; ++n) {
        float ang = (float)n * step;
        w[n] = (float)(0.5 - 0.5 * std::cos((double)ang));
    }
    return w;
}

// dense (n_freqs x n_mels) row-major, double evaluation of the published formula rounded to float
inline std::vector<float> make_mel_fbank(int n_freqs, int n_mels, int sample_rate, double f_min, double f_max) {
    auto hz2mel = [](double f) { return 2595.0 * std::log10(1.0 + f / 700.0); };
    const double m_min = hz2mel(f_min), m_max = hz2mel(f_max);
    std::vector<double> f_pts(n_mels + 2);
    for (int i = 0; i < n_mels + 2; ++i) {
        double m = m_min + (m_max - m_min) * (double)i / (double)(n_mels + 1);
        f_pts[i] = 700.0 * (std::pow(10.0, m / 2595.0) - 1.0);
    }
    std::vector<float> fb((size_t)n_freqs * n_mels, 0.f);
    for (int k = 0; k < n_freqs; ++k) {
        double f = (double)(sample_rate / 2) * (double)k / (double)(n_freqs - 1);
        for (int m = 0; m < n_mels; ++m) {
            double rising = (f - f_pts[m]) / (f_pts[m + 1] - f_pts[m]);
            double falling = (f_pts[m + 2] - f) / (f_pts[m + 2] - f_pts[m + 1]);
            double v = std::fmax(0.0, std::fmin(rising, falling));
            fb[(size_t)k * n_mels + m] = (float)v;
        }
    }
    return fb;
}

// pk2 slot of real-FFT bin k (0..Nc) inside one frame pair's power tile (must equal Geo<R>::bin_pos):
// rows are k mod 25 with stride 2*(R+1), columns k mod R; the Nyquist bin k = Nc sits in the spare slot R of row 0
inline int power_tile_pos(int k, int R) {
    const int Nc = R * 25;
    if (k == Nc) return R;
    return (k % 25) * (2 * (R + 1)) + (k % R);
}

// tap lists per mel band: band_ptr[m]..band_ptr[m+1] index into taps (k ascending); weights carry the 1/4 of
// the kernel's 4|X|^2 convention
inline void make_mel_taps(int n_fft, int n_mels, int sample_rate, std::vector<int32_t>& band_ptr,
                          std::vector<MelTap>& taps) {
    const int n_freqs = n_fft / 2 + 1, R = n_fft / 50;
    std::vector<float> fb = make_mel_fbank(n_freqs, n_mels, sample_rate, 0.0, (double)(sample_rate / 2));
    band_ptr.assign(n_mels + 1, 0);
    taps.clear();
    for (int m = 0; m < n_mels; ++m) {
        band_ptr[m] = (int32_t)taps.size();
        for (int k = 0; k < n_freqs; ++k) {
            float v = fb[(size_t)k * n_mels + m];
            if (v != 0.f) taps.push_back({power_tile_pos(k, R), v * 0.25f});
        }
    }
    band_ptr[n_mels] = (int32_t)taps.size();
}

// split twiddles W_{n_fft}^{k}, k = CRT(k1, k2) for rows k2 = 0..12, columns k1 = 0..R-1, splatted for the
// packed arithmetic: 4 floats (cos, cos, -sin, -sin) per entry, row stride R
inline std::vector<float> make_split_twiddles(int n_fft) {
    const int R = n_fft / 50, Nc = R * 25;
    std::vector<float> tw((size_t)13 * R * 4, 0.f);
    for (int k2 = 0; k2 <= 12; ++k2)
        for (int k1 = 0; k1 < R; ++k1) {
            int k = -1;
            for (int c = 0; c < Nc; ++c) if (c % R == k1 && c % 25 == k2) { k = c; break; }
            double ang = -2.0 * M_PI * (double)k / (double)n_fft;
            float* e = &tw[((size_t)k2 * R + k1) * 4];
            e[0] = e[1] = (float)std::cos(ang);
            e[2] = e[3] = (float)std::sin(ang);
        }
    return tw;
}

// (n_mels x n_mfcc) row-major.  torch evaluates the angle in float32; mirror that, cosine exactly.
inline std::vector<float> make_dct_ortho(int n_mfcc, int n_mels) {
    std::vector<float> d((size_t)n_mels * n_mfcc);
    const float scale0 = (float)(M_PI / (double)n_mels);
    for (int c = 0; c < n_mfcc; ++c)
        for (int m = 0; m < n_mels; ++m) {
            float ang = scale0 * ((float)m + 0.5f) * (float)c;
            float v = (float)std::cos((double)ang);
            if (c == 0) v = v * (float)(1.0 / std::sqrt(2.0));
            v = v * (float)std::sqrt(2.0 / (double)n_mels);
            d[(size_t)m * n_mfcc + c] = v;
        }
    return d;
}

}  // namespace sept
