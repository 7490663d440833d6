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

// mel bands as contiguous bin runs: band m covers bins k0 .. k0 + 4*nq - 1 with weights w[4*w4 ..]; k0 is a multiple
// of 4, runs are padded with zero weights to whole quads and never reach past bin Nc + 3 (the kernel's power tile keeps
// bins Nc+1 .. Nc+3 at zero).  Weights carry the 1/4 of the kernel's 4|X|^2 convention.
struct MelBand { int32_t k0, w4, nq, pad; };

inline void make_mel_bands(int n_fft, int n_mels, int sample_rate, std::vector<MelBand>& bands, std::vector<float>& weights) {
    const int n_freqs = n_fft / 2 + 1, Nc = n_fft / 2;
    std::vector<float> fb = make_mel_fbank(n_freqs, n_mels, sample_rate, 0.0, (double)(sample_rate / 2));
    bands.assign(n_mels, MelBand{0, 0, 0, 0});
    weights.clear();
    std::vector<int> lo(n_mels, -1), hi(n_mels, -1);
    for (int m = 0; m < n_mels; ++m)
        for (int k = 0; k < n_freqs; ++k)
            if (fb[(size_t)k * n_mels + m] != 0.f) { if (lo[m] < 0) lo[m] = k; hi[m] = k; }
    // the kernel gives band m to lane m % 32 in round m / 32 and runs every lane of a round for the round's widest band:
    // narrower bands are padded with zero-weight quads (their reads must stay inside bins 0 .. Nc + 3)
    for (int r0 = 0; r0 < n_mels; r0 += 32) {
        int round_nq = 1;
        for (int m = r0; m < n_mels && m < r0 + 32; ++m)
            if (lo[m] >= 0) { const int nq = (hi[m] - (lo[m] & ~3) + 4) / 4; if (nq > round_nq) round_nq = nq; }
        for (int m = r0; m < n_mels && m < r0 + 32; ++m) {
            int k0 = lo[m] >= 0 ? (lo[m] & ~3) : 0;                  // quads never straddle the tile's 16-bin blocks
            if (k0 + 4 * round_nq > Nc + 4) k0 = Nc + 4 - 4 * round_nq;
            MelBand b{k0, (int32_t)(weights.size() / 4), round_nq, 0};
            for (int i = 0; i < 4 * round_nq; ++i) {
                const int k = k0 + i;
                weights.push_back(k < n_freqs ? fb[(size_t)k * n_mels + m] * 0.25f : 0.f);
            }
            bands[m] = b;
        }
    }
}

// split twiddles W_{n_fft}^{k}, k = CRT(k1, k2) for rows k2 = 0..12, columns k1 = 0..R-1: (cos, -sin) pairs, row
// stride R + 1 pairs (Geo<R>::TWS)
inline std::vector<float> make_split_twiddles(int n_fft) {
    const int R = n_fft / 50, Nc = R * 25, TWS = R + 1;
    std::vector<float> tw((size_t)13 * TWS * 2, 0.f);
    for (int k2 = 0; k2 <= 12; ++k2)
        for (int k1 = 0; k1 < R; ++k1) {
            int k = -1;
            for (int c = 0; c < Nc; ++c) if (c % R == k1 && c % 25 == k2) { k = c; break; }
            double ang = -2.0 * M_PI * (double)k / (double)n_fft;
            tw[((size_t)k2 * TWS + k1) * 2 + 0] = (float)std::cos(ang);
            tw[((size_t)k2 * TWS + k1) * 2 + 1] = (float)std::sin(ang);
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

// Polyphase rows of torchaudio's sinc_interp_hann resampling kernel (functional.py: _get_sinc_resample_kernel,
// lowpass_filter_width 6, rolloff 0.99), evaluated in double like torchaudio (dtype None) and rounded to float, each row
// cut to a common-length window [k_lo[j], k_lo[j] + taps) that covers its non-zero support.  orig/up are the
// gcd-reduced rates.  Returns torchaudio's `width`.
inline int make_resample_rows(int orig, int up, std::vector<int32_t>& k_lo, std::vector<float>& rows, int& taps) {
    const int lpw = 6;
    const double rolloff = 0.99;
    const double base_freq = (double)(orig < up ? orig : up) * rolloff;
    const int width = (int)std::ceil((double)lpw * orig / base_freq);
    const int K = 2 * width + orig;
    std::vector<double> full((size_t)up * K);
    std::vector<int> first(up, K), last(up, -1);
    for (int j = 0; j < up; ++j)
        for (int k = 0; k < K; ++k) {
            double t = ((double)(-j) / up + (double)(k - width) / orig) * base_freq;
            const bool inside = t > -lpw && t < lpw;
            if (t < -lpw) t = -lpw;
            if (t > lpw) t = lpw;
            const double window = std::pow(std::cos(t * M_PI / lpw / 2.0), 2.0);
            const double tp = t * M_PI;
            const double v = (tp == 0.0 ? 1.0 : std::sin(tp) / tp) * window * (base_freq / orig);
            full[(size_t)j * K + k] = v;
            if (inside) { if (k < first[j]) first[j] = k; if (k > last[j]) last[j] = k; }
        }
    taps = 0;
    for (int j = 0; j < up; ++j) if (last[j] - first[j] + 1 > taps) taps = last[j] - first[j] + 1;
    k_lo.assign(up, 0);
    rows.assign((size_t)up * taps, 0.f);
    for (int j = 0; j < up; ++j) {
        int lo = first[j];
        if (lo + taps > K) lo = K - taps;
        k_lo[j] = lo;
        for (int k = 0; k < taps; ++k) rows[(size_t)j * taps + k] = (float)full[(size_t)j * K + lo + k];
    }
    return width;
}

}  // namespace sept
