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
#include <algorithm>
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

// ---- mel filterbank as a per-lane gather program -------------------------------------------------------------------
// The HTK triangles (norm=None) partition the bins into INTERVALS between neighbouring mel points: a bin in interval
// i (f_pts[i] <= f < f_pts[i+1]) lies on the rising slope of band i and on the falling slope of band i-1 and nowhere
// else, so   mel[b] = U[b] + D[b+1],   U[i] = sum_{k in interval i} fb[k][i] P[k],   D[i] = sum fb[k][i-1] P[k],
// and every power value is read ONCE for both of its bands.  The kernel gives interval 1 + 32 r + l to lane l in
// round r; a round runs for as many steps as its widest interval has bins.  One MelStep = {rising weight, falling
// weight, byte offset of the bin inside a frame pair's power tile}; rows of `width` entries per step.  Interval 0
// (only its rising side is used, by band 0) is a short broadcast list in front (`head`).
// The order in which a lane visits its bins is scheduled so that the 16 lanes of a half warp hit 16 different
// 8-byte bank pairs in (nearly) every step; slots without a bin read the tile's zero slot (one broadcast word).  Weights
// carry the 1/4 of the kernel's 4|X|^2, so rising + falling = 1/4 for every bin.
// pk2 slot of bin k in a frame pair's power tile (the same function as Geo<R>::bin_pos in extract_core.cuh, R = n_fft / 50;
// hostsim checks that they agree)
inline int power_tile_pos(int n_fft, int k) {
    return n_fft == 1600 ? k : n_fft == 800 ? k + ((9 * (k >> 4)) >> 1) : k + ((9 * (k >> 3)) >> 2);
}

// first of the kZeroSlots consecutive slots of every pair's power tile that the split phase keeps at zero: what idle
// program slots read (three bank pairs to choose from, so that the broadcast word rarely collides with a real read)
constexpr int kZeroSlots = 3;
inline int power_tile_zero_slot(int n_fft) { return power_tile_pos(n_fft, n_fft / 2) + 1; }

struct MelStep { float up, dn; int32_t off, pad; };   // pad: step count of the round in the entries of its first step

struct MelProgram {
    std::vector<MelStep> entries;          // [n_head] broadcast entries, then [total_steps][width]
    std::vector<int> round_steps;          // steps of every round, ceil(n_mels / 32) rounds
    int n_head = 0, width = 32, total_steps = 0;
};

inline void make_mel_program(int n_fft, int n_mels, int sample_rate, MelProgram& prog) {
    auto pos = [n_fft](int k) { return power_tile_pos(n_fft, k); };
    const int zslot = power_tile_zero_slot(n_fft);
    const int n_freqs = n_fft / 2 + 1;
    const double f_max = (double)(sample_rate / 2);
    std::vector<float> fb = make_mel_fbank(n_freqs, n_mels, sample_rate, 0.0, f_max);
    auto hz2mel = [](double f) { return 2595.0 * std::log10(1.0 + f / 700.0); };
    std::vector<double> f_pts(n_mels + 2);
    for (int i = 0; i < n_mels + 2; ++i)
        f_pts[i] = 700.0 * (std::pow(10.0, hz2mel(f_max) * (double)i / (double)(n_mels + 1) / 2595.0) - 1.0);
    // bins of every interval: the bands with a non-zero weight at bin k are a subset of {i-1, i} for ONE i; take i from
    // the weights themselves (robust against rounding at the mel points)
    std::vector<std::vector<int>> bins(n_mels + 1);
    for (int k = 0; k < n_freqs; ++k) {
        int first = -1, last = -1;
        for (int m = 0; m < n_mels; ++m)
            if (fb[(size_t)k * n_mels + m] != 0.f) { if (first < 0) first = m; last = m; }
        if (first < 0) continue;                                     // DC, Nyquist: no band
        int i;
        if (last == first + 1) i = last;
        else {                                                       // one band only: which slope?
            const double f = f_max * (double)k / (double)(n_freqs - 1);
            i = f < f_pts[first + 1] ? first : first + 1;
        }
        bins[i].push_back(k);
    }
    auto up_w = [&](int k, int i) { return i < n_mels ? 0.25f * fb[(size_t)k * n_mels + i] : 0.f; };
    auto dn_w = [&](int k, int i) { return i >= 1 ? 0.25f * fb[(size_t)k * n_mels + i - 1] : 0.f; };
    prog.entries.clear();
    prog.round_steps.clear();
    prog.width = n_mels < 32 ? n_mels : 32;
    for (int k : bins[0]) prog.entries.push_back(MelStep{up_w(k, 0), 0.f, 8 * pos(k), 0});
    prog.n_head = (int)prog.entries.size();
    prog.total_steps = 0;
    const int W = prog.width;
    for (int r0 = 0; r0 < n_mels; r0 += 32) {
        std::vector<std::vector<int>> rem(W);
        int n_steps = 1;
        for (int l = 0; l < W; ++l)
            if (r0 + l < n_mels) { rem[l] = bins[1 + r0 + l]; if ((int)rem[l].size() > n_steps) n_steps = (int)rem[l].size(); }
        for (int s = 0; s < n_steps; ++s) {
            const int left = n_steps - s;
            std::vector<int> pick(W, -1);
            for (int h = 0; h < W; h += 16) {
                const int h_end = h + 16 < W ? h + 16 : W;
                int used[16] = {0};
                bool may_idle = false;                               // some lane of this half warp may idle: keep one zero slot's bank pair free
                for (int l = h; l < h_end; ++l) if ((int)rem[l].size() < left) may_idle = true;
                // lanes with the least slack choose first
                std::vector<int> order;
                for (int l = h; l < h_end; ++l) order.push_back(l);
                std::stable_sort(order.begin(), order.end(), [&](int a, int b) {
                    return left - (int)rem[a].size() < left - (int)rem[b].size();
                });
                auto free_zero_banks = [&]() { int n = 0; for (int z = 0; z < kZeroSlots; ++z) n += used[(zslot + z) & 15] == 0; return n; };
                auto load = [&](int k) {
                    const int b = pos(k) & 15;
                    const bool last_zero = may_idle && used[b] == 0 && ((b - zslot) & 15) < kZeroSlots && free_zero_banks() == 1;
                    return used[b] + (last_zero ? 1 : 0);
                };
                for (int l : order) {
                    if (rem[l].empty()) continue;
                    const int slack = left - (int)rem[l].size();
                    int best = -1;
                    for (size_t c = 0; c < rem[l].size(); ++c)
                        if (load(rem[l][c]) == 0) { best = (int)c; break; }
                    if (best < 0 && slack > 0) continue;             // wait for a free bank pair
                    if (best < 0) {
                        best = 0;
                        for (size_t c = 1; c < rem[l].size(); ++c)
                            if (load(rem[l][c]) < load(rem[l][best])) best = (int)c;
                    }
                    pick[l] = rem[l][best];
                    ++used[pos(pick[l]) & 15];
                    rem[l].erase(rem[l].begin() + best);
                }
                int z_best = 0;                                      // idle lanes of the half warp share the emptiest zero slot
                for (int z = 1; z < kZeroSlots; ++z)
                    if (used[(zslot + z) & 15] < used[(zslot + z_best) & 15]) z_best = z;
                for (int l = h; l < h_end; ++l) if (pick[l] < 0) pick[l] = -2 - z_best;
            }
            for (int l = 0; l < W; ++l) {
                const int i = 1 + r0 + l;
                if (pick[l] >= 0) {
                    const int k = pick[l];
                    // the last interval has no band above it: its rising weight is never used, store the complement of
                    // the falling one so that "falling = 1/4 - rising" holds for every entry (unrolled path, extract.cu)
                    const float dn = dn_w(k, i), up = i < n_mels ? up_w(k, i) : 0.25f - dn;
                    prog.entries.push_back(MelStep{up, dn, 8 * pos(k), 0});
                } else {
                    prog.entries.push_back(MelStep{0.25f, 0.f, 8 * (zslot + (-2 - pick[l])), 0});   // idle: 1/4 x a zero slot
                }
            }
        }
        for (int l = 0; l < W; ++l) prog.entries[prog.entries.size() - (size_t)n_steps * W + l].pad = n_steps;
        prog.round_steps.push_back(n_steps);
        prog.total_steps += n_steps;
    }
}

// shared-memory wavefronts of the program's power-tile gathers for ONE frame pair (8-byte reads: two half warps per
// step, each costing the largest number of distinct 8-byte words that share a bank pair); lower bound 2 per step
inline int mel_program_wavefronts(const MelProgram& prog) {
    int total = 0;
    const int W = prog.width;
    for (int s = 0; s < prog.total_steps; ++s)
        for (int h = 0; h < W; h += 16) {
            int worst = 0;
            for (int b = 0; b < 16; ++b) {
                std::vector<int> words;
                for (int l = h; l < h + 16 && l < W; ++l) {
                    const int w = prog.entries[prog.n_head + (size_t)s * W + l].off / 8;
                    if ((w & 15) == b && std::find(words.begin(), words.end(), w) == words.end()) words.push_back(w);
                }
                if ((int)words.size() > worst) worst = (int)words.size();
            }
            total += worst;
        }
    return total;
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


// Register-tile form of the same rows for the tiled resampling kernel: phases are grouped four at a time; group g reads
// the input window that starts at tap base[g] = min k_lo of its phases and is TG taps long (TG odd: conflict-free 16-byte
// reads), and wt[(g * TG + i) * 4 + jj] is the weight of phase 4 g + jj at window position i (zero outside the phase's own
// support, and for the phases past `up` in the last group).
struct ResampleTiles {
    std::vector<int32_t> base;   // [n_groups]
    std::vector<float> wt;       // [n_groups][TG][4]
    int n_groups = 0, tg = 0, base_min = 0, base_max = 0;
};

inline void make_resample_tiles(int up, int taps, const std::vector<int32_t>& k_lo, const std::vector<float>& rows, ResampleTiles& t) {
    t.n_groups = (up + 3) / 4;
    t.base.assign(t.n_groups, 0);
    int span = 0;
    for (int g = 0; g < t.n_groups; ++g) {
        int lo = k_lo[4 * g], hi = lo;
        for (int j = 4 * g; j < 4 * g + 4 && j < up; ++j) { if (k_lo[j] < lo) lo = k_lo[j]; if (k_lo[j] > hi) hi = k_lo[j]; }
        t.base[g] = lo;
        if (hi - lo > span) span = hi - lo;
    }
    t.tg = (taps + span) | 1;
    t.wt.assign((size_t)t.n_groups * t.tg * 4, 0.f);
    t.base_min = t.base[0]; t.base_max = t.base[0];
    for (int g = 0; g < t.n_groups; ++g) {
        if (t.base[g] < t.base_min) t.base_min = t.base[g];
        if (t.base[g] > t.base_max) t.base_max = t.base[g];
        for (int jj = 0; jj < 4; ++jj) {
            const int j = 4 * g + jj;
            if (j >= up) continue;
            for (int k = 0; k < taps; ++k) t.wt[((size_t)g * t.tg + (k_lo[j] - t.base[g]) + k) * 4 + jj] = rows[(size_t)j * taps + k];
        }
    }
}

}  // namespace sept
