// Host emulation of one warp of extract_kernel (TEST CODE, not a fallback): runs the very same per-lane phase
// functions the CUDA kernel calls (csrc/extract_core.cuh) in a loop over lane ids, with plain arrays standing in
// for shared memory, so the index algebra is checked on a CPU-only box.
//   hostsim <n_fft> <hop> <n_mels> <deriv> <wav.f32> <out.f32>      writes (T, n_mels) log-mel dB
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "extract_core.cuh"
#include "tables.h"

using namespace sept;

template <int R>
int run(int hop, int n_mels, int deriv, const std::vector<float>& wav, const char* out_path) {
    using G = Geo<R>;
    const int n = (int)wav.size(), n_fft = G::NFFT;
    const int T = 1 + n / hop;
    std::vector<float> win = make_hann_periodic(n_fft), tws = make_split_twiddles(n_fft);
    std::vector<int32_t> bp;
    std::vector<MelTap> taps;
    make_mel_taps(n_fft, n_mels, 16000, bp, taps);
    std::vector<float> stage(G::span(hop) + 4);
    std::vector<pk4> Y(G::Y_PK4);
    pk2* P = reinterpret_cast<pk2*>(Y.data());
    std::vector<float> out((size_t)T * n_mels);
    for (int t0 = 0; t0 < T; t0 += G::FPW) {
        for (int lane = 0; lane < 32; ++lane) stage_item<G>(lane, wav.data(), n, t0, hop, deriv, stage.data());
        for (int lane = 0; lane < 32; ++lane)
            pass1<G>(lane, stage.data(), hop, reinterpret_cast<const f2*>(win.data()), Y.data());
        for (int task = 0; task < G::P2_TASKS; ++task) pass2_row<G>(task, Y.data());
        for (int k2 = 0; k2 <= 12; ++k2) {
            pk2 a[32], b[32];
            bool on[32];
            for (int lane = 0; lane < 32; ++lane)
                on[lane] = split_load<G>(lane, k2, Y.data(), reinterpret_cast<const f4*>(tws.data()), a[lane], b[lane]);
            for (int lane = 0; lane < 32; ++lane)
                if (on[lane]) split_store<G>(lane, k2, P, a[lane], b[lane]);
        }
        for (int task = 0; task < G::PPW * n_mels; ++task) {
            const int p = task / n_mels, m = task % n_mels;
            const pk2 acc = mel_band(P + p * (2 * G::YP), reinterpret_cast<const tap_t*>(taps.data()), bp[m], bp[m + 1]);
            const int ta = t0 + 2 * p;
            if (ta < T) out[(size_t)ta * n_mels + m] = 10.0f * std::log10(std::fmax(lo(acc), 1e-10f));
            if (ta + 1 < T) out[(size_t)(ta + 1) * n_mels + m] = 10.0f * std::log10(std::fmax(hi(acc), 1e-10f));
        }
    }
    FILE* f = std::fopen(out_path, "wb");
    if (!f) return 2;
    std::fwrite(out.data(), sizeof(float), out.size(), f);
    std::fclose(f);
    return 0;
}

int main(int argc, char** argv) {
    if (argc != 7) return 1;
    const int n_fft = std::atoi(argv[1]), hop = std::atoi(argv[2]), n_mels = std::atoi(argv[3]), deriv = std::atoi(argv[4]);
    FILE* f = std::fopen(argv[5], "rb");
    if (!f) return 2;
    std::fseek(f, 0, SEEK_END);
    const long bytes = std::ftell(f);
    std::fseek(f, 0, SEEK_SET);
    std::vector<float> wav(bytes / 4);
    if (std::fread(wav.data(), 4, wav.size(), f) != wav.size()) return 2;
    std::fclose(f);
    switch (n_fft) {
        case 400: return run<8>(hop, n_mels, deriv, wav, argv[6]);
        case 800: return run<16>(hop, n_mels, deriv, wav, argv[6]);
        case 1600: return run<32>(hop, n_mels, deriv, wav, argv[6]);
    }
    return 3;
}
