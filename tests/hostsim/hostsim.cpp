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
    MelProgram prog;
    make_mel_program(n_fft, n_mels, 16000, prog);
    for (int k = 0; k <= G::NC; ++k)
        if (power_tile_pos(n_fft, k) != G::bin_pos(k)) return 4;
    if (power_tile_zero_slot(n_fft) != G::ZSLOT || kZeroSlots != G::ZSLOTS) return 4;
    static_assert(sizeof(MelStep) == sizeof(mel_step), "table entry layout");
    const mel_step* mprog = reinterpret_cast<const mel_step*>(prog.entries.data());
    std::vector<float> stage(G::stage_floats(hop), 0.f);
    std::vector<pk4> Ybuf(G::Y_PK4 + 1);
    pk2* Yp = reinterpret_cast<pk2*>(Ybuf.data());
    pk2* P = Yp;
    std::vector<float> out((size_t)T * n_mels);
    const f2* win2 = reinterpret_cast<const f2*>(win.data());
    for (int t0 = 0; t0 < T; t0 += G::FPW) {
        // interior items: the kernel copies the RAW span asynchronously and differentiates on the fly in pass 1;
        // edge items: generic staging of the wanted stream
        const bool interior = item_is_interior<G>(n, t0, hop);
        for (int lane = 0; lane < 32; ++lane)
            stage_item<G>(lane, wav.data(), n, t0, hop, interior ? 0 : deriv, stage.data());
        for (int lane = 0; lane < 32; ++lane) {
            pass1_shared<G>(lane, stage.data(), hop, WinShared{win2}, Yp, interior && deriv);
        }
        const f2* tw2 = reinterpret_cast<const f2*>(tws.data());
        if constexpr (R <= 16) {
            static pk2 pu[G::PS_ROUNDS][32][R], pv[G::PS_ROUNDS][32][R];
            int p, j;
            for (int r = 0; r < G::PS_ROUNDS; ++r)
                for (int lane = 0; lane < 32; ++lane)
                    if (G::ps_task(lane, r, p, j)) {
                        TwShared tw{tw2 + j * G::TWS};
                        pass2_split<G>(p, j, Yp, tw, pu[r][lane], pv[r][lane]);
                    }
            for (int r = 0; r < G::PS_ROUNDS; ++r)
                for (int lane = 0; lane < 32; ++lane)
                    if (G::ps_task(lane, r, p, j)) pass2_split_store<G>(p, j, P, pu[r][lane], pv[r][lane]);
        } else {
            for (int task = 0; task < G::P2_TASKS; ++task) pass2_row<G>(task, Yp);
            pk2 a[32][13], b[32][13];
            bool on0[32];
            for (int lane = 0; lane < 32; ++lane)
                for (int k2 = 0; k2 <= 12; ++k2) {
                    TwStrided tw{tw2 + lane % R, G::TWS};
                    const bool on = split_load<G>(lane, k2, Yp, tw, a[lane][k2], b[lane][k2]);
                    if (k2 == 0) on0[lane] = on;
                }
            for (int lane = 0; lane < 32; ++lane) split_store_all<G>(lane, P, a[lane], b[lane], on0[lane]);
        }
        // mel: interval 0 (head), then rounds of 32 intervals; band b = U[interval b] + D[interval b + 1]
        {
            std::vector<pk2> Uall((size_t)(n_mels + 1) * G::PPW), Dall((size_t)(n_mels + 1) * G::PPW);
            pk2 U[G::PPW], D[G::PPW];
            mel_head<G>(P, mprog, prog.n_head, U);
            for (int p = 0; p < G::PPW; ++p) Uall[p] = U[p];
            const mel_step* e = mprog + prog.n_head;
            for (size_t r = 0; r < prog.round_steps.size(); ++r) {
                if (e->pad != prog.round_steps[r]) return 5;
                for (int lane = 0; lane < prog.width; ++lane) {
                    const int i = 1 + 32 * (int)r + lane;
                    if (i > n_mels) break;
                    mel_round<G>(P, e + lane, prog.round_steps[r], prog.width, U, D);
                    for (int p = 0; p < G::PPW; ++p) { Uall[(size_t)i * G::PPW + p] = U[p]; Dall[(size_t)i * G::PPW + p] = D[p]; }
                }
                e += (size_t)prog.round_steps[r] * prog.width;
            }
            for (int m = 0; m < n_mels; ++m)
                for (int p = 0; p < G::PPW; ++p) {
                    const pk2 acc = Uall[(size_t)m * G::PPW + p] + Dall[(size_t)(m + 1) * G::PPW + p];
                    const int ta = t0 + 2 * p;
                    if (ta < T) out[(size_t)ta * n_mels + m] = 10.0f * std::log10(std::fmax(lo(acc), 1e-10f));
                    if (ta + 1 < T) out[(size_t)(ta + 1) * n_mels + m] = 10.0f * std::log10(std::fmax(hi(acc), 1e-10f));
                }
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
