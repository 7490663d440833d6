// Checks the mel gather program (csrc/tables.h: make_mel_program) against the dense filterbank it encodes (TEST CODE).
//   mel_program <n_fft> <n_mels>   prints "ok <steps> <wavefronts per frame pair> <head entries>", exit 0 when the program
//   reproduces every non-zero weight of melscale_fbanks exactly once and nothing else.
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <map>

#include "tables.h"

using namespace sept;

int main(int argc, char** argv) {
    if (argc != 3) return 1;
    const int n_fft = std::atoi(argv[1]), n_mels = std::atoi(argv[2]), n_freqs = n_fft / 2 + 1;
    MelProgram prog;
    make_mel_program(n_fft, n_mels, 16000, prog);
    std::vector<float> fb = make_mel_fbank(n_freqs, n_mels, 16000, 0.0, 8000.0);
    std::map<int, int> bin_of;                                       // byte offset -> bin
    for (int k = 0; k < n_freqs; ++k) bin_of[8 * power_tile_pos(n_fft, k)] = k;
    std::vector<float> dense((size_t)n_freqs * n_mels, 0.f);
    std::vector<int> hits((size_t)n_freqs * n_mels, 0);
    auto put = [&](int off, int band, float w) {
        if (w == 0.f) return true;
        auto it = bin_of.find(off);
        if (it == bin_of.end() || band < 0 || band >= n_mels) return false;
        dense[(size_t)it->second * n_mels + band] += w * 4.0f;
        ++hits[(size_t)it->second * n_mels + band];
        return true;
    };
    for (int s = 0; s < prog.n_head; ++s)
        if (!put(prog.entries[s].off, 0, prog.entries[s].up) || prog.entries[s].dn != 0.f) return 2;
    size_t e = prog.n_head;
    if ((int)prog.round_steps.size() != (n_mels + 31) / 32) return 3;
    for (size_t r = 0; r < prog.round_steps.size(); ++r)
        for (int s = 0; s < prog.round_steps[r]; ++s)
            for (int l = 0; l < prog.width; ++l, ++e) {
                const MelStep& st = prog.entries[e];
                const int i = 1 + 32 * (int)r + l;                   // interval: rising side of band i, falling side of i - 1
                if (st.off >= 8 * power_tile_zero_slot(n_fft) && st.off < 8 * (power_tile_zero_slot(n_fft) + kZeroSlots)) {   // idle slot
                    if (st.dn != 0.f || st.up != 0.25f) return 4;
                    continue;
                }
                if (!bin_of.count(st.off)) return 4;
                if (std::fabs(0.25f - st.up - st.dn) > 3e-8f) return 9;   // the unrolled path derives dn from up
                if (i < n_mels && !put(st.off, i, st.up)) return 5;
                if (!put(st.off, i - 1, st.dn)) return 5;
                if (s == 0 && st.pad != prog.round_steps[r]) return 6;
            }
    if (e != prog.entries.size()) return 7;
    for (size_t i = 0; i < dense.size(); ++i)
        if (dense[i] != fb[i] || hits[i] > 1) return 8;
    std::printf("ok %d %d %d\n", prog.total_steps, mel_program_wavefronts(prog), prog.n_head);
    return 0;
}
