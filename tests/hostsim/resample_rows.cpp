// Dumps the polyphase row table the library builds (csrc/tables.h: make_resample_rows) so that the CPU tests can
// compare it with the oracle's restatement of torchaudio's kernel.   resample_rows <orig> <up> <out.bin>
#include <cstdio>
#include <cstdlib>
#include "tables.h"
int main(int argc, char** argv) {
    if (argc != 4) return 1;
    std::vector<int32_t> k_lo;
    std::vector<float> rows;
    int taps = 0;
    const int width = sept::make_resample_rows(std::atoi(argv[1]), std::atoi(argv[2]), k_lo, rows, taps);
    // the register-tile form of the same table (make_resample_tiles) must hold every row exactly once, at the window
    // position the tiled kernel reads it from, and zeros elsewhere
    {
        const int up = std::atoi(argv[2]);
        sept::ResampleTiles t;
        sept::make_resample_tiles(up, taps, k_lo, rows, t);
        if (t.n_groups != (up + 3) / 4 || t.tg % 2 != 1) return 3;
        for (int g = 0; g < t.n_groups; ++g)
            for (int jj = 0; jj < 4; ++jj)
                for (int i = 0; i < t.tg; ++i) {
                    const int j = 4 * g + jj;
                    const int k = j < up ? t.base[g] + i - k_lo[j] : -1;       // tap of phase j at window position i
                    const float want = (j < up && k >= 0 && k < taps) ? rows[(size_t)j * taps + k] : 0.f;
                    if (t.wt[((size_t)g * t.tg + i) * 4 + jj] != want) return 4;
                    if (t.base[g] < t.base_min || t.base[g] > t.base_max) return 5;
                }
    }
    FILE* f = std::fopen(argv[3], "wb");
    if (!f) return 2;
    const int32_t hdr[2] = {width, taps};
    std::fwrite(hdr, 4, 2, f);
    std::fwrite(k_lo.data(), 4, k_lo.size(), f);
    std::fwrite(rows.data(), 4, rows.size(), f);
    std::fclose(f);
    return 0;
}
