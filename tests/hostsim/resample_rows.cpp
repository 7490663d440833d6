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
    FILE* f = std::fopen(argv[3], "wb");
    if (!f) return 2;
    const int32_t hdr[2] = {width, taps};
    std::fwrite(hdr, 4, 2, f);
    std::fwrite(k_lo.data(), 4, k_lo.size(), f);
    std::fwrite(rows.data(), 4, rows.size(), f);
    std::fclose(f);
    return 0;
}
