// Per-lane phase functions of the fused frame -> window -> real FFT -> |X|^2 -> mel kernel.
//
// Work unit: one warp owns one ITEM = FPW consecutive frames of one utterance (FPW = 2 * 32/R: 8 / 4 / 2 frames
// for n_fft 400 / 800 / 1600) and takes it through every phase with only __syncwarp() in between, so the warps
// of an SM drift apart and the FP-heavy and the shared-memory-heavy phases of different warps overlap.
// Frames are processed in adjacent PAIRS packed into pk2 (vec.cuh).
//
// The functions are __host__ __device__ on purpose: the CUDA kernel (extract.cu) calls them with its lane id,
// and tests/hostsim calls the very same functions from a loop over lane ids with plain arrays standing in for
// shared memory, so the index algebra (prime-factor maps, real split, tile positions, reflect staging) is
// verified on a CPU-only box.  The host emulation is test code, not a fallback: the library has no CPU path.
//
// Warp-private shared memory:
//   stage  float [SPAN = (FPW-1)*hop + n_fft]   reflect-padded waveform span of the item
//   Y      pk4   [PPW][25 rows k2][YS = R+1]     pass 1 output / pass 2 in place; (re, im) x (frame a, frame b)
//   P      pk2   aliases Y row by row            4|X[k]|^2 of the frame pair at tile position bin_pos(k)
// Replaces, for one item: torch.stft framing/window/rFFT + abs().pow(2) (torchaudio functional.py:123-144)
// and MelScale's matmul (transforms/_transforms.py:417).
#pragma once
#include "dft.cuh"

namespace sept {

struct alignas(8) f2 { float x, y; };
struct alignas(16) f4 { float x, y, z, w; };
struct alignas(8) tap_t { int pos; float w; };
struct alignas(16) pk4 { pk2 re, im; };

template <int R_>
struct Geo {
    static constexpr int R = R_;
    static constexpr int NC = R * 25, NFFT = 2 * NC, PAD = NFFT / 2;
    static constexpr int PPW = 32 / R;                           // packed frame pairs per warp
    static constexpr int FPW = 2 * PPW;                           // frames per item
    static constexpr int YS = R + 1;                              // pk4 per k2 row (odd: rows hit distinct bank groups)
    static constexpr int YP = 25 * YS + 2;                        // pk4 per pair
    static constexpr int Y_PK4 = PPW * YP;                        // pk4 per warp
    static constexpr int NYQ_POS = R;                             // pk2 slot of bin NC (spare tail of row 0)
    static constexpr int P2_TASKS = PPW * 25;                     // pass-2 row tasks per item
    static SEPT_HD int span(int hop) { return (FPW - 1) * hop + NFFT; }
    // pk2 slot (relative to the pair's Y base) of real-FFT bin k, 0 <= k <= NC
    static SEPT_HD int bin_pos(int k) { return k == NC ? NYQ_POS : (k % 25) * (2 * YS) + (k % R); }
};

// ---- staging: reflect-padded span of the item starting at frame t0 ----------------------------------------
// padded index q = t0*hop + i maps to source sample q - pad, reflected without edge repeat
// (torch/functional.py:675-680).  deriv != 0 stages np.gradient(x) (audio_feature_extraction.py:20) instead.
SEPT_HD int reflect_src(long long q, int pad, int n) {
    long long s = q - pad;
    if (s < 0) s = -s;
    else if (s >= n) s = 2LL * (n - 1) - s;
    return (s >= 0 && s < n) ? (int)s : -1;
}

SEPT_HD float staged_sample(const float* wav, int n, int src, int deriv) {
    if (src < 0) return 0.f;
    if (!deriv) return wav[src];
    if (src == 0) return wav[1] - wav[0];
    if (src == n - 1) return wav[n - 1] - wav[n - 2];
    return (wav[src + 1] - wav[src - 1]) * 0.5f;
}

template <class G>
SEPT_HD void stage_item(int lane, const float* wav, int n, int t0, int hop, int deriv, float* stage) {
    const long long q0 = (long long)t0 * hop;
    const int span = G::span(hop);
    for (int i = lane; i < span; i += 32) stage[i] = staged_sample(wav, n, reflect_src(q0 + i, G::PAD, n), deriv);
}

// ---- pass 1: lane (p, n1) windows the 25 complex samples z[n] = xw[2n] + i xw[2n+1], n = (25 n1 + R n2) mod NC,
// of frames 2p and 2p+1 and transforms them over n2 ------------------------------------------------------------
template <class G>
SEPT_HD void pass1(int lane, const float* stage, int hop, const f2* win2, pk4* Y) {
    constexpr int R = G::R;
    const int p = lane / R, n1 = lane % R;
    const f2* xa = reinterpret_cast<const f2*>(stage + (2 * p) * hop);
    const f2* xb = reinterpret_cast<const f2*>(stage + (2 * p + 1) * hop);
    pk2 re[25], im[25];
#pragma unroll
    for (int n2 = 0; n2 < 25; ++n2) {
        const int idx = Pfa<R>::in_index(n1, n2);
        const f2 a = xa[idx], b = xb[idx], w = win2[idx];
        re[n2] = pk(a.x * w.x, b.x * w.x);
        im[n2] = pk(a.y * w.y, b.y * w.y);
    }
    Dft<25>::run(re, im);
    pk4* y = Y + p * G::YP + n1;
#pragma unroll
    for (int k2 = 0; k2 < 25; ++k2) y[k2 * G::YS] = pk4{re[k2], im[k2]};
}

// ---- pass 2: row task (p, k2) transforms its R samples over n1, in place ----------------------------------------
template <class G>
SEPT_HD void pass2_row(int task, pk4* Y) {
    constexpr int R = G::R;
    const int p = task / 25, k2 = task % 25;
    pk4* y = Y + p * G::YP + k2 * G::YS;
    pk2 re[R], im[R];
#pragma unroll
    for (int i = 0; i < R; ++i) { const pk4 v = y[i]; re[i] = v.re; im[i] = v.im; }
    Dft<R>::run(re, im);
#pragma unroll
    for (int i = 0; i < R; ++i) y[i] = pk4{re[i], im[i]};
}

// ---- real split of one conjugate pair: Zk = Z[k], Zm = Z[NC-k], tw = W_{NFFT}^k.  Returns 4|X[k]|^2 and
// 4|X[NC-k]|^2 (the 1/4 is folded into the mel weights).  X[k] = E + W^k O with E = (Zk + conj Zm)/2,
// O = (Zk - conj Zm)/(2i); X[NC-k] = conj(E - W^k O). ------------------------------------------------------
SEPT_HD void split_pair(pk4 zk, pk4 zm, pk2 twr, pk2 twi, pk2& pk_, pk2& pm_) {
    const pk2 er = zk.re + zm.re, ei = zk.im - zm.im;
    const pk2 o_r = zk.im + zm.im, o_i = zm.re - zk.re;          // (Zk - conj Zm) / i
    const pk2 tr = fnma2(o_i, twi, o_r * twr), ti = fma2(o_i, twr, o_r * twi);
    const pk2 xr = er + tr, xi = ei + ti, yr = er - tr, yi = ei - ti;
    pk_ = fma2(xi, xi, xr * xr);
    pm_ = fma2(yi, yi, yr * yr);
}

// iteration k2 (0..12) of the split: lane (p, k1) pairs Z at (k1, k2) with Z at (R-k1, 25-k2).
// Returns false when the lane has nothing to do (row 0 is its own partner: only k1 <= R/2 work).
template <class G>
SEPT_HD bool split_load(int lane, int k2, const pk4* Y, const f4* tws, pk2& pk_, pk2& pm_) {
    constexpr int R = G::R;
    const int p = lane / R, k1 = lane % R, km = (R - k1) % R;
    if (k2 == 0 && k1 > R / 2) return false;
    const int rb = (25 - k2) % 25;
    const pk4 zk = Y[p * G::YP + k2 * G::YS + k1];
    const pk4 zm = Y[p * G::YP + rb * G::YS + km];
    const f4 tw = tws[k2 * R + k1];
    split_pair(zk, zm, pk(tw.x, tw.y), pk(tw.z, tw.w), pk_, pm_);
    return true;
}

template <class G>
SEPT_HD void split_store(int lane, int k2, pk2* P, pk2 pk_, pk2 pm_) {
    constexpr int R = G::R;
    const int p = lane / R, k1 = lane % R, km = (R - k1) % R;
    const int rb = (25 - k2) % 25;
    pk2* base = P + p * (2 * G::YP);
    base[k2 * (2 * G::YS) + k1] = pk_;
    if (k2 == 0 && k1 == 0) base[G::NYQ_POS] = pm_;              // bin NC
    else if (!(k2 == 0 && 2 * k1 == R)) base[rb * (2 * G::YS) + km] = pm_;
}

// ---- mel: one (frame pair, band) dot product over the band's taps (MelScale, _transforms.py:417) ---------------
SEPT_HD pk2 mel_band(const pk2* Ppair, const tap_t* taps, int lo_, int hi_) {
    pk2 acc = splat(0.f);
    for (int i = lo_; i < hi_; ++i) {
        const tap_t e = taps[i];
        acc = fma2(Ppair[e.pos], splat(e.w), acc);
    }
    return acc;
}

}  // namespace sept
