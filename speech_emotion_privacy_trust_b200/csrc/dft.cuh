// Register-resident small DFTs for the mixed-radix real FFT (sm_100a; also compiled for the host by
// tests/hostsim so the index algebra is checked on the CPU).
//
// Every reference FFT size is 2^a * 5^2 (n_fft 400/800/1600).  A real n_fft-point transform is
// computed as a complex Nc = n_fft/2 = R*25 point transform (R = 8/16/32) by the prime-factor algorithm
// (gcd(R,25)=1, so there are NO twiddles between the 25-point pass and the R-point pass), followed by the
// usual real-input split.  This header holds the in-register pieces: radix-2/4/5 butterflies and the
// composite 8/16/32/25-point transforms built from them by Cooley-Tukey with compile-time twiddles
// (emitted as FFMA immediates).  All loops are fully unrolled; arrays live in registers.
//
// All arithmetic is on pk2 (vec.cuh): each value carries the same sample of two adjacent frames, so every
// butterfly below is one packed FADD2/FMUL2/FFMA2 for two frames.
#pragma once
#include "vec.cuh"

namespace sept {

// ---- compile-time cos/sin of 2*pi*j/n (constexpr Taylor series after octant reduction) ----------------
namespace detail {
constexpr double kPi = 3.14159265358979323846264338327950288;
constexpr double taylor_sin(double x) {  // |x| <= pi/4
    double x2 = x * x, term = x, sum = x;
    for (int i = 1; i < 12; ++i) { term *= -x2 / ((2 * i) * (2 * i + 1)); sum += term; }
    return sum;
}
constexpr double taylor_cos(double x) {
    double x2 = x * x, term = 1.0, sum = 1.0;
    for (int i = 1; i < 12; ++i) { term *= -x2 / ((2 * i - 1) * (2 * i)); sum += term; }
    return sum;
}
// cos(2*pi*j/n), exact symmetries first so that 0, +-1, +-sqrt(1/2) come out exactly representable
constexpr double cos2pi(long long j, long long n) {
    j %= n; if (j < 0) j += n;
    if (2 * j > n) j = n - j;                   // cos is even about pi
    if (4 * j > n) return -cos2pi(n - 2 * j, 2 * n);  // cos(x) = -cos(pi - x), with pi - x = 2*pi*(n-2j)/(2n)
    if (j == 0) return 1.0;
    if (4 * j == n) return 0.0;
    if (8 * j > n) return taylor_sin(kPi / 2 - 2 * kPi * (double)j / (double)n);
    return taylor_cos(2 * kPi * (double)j / (double)n);
}
constexpr double sin2pi(long long j, long long n) { return cos2pi(4 * j - n, 4 * n); }  // sin x = cos(x - pi/2)
}  // namespace detail

// W_n^j = exp(-2*pi*i*j/n) as float constants
template <int J, int N> struct Tw {
    static constexpr float re = (float)detail::cos2pi(J, N);
    static constexpr float im = (float)(-detail::sin2pi(J, N));
};

// (ar + i ai) *= W_N^J with the trivial rotations folded away
template <int J, int N>
SEPT_HD void twiddle(pk2& ar, pk2& ai) {
    constexpr int j = ((J % N) + N) % N;
    if constexpr (j == 0) {
    } else if constexpr (4 * j == N) {          // -i
        pk2 t = ar; ar = ai; ai = neg(t);
    } else if constexpr (2 * j == N) {          // -1
        ar = neg(ar); ai = neg(ai);
    } else if constexpr (4 * j == 3 * N) {      // +i
        pk2 t = ar; ar = neg(ai); ai = t;
    } else if constexpr (8 * j == N) {          // (1 - i) / sqrt 2
        constexpr float c = Tw<j, N>::re;
        pk2 tr = (ar + ai) * splat(c), ti = (ai - ar) * splat(c);
        ar = tr; ai = ti;
    } else {
        constexpr float cr = Tw<j, N>::re, ci = Tw<j, N>::im;
        pk2 tr = fma2(ai, splat(-ci), ar * splat(cr));
        pk2 ti = fma2(ai, splat(cr), ar * splat(ci));
        ar = tr; ai = ti;
    }
}

// ---- radix butterflies (forward transform, in place, natural order) ------------------------------------
SEPT_HD void dft2(pk2& r0, pk2& i0, pk2& r1, pk2& i1) {
    pk2 tr = r0 - r1, ti = i0 - i1;
    r0 = r0 + r1; i0 = i0 + i1; r1 = tr; i1 = ti;
}

SEPT_HD void dft4(pk2& r0, pk2& i0, pk2& r1, pk2& i1, pk2& r2, pk2& i2, pk2& r3, pk2& i3) {
    pk2 ar = r0 + r2, ai = i0 + i2, br = r0 - r2, bi = i0 - i2;
    pk2 cr = r1 + r3, ci = i1 + i3, dr = r1 - r3, di = i1 - i3;
    r0 = ar + cr; i0 = ai + ci;
    r2 = ar - cr; i2 = ai - ci;
    r1 = br + di; i1 = bi - dr;   // b - i*d
    r3 = br - di; i3 = bi + dr;   // b + i*d
}

SEPT_HD void dft5(pk2& r0, pk2& i0, pk2& r1, pk2& i1, pk2& r2, pk2& i2, pk2& r3, pk2& i3,
                  pk2& r4, pk2& i4) {
    constexpr float c1 = (float)detail::cos2pi(1, 5), c2 = (float)detail::cos2pi(2, 5);
    constexpr float s1 = (float)detail::sin2pi(1, 5), s2 = (float)detail::sin2pi(2, 5);
    pk2 t1r = r1 + r4, t1i = i1 + i4, t3r = r1 - r4, t3i = i1 - i4;
    pk2 t2r = r2 + r3, t2i = i2 + i3, t4r = r2 - r3, t4i = i2 - i3;
    pk2 a1r = fma2(t2r, splat(c2), fma2(t1r, splat(c1), r0)), a1i = fma2(t2i, splat(c2), fma2(t1i, splat(c1), i0));
    pk2 a2r = fma2(t2r, splat(c1), fma2(t1r, splat(c2), r0)), a2i = fma2(t2i, splat(c1), fma2(t1i, splat(c2), i0));
    pk2 b1r = fma2(t4r, splat(s2), t3r * splat(s1)), b1i = fma2(t4i, splat(s2), t3i * splat(s1));
    pk2 b2r = fma2(t4r, splat(-s1), t3r * splat(s2)), b2i = fma2(t4i, splat(-s1), t3i * splat(s2));
    r0 = r0 + t1r + t2r; i0 = i0 + t1i + t2i;
    r1 = a1r + b1i; i1 = a1i - b1r;   // a1 - i*b1
    r4 = a1r - b1i; i4 = a1i + b1r;   // a1 + i*b1
    r2 = a2r + b2i; i2 = a2i - b2r;
    r3 = a2r - b2i; i3 = a2i + b2r;
}

// ---- composite transforms: N = N1*N2 Cooley-Tukey, n = N2*n1 + n2, k = k1 + N1*k2 --------------------
template <int N> struct Dft;

template <> struct Dft<2> {
    static SEPT_HD void run(pk2 (&re)[2], pk2 (&im)[2]) { dft2(re[0], im[0], re[1], im[1]); }
};
template <> struct Dft<4> {
    static SEPT_HD void run(pk2 (&re)[4], pk2 (&im)[4]) {
        dft4(re[0], im[0], re[1], im[1], re[2], im[2], re[3], im[3]);
    }
};
template <> struct Dft<5> {
    static SEPT_HD void run(pk2 (&re)[5], pk2 (&im)[5]) {
        dft5(re[0], im[0], re[1], im[1], re[2], im[2], re[3], im[3], re[4], im[4]);
    }
};

template <int N1, int N2, int K1, int N2I> struct TwiddleRow {  // A[k1][n2] *= W_N^{n2*k1}, n2 = 0..N2-1
    template <int N>
    static SEPT_HD void apply(pk2 (&re)[N], pk2 (&im)[N]) {
        if constexpr (N2I < N2) {
            twiddle<K1 * N2I, N1 * N2>(re[K1 * N2 + N2I], im[K1 * N2 + N2I]);
            TwiddleRow<N1, N2, K1, N2I + 1>::apply(re, im);
        }
    }
};
template <int N1, int N2, int K1> struct TwiddleAll {
    template <int N>
    static SEPT_HD void apply(pk2 (&re)[N], pk2 (&im)[N]) {
        if constexpr (K1 < N1) {
            TwiddleRow<N1, N2, K1, 1>::apply(re, im);
            TwiddleAll<N1, N2, K1 + 1>::apply(re, im);
        }
    }
};

template <int N1, int N2>
SEPT_HD void dft_ct(pk2 (&re)[N1 * N2], pk2 (&im)[N1 * N2]) {
    constexpr int N = N1 * N2;
    pk2 ar[N], ai[N];  // A[k1][n2] at k1*N2 + n2
#pragma unroll
    for (int n2 = 0; n2 < N2; ++n2) {
        pk2 yr[N1], yi[N1];
#pragma unroll
        for (int n1 = 0; n1 < N1; ++n1) { yr[n1] = re[N2 * n1 + n2]; yi[n1] = im[N2 * n1 + n2]; }
        Dft<N1>::run(yr, yi);
#pragma unroll
        for (int k1 = 0; k1 < N1; ++k1) { ar[k1 * N2 + n2] = yr[k1]; ai[k1 * N2 + n2] = yi[k1]; }
    }
    TwiddleAll<N1, N2, 1>::apply(ar, ai);
#pragma unroll
    for (int k1 = 0; k1 < N1; ++k1) {
        pk2 zr[N2], zi[N2];
#pragma unroll
        for (int n2 = 0; n2 < N2; ++n2) { zr[n2] = ar[k1 * N2 + n2]; zi[n2] = ai[k1 * N2 + n2]; }
        Dft<N2>::run(zr, zi);
#pragma unroll
        for (int k2 = 0; k2 < N2; ++k2) { re[k1 + N1 * k2] = zr[k2]; im[k1 + N1 * k2] = zi[k2]; }
    }
}

template <> struct Dft<8>  { static SEPT_HD void run(pk2 (&re)[8],  pk2 (&im)[8])  { dft_ct<2, 4>(re, im); } };
template <> struct Dft<16> { static SEPT_HD void run(pk2 (&re)[16], pk2 (&im)[16]) { dft_ct<4, 4>(re, im); } };
template <> struct Dft<32> { static SEPT_HD void run(pk2 (&re)[32], pk2 (&im)[32]) { dft_ct<4, 8>(re, im); } };
template <> struct Dft<25> { static SEPT_HD void run(pk2 (&re)[25], pk2 (&im)[25]) { dft_ct<5, 5>(re, im); } };

// ---- prime-factor index maps for Nc = R * 25 ------------------------------------------------------------
//   input  n = (25*n1 + R*n2) mod Nc          (n1 < R, n2 < 25)
//   output k with k mod R = k1, k mod 25 = k2 (CRT);  k = (25*a*k1 + R*b*k2) mod Nc,
//          a = 25^-1 mod R, b = R^-1 mod 25
template <int R> struct Pfa {
    static constexpr int Nc = R * 25;
    static constexpr int inv(int x, int m) { for (int i = 1; i < m; ++i) if ((x * i) % m == 1) return i; return 0; }
    static constexpr int a = inv(25 % R, R);
    static constexpr int b = inv(R % 25, 25);
    static constexpr int cK1 = (25 * a) % Nc;
    static constexpr int cK2 = (R * b) % Nc;
    static SEPT_HD int in_index(int n1, int n2) { int v = 25 * n1 + R * n2; return v >= Nc ? v - Nc : v; }
    static SEPT_HD int out_index(int k1, int k2) { return (cK1 * k1 + cK2 * k2) % Nc; }
};

}  // namespace sept
