// MFCC phase 2 on the FP32 pipes: dB with the per-utterance top_db floor, then the 128 x 40 ortho DCT-II
// (torchaudio transforms/_transforms.py:714-717: amplitude_to_DB(mel, 10, amin 1e-10, top_db 80), then matmul with the basis).
//
// One CTA handles kFrames = 64 consecutive global frames (they may straddle utterances) as 32 frame PAIRS (f, f + 32).
//
//   load phase   a thread fetches the same 16-byte band quad of both frames of a pair (two 64-byte row segments per four
//                lanes; every load that depends on no other is issued up front, later rounds are in flight while the round
//                before is converted), converts 8 values to clamped dB and writes four 8-byte (frame f, frame f + 32) slots: the dB tile is
//                pair-interleaved, band m of pair p at slot m ^ (p & 15) of row p (conflict free for the 64-bit stores of the
//                load phase and for the 64-bit loads of the contraction)
//   contraction  warp (s, q): stream s (0 waveform, 1 np.gradient), coefficients [20 q, 20 q + 20).  A lane owns one pair,
//                packed in pk2.  The basis is symmetric, D[127 - m][c] = (-1)^c D[m][c], so bands m and 127 - m are folded
//                first (even coefficients take the sum, odd ones the difference): per band pair 2 packed adds, 20 FFMA2
//                with the weight as a broadcast scalar operand, 2 tile loads and 5 broadcast 16-byte weight loads.  The
//                16 slot addresses a lane ever needs are loop invariant (the swizzle only touches the low four band bits).
//   third stream np.gradient(x, 2) == np.gradient(x) / 2 exactly, a quarter of stream 1's power, so its dB is stream 1's minus
//                10 log10 4 and so is its floor: dB2 = max(clamped dB1 - 10 log10 4, -100) = clamped dB1 - 10 log10 4 + r
//                with r = 0 unless clamped dB1 < -93.98.  The DCT of a constant is a pure c0 term (sum_m D[m][0] = sqrt 128),
//                so stream 2 is stream 1's result with c0 shifted.  Only a tile that holds a frame whose floor lies below
//                -93.98 dB (a nearly silent utterance) runs the contraction again on the re-clamped values.
//
// Per frame: 1 KB read, 480 B written, 2 x 128 x 40 / 2 = 5 120 packed-pair FMAs' worth of work (2 560 FFMA2 lanes).
#include "extract.h"
#include "vec.cuh"

namespace sept {

namespace {

constexpr int kFrames = 64, kPairs = 32, kThreads = 128;
constexpr int kNM = 128, kNC = 40, kHalf = kNM / 2, kCoefPerWarp = 20;
constexpr float kDbQuarter = 6.02059991327962390f;     // 10 log10(4)
constexpr float kSqrtBands = 11.3137084989847604f;     // sum_m D[m][0] = 128 / sqrt(128)
constexpr float kDbPerLog2 = 3.01029995663981195f;     // 10 log10(2)
constexpr float kDbMin = -100.0f;                      // 10 log10(amin)

struct Smem {
    float D[kHalf * kNC];              // basis rows 0..63
    float2 X[2][kPairs][kNM];          // clamped dB: .x = frame p, .y = frame p + 32; band m at slot m ^ (p & 15)
    int frame_utt[kFrames];
    float frame_floor[2][kFrames];     // max(dB(max of the utterance) - top_db, -100) per stream
};
static_assert(sizeof(Smem) <= 76800, "three CTAs must fit one SM (3 x (smem + 1 KB) <= 228 KB)");

__device__ __forceinline__ float db_clamped(float p, float floor_db) {      // floor_db >= -100 carries the amin clamp
    float l;
    asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(l) : "f"(p));
    return fmaxf(kDbPerLog2 * l, floor_db);
}

// acc[c] += sum over the 64 folded band pairs of (e or o) * D[m][20 q + c] for the lane's frame pair.
// THIRD: the tile values are re-clamped to stream 2 (max(x - 10 log10 4, -100)) on the way in.
template <bool THIRD>
__device__ __forceinline__ void contract(const float2* __restrict__ row, unsigned px, const float* __restrict__ dq, pk2 (&acc)[kCoefPerWarp]) {
    const float2* a[16];
#pragma unroll
    for (int j = 0; j < 16; ++j) a[j] = row + (j ^ px);
    auto fetch = [](const float2* p) {
        const float2 v = *p;
        if constexpr (THIRD) return pk(fmaxf(v.x - kDbQuarter, kDbMin), fmaxf(v.y - kDbQuarter, kDbMin));
        else return pk(v.x, v.y);
    };
#pragma unroll 1
    for (int mb = 0; mb < kHalf / 16; ++mb) {
        const int fwd = 16 * mb, mir = kNM - 16 - 16 * mb;                   // band 16 mb + j, mirror 127 - (16 mb + j) = mir + (15 - j)
#pragma unroll
        for (int j = 0; j < 16; ++j) {
            const pk2 x = fetch(a[j] + fwd), y = fetch(a[15 - j] + mir);    // (15 - j) ^ px == (j ^ px) ^ 15
            const pk2 e = x + y, o = x - y;
            const float4* wrow = reinterpret_cast<const float4*>(dq + (fwd + j) * kNC);
            float w[kCoefPerWarp];
#pragma unroll
            for (int i = 0; i < kCoefPerWarp / 4; ++i) {
                const float4 t = wrow[i];
                w[4 * i] = t.x; w[4 * i + 1] = t.y; w[4 * i + 2] = t.z; w[4 * i + 3] = t.w;
            }
#pragma unroll
            for (int c = 0; c < kCoefPerWarp; ++c) acc[c] = fma2((c & 1) ? o : e, splat(w[c]), acc[c]);   // 20 q is even
        }
    }
}

__global__ void __launch_bounds__(kThreads, 3) mfcc_dct_kernel(const MfccDctParams prm) {
    extern __shared__ __align__(16) unsigned char dct_smem[];
    Smem& sm = *reinterpret_cast<Smem*>(dct_smem);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const long long g0 = (long long)blockIdx.x * kFrames;

    // Everything that does not depend on another load is issued first: the frame -> utterance map, the basis, the first
    // round of mel power.  The per-utterance maxima (second link of the only dependent chain) follow, and each later
    // round of power loads is in flight while the round before it is converted.
    constexpr int NIT = 4, ROUNDS = 2 * kPairs * (kNM / 4) / kThreads / NIT;       // 64 warp slots: 4 rounds x 4 per warp
    const int p_lo = lane >> 2, q_lo = lane & 3;
    float4 pa[2][NIT], pb[2][NIT];
    auto load_round = [&](int round, float4 (&a)[NIT], float4 (&b)[NIT]) {
#pragma unroll
        for (int it = 0; it < NIT; ++it) {
            const int slot = (round * NIT + it) * 4 + warp, s = slot >> 5, p = ((slot >> 3) & 3) * 8 + p_lo, m4 = (slot & 7) * 4 + q_lo;
            const long long ga = g0 + p, gb = ga + kPairs;
            const float4* src = reinterpret_cast<const float4*>(prm.power + ((long long)s * prm.total_frames + ga) * kNM) + m4;
            a[it] = ga < prm.total_frames ? __ldg(src) : make_float4(0.f, 0.f, 0.f, 0.f);
            b[it] = gb < prm.total_frames ? __ldg(src + kPairs * (kNM / 4)) : make_float4(0.f, 0.f, 0.f, 0.f);
        }
    };
    auto convert_round = [&](int round, const float4 (&a)[NIT], const float4 (&b)[NIT]) {
#pragma unroll
        for (int it = 0; it < NIT; ++it) {
            const int slot = (round * NIT + it) * 4 + warp, s = slot >> 5, p = ((slot >> 3) & 3) * 8 + p_lo, m4 = (slot & 7) * 4 + q_lo;
            const float fa = sm.frame_floor[s][p], fb = sm.frame_floor[s][p + kPairs];
            float2* row = sm.X[s][p];
            const int px = p & 15, m = 4 * m4;
            row[m ^ px] = make_float2(db_clamped(a[it].x, fa), db_clamped(b[it].x, fb));
            row[(m + 1) ^ px] = make_float2(db_clamped(a[it].y, fa), db_clamped(b[it].y, fb));
            row[(m + 2) ^ px] = make_float2(db_clamped(a[it].z, fa), db_clamped(b[it].z, fb));
            row[(m + 3) ^ px] = make_float2(db_clamped(a[it].w, fa), db_clamped(b[it].w, fb));
        }
    };
    const bool live = tid < kFrames && g0 + tid < prm.total_frames;
    const int u_mine = live ? __ldg(prm.frame_utt + g0 + tid) : 0;
    constexpr int DQ = kHalf * kNC / 4 / kThreads;                                 // 5 basis quads per thread
    float4 dreg[DQ];
#pragma unroll
    for (int i = 0; i < DQ; ++i) dreg[i] = __ldg(reinterpret_cast<const float4*>(prm.dct) + tid + kThreads * i);
    load_round(0, pa[0], pb[0]);
    bool quiet = false;
    if (tid < kFrames) {
        const float max0 = __int_as_float(__ldg(prm.utt_max + u_mine)), max1 = __int_as_float(__ldg(prm.utt_max + prm.n_utts + u_mine));
        const float fl0 = fmaxf(db_clamped(max0, kDbMin) - prm.top_db, kDbMin);
        const float fl1 = fmaxf(db_clamped(max1, kDbMin) - prm.top_db, kDbMin);
        sm.frame_utt[tid] = u_mine;
        sm.frame_floor[0][tid] = fl0;
        sm.frame_floor[1][tid] = fl1;
        quiet = live && fl1 < kDbMin + kDbQuarter;
    }
#pragma unroll
    for (int i = 0; i < DQ; ++i) reinterpret_cast<float4*>(sm.D)[tid + kThreads * i] = dreg[i];
    const bool need_third = __syncthreads_or(quiet) != 0;                           // also: floors and basis are in place
#pragma unroll
    for (int round = 0; round < ROUNDS; ++round) {
        if (round + 1 < ROUNDS) load_round(round + 1, pa[(round + 1) & 1], pb[(round + 1) & 1]);
        convert_round(round, pa[round & 1], pb[round & 1]);
    }
    __syncthreads();

    // ---- contraction ----------------------------------------------------------------------------------------------------
    const int s = warp & 1, q = warp >> 1, p = lane;
    const float2* row = sm.X[s][p];
    const float* dq = sm.D + kCoefPerWarp * q;
    pk2 acc[kCoefPerWarp];
#pragma unroll
    for (int c = 0; c < kCoefPerWarp; ++c) acc[c] = splat(0.f);
    contract<false>(row, (unsigned)(p & 15), dq, acc);

    // ---- store: lanes = consecutive frames, coefficient rows T apart --------------------------------------------------------
    const bool third = s == 1 && need_third;
    float* dst[2];
    long long stride[2];
#pragma unroll
    for (int half = 0; half < 2; ++half) {
        const int fr = p + kPairs * half;
        const long long g = g0 + fr;
        dst[half] = nullptr;
        stride[half] = 0;
        if (g >= prm.total_frames) continue;
        const int u = sm.frame_utt[fr];
        const long long f0 = prm.frame_off[u];
        const long long T = prm.frame_off[u + 1] - f0;
        stride[half] = T;
        dst[half] = prm.out + f0 * (3 * kNC) + (g - f0) + (long long)(s * kNC + kCoefPerWarp * q) * T;
    }
#pragma unroll
    for (int half = 0; half < 2; ++half) {
        if (!dst[half]) continue;
        float* o1 = dst[half];
        const long long T = stride[half];
#pragma unroll
        for (int c = 0; c < kCoefPerWarp; ++c) {
            const float v = half ? hi(acc[c]) : lo(acc[c]);
            o1[c * T] = v;
            if (s == 1 && !third) o1[(kNC + c) * T] = (c == 0 && q == 0) ? v - kDbQuarter * kSqrtBands : v;
        }
    }
    if (third) {                                                             // rare: a nearly silent utterance in the tile
#pragma unroll
        for (int c = 0; c < kCoefPerWarp; ++c) acc[c] = splat(0.f);
        contract<true>(row, (unsigned)(p & 15), dq, acc);
#pragma unroll
        for (int half = 0; half < 2; ++half) {
            if (!dst[half]) continue;
            float* o2 = dst[half] + kNC * stride[half];
#pragma unroll
            for (int c = 0; c < kCoefPerWarp; ++c) o2[c * stride[half]] = half ? hi(acc[c]) : lo(acc[c]);
        }
    }
}

}  // namespace

cudaError_t launch_mfcc_dct(const MfccDctParams& prm, cudaStream_t stream) {
    const long long blocks = (prm.total_frames + kFrames - 1) / kFrames;
    if (blocks == 0) return cudaSuccess;
    cudaError_t e = cudaFuncSetAttribute(mfcc_dct_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(Smem));
    if (e != cudaSuccess) return e;
    mfcc_dct_kernel<<<(unsigned)blocks, kThreads, sizeof(Smem), stream>>>(prm);
    return cudaGetLastError();
}

}  // namespace sept
