// MFCC phase 2 on the 5th-generation tensor cores (sm_100a): dB + per-utterance top_db floor, then the 128 x 40 orthonormal
// DCT-II (torchaudio/transforms/_transforms.py:714-717) as tcgen05.mma with the accumulators in tensor memory.
//
// Why tensor cores here and nowhere else on this path: the DCT is the one dense contraction of the extraction chain
// ([frames x 128 bands] x [128 x 40], 49 % of an MFCC frame's flops) and ncu showed round 1's FMA version with its FMA
// pipe 66 % busy at 184 us per audio-hour against 66 us of HBM time -- north_star's condition for trying it ("only if ncu
// shows that small GEMM actually limits throughput").
// Outcome (round 2, measured): parity green, 197 us per audio-hour -- slower than the FMA kernel (mfcc_dct.cu: 125 us after
// its rewrite).  Stripped variants (-DSEPT_TC_NO_MMA / NO_STORE / NO_CVT / NO_LOAD) put the cost where it is: the MMAs
// 11 us, the output stores 24 us, the dB conversion 5 us, the global loads of the mel power 96 us, barriers + metadata +
// operand stores 63 us.  The phase is bound by feeding the contraction, not by the contraction; the kernel is kept as an
// opt-in (SEPT_MFCC_DCT=tc) and as the starting point for a version whose loads are asynchronous (a ring of raw tiles filled
// by TMA / cp.async, converter warps, one MMA thread, epilogue warps) instead of one CTA-wide sequence of barriers.
//
// Precision: kind::tf32 multiplies 10-bit mantissas, far too coarse for the 1e-4 MFCC tolerance (dB values of +-100
// against a basis of 0.09).  Both operands are split into a tf32-exact high part and the remainder,
//     A = A_hi + A_lo,  B = B_hi + B_lo,      D = A_hi B_hi + A_lo B_hi + A_hi B_lo      (fp32 accumulate in TMEM)
// which leaves only A_lo B_lo ~ 2^-22 of a term: the result is as close to the reference as the FMA kernel is.
//
// One persistent CTA per SM (512 threads) walks tiles of 128 frames:
//   load    two mel-power streams (waveform, its gradient), 64-byte segments per frame row, 8 + 8 float4 in flight per thread
//   convert 10 log10(max(p, 1e-10)) clamped at the per-utterance floor; split hi / lo; 16-byte stores straight into the
//           canonical no-swizzle K-major operand layout (8 x 16-byte core matrices, K-adjacent cores contiguous)
//   mma     one elected thread: per 64-band half 8 K-steps x 3 products of M128 N48 K8 from one of TWO operand buffers,
//           committed to that buffer's mbarrier -- the next half is converted while the tensor core works on this one
//   stream 3 (np.gradient(x, 2): a quarter of stream 2's power) is max(dB2 - 10 log10 4, -100) = dB2 - 10 log10 4 + r with
//           r = max(0, -(100 - 10 log10 4) - dB2) -- and the DCT of a constant is a pure c0 term, so
//           MFCC3 = MFCC2 - 10 log10 4 * sqrt(128) * [c = 0] + DCT(r).  r is zero unless the gradient stream's floor lies
//           below -93.98 dB; only tiles that hold such a frame run a third (small) product.
//   epilogue tcgen05.ld (thread = frame, 10 coefficients per warp) and the (120, T) band-major stores.
// The next tile's metadata and first stream are fetched while the current tile is converted and stored.
#include <cuda_runtime.h>
#include <cstdint>

#include "extract.h"
#include "tmem.cuh"

namespace sept {

namespace {

constexpr int kM = 128;            // frames per tile = MMA M = TMEM lanes
constexpr int kK = 128;            // mel bands
constexpr int kN = 48;             // 40 coefficients padded to a legal MMA N (multiple of 16 at M = 128)
constexpr int kNC = 40;
constexpr int kThreads = 512;
constexpr int kWarps = kThreads / 32;
constexpr int kIt = 64 / kWarps;                          // 16-byte loads per thread and half tile
constexpr int kCoefPerWarp = kNC / (kWarps / 4);          // every TMEM lane quarter is served by kWarps / 4 warps
constexpr int kKH = 64;            // bands per operand buffer: a stream's tile goes through the tensor core in two halves
constexpr int kTileFloats = kM * kKH;                      // one operand plane (hi or lo) of one A buffer
constexpr int kBFloats = kN * kK;
constexpr uint32_t kTmemCols = 256;                        // D0 @ 0, D1 @ 64, D(r) @ 128
constexpr float kDbPerLog2 = 3.01029995663981195f;
constexpr float kAmin = 1e-10f;
constexpr float kDbQuarter = 6.02059991327962390f;         // 10 log10(4)
constexpr float kSqrtBands = 11.3137084989847604f;         // sum_m D[m][0] = 128 / sqrt(128)
constexpr size_t kSmemBytes = (size_t)(4 * kTileFloats + 2 * kBFloats) * 4 + 4 * kM * 4 + 64;   // 2 A buffers (hi, lo), B (hi, lo), 2 x 2 floor rows, barriers

__device__ __forceinline__ float power_to_db(float p) {
    float l;
    asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(l) : "f"(fmaxf(p, kAmin)));
    return kDbPerLog2 * l;
}

// float index of element (row, k) in the canonical K-major, no-swizzle operand layout: core matrix = 8 rows x 16 bytes,
// the K/4 cores of an 8-row group are contiguous (leading byte offset 128), groups follow each other (stride K/4 * 128 B)
__device__ __forceinline__ int canon(int row, int k, int k_extent) {
    return ((row >> 3) * (k_extent / 4) + (k >> 2)) * 32 + (row & 7) * 4 + (k & 3);
}

__device__ __forceinline__ uint64_t smem_desc(const void* p, int k_extent) {
    const uint32_t a = (uint32_t)__cvta_generic_to_shared(p);
    return (uint64_t)((a >> 4) & 0x3FFFu)                  // start address / 16
           | ((uint64_t)(128 >> 4) << 16)                  // leading byte offset: next core matrix along K
           | ((uint64_t)((k_extent / 4) * 128 >> 4) << 32) // stride byte offset: next 8-row group
           | (1ull << 46);                                 // descriptor version (sm_100); layout type 0 = no swizzle
}

// D[tmem] (+)= A[smem] * B[smem]^T, M128 N48 K8, tf32 inputs, fp32 accumulate
__device__ __forceinline__ void mma_tf32(uint32_t d_tmem, uint64_t da, uint64_t db, uint32_t accumulate) {
    constexpr uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(kN >> 3) << 17) | ((uint32_t)(kM >> 4) << 24);
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, {%5, %5, %5, %5}, p;\n\t"
        "}\n" ::"r"(d_tmem), "l"(da), "l"(db), "r"(idesc), "r"(accumulate), "r"(0u)
        : "memory");
}

__device__ __forceinline__ void mma_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(
                     (uint32_t)__cvta_generic_to_shared(bar))
                 : "memory");
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"((uint32_t)__cvta_generic_to_shared(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    const uint32_t a = (uint32_t)__cvta_generic_to_shared(bar);
    uint32_t done = 0, spins = 0;
    while (!done) {
        if (++spins > (1u << 24)) __trap();                       // a lost arrival becomes an error, never a hung GPU
        asm volatile(
            "{\n\t"
            ".reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t"
            "}\n"
            : "=r"(done)
            : "r"(a), "r"(parity)
            : "memory");
    }
}
__device__ __forceinline__ void fence_async_proxy() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

__device__ __forceinline__ float tf32_hi(float x) { return __uint_as_float(__float_as_uint(x) & 0xFFFFE000u); }

}  // namespace

__global__ void __launch_bounds__(kThreads, 1) mfcc_dct_tc_kernel(const MfccDctParams prm, int n_tiles) {
    extern __shared__ __align__(128) unsigned char smem[];
    float* a_base = reinterpret_cast<float*>(smem);               // [2 buffers][hi, lo][kM x kKH]
    float* b_hi = a_base + 4 * kTileFloats;
    float* b_lo = b_hi + kBFloats;
    float* floors = b_lo + kBFloats;                              // [2 tile buffers][2 streams][kM] top_db floors
    uint64_t* bars = reinterpret_cast<uint64_t*>(floors + 4 * kM);   // [2]: one per A buffer
    __shared__ uint32_t tmem_base;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;

    // ---- once per CTA: tensor memory, barriers, the basis as B[n = coefficient][k = band] split into hi / lo ---------
    if (warp == 0) tmem::alloc(&tmem_base, kTmemCols);
    if (tid == 0) { mbar_init(bars + 0, 1); mbar_init(bars + 1, 1); }
    for (int i = tid; i < kBFloats; i += kThreads) {
        const int n = i / kK, k = i % kK;
        const float v = n < kNC ? __ldg(prm.dct + k * kNC + n) : 0.f;
        const float h = tf32_hi(v);
        b_hi[canon(n, k, kK)] = h;
        b_lo[canon(n, k, kK)] = v - h;
    }
    fence_async_proxy();
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    tmem::fence_before_sync();
    __syncthreads();
    tmem::fence_after_sync();
    const uint32_t d_tmem = tmem_base;
    const uint64_t db_hi = smem_desc(b_hi, kK), db_lo = smem_desc(b_lo, kK);

    // the three products of one half tile (operand buffer `buf`, bands [64 half, 64 half + 64)) into the accumulator at `col`
    auto issue = [&](uint32_t col, int buf, int half, bool first) {
        const uint64_t da_hi = smem_desc(a_base + (2 * buf) * kTileFloats, kKH), da_lo = smem_desc(a_base + (2 * buf + 1) * kTileFloats, kKH);
        const uint64_t b_adv = (uint64_t)(half * (kKH / 4) * 128 >> 4);
#ifndef SEPT_TC_NO_MMA
#pragma unroll 1
        for (int ks = 0; ks < kKH / 8; ++ks) {
            const uint64_t adv = (uint64_t)(ks * 2 * 128 >> 4);   // two core matrices along K per step
            mma_tf32(d_tmem + col, da_hi + adv, db_hi + b_adv + adv, (first && ks == 0) ? 0u : 1u);
            mma_tf32(d_tmem + col, da_lo + adv, db_hi + b_adv + adv, 1u);
            mma_tf32(d_tmem + col, da_hi + adv, db_lo + b_adv + adv, 1u);
        }
#endif
        mma_commit(bars + buf);
    };
    // completions of each buffer's barrier: committed by the issuing thread's program order (the same for every thread),
    // observed at most once each -- a buffer is only refilled, and an accumulator only read, after `ensure`
    uint32_t committed[2] = {0, 0}, observed[2] = {0, 0};
    auto ensure = [&](int buf) {
        while (observed[buf] < committed[buf]) { mbar_wait(bars + buf, observed[buf] & 1u); ++observed[buf]; }
    };

    // A half tile is 64 units of (8 frames x 4 four-band chunks); unit = it * kWarps + warp, so a warp's 16-byte stores
    // of one iteration cover 512 contiguous bytes of the operand layout and its loads eight 64-byte row segments
    const int fr = lane & 7, kc4 = lane >> 3;
    const int m = 32 * (warp & 3) + lane;                          // epilogue: this thread's frame = its TMEM lane ...
    const int c0 = kCoefPerWarp * (warp >> 2);                     // ... and its share of the 40 coefficients
    const uint32_t lane_addr = d_tmem + ((uint32_t)(32 * (warp & 3)) << 16) + (uint32_t)c0;

    auto load = [&](float4 (&pw)[2][kIt], long long g0, int s) {
#pragma unroll
        for (int half = 0; half < 2; ++half)
#pragma unroll
            for (int it = 0; it < kIt; ++it) {
                const int unit = it * kWarps + warp, f = 8 * (unit >> 2) + fr, kc = 16 * half + 4 * (unit & 3) + kc4;
                const long long g = g0 + f;
#ifdef SEPT_TC_NO_LOAD
                pw[half][it] = make_float4((float)g, (float)kc, 1.f, 2.f);
#else
                pw[half][it] = g < prm.total_frames
                                   ? __ldcs(reinterpret_cast<const float4*>(prm.power + ((long long)s * prm.total_frames + g) * kK) + kc)
                                   : make_float4(0.f, 0.f, 0.f, 0.f);
#endif
            }
    };
    // mode 0: clamped dB of the stream; mode 1: the residual r of stream 3
    auto fill = [&](int buf, const float4 (&pw)[kIt], const float* floor_of, int mode) {
        ensure(buf);
        float* hi = a_base + (2 * buf) * kTileFloats;
        float* lo = hi + kTileFloats;
#pragma unroll
        for (int it = 0; it < kIt; ++it) {
            const int unit = it * kWarps + warp, f = 8 * (unit >> 2) + fr, kcl = 4 * (unit & 3) + kc4;   // chunk within the half
            const float fl = floor_of[f];
            const float v[4] = {pw[it].x, pw[it].y, pw[it].z, pw[it].w};
            float4 h, l;
            float* hp = &h.x;
            float* lp = &l.x;
#pragma unroll
            for (int c = 0; c < 4; ++c) {
#ifdef SEPT_TC_NO_CVT
                float d = fmaxf(v[c], fl);
#else
                float d = fmaxf(power_to_db(v[c]), fl);
#endif
                if (mode == 1) d = fmaxf(0.f, -(100.0f - kDbQuarter) - d);
                hp[c] = tf32_hi(d);
                lp[c] = d - hp[c];
            }
            const int at = ((unit >> 2) * (kKH / 4) + kcl) * 32 + fr * 4;            // canon(f, 4 kcl, kKH)
            *reinterpret_cast<float4*>(hi + at) = h;
            *reinterpret_cast<float4*>(lo + at) = l;
        }
        fence_async_proxy();
        tmem::fence_before_sync();
    };
    auto launch_mma = [&](uint32_t col, int buf, int half, bool first) {       // after the barrier that follows fill()
        if (tid == 0) { tmem::fence_after_sync(); issue(col, buf, half, first); }
        ++committed[buf];
    };
    // per-tile metadata: the floors of the tile's 128 frames (threads 0-127, into buffer `tb`) and this thread's output slot
    struct Slot { long long f0; int T, t; bool live; };
    auto load_meta = [&](long long g0, int tb) {
        if (tid < kM) {
            const long long g = g0 + tid;
            const int u = g < prm.total_frames ? __ldg(prm.frame_utt + g) : 0;
            floors[(2 * tb) * kM + tid] = power_to_db(__int_as_float(__ldg(prm.utt_max + u))) - prm.top_db;
            floors[(2 * tb + 1) * kM + tid] = power_to_db(__int_as_float(__ldg(prm.utt_max + prm.n_utts + u))) - prm.top_db;
        }
        Slot sl{0, 1, 0, false};
        const long long g = g0 + m;
        if (g < prm.total_frames) {
            const int u = __ldg(prm.frame_utt + g);
            sl.f0 = __ldg(prm.frame_off + u);
            sl.T = (int)(__ldg(prm.frame_off + u + 1) - sl.f0);
            sl.t = (int)(g - sl.f0);
            sl.live = true;
        }
        return sl;
    };

    float4 pw0[2][kIt], pw1[2][kIt];
    int tile = blockIdx.x, tb = 0;
    Slot cur = load_meta((long long)tile * kM, 0);
    load(pw0, (long long)tile * kM, 0);
    for (; tile < n_tiles; tile += gridDim.x, tb ^= 1) {
        const long long g0 = (long long)tile * kM;
        const float* floor0 = floors + (2 * tb) * kM;
        const float* floor1 = floor0 + kM;
        __syncthreads();                                           // this tile's floors are in place; the previous epilogue is done
        load(pw1, g0, 1);                                          // stream 1 flies while stream 0 is converted
        fill(0, pw0[0], floor0, 0);
        __syncthreads();
        launch_mma(0, 0, 0, true);
        fill(1, pw0[1], floor0, 0);
        __syncthreads();
        launch_mma(0, 1, 1, false);
        // next tile: stream 0 (in flight across the rest of this tile) and, below, its metadata
        const int nxt = tile + gridDim.x;
        if (nxt < n_tiles) load(pw0, (long long)nxt * kM, 0);
        bool need_r = false;
#pragma unroll
        for (int half = 0; half < 2; ++half)
#pragma unroll
            for (int it = 0; it < kIt; ++it) {                     // does any frame of the tile dip below -93.98 dB?
                const int unit = it * kWarps + warp, f = 8 * (unit >> 2) + fr;
                const float lim = -(100.0f - kDbQuarter);
                if (floor1[f] < lim) {
                    const float mn = fminf(fminf(pw1[half][it].x, pw1[half][it].y), fminf(pw1[half][it].z, pw1[half][it].w));
                    need_r |= fmaxf(power_to_db(mn), floor1[f]) < lim;
                }
            }
        fill(0, pw1[0], floor1, 0);
        __syncthreads();
        launch_mma(64, 0, 0, true);
        Slot next_slot{0, 1, 0, false};
        if (nxt < n_tiles) next_slot = load_meta((long long)nxt * kM, tb ^ 1);
        fill(1, pw1[1], floor1, 0);                                // ensure(1): stream 0 is complete in tensor memory
        const int any_r = __syncthreads_or(need_r ? 1 : 0);
        launch_mma(64, 1, 1, false);

        // ---- epilogue of stream 0 (the tensor core is busy with stream 1) ------------------------------------------
        float* out = prm.out + cur.f0 * (3 * kNC) + cur.t;
        uint32_t v[kCoefPerWarp];
        tmem::fence_after_sync();
        tmem::ld<kCoefPerWarp>(lane_addr, v);
        tmem::wait_ld();
#ifndef SEPT_TC_NO_STORE
        if (cur.live) {
#pragma unroll
            for (int c = 0; c < kCoefPerWarp; ++c) out[(long long)(c0 + c) * cur.T] = __uint_as_float(v[c]);
        }
#endif
        // ---- streams 1 and 2 --------------------------------------------------------------------------------------------
        if (any_r) {                                               // rare: some band of stream 3 sits on the -100 dB clamp
            load(pw1, g0, 1);
            fill(0, pw1[0], floor1, 1);
            __syncthreads();
            launch_mma(128, 0, 0, true);
            fill(1, pw1[1], floor1, 1);
            __syncthreads();
            launch_mma(128, 1, 1, false);
            ensure(0);
        }
        ensure(1);
        tmem::fence_after_sync();
        tmem::ld<kCoefPerWarp>(lane_addr + 64, v);
        uint32_t r[kCoefPerWarp];
        if (any_r) tmem::ld<kCoefPerWarp>(lane_addr + 128, r);
        tmem::wait_ld();
#ifndef SEPT_TC_NO_STORE
        if (cur.live)
#else
        if (cur.live && v[0] == 0x12345678u)
#endif
        {
#pragma unroll
            for (int c = 0; c < kCoefPerWarp; ++c) {
                const float d1 = __uint_as_float(v[c]);
                float d2 = d1 + (any_r ? __uint_as_float(r[c]) : 0.f);
                if (c0 + c == 0) d2 -= kDbQuarter * kSqrtBands;
                out[(long long)(kNC + c0 + c) * cur.T] = d1;
                out[(long long)(2 * kNC + c0 + c) * cur.T] = d2;
            }
        }
        tmem::fence_before_sync();
        cur = next_slot;
    }
    tmem::fence_before_sync();
    __syncthreads();
    if (warp == 0) tmem::dealloc(tmem_base, kTmemCols);
}

cudaError_t launch_mfcc_dct_tc(const MfccDctParams& prm, int sms, cudaStream_t stream) {
    const long long tiles = (prm.total_frames + kM - 1) / kM;
    if (tiles == 0) return cudaSuccess;
    cudaError_t e = cudaFuncSetAttribute(mfcc_dct_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmemBytes);
    if (e != cudaSuccess) return e;
    const int grid = (int)(tiles < sms ? tiles : sms);
    mfcc_dct_tc_kernel<<<grid, kThreads, kSmemBytes, stream>>>(prm, (int)tiles);
    return cudaGetLastError();
}

}  // namespace sept
