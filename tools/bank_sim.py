#!/usr/bin/env python3
"""[Historical: models the tile layouts and lane->task maps of the kernel generations before the mel gather program;
the current layout search is tools/tile_layout_search.py and the gather schedule is checked by tests/hostsim/mel_program.cpp.]
Shared-memory wavefront simulator for the extraction kernel's access patterns (design aid, CPU only).
Model: 32 banks x 4 B; a warp access of 8 B/lane is served as 2 half-warps, 16 B/lane as 4 quarter-warps; within a
group the wavefront count is the max over banks of DISTINCT 4-byte words touched (same word = broadcast)."""
import sys
import numpy as np

def wavefronts(byte_addrs, size, active=None):
    """byte_addrs: per-lane start byte address (len 32, None = inactive); size in {4,8,16}."""
    lanes = [(l, a) for l, a in enumerate(byte_addrs) if a is not None]
    group = {4: 32, 8: 16, 16: 8}[size]
    total = 0
    for g0 in range(0, 32, group):
        words = {}
        for l, a in lanes:
            if g0 <= l < g0 + group:
                for w in range(a // 4, (a + size) // 4):
                    words.setdefault(w % 32, set()).add(w)
        if words:
            total += max(len(v) for v in words.values())
    return total

def crt(R, k1, k2):
    NC = 25 * R
    for c in range(NC):
        if c % R == k1 and c % 25 == k2:
            return c

def fused_store_wf(R, pos, PP):
    """fused pass2+split stores: lane = task (p, j), per k1 two STS.64"""
    NC, PPW = 25 * R, 32 // R
    tot = 0
    n_tasks = PPW * 13
    for rnd in range((n_tasks + 31) // 32):
        for k1 in range(R):
            a1, a2 = [], []
            for lane in range(32):
                t = lane + 32 * rnd
                if t >= n_tasks:
                    a1.append(None); a2.append(None); continue
                p, j = divmod(t, 13)
                rb = (25 - j) % 25
                a1.append(8 * (p * PP + pos(crt(R, k1, j))))
                if j > 0:
                    a2.append(8 * (p * PP + pos(crt(R, k1, rb))))
                else:
                    a2.append(8 * (p * PP + pos(NC)) if k1 == 0 else None)
            tot += wavefronts(a1, 8) + wavefronts(a2, 8)
    return tot

def fused_load_wf(R, YS, YP):
    PPW = 32 // R
    tot = 0
    n_tasks = PPW * 13
    for rnd in range((n_tasks + 31) // 32):
        for which in (0, 1):
            for i in range(R):
                a = []
                for lane in range(32):
                    t = lane + 32 * rnd
                    if t >= n_tasks:
                        a.append(None); continue
                    p, j = divmod(t, 13)
                    row = j if which == 0 else (25 - j) % 25
                    a.append(16 * (p * YP + row * YS + i))
                tot += wavefronts(a, 16)
    return tot

if __name__ == "__main__":
    R = int(sys.argv[1]) if len(sys.argv) > 1 else 16
    NC = 25 * R
    cur = lambda k: k + ((k >> 4) << 1)
    print("current fused stores wf/item:", fused_store_wf(R, cur, NC + 2 * (NC // 16) + 8))
    print("current fused loads wf/item:", fused_load_wf(R, R + 1, 25 * (R + 1) + 2))

def search(R=16):
    NC = 25 * R
    best = []
    for YS in range(R, R + 8):
        for pad in range(0, 16):
            YP = 25 * YS + pad
            best.append((fused_load_wf(R, YS, YP), YS, pad))
    best.sort()
    print("loads best:", best[:6])
    res = []
    for blk in (4, 8, 16, 32):
        for padq in (0, 2, 4, 6, 10, 14, 18):
            sh = blk.bit_length() - 1
            pos = lambda k, sh=sh, padq=padq: k + (k >> sh) * padq
            for ppad in (0, 2, 4, 6, 8, 10, 12, 14):
                PP = pos(NC + 4) + 4 + ppad
                res.append((fused_store_wf(R, pos, PP), blk, padq, ppad))
    res.sort()
    print("stores best (wf, block, pad slots per block, pair pad):", res[:8])

def band_runs(n_fft, n_mels=128, align=4):
    sys.path.insert(0, str(__import__("pathlib").Path(__file__).resolve().parents[1]))
    from oracle import restate
    fb = restate.melscale_fbanks_htk(n_fft // 2 + 1, n_mels, dtype=np.float64).astype(np.float32)
    runs = []
    for m in range(n_mels):
        nz = np.nonzero(fb[:, m])[0]
        if len(nz) == 0:
            runs.append((0, 0)); continue
        k0 = int(nz[0]) // align * align
        n = -(-(int(nz[-1]) + 1 - k0) // align) * align
        runs.append((k0, n))
    return runs

def fused_store_wf2(R, pos, PP):
    """lane = 16 * (p % 2) + j (j < 13); rounds over pair groups"""
    NC, PPW = 25 * R, 32 // R
    tot = 0
    for rnd in range((PPW + 1) // 2):
        for k1 in range(R):
            a1, a2 = [], []
            for lane in range(32):
                p, j = 2 * rnd + lane // 16, lane % 16
                if j >= 13 or p >= PPW:
                    a1.append(None); a2.append(None); continue
                rb = (25 - j) % 25
                a1.append(8 * (p * PP + pos(crt(R, k1, j))))
                if j > 0:
                    a2.append(8 * (p * PP + pos(crt(R, k1, rb))))
                else:
                    a2.append(8 * (p * PP + pos(NC)) if k1 == 0 else None)
            tot += wavefronts(a1, 8) + wavefronts(a2, 8)
    return tot

def mel_read_wf(R, pos, PP, runs, width, band_of=None):
    """lane = band (m = band_of[round][lane]); per step every lane reads `width` bytes (8: one bin, 16: two bins) of
    each pair"""
    PPW = 32 // R
    n_mels = len(runs)
    tot = 0
    per = width // 8
    for rnd in range((n_mels + 31) // 32):
        ms = [band_of[rnd][l] if band_of else rnd * 32 + l for l in range(32)]
        steps = max((runs[m][1] for m in ms if m < n_mels), default=0) // per
        for t in range(steps):
            for p in range(PPW):
                a = []
                for m in ms:
                    if m >= n_mels or t * per >= runs[m][1]:
                        a.append(None)
                    else:
                        a.append(8 * (p * PP + pos(runs[m][0] + t * per)))
                tot += wavefronts(a, width)
    return tot

def search2(R=16):
    NC = 25 * R
    n_fft = 2 * NC
    out = []
    for align, width in ((4, 16), (2, 16), (4, 8), (2, 8), (1, 8)):
        runs = band_runs(n_fft, 128, align)
        for sh in (2, 3, 4, 5):
            for a in range(0, 20):
                if width == 16 and a % 2:
                    continue
                pos = lambda k, sh=sh, a=a: k + (k >> sh) * a
                PP = pos(NC + 4) + 8
                PP += PP & 1
                st = fused_store_wf2(R, pos, PP)
                ml = mel_read_wf(R, pos, PP, runs, width)
                out.append((st + ml, st, ml, align, width, sh, a))
    out.sort()
    for o in out[:12]:
        print("total %d stores %d mel %d | align %d width %d  pos = k + (k>>%d)*%d" % o)


def segments(n_fft, n_mels=128):
    """bins grouped by the mel segment [f_s, f_{s+1}) they fall in (s = 0 .. n_mels); returns (k_start, n_bins) per
    segment, derived from the dense filterbank: bin k has rising weight in band s and falling weight in band s-1"""
    sys.path.insert(0, str(__import__("pathlib").Path(__file__).resolve().parents[1]))
    from oracle import restate
    import math
    n_freqs = n_fft // 2 + 1
    m_max = 2595.0 * math.log10(1.0 + 8000.0 / 700.0)
    m_pts = np.linspace(0.0, m_max, n_mels + 2)
    f_pts = 700.0 * (10.0 ** (m_pts / 2595.0) - 1.0)
    freqs = np.linspace(0, 8000, n_freqs)
    seg = np.searchsorted(f_pts, freqs, side="right") - 1          # f_pts[seg] <= f < f_pts[seg+1]
    seg = np.clip(seg, 0, n_mels)
    out = []
    for s in range(n_mels + 1):
        ks = np.nonzero(seg == s)[0]
        out.append((int(ks[0]), len(ks)) if len(ks) else (0, 0))
    return out


def mel_segment_wf(R, pos, PP, segs, lanes_per_round=26):
    PPW = 32 // R
    tot_wf = tot_steps = 0
    for r0 in range(0, len(segs), lanes_per_round):
        grp = segs[r0:r0 + lanes_per_round]
        steps = max(n for _, n in grp)
        tot_steps += steps
        for t in range(steps):
            aw = [8 * (k0 + t) if (l < len(grp) and t < grp[l][1]) else None for l, (k0, _) in enumerate(grp + [(0, 0)] * (32 - len(grp)))]
            tot_wf += wavefronts(aw, 8)                                  # (u, d) weight pair of the bin
            for p in range(PPW):
                a = [8 * (p * PP + pos(k0 + t)) if t < n else None for (k0, n) in grp] + [None] * (32 - len(grp))
                tot_wf += wavefronts(a, 8)
    return tot_wf, tot_steps
