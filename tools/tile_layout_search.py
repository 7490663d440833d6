#!/usr/bin/env python3
"""Design aid (CPU only): exhaustive search of the power-tile layout `bin_pos(k) = k + ((c * (k >> s)) >> d)` against the
extraction kernel's own 8-byte store patterns (csrc/extract_core.cuh: pass2_split_store for R <= 16, split_store_all for
R = 32).  Model: 8-byte accesses are served per half warp; a half warp costs the largest number of distinct 8-byte words
that share a bank pair (16 bank pairs).  Prints, per R, the conflict-free ideal, the cost of the old k + k/16 layout and
the best layouts found -- the constants in Geo<R>::bin_pos / tables.h: power_tile_pos come from here.
(tools/bank_sim.py models the layouts of the earlier kernel generations and is kept for the record.)"""


def wf_half(slots):
    banks = {}
    for s in slots:
        banks.setdefault(s % 16, set()).add(s)
    return max(len(v) for v in banks.values()) if banks else 0


def store_instructions(R):
    NC = 25 * R
    crt = {(c % R, c % 25): c for c in range(NC)}
    instrs = []
    if R <= 16:  # fused pass: lanes of a half warp = rows j = 0..12 of one frame pair, one k1 per instruction
        for k1 in range(R):
            a = [crt[(k1, j)] for j in range(13)]
            b = [crt[(k1, (25 - j) % 25)] for j in range(1, 13)]
            if k1 == 0:
                b.append(NC)
            instrs += [a, b]
        mult = 2 * ((32 // R + 1) // 2)  # half warps x rounds per item
    else:  # unfused: lane = k1, one row k2 per instruction, bins k and NC - k
        for k2 in range(13):
            for h in (0, 16):
                a, b = [], []
                for k1 in range(h, h + 16):
                    if k2 == 0 and k1 > R // 2:
                        continue
                    k = crt[(k1, k2)]
                    a.append(k)
                    if k != NC - k:
                        b.append(NC - k)
                instrs += [a, b]
        mult = 1
    return instrs, mult


if __name__ == "__main__":
    for R in (8, 16, 32):
        instrs, mult = store_instructions(R)
        ideal = sum(1 for i in instrs if i) * mult
        old = sum(wf_half([k + (k >> 4) for k in i]) for i in instrs) * mult
        res = []
        for s in (3, 4, 5):
            for c in range(0, 33):
                for d in range(0, 4):
                    cost = sum(wf_half([k + ((c * (k >> s)) >> d) for k in i]) for i in instrs) * mult
                    res.append((cost, s, c, d))
        res.sort()
        print(f"R={R}: conflict-free {ideal}, old layout {old}, best (cost, s, c, d): {res[:4]}")
