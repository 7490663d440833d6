#!/usr/bin/env python3
"""Condense an .ncu-rep (one profiled kernel launch) into the text summary kept under profiles/.

    python tools/ncu_summary.py gpurun_out/prof.ncu-rep profiles/extract_r01.txt [kernel-name regex, for reports with several kernels]
"""
import csv
import io
import subprocess
import sys

KEYS = ["gpu__time_duration.sum", "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic", "launch__grid_size",
        "launch__block_size", "sm__warps_active.avg.per_cycle_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fmaheavy.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared_op_ld.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared_op_st.sum",
        "memory_l1_wavefronts_shared", "memory_l1_wavefronts_shared_ideal", "sass__inst_executed_shared_loads", "sass__inst_executed_shared_stores",
        "smsp__average_warp_latency_per_inst_issued.ratio",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_dispatch_stall_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio"]


def page(rep, name, kernel=None):
    cmd = ["ncu", "-i", rep, "--page", name, "--csv"] + (["--kernel-name", "regex:" + kernel] if kernel else [])
    out = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True).stdout
    return list(csv.reader(io.StringIO(out)))


def main(rep, dst, kernel=None):
    lines = [f"ncu summary of {rep} (one launch, --set full --clock-control none)"]
    raw = page(rep, "raw", kernel)
    hdr, units, vals = raw[0], raw[1], raw[2]
    col = {h: (u, v) for h, u, v in zip(hdr, units, vals)}
    lines.append(f"kernel: {col.get('Kernel Name', ('', '?'))[1]}")
    for k in KEYS:
        if k in col:
            lines.append(f"  {k} [{col[k][0]}] = {col[k][1]}")
    src = [r for r in page(rep, "source", kernel) if len(r) > 45 and r[0].startswith("0x")]       # SASS rows only
    tot = sum(int(r[4]) for r in src) or 1
    lines.append(f"\nSASS regions (100 instructions each): share of {tot} warp-stall samples, instructions executed, excess smem wavefronts, opcode mix")
    for b in range(0, len(src), 100):
        seg = src[b:b + 100]
        s, ie, exc = sum(int(r[4]) for r in seg), sum(int(r[5]) for r in seg), sum(int(r[16]) for r in seg)
        ops = {}
        for r in seg:
            t = r[1].split()
            op = (t[1] if t[0].startswith("@") else t[0]).split(".")[0]
            ops[op] = ops.get(op, 0) + 1
        top = " ".join(f"{k}:{v}" for k, v in sorted(ops.items(), key=lambda kv: -kv[1])[:6])
        if ie:
            lines.append(f"  [{b:5d}] {100 * s / tot:5.1f}%  inst {ie / 1e6:8.2f}M  smem_excess {exc / 1e6:7.2f}M  {top}")
    lines.append("\ntop 20 instructions by stall samples (samples, executed, SASS | long_sb short_sb wait math mio)")
    for r in sorted(src, key=lambda r: -int(r[4]))[:20]:
        lines.append(f"  {r[4]:>6} {r[5]:>9} {r[1].strip()[:80]:80s} | {r[34]} {r[42]} {r[45]} {r[35]} {r[37]}")
    open(dst, "w").write("\n".join(lines) + "\n")
    print("\n".join(lines))


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2], sys.argv[3] if len(sys.argv) > 3 else None)
