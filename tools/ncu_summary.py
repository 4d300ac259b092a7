#!/usr/bin/env python
"""Summarise an .ncu-rep (read here, no GPU needed) into a small markdown file for profiles/.

    python tools/ncu_summary.py gpurun_out/prof.ncu-rep profiles/r01_xxx.md ["title"]
"""
import collections
import csv
import io
import subprocess
import sys

KEYS = [
    "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
    "launch__shared_mem_per_block_dynamic", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
    "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_tensor_subpipe_dmma.avg.pct_of_peak_sustained_active",
    "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_tma.avg.pct_of_peak_sustained_active",
    "sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed",
    "sm__cycles_elapsed.max", "smsp__cycles_active.avg",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_bytes.sum",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
    "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_membar_per_issue_active.ratio",
]


def ncu(rep, page):
    out = subprocess.run(["ncu", "-i", rep, "--page", page, "--csv"], capture_output=True, text=True).stdout
    return list(csv.reader(io.StringIO(out)))


def main():
    rep, dst = sys.argv[1], sys.argv[2]
    title = sys.argv[3] if len(sys.argv) > 3 else rep
    raw = ncu(rep, "raw")
    hdr, units = raw[0], raw[1]
    lines = [f"# {title}", "", f"source: `{rep}` (`ncu --set full --clock-control none --import-source on`)", ""]
    name_col = hdr.index("Kernel Name") if "Kernel Name" in hdr else None
    for row in raw[2:]:
        if len(row) < len(hdr):
            continue
        lines.append(f"## {row[name_col][:110] if name_col is not None else 'kernel'}")
        lines.append("")
        lines.append("| metric | value | unit |")
        lines.append("|---|---|---|")
        for k in KEYS:
            if k in hdr:
                i = hdr.index(k)
                lines.append(f"| `{k}` | {row[i]} | {units[i]} |")
        lines.append("")
    src = ncu(rep, "source")
    # the source page lists one kernel; aggregate stall samples by opcode
    for start in range(len(src)):
        if src[start] and src[start][0] == "Address":
            break
    else:
        start = None
    if start is not None:
        h = src[start]
        idx = {x: i for i, x in enumerate(h)}
        stalls = [x for x in h if x.startswith("stall_") and "Not Issued" not in x]
        tot, ex = collections.Counter(), collections.Counter()
        st = collections.defaultdict(collections.Counter)
        for r in src[start + 1:]:
            if len(r) < len(h) or r[0] == "Address":
                continue
            s = r[idx["Source"]].strip()
            if not s:
                continue
            toks = s.split()
            op = toks[1] if toks[0].startswith("@") and len(toks) > 1 else toks[0]
            op = op.split(".")[0]
            n = int(r[idx["# Samples"]] or 0)
            tot[op] += n
            ex[op] += int(r[idx["Instructions Executed"]] or 0)
            for q in stalls:
                st[op][q] += int(r[idx[q]] or 0)
        T, E = sum(tot.values()) or 1, sum(ex.values()) or 1
        lines += ["## warp-stall samples by opcode (first kernel of the source page)", "",
                  f"total samples {T}, warp instructions executed {E}", "",
                  "| opcode | samples % | executed % | top stall reasons |", "|---|---|---|---|"]
        for op, n in tot.most_common(16):
            top = ", ".join(f"{k[6:]} {v}" for k, v in st[op].most_common(3))
            lines.append(f"| {op} | {100 * n / T:.1f} | {100 * ex[op] / E:.2f} | {top} |")
        lines.append("")
    open(dst, "w").write("\n".join(lines))
    print("wrote", dst)


if __name__ == "__main__":
    main()
