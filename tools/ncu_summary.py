#!/usr/bin/env python
"""Summarise an ncu report: key raw metrics per launch plus the source-level stall
breakdown split at the kernel's BAR.SYNC instructions.  Usage: ncu_summary.py rep [out.txt]"""
import csv
import io
import subprocess
import sys

RAW = [
    "gpu__time_duration.sum", "launch__registers_per_thread", "launch__grid_size",
    "launch__block_size", "launch__shared_mem_per_block_dynamic", "launch__waves_per_multiprocessor",
    "sm__warps_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
    "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__throughput.avg.pct_of_peak_sustained_elapsed",
    "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "lts__t_bytes.sum",
    "sm__cycles_elapsed.avg", "sm__cycles_active.avg", "gpc__cycles_elapsed.avg.per_second",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
    "smsp__inst_executed.sum", "sm__sass_thread_inst_executed_op_dfma_pred_on.sum",
    "sm__sass_thread_inst_executed_op_dmul_pred_on.sum", "sm__sass_thread_inst_executed_op_dadd_pred_on.sum",
    "sm__sass_thread_inst_executed_op_ffma_pred_on.sum",
]
STALLS = ["long_sb", "math", "wait", "short_sb", "barrier", "not_selected", "selected", "dispatch",
          "branch_resolving", "no_inst", "mio", "lg", "membar", "drain", "tex", "sleep", "misc"]


def run(args):
    return subprocess.run(["ncu"] + args, capture_output=True, text=True).stdout


def main():
    rep = sys.argv[1]
    out = open(sys.argv[2], "w") if len(sys.argv) > 2 else sys.stdout
    rows = list(csv.reader(io.StringIO(run(["-i", rep, "--page", "raw", "--csv"]))))
    hdr, units = rows[0], rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    names = []
    for r in rows[2:]:
        name = r[idx["Kernel Name"]]
        names.append(name)
        print("==== launch:", name[:150], file=out)
        for m in RAW:
            if m in idx:
                print("  %-75s %s %s" % (m, r[idx[m]], units[idx[m]]), file=out)
        for h in hdr:
            if "issue_stalled" in h and h.endswith("_per_issue_active.ratio") and "not_issued" not in h:
                v = float(r[idx[h]] or 0)
                if v > 0.05:
                    print("  %-75s %.3f" % (h.replace("smsp__average_warps_issue_stalled_", "stall:"), v), file=out)
    # source-level view of the first launch
    txt = run(["-i", rep, "--page", "source", "--csv", "--launch-skip", "0", "--launch-count", "1"])
    rows = list(csv.reader(io.StringIO(txt)))
    hrow = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
    hdr = rows[hrow]
    idx = {h: i for i, h in enumerate(hdr)}
    data = [r for r in rows[hrow + 1:] if r and r[0].startswith("0x") and len(r) >= len(hdr) - 2]
    # the listing can repeat the function; keep the first copy
    seen, uniq = set(), []
    for r in data:
        if r[0] in seen:
            break
        seen.add(r[0])
        uniq.append(r)
    data = uniq
    S = idx["# Samples"]
    if len(sys.argv) > 3:  # per-instruction table for offline analysis
        with open(sys.argv[3], "w") as instr_csv:
            w = csv.writer(instr_csv)
            keep = ["Address", "Source", "# Samples", "Instructions Executed"] + ["stall_" + k for k in STALLS if "stall_" + k in idx]
            w.writerow(keep)
            for r in data:
                w.writerow([r[idx[k]].strip() for k in keep])
    tot = sum(int(r[S]) for r in data) or 1
    bars = [i for i, r in enumerate(data) if "BAR.SYNC" in r[idx["Source"]]]
    bounds = [0] + bars + [len(data)]
    print("\n==== source-level regions split at BAR.SYNC (first launch); total samples %d" % tot, file=out)
    for a, b in zip(bounds[:-1], bounds[1:]):
        seg = data[a:b]
        s = sum(int(r[S]) for r in seg)
        ie = sum(int(r[idx["Instructions Executed"]]) for r in seg)
        dp = sum(int(r[idx["Instructions Executed"]]) for r in seg
                 if any(x in r[idx["Source"]] for x in ("DFMA", "DMUL", "DADD", "DSETP")))
        st = {k: sum(int(r[idx["stall_" + k]]) for r in seg) for k in STALLS if "stall_" + k in idx}
        st = {k: v for k, v in st.items() if v > 0.02 * max(s, 1)}
        print("  rows %5d-%5d  samples %8d (%5.1f%%)  warp-inst %11d  fp64-inst %11d  %s"
              % (a, b, s, 100.0 * s / tot, ie, dp, st), file=out)
    print("\n==== top 40 instructions by samples", file=out)
    for r in sorted(data, key=lambda r: -int(r[S]))[:40]:
        st = {k: int(r[idx["stall_" + k]]) for k in STALLS if "stall_" + k in idx}
        st = {k: v for k, v in st.items() if v > 0.15 * max(int(r[S]), 1)}
        print("  %7s  %-72s %s" % (r[S], r[idx["Source"]].strip()[:72], st), file=out)


if __name__ == "__main__":
    main()
