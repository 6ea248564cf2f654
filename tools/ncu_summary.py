#!/usr/bin/env python
"""Summarises an exported ncu report: key raw metrics + the hottest source lines (needs -lineinfo).
usage: ncu_summary.py raw.csv source.csv [n_lines] [--json out.json segments_per_launch "kernel description" "source note"]"""
import csv
import sys


def fl(x):
    try:
        return float(x.replace(",", ""))
    except (ValueError, AttributeError):
        return 0.0


def sass_regions(rows, total_w, min_share=0.008):
    """SASS view of the source page (no source correlation): runs of consecutive instructions with the same execution
    count = one basic-block region; prints each region's share of the warp instructions and its mean active lanes."""
    H = rows[1]
    ia, isrc, ie, it = H.index("Address"), H.index("Source"), H.index("Instructions Executed"), H.index("Thread Instructions Executed")
    data = [(r[ia], r[isrc], fl(r[ie]), fl(r[it])) for r in rows[2:] if len(r) > it]
    tot = sum(x[2] for x in data)
    print("--- SASS regions (runs of equal execution count): instruction range, instructions, executions per instruction, share of warp "
          "instructions, mean active lanes, first instruction")
    i = 0
    while i < len(data):
        j, s, st = i, 0.0, 0.0
        while j < len(data) and abs(data[j][2] - data[i][2]) <= 0.03 * max(data[i][2], 1):
            s += data[j][2]; st += data[j][3]; j += 1
        if s / max(tot, 1) > min_share:
            print("%4d-%4d n=%3d exec %.3e share %5.2f%% lanes %4.1f | %s" % (i, j - 1, j - i, data[i][2], 100 * s / tot, st / max(s, 1), data[i][1].strip()[:60]))
        i = j
    print("SASS instructions %d, warp instructions %.3e (kernel total %.3e)" % (len(data), tot, total_w))


def write_json(d, out, segments, kernel, source):
    """The numbers bench.py quotes in its roofline object (profiles/rNN_trace_kernel.json)."""
    import json
    g = lambda k: fl(d[k][0]) if k in d else 0.0
    unit = lambda k: d[k][1] if k in d else ""
    def nbytes(k):
        v, u = g(k), unit(k).lower()
        return v * {"byte": 1, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9}.get(u, 1)
    ms = g("gpu__time_duration.sum") * {"ms": 1, "us": 1e-3, "ns": 1e-6, "s": 1e3}.get(unit("gpu__time_duration.sum"), 1)
    w = g("smsp__inst_executed.sum")
    lanes = g("smsp__thread_inst_executed_per_inst_executed.ratio")
    cyc = g("sm__cycles_elapsed.avg")
    shw = g("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum")
    j = {"kernel": kernel, "source": source, "duration_ms": ms, "warp_inst": w, "avg_active_threads_per_inst": lanes,
         "segments_per_launch": segments, "thread_inst_per_segment": w * lanes / segments, "warp_inst_per_segment": w / segments,
         "issue_slot_utilisation": g("smsp__issue_active.avg.pct_of_peak_sustained_active") / 100.0, "lane_efficiency": lanes / 32.0,
         "alu_pipe_utilisation": g("sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active") / 100.0,
         "fma_pipe_utilisation": g("sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active") / 100.0,
         "dram_bytes_per_launch": nbytes("dram__bytes_read.sum") + nbytes("dram__bytes_write.sum"),
         "shared_wavefronts": shw, "shared_wavefronts_per_cycle_per_sm": shw / max(cyc, 1) / max(g("launch__grid_size"), 1) if g("launch__grid_size") <= 148 else shw / max(cyc, 1) / 148,
         "registers": int(g("launch__registers_per_thread")), "threads_per_cta": int(g("launch__block_size")), "ctas_per_sm": 1}
    json.dump(j, open(out, "w"), indent=1)


def main():
    raw, src = sys.argv[1], sys.argv[2]
    nlines = int(sys.argv[3]) if len(sys.argv) > 3 and not sys.argv[3].startswith("--") else 30
    rows = list(csv.reader(open(raw)))
    hdr, units, vals = rows[0], rows[1], rows[2]
    d = {h: (vals[i], units[i]) for i, h in enumerate(hdr)}
    if "--json" in sys.argv:
        k = sys.argv.index("--json")
        write_json(d, sys.argv[k + 1], float(sys.argv[k + 2]), sys.argv[k + 3], sys.argv[k + 4])
    keys = ["Kernel Name", "gpu__time_duration.sum", "smsp__inst_executed.sum", "smsp__thread_inst_executed_per_inst_executed.ratio",
            "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
            "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "launch__occupancy_limit_registers",
            "launch__occupancy_limit_shared_mem", "dram__bytes_read.sum", "dram__bytes_write.sum", "sm__cycles_elapsed.avg",
            "smsp__sass_average_branch_targets_threads_uniform.pct", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
            "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
            "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
            "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "smsp__inst_executed_op_local_ld.sum", "smsp__inst_executed_op_local_st.sum",
            "smsp__inst_executed_op_shared_ld.sum"]
    for k in keys:
        if k in d:
            print("%-70s %s %s" % (k, d[k][0], d[k][1]))
    print("--- warp stall reasons (warps per issue-active cycle)")
    for h in hdr:
        if h.startswith("smsp__average_warps_issue_stalled_") and h.endswith("_per_issue_active.ratio") and "not_issued" not in h:
            print("  %-24s %s" % (h[len("smsp__average_warps_issue_stalled_"):-len("_per_issue_active.ratio")], d[h][0]))
    total_w = fl(d["smsp__inst_executed.sum"][0])
    rows = list(csv.reader(open(src)))
    if len(rows) > 2 and rows[1] and rows[1][0] == "Address":
        sass_regions(rows, total_w)
        return
    hidx = [i for i, r in enumerate(rows) if r and r[0] == "Line No"]
    allrows = []
    for h in hidx:
        fname = rows[h - 2][1] if h >= 2 else ""
        if not fname.endswith((".cu", ".cuh")):
            continue
        H = rows[h]
        ci, ti = H.index("Instructions Executed"), H.index("Thread Instructions Executed")
        for r in rows[h + 1:]:
            if not r or r[0] in ("File Path", "Function Name", "Line No"):
                break
            allrows.append((fl(r[ci]), fl(r[ti]), fname.split("/")[-1], r[0], r[1].strip()))
    s_w = sum(a for a, *_ in allrows)
    print("--- hottest source lines (share of warp instructions in .cu/.cuh sources, avg active threads)")
    for w, t, f, ln, text in sorted(allrows, key=lambda x: -x[0])[:nlines]:
        print("%6.2f%% act=%4.1f %s:%s  %s" % (100 * w / s_w, t / max(w, 1), f, ln, text[:100]))
    print("sum of source-attributed warp inst %.3e, kernel total %.3e" % (s_w, total_w))


if __name__ == "__main__":
    main()
