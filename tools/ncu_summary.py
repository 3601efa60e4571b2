"""Summarise an .ncu-rep (read here, no GPU): key counters + per-region sample breakdown of the trace kernel."""
import csv, subprocess, sys, io, json
rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, vals = rows[0], rows[1], rows[2]
M = {h: (v, u) for h, u, v in zip(hdr, units, vals)}
keys = ["gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "smsp__thread_inst_executed_per_inst_executed.ratio",
        "smsp__inst_executed.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "sm__cycles_active.avg",
        "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_dispatch_stall_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio"]
out = {}
for k in keys:
    if k in M:
        out[k] = M[k][0] + " " + M[k][1]
        print(f"{k:90s} {M[k][0]} {M[k][1]}")
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
hdr, data = rows[1], rows[2:]
iS, iSamp, iEx, iThr = hdr.index("Source"), hdr.index("# Samples"), hdr.index("Instructions Executed"), hdr.index("Avg. Threads Executed")
tot_s = sum(int(r[iSamp]) for r in data); tot_e = sum(int(r[iEx]) for r in data)
def opname(r):
    t = r[iS].split()
    return t[1] if t[0].startswith("@") else t[0]
# classify each instruction by the dominant opcode family of its 40-instruction neighbourhood
regions = {}
for k in range(0, len(data), 40):
    blk = data[k:k + 40]
    ops = [opname(r) for r in blk]
    packed = sum(o in ("FFMA2", "FADD2", "FMUL2") for o in ops)
    name = "sweep (packed FP32x2)" if packed >= 12 else f"other@{k}"
    s = sum(int(r[iSamp]) for r in blk); e = sum(int(r[iEx]) for r in blk)
    thr = sum(float(r[iThr]) * int(r[iEx]) for r in blk) / max(1, e)
    a = regions.setdefault(name, [0, 0, 0.0]); a[0] += s; a[1] += e; a[2] += thr * e
print(f"\ntotal samples {tot_s}, instructions {tot_e}")
for name, (s, e, t) in sorted(regions.items(), key=lambda x: -x[1][0])[:14]:
    print(f"{name:28s} samples {100 * s / tot_s:5.1f}%  inst {100 * e / tot_e:5.1f}%  avg threads {t / max(1, e):4.1f}")
