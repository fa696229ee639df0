"""Stall ratios, pipe utilisation and issue statistics of one kernel from an ncu report. usage: python tools/ncu_stalls.py <report> [<report2> ...]"""
import csv, io, subprocess, sys
def load(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    r = list(csv.reader(io.StringIO(out)))
    return dict(zip(r[0], r[-1]))
reps = [load(p) for p in sys.argv[1:]]
keys = [k for k in reps[0] if ("issue_stalled" in k and k.endswith("per_issue_active.ratio")) or
        ("sm__inst_executed_pipe_" in k and k.endswith("avg.pct_of_peak_sustained_active")) or
        ("sm__pipe_" in k and k.endswith("cycles_active.avg.pct_of_peak_sustained_active")) or
        k in ("sm__inst_executed.avg.per_cycle_active", "smsp__inst_executed.sum", "gpu__time_duration.sum", "smsp__thread_inst_executed_per_inst_executed.ratio",
              "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.per_cycle_active", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
              "l1tex__lsu_writeback_active.avg.pct_of_peak_sustained_elapsed", "l1tex__data_pipe_lsu_wavefronts.sum.pct_of_peak_sustained_elapsed", "sm__mio_inst_issued.sum.pct_of_peak_sustained_elapsed" )]
def f(v):
    try: return float(v.replace(",", ""))
    except Exception: return 0.0
for k in sorted(keys, key=lambda k: -f(reps[0][k])):
    if max(f(r.get(k, "0")) for r in reps) < 0.05: continue
    print("%-95s %s" % (k, "  ".join("%12s" % r.get(k, "-") for r in reps)))
