"""Per-source-line / per-function instruction and stall-sample totals from an ncu report (cuda,sass source view).
usage: python tools/ncu_src.py <report.ncu-rep> [kernel-regex] [env_steps_in_launch]"""
import csv, io, os, re, subprocess, sys
rep = sys.argv[1]; kname = sys.argv[2] if len(sys.argv) > 2 else "sf_rollout_kernel"
nsteps = float(sys.argv[3]) if len(sys.argv) > 3 else 0
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass", "-k", "regex:" + kname],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
cur = None; hdr = None; agg = {}
for r in rows:
    if not r: continue
    if r[0] == "File Path": cur = os.path.basename(r[1]); continue
    if r[0] == "Line No": hdr = r; continue
    if r[0] == "Function Name" or hdr is None or not r[0].isdigit(): continue
    def g(name):  # index from the end: an unescaped quote in the source text can split the early fields
        v = r[len(r) - (len(hdr) - hdr.index(name))]
        try: return float(v)
        except ValueError: return 0.0
    a = agg.setdefault((cur, int(r[0])), [0, 0, 0, r[1]])
    a[0] += g("Instructions Executed"); a[1] += g("# Samples"); a[2] += g("stall_no_inst")
tot = sum(a[0] for a in agg.values()); stot = sum(a[1] for a in agg.values()); ntot = sum(a[2] for a in agg.values())
print("warp-instr %.0f  samples %.0f  no_inst samples %.0f%s" % (tot, stot, ntot, ("  instr/env-step %.0f" % (tot / nsteps)) if nsteps else ""))
for (f, ln), a in sorted(agg.items(), key=lambda kv: -kv[1][0])[:int(os.environ.get("TOP", "40"))]:
    print("%5.1f%% inst %5.1f%% smp  %-14s:%-4d %s" % (a[0] / tot * 100, a[1] / stot * 100, f, ln, a[3].strip()[:95]))
ph = {}
for fn in ("sf_render.cuh", "sf_step.cuh", "sf_kernels.cu", "sf_geom.h"):
    p = os.path.join(ROOT, "spacefortress_b200/csrc", fn)
    src = open(p).read().split("\n")
    marks = [(i + 1, re.search(r"(sf_\w+)\s*\(", l).group(1)) for i, l in enumerate(src)
             if re.match(r"(template.*)?(__device__|__global__|SF_HD|static)", l) and re.search(r"(sf_\w+)\s*\(", l)]
    marks.append((len(src) + 1, "end"))
    for (f, ln), a in agg.items():
        if f != fn: continue
        name = fn + ":?"
        for (s0, nm), (s1, _) in zip(marks, marks[1:]):
            if s0 <= ln < s1: name = fn.split(".")[0][3:] + ":" + nm; break
        q = ph.setdefault(name, [0, 0, 0]); q[0] += a[0]; q[1] += a[1]; q[2] += a[2]
for (f, ln), a in agg.items():
    if f not in ("sf_render.cuh", "sf_step.cuh", "sf_kernels.cu", "sf_geom.h"):
        q = ph.setdefault(f, [0, 0, 0]); q[0] += a[0]; q[1] += a[1]; q[2] += a[2]
for name, (n, s, ni) in sorted(ph.items(), key=lambda kv: -kv[1][0]):
    if n / tot > 0.001:
        print("%6.1f%% inst %6.1f%% smp %6.1f%% no_inst   %s%s" % (n / tot * 100, s / stot * 100, ni / max(ntot, 1) * 100, name, ("   %.0f/env-step" % (n / nsteps)) if nsteps else ""))
