"""Map an ncu SASS source page to CUDA source lines via nvdisasm line info; print hot lines and phase totals.
usage: python tools/ncu_lines.py <report.ncu-rep> [kernel-substring]"""
import csv, os, re, subprocess, sys, tempfile
rep = sys.argv[1]; kname = sys.argv[2] if len(sys.argv) > 2 else "sf_rollout_kernel"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tmp = tempfile.mkdtemp()
subprocess.run("cd %s && cuobjdump -xelf all %s/spacefortress_b200/libsf_b200.so >/dev/null 2>&1 && nvdisasm -g -c sf_kernels.sm_100a.cubin > dis.txt 2>/dev/null" % (tmp, ROOT), shell=True)
lines = open(tmp + "/dis.txt").read().split("\n")
start = [i for i, l in enumerate(lines) if l.startswith(".text.") and kname in l][0]
end = [i for i, l in enumerate(lines) if l.startswith("//---------------------") and i > start]
end = end[0] if end else len(lines)
cur = None; seq = []
for l in lines[start:end]:
    m = re.search(r'//## File "([^"]+)", line (\d+)', l)
    if m: cur = (m.group(1).split("/")[-1], int(m.group(2))); continue
    m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(.*?);", l)
    if m: seq.append((int(m.group(1), 16), cur, m.group(2)))
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "-k", "regex:" + kname], capture_output=True, text=True).stdout
rows = list(csv.reader(out.split("\n")))
hi = [i for i, r in enumerate(rows) if r and r[0] == "Address"][0]
hdr = rows[hi]; data = [r for r in rows[hi + 1:] if len(r) > 5]
ia, ii, isamp = hdr.index("Address"), hdr.index("Instructions Executed"), hdr.index("# Samples")
base = int(data[0][ia], 16)
agg = {}; tot = stot = 0
for r in data:
    idx = (int(r[ia], 16) - base) // 16
    if idx >= len(seq): continue
    n = float(r[ii] or 0); s = float(r[isamp] or 0)
    a = agg.setdefault(seq[idx][1], [0, 0]); a[0] += n; a[1] += s; tot += n; stot += s
print("sass instrs", len(seq), "executed warp-instr", tot, "samples", stot)
src = {}
def getsrc(f, ln):
    for d in ("spacefortress_b200/csrc/", "include/"):
        p = os.path.join(ROOT, d, f)
        if os.path.exists(p):
            if p not in src: src[p] = open(p).read().split("\n")
            return src[p][ln - 1].strip()[:100] if ln - 1 < len(src[p]) else ""
    return ""
for k, (n, s) in sorted(agg.items(), key=lambda kv: -kv[1][0])[:int(os.environ.get("TOP", "40"))]:
    if k: print("%5.1f%% inst %5.1f%% smp  %-14s:%-4d %s" % (n / tot * 100, s / stot * 100, k[0], k[1], getsrc(*k)))
# phase totals: functions of sf_render.cuh by line ranges (found from the source)
rs = open(os.path.join(ROOT, "spacefortress_b200/csrc/sf_render.cuh")).read().split("\n")
marks = [(i + 1, re.search(r"(sf_\w+)\s*\(", l).group(1)) for i, l in enumerate(rs) if l.startswith("__device__") and re.search(r"(sf_\w+)\s*\(", l)]
marks.append((len(rs) + 1, "end"))
ph = {}
for k, (n, s) in agg.items():
    if not k: continue
    name = k[0]
    if k[0] == "sf_render.cuh":
        for (a, nm), (b, _) in zip(marks, marks[1:]):
            if a <= k[1] < b: name = "render:" + nm; break
    p = ph.setdefault(name, [0, 0]); p[0] += n; p[1] += s
for name, (n, s) in sorted(ph.items(), key=lambda kv: -kv[1][0]):
    print("%6.1f%% inst %6.1f%% smp   %s" % (n / tot * 100, s / stot * 100, name))
