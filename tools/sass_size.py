"""Static SASS size of a kernel by source function (from nvdisasm line info). usage: python tools/sass_size.py [kernel-substring]"""
import os, re, subprocess, sys, tempfile
kname = sys.argv[1] if len(sys.argv) > 1 else "sf_rollout_kernelILb1"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tmp = tempfile.mkdtemp()
subprocess.run("cd %s && cuobjdump -xelf all %s/spacefortress_b200/libsf_b200.so >/dev/null 2>&1 && nvdisasm -g -c sf_kernels.sm_100a.cubin > dis.txt 2>/dev/null" % (tmp, ROOT), shell=True)
lines = open(tmp + "/dis.txt").read().split("\n")
start = [i for i, l in enumerate(lines) if l.startswith(".text.") and kname in l][0]
end = [i for i, l in enumerate(lines) if l.startswith("//---------------------") and i > start]
end = end[0] if end else len(lines)
cur = None; cnt = {}
for l in lines[start:end]:
    m = re.search(r'//## File "([^"]+)", line (\d+)', l)
    if m: cur = (os.path.basename(m.group(1)), int(m.group(2))); continue
    if re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+", l): cnt[cur] = cnt.get(cur, 0) + 1
tot = sum(cnt.values())
print("total SASS instructions", tot, "=", tot * 16 // 1024, "KB")
ph = {}
for fn in ("sf_render.cuh", "sf_step.cuh", "sf_kernels.cu", "sf_geom.h"):
    src = open(os.path.join(ROOT, "spacefortress_b200/csrc", fn)).read().split("\n")
    marks = [(i + 1, re.search(r"(sf_\w+)\s*\(", l).group(1)) for i, l in enumerate(src)
             if re.match(r"(template.*)?(__device__|__global__|SF_HD|static)", l) and re.search(r"(sf_\w+)\s*\(", l)]
    marks.append((len(src) + 1, "end"))
    for (f, ln), n in cnt.items() if True else []:
        if f != fn: continue
        name = fn + ":?"
        for (s0, nm), (s1, _) in zip(marks, marks[1:]):
            if s0 <= ln < s1: name = fn.split(".")[0][3:] + ":" + nm; break
        ph[name] = ph.get(name, 0) + n
for k, n in cnt.items():
    if k is None or k[0] not in ("sf_render.cuh", "sf_step.cuh", "sf_kernels.cu", "sf_geom.h"):
        nm = k[0] if k else "none"; ph[nm] = ph.get(nm, 0) + n
for name, n in sorted(ph.items(), key=lambda kv: -kv[1]):
    if n >= 20: print("%6d  %5.1f%%  %s" % (n, n / tot * 100, name))
