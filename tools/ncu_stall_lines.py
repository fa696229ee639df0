"""Source lines of a kernel by warp-stall SAMPLES (not instructions), with the top stall reasons of each line — the view
that shows where warps WAIT (copy-engine queue, barriers, scoreboards). usage: python tools/ncu_stall_lines.py report.ncu-rep [kernel-regex] [top]"""
import csv, io, os, subprocess, sys
rep = sys.argv[1]; kname = sys.argv[2] if len(sys.argv) > 2 else "sf_rollout_kernel"; top = int(sys.argv[3]) if len(sys.argv) > 3 else 30
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass", "-k", "regex:" + kname], capture_output=True, text=True).stdout
hdr = cur = stalls = None
agg, tot = {}, 0.0
for r in csv.reader(io.StringIO(out)):
    if not r: continue
    if r[0] == "File Path": cur = os.path.basename(r[1]); continue
    if r[0] == "Line No":
        hdr = r; stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]; continue
    if hdr is None or cur is None or not r[0].isdigit(): continue
    def g(name):  # index from the end: an unescaped quote in the source text can split the early fields
        try: return float(r[len(r) - (len(hdr) - hdr.index(name))])
        except ValueError: return 0.0
    a = agg.setdefault((cur, int(r[0])), {"smp": 0.0, "src": r[1]})
    a["smp"] += g("# Samples"); tot += g("# Samples")
    for s in stalls: a[s] = a.get(s, 0.0) + g(s)
print("samples %d (a SASS instruction inlined from several lines is counted under each of them)" % tot)
by_reason = {s: sum(a.get(s, 0.0) for a in agg.values()) for s in stalls}
print("by reason: " + "  ".join("%s %.1f%%" % (s[6:], v / tot * 100) for s, v in sorted(by_reason.items(), key=lambda kv: -kv[1])[:10]))
for k, a in sorted(agg.items(), key=lambda kv: -kv[1]["smp"])[:top]:
    t3 = sorted(((a[s], s) for s in stalls if a.get(s, 0) > 0), reverse=True)[:3]
    print("%5.2f%% %-16s:%-5d %-72s %s" % (a["smp"] / tot * 100, k[0], k[1], a["src"].strip()[:72], " ".join("%s=%.0f" % (s[6:], v) for v, s in t3)))
