"""Copy the judged artefacts from gpurun_out/ (scratch) to profiles/ (tracked): bench lines, ncu launch list, key metrics,
DRAM traffic of the rollout kernel, source hot spots. usage: python tools/export_profiles.py [round-tag]"""
import csv, io, json, os, shutil, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag = sys.argv[1] if len(sys.argv) > 1 else "r01"
K = int(sys.argv[2]) if len(sys.argv) > 2 else 16   # steps per launch of the profiled command
G, P = os.path.join(ROOT, "gpurun_out"), os.path.join(ROOT, "profiles")
os.makedirs(P, exist_ok=True)
for src, dst in (("bench_%s.json" % tag, "%s_bench_n1.json" % tag), ("bench_ref_%s.json" % tag, "%s_bench_reference_arm.json" % tag),
                 ("launches_%s.csv" % tag, "%s_launches_bench.csv" % tag)):
    if os.path.exists(os.path.join(G, src)): shutil.copy(os.path.join(G, src), os.path.join(P, dst))
rep = os.path.join(G, "prof_%s_bench.ncu-rep" % tag)
out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
r = list(csv.reader(io.StringIO(out))); h, u, d = r[0], r[1], r[2]
def val(k):
    v = float(d[h.index(k)].replace(",", "")); unit = u[h.index(k)]
    return v * {"Mbyte": 1e6, "Kbyte": 1e3, "Gbyte": 1e9, "byte": 1, "us": 1e-6, "ms": 1e-3, "ns": 1e-9, "s": 1}.get(unit, 1)
keys = [k for k in h if any(t in k for t in ("dram__bytes", "gpu__time_duration", "launch__", "sm__inst_executed.avg.per_cycle_elapsed", "smsp__inst_executed.sum",
        "sm__warps_active", "issue_stalled", "thread_inst_executed_per_inst", "sm__pipe_tensor", "gpu__dram_throughput", "smsp__issue_active", "l1tex__t_sector_hit_rate", "lts__t_sector_hit_rate"))]
with open(os.path.join(P, "%s_ncu_key_metrics_rollout_kernel.txt" % tag), "w") as f:
    f.write("# ncu --set full --clock-control none -k regex:sf_rollout_kernel, launch of `python bench.py --steps %d ...` (4096 envs x %d steps)\n" % (K, K))
    for k in sorted(keys): f.write("%-90s %-16s %s\n" % (k, u[h.index(k)], d[h.index(k)]))
steps = 4096 * K
traffic = {"source": "profiles/%s_ncu_key_metrics_rollout_kernel.txt" % tag, "env_steps_in_launch": steps,
           "dram_bytes_read": val("dram__bytes_read.sum"), "dram_bytes_write": val("dram__bytes_write.sum"),
           "dram_bytes_per_env_step": (val("dram__bytes_read.sum") + val("dram__bytes_write.sum")) / steps,
           "launch_seconds_under_ncu": val("gpu__time_duration.sum"), "warp_instructions_per_env_step": val("smsp__inst_executed.sum") / steps}
json.dump(traffic, open(os.path.join(P, "%s_ncu_traffic.json" % tag), "w"), indent=1)
print(json.dumps(traffic, indent=1))
det = subprocess.run(["ncu", "-i", rep, "--page", "details"], capture_output=True, text=True).stdout
open(os.path.join(P, "%s_ncu_full_details_rollout_kernel.txt" % tag), "w").write(det)
hot = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "ncu_src.py"), rep, "sf_rollout_kernel", str(steps)], capture_output=True, text=True, env=dict(os.environ, TOP="60")).stdout
open(os.path.join(P, "%s_ncu_source_hotspots.txt" % tag), "w").write(hot)
