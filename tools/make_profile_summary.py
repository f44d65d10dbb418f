"""Turn an ncu report (+ optional launch-list csv) into the tracked summaries under profiles/.
usage: python tools/make_profile_summary.py <k2.ncu-rep> <tag> [launches.csv]"""
import csv, io, json, os, subprocess, sys
rep, tag = sys.argv[1], sys.argv[2]
ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
out_dir = os.path.join(ROOT, "profiles")
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, vals = rows[0], rows[1], rows[2]
m = {h: (vals[i], units[i]) for i, h in enumerate(hdr)}
def f(name):
    v, u = m[name]
    v = float(v.replace(",", ""))
    scale = {"Mbyte": 1e6, "Kbyte": 1e3, "Gbyte": 1e9, "byte": 1, "ms": 1e-3, "us": 1e-6, "s": 1, "ns": 1e-9}.get(u, 1)
    return v * scale
summary = {
    "kernel": m["Kernel Name"][0] if "Kernel Name" in m else "spcos_dense_kernel",
    "source_report": os.path.basename(rep),
    "duration_s_under_ncu": f("gpu__time_duration.sum"),
    "dram_bytes_read": f("dram__bytes_read.sum"), "dram_bytes_write": f("dram__bytes_write.sum"),
    "dram_bytes_per_launch": f("dram__bytes_read.sum") + f("dram__bytes_write.sum"),
    "l2_sector_hit_rate_pct": f("lts__t_sector_hit_rate.pct"),
    "lsu_data_pipe_wavefronts_pct_of_peak": f("l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed"),
    "smem_wavefronts": f("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum"),
    "smem_bank_conflict_wavefronts": f("l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum"),
    "issue_active_pct": f("smsp__issue_active.avg.pct_of_peak_sustained_active"),
    "warps_active_pct": f("sm__warps_active.avg.pct_of_peak_sustained_active"),
    "inst_executed": f("smsp__inst_executed.sum"),
    "registers_per_thread": f("launch__registers_per_thread"),
    "grid_size": f("launch__grid_size"), "block_size": f("launch__block_size"),
}
json.dump(summary, open(os.path.join(out_dir, "k2_ncu_summary.json"), "w"), indent=1)
with open(os.path.join(out_dir, f"k2_{tag}_metrics.txt"), "w") as fo:
    fo.write(subprocess.run([sys.executable, os.path.join(ROOT, "tools", "ncu_summary.py"), rep, "0.01"], capture_output=True, text=True).stdout)
if len(sys.argv) > 3:
    # launch list: aggregate per kernel name
    agg = {}
    tot = 0.0
    with open(sys.argv[3]) as fh:
        lines = [l for l in fh if not l.startswith("==")]
    rd = csv.DictReader(lines)
    for r in rd:
        if r.get("Metric Name") != "gpu__time_duration.sum":
            continue
        v = float(r["Metric Value"].replace(",", "")); u = r["Metric Unit"]
        v *= {"ns": 1e-9, "us": 1e-6, "ms": 1e-3, "s": 1, "usecond": 1e-6, "nsecond": 1e-9, "msecond": 1e-3}.get(u, 1e-9)
        k = r["Kernel Name"].split("(")[0]
        a = agg.setdefault(k, [0, 0.0]); a[0] += 1; a[1] += v; tot += v
    with open(os.path.join(out_dir, f"launches_{tag}_summary.txt"), "w") as fo:
        fo.write(f"# per-kernel totals from {os.path.basename(sys.argv[3])} (ncu --metrics gpu__time_duration.sum, cold-cache serialised: compare SHARES)\n")
        for k, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            fo.write(f"{100 * t / tot:6.2f}%  {t * 1e3:10.3f} ms  x{c:4d}  {k}\n")
print(json.dumps(summary, indent=1))
