"""Summarise an .ncu-rep (read here, no GPU): key raw metrics + per-source-line hot spots."""
import csv, subprocess, sys, io, json
rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, vals = rows[0], rows[1], rows[2]
want = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "lts__t_bytes.sum", "lts__t_sector_hit_rate.pct",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
        "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_registers", "smsp__inst_executed.sum",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "l1tex__t_sector_hit_rate.pct", "smsp__thread_inst_executed_per_inst_executed.ratio",
        "l1tex__data_pipe_lsu_wavefronts.sum", "sm__inst_executed_pipe_lsu.sum", "l1tex__lsu_writeback_active.avg.pct_of_peak_sustained_elapsed",
        "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed", "lts__t_sectors_srcunit_tex_op_read.sum", "sm__cycles_elapsed.avg"]
out = {}
for i, h in enumerate(hdr):
    if h in want or ("issue_stalled" in h and h.endswith("per_issue_active.ratio")):
        out[h] = (vals[i], units[i])
for k, v in out.items():
    print(f"{k:90s} {v[0]:>18s} {v[1]}")
if len(sys.argv) > 2:
    src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--print-source", "cuda,sass", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(src)))
    cur = None; h2 = None; agg = []
    for r in rows:
        if not r: continue
        if r[0] == "File Path": cur = r[1].split("/")[-1]; continue
        if r[0] == "Function Name": continue
        if r[0] == "Line No": h2 = r; continue
        if r[0] != "" and h2:
            d = dict(zip(h2, r))
            try: smp = int(d["# Samples"] or 0); ex = int(d["Instructions Executed"] or 0)
            except Exception: continue
            agg.append((cur, int(r[0]), r[1].strip()[:110], smp, ex, d.get("L1 Wavefronts Shared", "0"), d.get("L1 Wavefronts Shared Ideal", "0")))
    ts = sum(a[3] for a in agg) or 1; te = sum(a[4] for a in agg) or 1
    print("total samples", ts, "instructions", te)
    thr = float(sys.argv[2])
    for a in sorted(agg, key=lambda a: (a[0], a[1])):
        if a[3] / ts > thr or a[4] / te > thr:
            print(f"{a[0][:14]:14s}:{a[1]:4d} smp {100*a[3]/ts:5.1f}% ex {100*a[4]/te:5.1f}% shw {a[5]:>11}/{a[6]:>11} | {a[2]}")
