"""Aggregate an ncu report's per-source-line instruction counts:  python tools/ncu_lines.py report.ncu-rep [topN]"""
import csv, subprocess, sys, io
rep = sys.argv[1]; top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True).stdout
rows = list(csv.reader(io.StringIO(raw))); hdr, units, data = rows[0], rows[1], rows[2:]
for w in ["gpu__time_duration.sum", "smsp__inst_executed.sum", "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "sm__warps_active.avg.pct_of_peak_sustained_active",
          "dram__bytes_read.sum", "dram__bytes_write.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__thread_inst_executed_per_inst_executed.ratio",
          "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct", "sm__cycles_elapsed.max", "launch__occupancy_limit_registers", "launch__waves_per_multiprocessor", "sm__inst_executed_pipe_fp64.sum",
          "smsp__inst_executed_pipe_fp64.sum", "l1tex__data_pipe_lsu_wavefronts_mem_lg.sum", "smsp__inst_executed_op_local_ld.sum", "smsp__inst_executed_op_local_st.sum"]:
    if w in hdr:
        i = hdr.index(w); print("%-62s %-10s %s" % (w, units[i], [r[i] for r in data]))
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True).stdout
cur = None; agg = {}; hdr = None
for r in csv.reader(io.StringIO(src)):
    if len(r) >= 2 and r[0] == "File Path": cur = r[1].split("/")[-1]; hdr = None; continue
    if len(r) >= 2 and r[0] == "Line No": hdr = r; ie = hdr.index("Instructions Executed"); it = hdr.index("Thread Instructions Executed"); isamp = hdr.index("# Samples"); continue
    if hdr and cur and len(r) == len(hdr) and r[2] == "-":
        try: ln = int(r[0]); v = float(r[ie] or 0); t = float(r[it] or 0); s = float(r[isamp] or 0)
        except Exception: continue
        a = agg.setdefault((cur, ln), [0, 0, 0, r[1]]); a[0] += v; a[1] += t; a[2] += s
tot = sum(a[0] for a in agg.values()); tots = sum(a[2] for a in agg.values())
print("total warp inst (all launches)", tot, "samples", tots)
byfile = {}
for (f, l), a in agg.items(): byfile[f] = byfile.get(f, 0) + a[0]
print({k: round(100 * v / tot, 1) for k, v in byfile.items()})
for k, a in sorted(agg.items(), key=lambda kv: -kv[1][2])[:top]:
    print("%-14s %4d inst %5.2f%% smp %5.2f%% thr/inst %4.1f | %s" % (k[0], k[1], 100 * a[0] / tot, 100 * a[2] / max(tots, 1), a[1] / max(a[0], 1), a[3].strip()[:95]))
