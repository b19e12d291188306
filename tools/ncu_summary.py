"""Markdown summary of one ncu --set full report for profiles/:  python tools/ncu_summary.py rep.ncu-rep "title" > profiles/x.md"""
import csv, io, subprocess, sys
rep, title = sys.argv[1], (sys.argv[2] if len(sys.argv) > 2 else sys.argv[1])
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True).stdout
rows = list(csv.reader(io.StringIO(raw))); hdr, units, data = rows[0], rows[1], rows[2:]
want = ["Kernel Name", "launch__grid_size", "launch__block_size", "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic", "launch__waves_per_multiprocessor",
        "gpu__time_duration.sum", "sm__cycles_elapsed.max", "smsp__inst_executed.sum", "sm__inst_executed.avg.per_cycle_elapsed", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "smsp__thread_inst_executed_per_inst_executed.ratio", "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__warps_active.avg.per_cycle_active",
        "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "dram__bytes_read.sum", "dram__bytes_write.sum", "lts__t_bytes.sum",
        "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct", "sm__throughput.avg.pct_of_peak_sustained_elapsed", "gpu__compute_memory_throughput.avg.pct_of_peak_sustained_elapsed",
        # divergence / branch efficiency and LOCAL-memory traffic (the share of the DRAM bytes that is not state: spilled bodies / solver groups)
        "smsp__branch_targets_threads_divergent", "smsp__sass_branch_targets.sum", "smsp__sass_branch_targets_threads_divergent.sum", "smsp__sass_branch_targets_threads_uniform.sum",
        "sass__inst_executed_local_loads", "sass__inst_executed_local_stores", "sass__inst_executed_global_loads", "sass__inst_executed_global_stores",
        "l1tex__t_sectors_pipe_lsu_mem_local_op_ld.sum", "l1tex__t_sectors_pipe_lsu_mem_local_op_st.sum", "l1tex__t_sector_pipe_lsu_mem_local_op_ld_hit_rate.pct",
        "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum", "l1tex__t_sectors_pipe_lsu_mem_global_op_st.sum", "l1tex__t_sector_pipe_lsu_mem_global_op_ld_hit_rate.pct",
        "smsp__inst_executed_op_local_ld.sum", "smsp__inst_executed_op_local_st.sum", "smsp__sass_inst_executed_op_shared_ld.sum", "smsp__sass_inst_executed_op_shared_st.sum",
        "smsp__sass_inst_executed_op_global_ld.sum", "smsp__sass_inst_executed_op_global_st.sum", "derived__smsp__sass_thread_inst_executed_op_ffma_pred_on_x2", "smsp__sass_thread_inst_executed_op_ffma_pred_on.sum",
        "smsp__sass_thread_inst_executed_op_fadd_pred_on.sum", "smsp__sass_thread_inst_executed_op_fmul_pred_on.sum"]
print("# %s\n" % title)
print("Source: `%s` (ncu --set full --clock-control none --import-source on, one launch; times under ncu are serialised / cold and are NOT bench values).\n" % rep.split("/")[-1])
print("| metric | value | unit |\n|---|---|---|")
for w in want:
    if w in hdr:
        i = hdr.index(w); print("| %s | %s | %s |" % (w, data[0][i], units[i]))
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True).stdout
cur = None; h = None; agg = {}; lines = {}; st_tot = {}
for r in csv.reader(io.StringIO(src)):
    if len(r) >= 2 and r[0] == "File Path": cur = r[1].split("/")[-1]; h = None; continue
    if len(r) >= 2 and r[0] == "Line No":
        h = r; isamp = h.index("# Samples"); ie = h.index("Instructions Executed"); stalls = [(i, x) for i, x in enumerate(h) if x.startswith("stall_") and "Not Issued" not in x]; continue
    if h and cur and len(r) == len(h) and r[2] == "-":
        try: ln = int(r[0])
        except Exception: continue
        s = float(r[isamp] or 0); v = float(r[ie] or 0)
        a = agg.setdefault(cur, [0, 0, {}]); a[0] += s; a[1] += v
        l = lines.setdefault((cur, ln), [0, 0, r[1]]); l[0] += s; l[1] += v
        for i, x in stalls:
            y = float(r[i] or 0); a[2][x] = a[2].get(x, 0) + y; st_tot[x] = st_tot.get(x, 0) + y
tot = sum(a[0] for a in agg.values()) or 1; toti = sum(a[1] for a in agg.values()) or 1
print("\n## Time (warp-state samples) and executed warp instructions per source file\n\n| file | time % | inst % | top stall reasons |\n|---|---|---|---|")
for f, a in sorted(agg.items(), key=lambda kv: -kv[1][0]):
    top = sorted(a[2].items(), key=lambda kv: -kv[1])[:4]
    print("| %s | %.1f | %.1f | %s |" % (f, 100 * a[0] / tot, 100 * a[1] / toti, ", ".join("%s %.0f%%" % (x[6:], 100 * v / max(a[0], 1)) for x, v in top)))
print("\nOverall stall reasons: " + ", ".join("%s %.1f%%" % (x[6:], 100 * v / tot) for x, v in sorted(st_tot.items(), key=lambda kv: -kv[1])[:8]))
print("\n## Hottest source lines (by samples)\n\n| file:line | time % | inst % | source |\n|---|---|---|---|")
for (f, ln), l in sorted(lines.items(), key=lambda kv: -kv[1][0])[:25]:
    print("| %s:%d | %.2f | %.2f | `%s` |" % (f, ln, 100 * l[0] / tot, 100 * l[1] / toti, l[2].strip()[:110].replace("|", "\\|")))
