"""Per-file time share (stall samples) and stall reasons of an ncu report: python tools/ncu_stalls.py report.ncu-rep"""
import csv, subprocess, io, sys
rep = sys.argv[1]
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True).stdout
cur = None; hdr = None; agg = {}; stall_tot = {}
for r in csv.reader(io.StringIO(src)):
    if len(r) >= 2 and r[0] == "File Path": cur = r[1].split("/")[-1]; hdr = None; continue
    if len(r) >= 2 and r[0] == "Line No":
        hdr = r; isamp = hdr.index("# Samples"); ie = hdr.index("Instructions Executed")
        stalls = [(i, h) for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
        continue
    if hdr and cur and len(r) == len(hdr) and r[2] == "-":
        try: int(r[0])
        except Exception: continue
        s = float(r[isamp] or 0); v = float(r[ie] or 0)
        a = agg.setdefault(cur, [0, 0, {}]); a[0] += s; a[1] += v
        for i, h in stalls:
            x = float(r[i] or 0); a[2][h] = a[2].get(h, 0) + x; stall_tot[h] = stall_tot.get(h, 0) + x
tot = sum(a[0] for a in agg.values()); toti = sum(a[1] for a in agg.values())
for f, a in sorted(agg.items(), key=lambda kv: -kv[1][0]):
    top = sorted(a[2].items(), key=lambda kv: -kv[1])[:4]
    print("%-22s time %5.1f%% inst %5.1f%% | %s" % (f, 100 * a[0] / tot, 100 * a[1] / toti, ", ".join("%s %.0f%%" % (h[6:], 100 * v / max(a[0], 1)) for h, v in top)))
print("overall stalls:", ", ".join("%s %.1f%%" % (h[6:], 100 * v / tot) for h, v in sorted(stall_tot.items(), key=lambda kv: -kv[1])[:8]))
