"""Summarise an ncu report per CUDA source line (stall samples + warp instructions executed).
usage: python profiles/ncu_by_line.py report.ncu-rep [top] [kernel-name-substring]   (needs -lineinfo and --import-source on)"""
import csv, io, os, subprocess, sys

rep = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 30
only = sys.argv[3] if len(sys.argv) > 3 else None
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
cur_file, hdr, agg, seen_fn = "?", None, [], None
for r in rows:
    if not r:
        continue
    if r[0] == "File Path":
        cur_file = os.path.basename(r[1]); continue
    if r[0] == "Function Name":
        if seen_fn is None and (only is None or only in r[1]):
            seen_fn = r[1]
        fn = r[1]; continue
    if r[0] == "Line No":
        hdr = {h: k for k, h in enumerate(r) if h not in ("Source",)}; src_col = 1; continue
    if hdr is None or r[0] == "" or fn != seen_fn:
        continue
    try:
        agg.append((cur_file, int(r[0]), r[src_col].strip()[:100], int(r[hdr["# Samples"]]), int(r[hdr["Instructions Executed"]])))
    except Exception:
        pass
print("kernel:", seen_fn)
ts, ti = sum(a[3] for a in agg), sum(a[4] for a in agg)
print("total samples %d, warp instructions %d" % (ts, ti))
for a in sorted(agg, key=lambda a: -a[3])[:top]:
    print("%5.1f%% smp %5.1f%% inst  %s:%d  %s" % (100.0 * a[3] / max(ts, 1), 100.0 * a[4] / max(ti, 1), a[0], a[1], a[2]))
