#!/usr/bin/env python
"""Top source lines by executed warp-instructions from an ncu report (needs -lineinfo + --import-source on).
usage: python profiles/hot_lines.py <report.ncu-rep> [steps] [topN]"""
import csv, io, os, subprocess, sys
rep = sys.argv[1]; steps = float(sys.argv[2]) if len(sys.argv) > 2 else 16384.0; top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--print-source", "cuda,sass", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
cur, hdr, res = None, None, []
for r in rows:
    if not r: continue
    if r[0] == "File Path": cur = os.path.basename(r[1]); continue
    if r[0] == "Line No": hdr = {h: i for i, h in enumerate(r)}; ii = r.index("Instructions Executed"); si = r.index("# Samples"); continue
    if hdr and r[0] and r[0].isdigit():
        try: res.append((float(r[ii] or 0), float(r[si] or 0), cur, int(r[0]), r[1].strip()))
        except (ValueError, IndexError): pass
tot = sum(x[0] for x in res); ts = sum(x[1] for x in res)
print("total attributed warp-instructions/step: %.0f" % (tot / steps))
for i, s, f, ln, src in sorted(res, key=lambda x: -x[0])[:top]:
    print("%7.0f/step %5.1f%% smp %4.1f%%  %s:%d  %s" % (i / steps, 100 * i / tot, 100 * s / max(ts, 1), f, ln, src[:110]))
