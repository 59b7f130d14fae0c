#!/usr/bin/env python
"""Divergence cost by source line: joins the per-instruction samples of an ncu report (rollout_kernel_quadruped) with
`nvdisasm -g` line info of the built cubin.  For every BSYNC (reconvergence point) the samples of the BSSY..BSYNC region
cannot be separated cheaply, so this reports samples ON the BSYNC / BRA instructions by source line.
usage: python profiles/bsync_lines.py <report.ncu-rep> <all.sass from nvdisasm -g -c> [top]"""
import collections, csv, io, re, subprocess, sys
rep, sass = sys.argv[1], sys.argv[2]; top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
txt = open(sass).read()
i = txt.index(".text.rollout_kernel_quadruped:")
j = txt.find("//--------------------- .text.", i + 10)
body = txt[i:j if j > 0 else len(txt)]
line_of = {}
cur = None
for l in body.splitlines():
    m = re.search(r'//## File "([^"]+)", line (\d+)', l)
    if m:
        cur = (m.group(1).split("/")[-1], int(m.group(2))); continue
    m = re.match(r"\s+/\*([0-9a-f]{4,6})\*/\s+(.*)", l)
    if m: line_of[int(m.group(1), 16)] = cur
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr = rows[1]; ix = {h: k for k, h in enumerate(hdr)}
data = rows[2:]
a0 = int(data[0][0], 16)
agg = collections.defaultdict(lambda: [0.0, 0.0, 0.0])
tot = 0.0
for r in data:
    src = r[ix["Source"]]
    s = float(r[ix["# Samples"]] or 0); tot += s
    if "BSYNC" in src or re.search(r"\bBRA\b", src) or "BSSY" in src:
        k = line_of.get(int(r[0], 16) - a0)
        agg[k][0] += s; agg[k][1] += float(r[ix["Instructions Executed"]] or 0); agg[k][2] += 1
print("total samples %d; on BSSY/BRA/BSYNC: %.1f%%" % (tot, 100 * sum(v[0] for v in agg.values()) / tot))
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
    print("%6.0f smp (%.2f%%)  %7.0f exec/step  %2d instr  %s" % (v[0], 100 * v[0] / tot, v[1] / 16384.0, v[2], k))
