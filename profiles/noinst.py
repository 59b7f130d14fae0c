#!/usr/bin/env python
"""Where the instruction-fetch stalls are: per device function, stall_no_inst samples vs executed instructions, and the
instructions with the most no_inst samples (branch / call targets whose line is not resident).
usage: python profiles/noinst.py <report.ncu-rep> [top]"""
import bisect, collections, csv, io, os, re, subprocess, sys
rep = sys.argv[1]; top = int(sys.argv[2]) if len(sys.argv) > 2 else 25
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr = rows[1]; ix = {h: i for i, h in enumerate(hdr)}
data = rows[2:]
addr = [int(r[0], 16) for r in data]
f = lambda r, k: float(r[ix[k]] or 0)
tot_s = sum(f(r, "# Samples") for r in data); tot_n = sum(f(r, "stall_no_inst") for r in data)
print("samples %d, no_inst %d (%.1f%%)" % (tot_s, tot_n, 100 * tot_n / tot_s))
# consecutive-address runs: how far apart are the no_inst hot spots (static code between them)
hot = sorted(range(len(data)), key=lambda i: -f(data[i], "stall_no_inst"))[:top]
for i in hot:
    prev = data[i - 1][ix["Source"]].strip() if i else ""
    print("%6.0f no_inst  %6.0f exec  +0x%05x  %-60s  prev: %s" % (f(data[i], "stall_no_inst"), f(data[i], "Instructions Executed"),
          addr[i] - addr[0], data[i][ix["Source"]].strip()[:60], prev[:50]))
# fraction of no_inst samples that sit on instructions following a taken-branch boundary (prev is BRA/CALL/RET/EXIT/BSYNC) 
ctl = re.compile(r"\b(BRA|CALL|RET|EXIT|BRX|JMP|BSYNC|WARPSYNC|BSSY)\b")
after_ctl = sum(f(data[i], "stall_no_inst") for i in range(1, len(data)) if ctl.search(data[i - 1][ix["Source"]]))
print("no_inst samples right after a control instruction: %.1f%%" % (100 * after_ctl / max(tot_n, 1)))
# 128-byte lines: executed-instruction-weighted count of distinct lines touched
lines = collections.Counter()
for a, r in zip(addr, data):
    if f(r, "Instructions Executed") > 0: lines[a >> 7] += 1
print("distinct 128 B instruction lines executed: %d (%.0f KB)" % (len(lines), len(lines) * 128 / 1024))
# aggregate no_inst samples by opcode of the instruction they sit on
byop = collections.Counter(); byop_all = collections.Counter()
for r in data:
    src = r[ix["Source"]].strip()
    op = re.sub(r"^@!?U?P\d+\s+", "", src).split()[0].split(".")[0] if src else "?"
    byop[op] += f(r, "stall_no_inst"); byop_all[op] += f(r, "# Samples")
print("no_inst samples by opcode:", [(k, int(v), "%.0f%%" % (100 * v / tot_n)) for k, v in byop.most_common(8)])
print("all samples by opcode:", [(k, int(v), "%.0f%%" % (100 * v / tot_s)) for k, v in byop_all.most_common(10)])
nb = sum(1 for r in data if "BSYNC" in r[ix["Source"]] and f(r, "Instructions Executed") > 0)
eb = sum(f(r, "Instructions Executed") for r in data if "BSYNC" in r[ix["Source"]])
print("BSYNC: %d static executed, %.0f executions per env-step" % (nb, eb / 16384.0))
