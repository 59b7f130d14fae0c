#!/usr/bin/env python
"""Executed code footprint of the rollout kernel from an ncu report: per device function, how many distinct SASS
instructions are executed at least once per env-step, and how many sit in the Newton loop (> 3 executions per env-step).
usage: python profiles/footprint.py <report.ncu-rep> [steps_per_launch]   (MJPC_B200_SO = the library the report was taken from)"""
import bisect, collections, csv, io, os, re, subprocess, sys
rep = sys.argv[1]; steps = float(sys.argv[2]) if len(sys.argv) > 2 else 16384.0
so = os.environ.get("MJPC_B200_SO") or os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "mujoco_mpc_b200", "csrc", "libmjpc_b200.so")
rows = list(csv.reader(io.StringIO(subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout)))
hdr = rows[1]; ix = {h: i for i, h in enumerate(hdr)}; data = rows[2:]
addr = [int(r[0], 16) for r in data]; base = addr[0]
ex = [float(r[ix["Instructions Executed"]] or 0) / steps for r in data]
# function boundaries: CALL targets inside the kernel; names from the ELF symbol table (local symbols "$kernel$mangled")
targets = sorted({int(re.search(r"0x[0-9a-f]+", r[ix["Source"]]).group(0), 16) for r in data if "CALL" in r[ix["Source"]] and "0x" in r[ix["Source"]]})
keys = sorted(set([base] + targets))
kname = "rollout_kernel_quadruped"
sym = {}
for line in subprocess.run(["cuobjdump", "-elf", so], capture_output=True, text=True).stdout.splitlines():
    mm = re.match(r"\s*0x[0-9a-f]+\s+(0x[0-9a-f]+)\s+0x[0-9a-f]+\s+0x\d+\s+\d+\s+0x[0-9a-f]+\s+\$" + re.escape(kname) + r"\$(\S+)", line)
    if mm: sym[int(mm.group(1), 16)] = re.sub(r"_ZN8mjpc_dev\d+|I(NS_|Li).*|E(RNS_3CtxE|Pf|RKNS).*", "", mm.group(2))
fn_at = [(k, "body" if k == base else sym.get(k - base, hex(k - base))) for k in keys]
agg = collections.OrderedDict()
for a, e in zip(addr, ex):
    i = bisect.bisect_right(keys, a) - 1
    fn = fn_at[i][1] if i >= 0 else "body"
    d = agg.setdefault(fn, [0, 0, 0, 0.0])
    d[0] += 1; d[1] += e > 1e-4; d[2] += e > 3.0; d[3] += e
print("%-28s %8s %10s %12s %12s" % ("function", "static", "executed", "in Newton", "instr/step"))
tot = [0, 0, 0, 0.0]
for fn, d in sorted(agg.items(), key=lambda kv: -kv[1][2]):
    print("%-28s %6.1f KB %7.1f KB %9.1f KB %12.0f" % (fn, d[0] / 64, d[1] / 64, d[2] / 64, d[3]))
    for k in range(4): tot[k] += d[k]
print("%-28s %6.1f KB %7.1f KB %9.1f KB %12.0f" % ("total", tot[0] / 64, tot[1] / 64, tot[2] / 64, tot[3]))
