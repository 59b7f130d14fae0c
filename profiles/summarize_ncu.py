#!/usr/bin/env python
"""Summarise an ncu report of the rollout kernel: headline counters + executed instructions per device function.

usage: python profiles/summarize_ncu.py gpurun_out/<report>.ncu-rep [steps_per_launch] > profiles/<name>.txt
Function names are recovered by matching SASS sizes against `cuobjdump -sass` of the built library.
"""
import bisect, collections, csv, io, os, re, subprocess, sys

rep = sys.argv[1]
steps = float(sys.argv[2]) if len(sys.argv) > 2 else 16384.0
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
so = os.path.join(root, "mujoco_mpc_b200", "csrc", "libmjpc_b200.so")


def run(cmd):
    return subprocess.run(cmd, capture_output=True, text=True).stdout


raw = list(csv.reader(io.StringIO(run(["ncu", "-i", rep, "--page", "raw", "--csv"]))))
hdr, vals = raw[0], raw[-1]
want = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "launch__registers_per_thread",
        "launch__shared_mem_per_block_dynamic", "launch__grid_size", "launch__block_size", "smsp__inst_executed.sum",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "sass__inst_executed_local_loads", "sass__inst_executed_shared_loads", "sass__inst_executed_shared_stores",
        "sm__cycles_elapsed.max", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active"]
print("== headline counters (%s)" % os.path.basename(rep))
units = raw[1] if len(raw) > 2 else [""] * len(hdr)
for h, u, v in zip(hdr, units, vals):
    if h in want:
        print("%-70s %s %s" % (h, v, u))

src = list(csv.reader(io.StringIO(run(["ncu", "-i", rep, "--page", "source", "--csv"]))))
h2 = src[1]; ix = {h: i for i, h in enumerate(h2)}
data = src[2:]
addr = [int(r[0], 16) for r in data]
inst = [float(r[ix["Instructions Executed"]] or 0) for r in data]
samp = [float(r[ix["# Samples"]] or 0) for r in data]
targets = sorted({int(re.search(r"0x[0-9a-f]+", r[ix["Source"]]).group(0), 16) for r in data
                  if "CALL" in r[ix["Source"]] and "0x" in r[ix["Source"]]})
starts = sorted(set([addr[0]] + targets))
agg = collections.defaultdict(lambda: [0, 0, 0]); calls = collections.Counter()
for r in data:
    if "CALL" in r[ix["Source"]] and "0x" in r[ix["Source"]]:
        calls[int(re.search(r"0x[0-9a-f]+", r[ix["Source"]]).group(0), 16)] += float(r[ix["Instructions Executed"]] or 0)
for a, i, s in zip(addr, inst, samp):
    k = starts[bisect.bisect_right(starts, a) - 1]
    agg[k][0] += i; agg[k][1] += s; agg[k][2] += 1
# names from the ELF symbol table: device functions are local symbols "$kernel$mangled" whose value is the
# offset from the kernel's first instruction
sym = {}
kname = vals[hdr.index("Kernel Name")].split("(")[0] if "Kernel Name" in hdr else "rollout_kernel"
for line in run(["cuobjdump", "-elf", so]).splitlines():
    mm = re.match(r"\s*0x[0-9a-f]+\s+(0x[0-9a-f]+)\s+0x[0-9a-f]+\s+0x\d+\s+\d+\s+0x[0-9a-f]+\s+\$" + re.escape(kname) + r"\$(\S+)", line)
    if mm:
        sym[int(mm.group(1), 16)] = re.sub(r"_ZN8mjpc_dev\d+|I(NS_|Li).*|E(RNS_3CtxE|Pf|RKNS).*", "", mm.group(2))


def name_of(k):
    off = k - addr[0]
    return kname + " (body)" if off == 0 else sym.get(off, hex(off))


tot, ts = sum(inst), sum(samp)
print("\n== executed warp-instructions per simulated env-step: %.0f (total %.3e over %.0f steps)" % (tot / steps, tot, steps))
print("%-38s %12s %7s %9s %11s %10s" % ("function", "inst/step", "inst%", "samples%", "calls/step", "inst/call"))
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][0]):
    if v[0] == 0: continue
    print("%-38s %12.0f %6.1f%% %8.1f%% %11.1f %10.0f" % (name_of(k)[:38], v[0] / steps, 100 * v[0] / tot, 100 * v[1] / max(ts, 1),
                                                       calls[k] / steps, v[0] / max(calls[k], 1)))
# stall reasons (sampling columns start with 'stall_')
st = collections.Counter()
for r in data:
    for h, i in ix.items():
        if h.startswith("stall_") and i < len(r) and r[i]:
            try: st[h] += float(r[i])
            except ValueError: pass
if st:
    tt = sum(st.values())
    print("\n== warp stall sampling (all samples)")
    for h, v in st.most_common(10):
        print("%-28s %5.1f%%" % (h, 100 * v / tt))
