import csv, io, subprocess, sys, re, collections, bisect
rep=sys.argv[1]
import os; so=os.environ.get('MJPC_B200_SO') or os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'mujoco_mpc_b200', 'csrc', 'libmjpc_b200.so')
elf=subprocess.run(["cuobjdump","-elf",so],capture_output=True,text=True).stdout
fns=[]
for l in elf.splitlines():
    m=re.match(r"\s*0x[0-9a-f]+\s+(0x[0-9a-f]+)\s+(0x[0-9a-f]+)\s+.*\$rollout_kernel_quadruped\$_ZN8mjpc_dev\d+(\w+?)I", l)
    if m: fns.append((int(m.group(1),16), int(m.group(2),16), m.group(3)))
fns.sort()
out=subprocess.run(["ncu","-i",rep,"--page","source","--csv"],capture_output=True,text=True).stdout
rows=list(csv.reader(io.StringIO(out))); hdr=rows[1]; ix={h:i for i,h in enumerate(hdr)}; data=rows[2:]
a0=int(data[0][0],16)
starts=[f[0] for f in fns]
agg=collections.defaultdict(lambda: collections.Counter())
keys=["# Samples","stall_no_inst","stall_wait","stall_short_sb","stall_selected","stall_branch_resolving","stall_barrier","stall_math_pipe_throttle","stall_dispatch","stall_long_sb","stall_lg_throttle","stall_mio_throttle"]
keys=[k for k in keys if k in ix]
for r in data:
    off=int(r[0],16)-a0
    i=bisect.bisect_right(starts,off)-1
    name=fns[i][2] if i>=0 and off<fns[i][0]+fns[i][1] else "body"
    for k in keys: agg[name][k]+=float(r[ix[k]] or 0)
    agg[name]["inst"]+=float(r[ix["Instructions Executed"]] or 0)
tot=sum(v["# Samples"] for v in agg.values())
print("%-22s %7s %6s %6s | %s"%("fn","inst/st","smp%","CPIrel"," ".join(k.replace('stall_','')[:8].rjust(8) for k in keys[1:])))
ti=sum(v["inst"] for v in agg.values())
for n,v in sorted(agg.items(), key=lambda kv:-kv[1]["# Samples"]):
    s=v["# Samples"]
    print("%-22s %7.0f %6.1f %6.2f | %s"%(n, v["inst"]/16384, 100*s/tot, (s/tot)/(v["inst"]/ti) if v["inst"] else 0, " ".join(("%7.1f%%"%(100*v[k]/s) if s else "").rjust(8) for k in keys[1:])))
