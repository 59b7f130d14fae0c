#!/usr/bin/env python
"""Print the SASS of one device function inside a kernel of the built library.
usage: python profiles/fn_sass.py <kernel> <device-function-substring> [max_lines]"""
import os, re, subprocess, sys
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
so = os.path.join(root, "mujoco_mpc_b200", "csrc", "libmjpc_b200.so")
kern, fn = sys.argv[1], sys.argv[2]
maxl = int(sys.argv[3]) if len(sys.argv) > 3 else 100000
elf = subprocess.run(["cuobjdump", "-elf", so], capture_output=True, text=True).stdout
off = size = None
for l in elf.splitlines():
    m = re.match(r"\s*0x[0-9a-f]+\s+(0x[0-9a-f]+)\s+(0x[0-9a-f]+)\s+0x\d+\s+\d+\s+0x[0-9a-f]+\s+\$" + re.escape(kern) + r"\$(\S+)", l)
    if m and fn in m.group(3):
        off, size = int(m.group(1), 16), int(m.group(2), 16)
        break
if off is None:
    sys.exit("function not found")
txt = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True).stdout
i = txt.index("Function : " + kern + "\n")
j = txt.find("Function :", i + 10)
body = txt[i:j if j > 0 else len(txt)]
n = 0
ops = {}
for l in body.splitlines():
    m = re.match(r"\s+/\*([0-9a-f]{4,6})\*/\s+(.*?)\s*/\*", l)
    if not m:
        continue
    a = int(m.group(1), 16)
    if off <= a < off + size:
        ins = m.group(2).rstrip(" ;")
        op = re.sub(r"^(@!?U?P\d+\s+)?", "", ins).split()[0].split(".")[0]
        ops[op] = ops.get(op, 0) + 1
        if n < maxl:
            print("%06x  %s" % (a, ins))
        n += 1
print("# %d instructions, %d bytes; opcode mix: %s" % (n, size, sorted(ops.items(), key=lambda kv: -kv[1])[:14]))
