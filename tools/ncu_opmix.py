"""Summarise an ncu report's SASS page: executed warp-instructions by opcode and stall samples (diagnostic)."""
import collections
import csv
import subprocess
import sys

rep = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 25
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
# first kernel only
hdr = None
ops = collections.Counter()
stall = collections.Counter()
total = 0
nk = 0
for r in rows:
    if r and r[0] == "Kernel Name":
        nk += 1
        if nk > 1:
            break
        continue
    if r and r[0] == "Address":
        hdr = r
        continue
    if hdr is None or len(r) < len(hdr):
        continue
    d = dict(zip(hdr, r))
    src = d["Source"].strip()
    toks = src.split()
    op = toks[1] if toks[0].startswith("@") else toks[0]
    op = op.split(".")[0] if not op.startswith(("LDG", "LDS", "STS", "STG", "HMMA", "SHFL", "MUFU", "I2F", "F2I", "F2F", "LDC")) else ".".join(op.split(".")[:3])
    n = int(d["Instructions Executed"])
    ops[op] += n
    stall[op] += int(d["Warp Stall Sampling (All Samples)"])
    total += n
print(f"total warp-instructions: {total}")
tot_stall = sum(stall.values())
for op, n in ops.most_common(top):
    print(f"{op:28s} {n:12d} {100*n/total:6.2f}%   stall-samples {100*stall[op]/max(tot_stall,1):6.2f}%")
