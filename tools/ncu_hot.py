"""Hot SASS lines of one kernel in an ncu report: python tools/ncu_hot.py rep kernel_regex [launch_skip]"""
import csv, subprocess, sys
from collections import Counter
rep, rx = sys.argv[1], sys.argv[2]
skip = sys.argv[3] if len(sys.argv) > 3 else "0"
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", "regex:" + rx],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
blocks, cur = [], None
for r in rows:
    if r and r[0] == "Kernel Name":
        cur = [r]; blocks.append(cur)
    elif cur is not None:
        cur.append(r)
rows = blocks[int(skip)] if blocks else rows
hdr = rows[1]
data = [r for r in rows[2:] if len(r) == len(hdr)]
iS, iI, isrc = hdr.index("# Samples"), hdr.index("Instructions Executed"), hdr.index("Source")
tot, toti = sum(int(r[iS]) for r in data), sum(int(r[iI]) for r in data)
print("kernel", rows[0][1][:80]); print("total samples", tot, "warp-instructions", toti, "SASS lines", len(data))
c, ci = Counter(), Counter()
for r in data:
    w = r[isrc].split()
    op = w[1] if w[0].startswith("@") else w[0]
    op = op.split(".")[0]
    c[op] += int(r[iS]); ci[op] += int(r[iI])
print("samples by opcode:", [(k, round(100 * v / max(tot, 1), 1)) for k, v in c.most_common(12)])
print("instr by opcode:  ", [(k, round(100 * v / max(toti, 1), 1)) for k, v in ci.most_common(12)])
top = sorted(range(len(data)), key=lambda k: -int(data[k][iS]))[:16]
for k in sorted(top):
    print(f"{k:5d} samples {100*int(data[k][iS])/max(tot,1):5.1f}%  inst {int(data[k][iI]):9d}  {data[k][isrc][:100]}")
