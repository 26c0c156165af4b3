"""Per-instruction shared-memory wavefronts and stall samples of one kernel in an ncu report (needs --import-source on):
python tools/ncu_lsu.py rep kernel_regex [block_index]"""
import csv, subprocess, sys
rep, rx = sys.argv[1], sys.argv[2]
skip = int(sys.argv[3]) if len(sys.argv) > 3 else 0
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", "regex:" + rx], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
blocks, cur = [], None
for r in rows:
    if r and r[0] == "Kernel Name":
        cur = [r]; blocks.append(cur)
    elif cur is not None:
        cur.append(r)
b = blocks[skip]
hdr = b[1]
data = [r for r in b[2:] if len(r) == len(hdr)]
ix = {h: i for i, h in enumerate(hdr)}
print("kernel", b[0][1][:90], " blocks in report:", len(blocks))
tot_s = sum(int(r[ix["# Samples"]]) for r in data)
cols = ["# Samples", "Instructions Executed", "L1 Wavefronts Shared", "L1 Wavefronts Shared Ideal", "L1 Tag Requests Global"]
tot = {c: sum(int(r[ix[c]]) for r in data) for c in cols}
print({c: tot[c] for c in cols})
stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
st = {h: sum(int(r[ix[h]]) for r in data) for h in stalls}
print("stalls:", sorted(((v, k) for k, v in st.items() if v), reverse=True)[:8])
for k, r in enumerate(data):
    w, s = int(r[ix["L1 Wavefronts Shared"]]), int(r[ix["# Samples"]])
    g = int(r[ix["L1 Tag Requests Global"]])
    if w > 0 or g > 0 or s > 0.01 * tot_s:
        print(f"{k:4d} smp {100*s/max(tot_s,1):5.1f}%  inst {int(r[ix['Instructions Executed']]):8d}  shW {w:8d} ideal {int(r[ix['L1 Wavefronts Shared Ideal']]):8d}  glob {g:7d}  {r[ix['Source']].strip()[:70]}")
