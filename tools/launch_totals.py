"""Per-kernel totals of an ncu launch list (csv with gpu__time_duration.sum): python tools/launch_totals.py file.csv [top]"""
import csv, sys, re
from collections import defaultdict
rows = [r for r in csv.reader(open(sys.argv[1], errors="replace")) if len(r) > 10]
hdr = rows[0]
iname, ival, imet = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Name")
iunit = hdr.index("Metric Unit")
tot, cnt = defaultdict(float), defaultdict(int)
order = []
for r in rows[1:]:
    if r[imet] != "gpu__time_duration.sum":
        continue
    v = float(r[ival].replace(",", ""))
    u = r[iunit]
    us = v / 1e3 if u in ("ns", "nsecond") else v if u in ("us", "usecond") else v * 1e3 if u in ("ms", "msecond") else v
    name = re.sub(r"\(.*", "", r[iname]).replace("void ", "").replace("fsb::<unnamed>::", "")[:70]
    tot[name] += us; cnt[name] += 1
    order.append((name, us))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
print(f"{len(order)} launches, {sum(tot.values())/1e3:.3f} ms GPU time")
for k, v in sorted(tot.items(), key=lambda kv: -kv[1])[:top]:
    print(f"{v/1e3:9.3f} ms {cnt[k]:6d} x  {k}")
