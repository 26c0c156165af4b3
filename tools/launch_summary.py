"""Per-kernel totals of an ncu launch list (--metrics gpu__time_duration.sum[,dram__bytes_*] --csv)."""
import csv, sys
from collections import defaultdict
rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 10]
hdr = rows[0]
iK, iM, iV, iG, iB = hdr.index("Kernel Name"), hdr.index("Metric Name"), hdr.index("Metric Value"), hdr.index("Grid Size"), hdr.index("Block Size")
iID = hdr.index("ID")
agg = defaultdict(lambda: defaultdict(float)); cnt = defaultdict(set)
for r in rows[1:]:
    name = r[iK].split("(")[0].replace("void fsb::<unnamed>::", "").replace("void fsb::", "")
    key = f"{name} grid={r[iG]} block={r[iB]}"
    agg[key][r[iM]] += float(r[iV].replace(",", "")); cnt[key].add(r[iID])
tot = sum(v["gpu__time_duration.sum"] for v in agg.values())
print(f"{'kernel':95s} {'launches':>8s} {'total_us':>10s} {'share':>6s} {'avg_us':>8s} {'dram_MB/launch':>14s} {'GB/s':>7s}")
for k, v in sorted(agg.items(), key=lambda kv: -kv[1]["gpu__time_duration.sum"]):
    n = len(cnt[k]); t = v["gpu__time_duration.sum"] / 1e3  # ns -> us
    b = v.get("dram__bytes_read.sum", 0) + v.get("dram__bytes_write.sum", 0)
    print(f"{k[:95]:95s} {n:8d} {t:10.1f} {100*v['gpu__time_duration.sum']/tot:5.1f}% {t/n:8.2f} {b/n/1e6:14.2f} {b/max(t,1e-9)/1e3:7.0f}")
