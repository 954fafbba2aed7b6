"""Per-launch table (and per-kernel aggregate) of an `ncu --metrics gpu__time_duration.sum[,dram__bytes_read.sum,dram__bytes_write.sum] --csv` log.
usage: python tools/launch_list.py <launches.csv> [agg|list]"""
import collections
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
start = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
hdr = rows[start]
idx = {h: i for i, h in enumerate(hdr)}
L = collections.OrderedDict()
for r in rows[start + 1:]:
    if len(r) < len(hdr):
        continue
    d = L.setdefault(int(r[idx["ID"]]), {"name": r[idx["Kernel Name"]].split("(")[0].replace("void ", "").replace("b200::", ""),
                                          "grid": r[idx["Grid Size"]]})
    d[r[idx["Metric Name"]]] = float(r[idx["Metric Value"]])
mode = sys.argv[2] if len(sys.argv) > 2 else "agg"
if mode == "list":
    for k, d in L.items():
        t = d["gpu__time_duration.sum"] / 1e3
        rd, wr = d.get("dram__bytes_read.sum", 0) / 1e6, d.get("dram__bytes_write.sum", 0) / 1e6
        print(f"{k:4d} {d['name'][:44]:44s} {d['grid']:16s} {t:8.1f} us  r {rd:8.1f}  w {wr:8.1f} MB")
else:
    agg = collections.OrderedDict()
    for d in L.values():
        a = agg.setdefault(d["name"][:50], [0, 0.0])
        a[0] += 1
        a[1] += d["gpu__time_duration.sum"] / 1e3
    tot = sum(a[1] for a in agg.values())
    for n, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"{n:52s} {c:4d} {t:9.1f} us {100 * t / tot:5.1f}%")
    print(f"{'total':52s} {len(L):4d} {tot:9.1f} us")
