"""Condense `ncu --page raw --csv` exports / launch lists into the small tables committed under profiles/.
usage: python tools/summarize_ncu.py raw <in.csv> <out.md> | launches <in.csv> <out.md>"""
import collections
import csv
import sys

COLS = [("gpu__time_duration.sum", "us"), ("dram__bytes_read.sum", "read"), ("dram__bytes_write.sum", "write"),
        ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram%"),
        ("sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active", "hmma%"),
        ("sm__inst_executed_pipe_tensor_subpipe_hmma.avg.pct_of_peak_sustained_active", "hmma_inst%"),
        ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps%"), ("launch__registers_per_thread", "regs"),
        ("lts__t_sector_hit_rate.pct", "L2hit%"), ("l1tex__t_sector_hit_rate.pct", "L1hit%")]


def _table(path):
    rows = list(csv.reader(open(path)))
    for i, r in enumerate(rows):
        if "Kernel Name" in r:
            return r, rows[i + 1], rows[i + 2:]
    raise SystemExit("no header in " + path)


def raw(inp, out):
    hdr, units, data = _table(inp)
    idx = {h: i for i, h in enumerate(hdr)}
    with open(out, "w") as f:
        f.write(f"# ncu --set full summary of `{inp}` (one row per captured launch, serialized, cold cache)\n\n")
        f.write("| kernel | grid | " + " | ".join(f"{n} [{units[idx[c]]}]" for c, n in COLS if c in idx) + " |\n")
        f.write("|---|---|" + "---|" * sum(c in idx for c, _ in COLS) + "\n")
        for r in data:
            if len(r) < len(hdr):
                continue
            name = r[idx["Kernel Name"]].split("(")[0].replace("void ", "")[:44]
            vals = []
            for c, _ in COLS:
                if c in idx:
                    v = r[idx[c]]
                    try:
                        v = f"{float(v):.3f}".rstrip("0").rstrip(".")
                    except ValueError:
                        pass
                    vals.append(v)
            f.write(f"| {name} | {r[idx['Grid Size']]} | " + " | ".join(vals) + " |\n")


def launches(inp, out):
    hdr, _, _ = _table(inp)
    rows = list(csv.reader(open(inp)))
    start = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
    idx = {h: i for i, h in enumerate(hdr)}
    agg, tot = collections.OrderedDict(), 0.0
    for r in rows[start + 1:]:
        if len(r) < len(hdr):
            continue
        n = r[idx["Kernel Name"]].split("(")[0].replace("void ", "")[:60]
        v = float(r[idx["Metric Value"]])
        u = r[idx["Metric Unit"]]
        v = v / 1000 if u == "ns" else (v * 1000 if u == "ms" else v)
        a = agg.setdefault(n, [0, 0.0])
        a[0] += 1
        a[1] += v
        tot += v
    with open(out, "w") as f:
        f.write(f"# ncu launch list of one training step (`{inp}`; gpu__time_duration, serialized, cold cache: compare SHARES)\n\n")
        f.write("| kernel | launches | total ms | share |\n|---|---|---|---|\n")
        for n, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            f.write(f"| {n} | {c} | {t / 1000:.3f} | {100 * t / tot:.1f}% |\n")
        f.write(f"| **total** | {sum(c for c, _ in agg.values())} | {tot / 1000:.3f} | 100% |\n")


if __name__ == "__main__":
    {"raw": raw, "launches": launches}[sys.argv[1]](sys.argv[2], sys.argv[3])
