"""Readable table of an `ncu --set full ... --page raw --csv` export: one row per launch with the columns the roofline
discussion uses, and a trimmed CSV of the same columns.

    python tools/ncu_summary.py gpurun_out/r2_full.csv profiles/r2_ncu_full_hot_kernels.csv > profiles/r2_ncu_summary.md
"""
import csv
import re
import sys

COLS = [
    ("Kernel Name", "kernel"),
    ("launch__grid_size", "grid"),
    ("launch__block_size", "block"),
    ("gpu__time_duration.sum", "us"),
    ("dram__bytes_read.sum", "DRAM read MB"),
    ("dram__bytes_write.sum", "DRAM write MB"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram %"),
    ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor %"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm %"),
    ("lts__t_sector_hit_rate.pct", "L2 hit %"),
    ("launch__registers_per_thread", "regs"),
    ("launch__shared_mem_config_size", "smem cfg KB"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "occupancy %"),
]


def to_float(s):
    try:
        return float(s.replace(",", ""))
    except ValueError:
        return None


def main():
    src = sys.argv[1]
    rows = list(csv.reader(open(src)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    idx = {h: i for i, h in enumerate(hdr)}
    have = [(c, t) for c, t in COLS if c in idx]
    out_rows = []
    for r in data:
        o = {}
        for c, t in have:
            v, u = r[idx[c]], units[idx[c]]
            f = to_float(v)
            if c == "Kernel Name":
                name = re.sub(r"\(.*", "", v).replace("void ", "")
                name = re.sub(r"b200mvs::\(anonymous namespace\)::|b200mvs::|<?unnamed>::", "", name)
                o[t] = name
            elif f is None:
                o[t] = v
            elif u in ("ns", "nsecond"):
                o[t] = f"{f / 1e3:.1f}"
            elif u in ("us", "usecond"):
                o[t] = f"{f:.1f}"
            elif u in ("ms", "msecond"):
                o[t] = f"{f * 1e3:.1f}"
            elif u == "byte" and "KB" in t:
                o[t] = f"{f / 1024:.0f}"
            elif u == "Kbyte" and "KB" in t:
                o[t] = f"{f:.0f}"
            elif u == "byte":
                o[t] = f"{f / 1e6:.2f}"
            elif u == "Kbyte":
                o[t] = f"{f / 1e3:.2f}"
            elif u == "Mbyte":
                o[t] = f"{f:.2f}"
            elif u == "Gbyte":
                o[t] = f"{f * 1e3:.2f}"
            elif u == "%":
                o[t] = f"{f:.1f}"
            else:
                o[t] = f"{f:g}"
        out_rows.append(o)
    titles = [t for _, t in have]
    if len(sys.argv) > 2:
        with open(sys.argv[2], "w", newline="") as f:
            w = csv.writer(f)
            w.writerow(titles)
            for o in out_rows:
                w.writerow([o[t] for t in titles])
    print("| # | " + " | ".join(titles) + " |")
    print("|---" * (len(titles) + 1) + "|")
    for i, o in enumerate(out_rows):
        print(f"| {i} | " + " | ".join(f"`{o[t]}`" if t == "kernel" else str(o[t]) for t in titles) + " |")
    # per-kernel totals
    agg = {}
    for o in out_rows:
        a = agg.setdefault(o["kernel"], [0, 0.0, 0.0, 0.0])
        a[0] += 1
        a[1] += float(o["us"])
        a[2] += float(o.get("DRAM read MB", 0) or 0) + float(o.get("DRAM write MB", 0) or 0)
        a[3] += float(o.get("tensor %", 0) or 0) * float(o["us"])
    print("\n| kernel | launches | total us (serialised, cold) | DRAM MB | time-weighted tensor % |")
    print("|---|---|---|---|---|")
    for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"| `{k}` | {a[0]} | {a[1]:.1f} | {a[2]:.1f} | {a[3] / max(a[1], 1e-9):.1f} |")


if __name__ == "__main__":
    main()
