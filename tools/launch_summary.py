"""Summarises an ncu `--metrics gpu__time_duration.sum --csv` launch list: per-kernel totals of the LAST forward.
usage: python tools/launch_summary.py gpurun_out/launches.csv [launches_per_forward]"""
import collections
import csv
import re
import sys


def main():
    path = sys.argv[1]
    with open(path) as f:
        lines = [l for l in f if not l.startswith("==")]
    rows = list(csv.DictReader(lines))
    per = int(sys.argv[2]) if len(sys.argv) > 2 else len(rows) // 2
    last = rows[-per:]
    tot = 0.0
    agg = collections.OrderedDict()
    for x in last:
        name = re.sub(r"\(.*", "", x["Kernel Name"]).replace("void ", "").replace("unnamed>::", "")
        v = float(x["Metric Value"].replace(",", ""))
        unit = x["Metric Unit"]
        v = v / 1000 if unit == "ns" else (v * 1000 if unit == "ms" else v)
        tot += v
        key = (name, x["Grid Size"], x["Block Size"])
        agg.setdefault(key, [0.0, 0])
        agg[key][0] += v
        agg[key][1] += 1
    print(f"{len(rows)} launches in file, last {per}: {tot:.1f} us (serialised, cold caches)")
    for k, v in agg.items():
        print(f"{v[0]:9.1f} us x{v[1]:3d}  {k[0]} grid={k[1]} block={k[2]}")


if __name__ == "__main__":
    main()
