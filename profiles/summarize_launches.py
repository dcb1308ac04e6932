"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list per kernel.
usage: python profiles/summarize_launches.py gpurun_out/launches.csv [skip_fraction]"""
import collections
import csv
import re
import sys


def main(path, skip=0.5):
    lines = [ln for ln in open(path) if not ln.startswith("==")]
    rows = list(csv.DictReader(lines))
    rows = rows[int(len(rows) * skip):]          # drop the warm-up step(s)
    tot, cnt = collections.Counter(), collections.Counter()
    for r in rows:
        name = re.sub(r"^void ", "", r["Kernel Name"])
        name = re.sub(r"\(.*$", "", name)
        v = float(r["Metric Value"].replace(",", ""))
        v = v / 1e3 if r["Metric Unit"] == "ns" else v * 1e3 if r["Metric Unit"] == "ms" else v
        tot[name] += v; cnt[name] += 1
    T = sum(tot.values())
    print(f"# {path}: {len(rows)} launches after skipping the first {skip:.0%}; total {T / 1e3:.2f} ms (cold-cache, serialised)")
    print(f"{'us':>10} {'share':>6} {'n':>5}  kernel")
    for k, v in tot.most_common():
        print(f"{v:10.1f} {100 * v / T:5.1f}% {cnt[k]:5d}  {k[:150]}")


if __name__ == "__main__":
    main(sys.argv[1], float(sys.argv[2]) if len(sys.argv) > 2 else 0.5)
