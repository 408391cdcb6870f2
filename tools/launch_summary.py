"""Per-kernel totals of an `ncu --metrics gpu__time_duration.sum --csv` launch list: python tools/launch_summary.py file.csv"""
import collections
import csv
import re
import sys


def main(path):
    rows = list(csv.reader(open(path)))
    start = next(i for i, r in enumerate(rows) if r and r[0] == "ID")
    hdr = rows[start]
    ki, mi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    tot, cnt = collections.defaultdict(float), collections.Counter()
    for r in rows[start + 1:]:
        if len(r) <= mi:
            continue
        name = re.sub(r"\(.*", "", r[ki]).replace("void ", "").replace("gpk::", "").replace("<unnamed>::", "").replace("unnamed>::", "")
        v = float(r[mi].replace(",", "")) * {"ns": 1e-3, "us": 1.0, "ms": 1e3}.get(r[ui], 1.0)
        tot[name] += v
        cnt[name] += 1
    T = sum(tot.values())
    print(f"# {path}: {sum(cnt.values())} launches, {T / 1e3:.3f} ms serialised kernel time (cold-cache, per-launch: compare SHARES)")
    for k, v in sorted(tot.items(), key=lambda kv: -kv[1]):
        print(f"{k[:70]:70s} n={cnt[k]:5d} total={v / 1e3:8.3f} ms  avg={v / cnt[k]:8.1f} us  share={v / T * 100:5.1f}%")


if __name__ == "__main__":
    main(sys.argv[1])
