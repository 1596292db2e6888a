"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per-kernel total device time and share.

    python profiles/summarize_launches.py gpurun_out/launches_rN.csv > profiles/rN_launches.md
"""
import collections
import csv
import sys


def main(path):
    rows = list(csv.reader(open(path)))
    start = next(i for i, r in enumerate(rows) if r and r[0] == "ID")
    hdr = rows[start]
    ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    tot, cnt = collections.defaultdict(float), collections.Counter()
    scale = {"ns": 1e-6, "nsecond": 1e-6, "us": 1e-3, "usecond": 1e-3, "ms": 1.0, "msecond": 1.0, "s": 1e3, "second": 1e3}
    for r in rows[start + 1:]:
        if len(r) <= vi:
            continue
        name = r[ki].split("(")[0].replace("void ", "").replace("cmoe::<unnamed>::", "").replace("cmoe::", "")
        tot[name] += float(r[vi].replace(",", "")) * scale.get(r[ui], 1.0)
        cnt[name] += 1
    total = sum(tot.values())
    print(f"| kernel | launches | total ms | share |\n|---|---:|---:|---:|")
    for k, v in sorted(tot.items(), key=lambda kv: -kv[1]):
        print(f"| `{k}` | {cnt[k]} | {v:.3f} | {100 * v / total:.2f} % |")
    print(f"| **all** | {sum(cnt.values())} | {total:.3f} | 100 % |")


if __name__ == "__main__":
    main(sys.argv[1])
