"""Turn an `ncu --metrics gpu__time_duration.sum --csv` launch list into a markdown summary (per kernel: launches, total,
share of the step).  Usage: python tools/summarize_launches.py gpurun_out/launches.csv [frames_in_capture] > profiles/...md"""
import collections
import csv
import sys


def main():
    path = sys.argv[1]
    with open(path) as f:
        lines = [l for l in f if not l.startswith("==")]
    rows = list(csv.DictReader(lines))
    # keep only full frames: from the first to the last input-conversion kernel
    idx = [i for i, r in enumerate(rows) if "to_s2d" in r["Kernel Name"]]
    if len(idx) >= 2:
        rows = rows[idx[0]:idx[-1]]
        frames = len(idx) - 1
    else:
        frames = 1
    agg = collections.OrderedDict()
    for r in rows:
        name = r["Kernel Name"].split("(")[0].replace("yp::<unnamed>::", "").replace("void ", "")
        if "at::" in name or "elementwise" in name:
            name = "torch copy/fill (counts, staging)"
        t = float(r["Metric Value"].replace(",", "")) / 1e3
        a = agg.setdefault(name, [0, 0.0])
        a[0] += 1
        a[1] += t
    total = sum(v[1] for v in agg.values())
    print(f"ncu launch list `{path}` ({frames} frame(s); per-launch times are cold-cache and serialised: compare shares)\n")
    print("| kernel | launches / frame | us / frame | share |")
    print("|---|---:|---:|---:|")
    for name, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"| `{name}` | {n / frames:.1f} | {t / frames:.1f} | {100 * t / total:.1f} % |")
    print(f"| **total** | {sum(v[0] for v in agg.values()) / frames:.1f} | {total / frames:.1f} | 100 % |")


if __name__ == "__main__":
    main()
