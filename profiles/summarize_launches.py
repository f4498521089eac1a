"""Per-kernel summary of an ncu launch list (`ncu --metrics gpu__time_duration.sum --csv --log-file X.csv ...`):
    python profiles/summarize_launches.py profiles/r02t_cfg4_launches.csv "title line" > profiles/r02t_cfg4_launches_summary.md
Kernel names are cut at the argument list; shares are of the summed per-launch durations (serialised, cold-cache under ncu)."""
import collections
import csv
import re
import sys


def main(path, title):
    lines = [l for l in open(path) if l.startswith('"')]
    agg, tot = collections.OrderedDict(), 0.0
    for row in csv.DictReader(lines):
        if row["Metric Name"] != "gpu__time_duration.sum":
            continue
        ns = float(row["Metric Value"].replace(",", ""))
        name = re.sub(r"\(.*$", "", row["Kernel Name"])[:110]
        a = agg.setdefault(name, [0, 0.0])
        a[0] += 1; a[1] += ns; tot += ns
    print("# %s\n" % title)
    print("%d launches, %.2f ms summed.\n" % (sum(a[0] for a in agg.values()), tot / 1e6))
    print("| kernel | launches | total us | avg us | share |\n|---|---|---|---|---|")
    for name, (n, ns) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print("| `%s` | %d | %.1f | %.2f | %.1f%% |" % (name, n, ns / 1e3, ns / 1e3 / n, 100 * ns / tot))


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2] if len(sys.argv) > 2 else "ncu launch list")
