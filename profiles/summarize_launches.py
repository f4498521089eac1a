"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list per kernel (shares of the step)."""
import collections
import csv
import sys


def main(path, launches_per_step):
    lines = [l for l in open(path) if not l.startswith("==")]
    rows = list(csv.DictReader(lines))
    agg = collections.defaultdict(lambda: [0, 0.0])
    for x in rows:
        k = x["Kernel Name"].split("(")[0][:80]
        agg[k][0] += 1
        agg[k][1] += float(x["Metric Value"].replace(",", "")) / 1e3
    tot = sum(v[1] for v in agg.values())
    print("| kernel | launches | total us | avg us | share |\n|---|---|---|---|---|")
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print("| `%s` | %d | %.1f | %.2f | %.1f%% |" % (k, v[0], v[1], v[1] / v[0], 100 * v[1] / tot))
    print("\nTotal %.1f us over %d launches" % (tot, len(rows)), end="")
    if launches_per_step:
        print(" (~%.2f ms GPU time per step at %d launches per step)." % (tot / 1e3 / (len(rows) / launches_per_step),
                                                                         launches_per_step))
    else:
        print(".")


if __name__ == "__main__":
    main(sys.argv[1], float(sys.argv[2]) if len(sys.argv) > 2 else 0)
