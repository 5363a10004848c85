"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list into markdown tables.

    python tools/summarize_launches.py launches.csv [top] [--last-pass MARKER]

Without --last-pass: one per-kernel table over the whole capture (warm-up passes and the one-off weight packing
included).  With it: the launches from the LAST launch whose name contains MARKER to the end, in order — one
inference pass when MARKER is the first kernel of the pass (e.g. sp_bucket_build_cluster_kernel).
"""
import collections
import csv
import re
import sys


def read(path):
    lines = [l for l in open(path) if not l.startswith("==")]
    out = []
    for row in csv.DictReader(lines):
        if row.get("Metric Name") != "gpu__time_duration.sum":
            continue
        v = float(row["Metric Value"].replace(",", ""))
        unit = row["Metric Unit"]
        v = v / 1e3 if unit == "ns" else v * 1e3 if unit == "ms" else v
        name = re.sub(r"\(.*", "", row["Kernel Name"])
        name = re.sub(r"^void ", "", name).replace("<unnamed>::", "")[:100]
        out.append((name, v))
    return out


def table(launches, top, title):
    tot, cnt = collections.defaultdict(float), collections.Counter()
    for name, v in launches:
        tot[name] += v
        cnt[name] += 1
    total = sum(tot.values())
    print(f"{title}: {total / 1e3:.3f} ms over {sum(cnt.values())} launches\n")
    print("| kernel | launches | total us | share | avg us |")
    print("|---|---:|---:|---:|---:|")
    for k, v in sorted(tot.items(), key=lambda kv: -kv[1])[:top]:
        print(f"| `{k}` | {cnt[k]} | {v:.0f} | {100 * v / total:.1f}% | {v / cnt[k]:.1f} |")


def main(argv):
    path = argv[1]
    top = int(argv[2]) if len(argv) > 2 and not argv[2].startswith("--") else 30
    launches = read(path)
    if "--last-pass" in argv:
        marker = argv[argv.index("--last-pass") + 1]
        starts = [i for i, (n, _) in enumerate(launches) if marker in n]
        launches = launches[starts[-1]:]
        table(launches, top, f"last pass of {path}")
        print("\nin launch order:\n")
        for name, v in launches:
            print(f"    {v:8.1f} us  {name}")
    else:
        table(launches, top, f"whole capture {path}")


if __name__ == "__main__":
    main(sys.argv)
