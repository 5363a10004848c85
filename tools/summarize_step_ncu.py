"""Aggregate an ncu --csv metrics log of a step (training or inference) per kernel: launches, total duration, DRAM
read / write, achieved DRAM GB/s, L2 bytes, warp instructions, tensor-pipe activity (duration-weighted).
    python tools/summarize_step_ncu.py LOG.csv "title" [--last-fraction F] [--lib-only] [--last-n N]
    --last-fraction F: keep the last 1/F of the launches (2 for a log holding a warm-up pass and the measured one);
    --lib-only: this library's kernels only (anonymous-namespace names); --last-n N: the last N launches after that;
    --from-last NAME: everything from the last launch whose kernel name contains NAME (the first kernel of a pass)."""
import collections
import csv
import sys


def main():
    path, title = sys.argv[1], sys.argv[2]
    frac = int(sys.argv[sys.argv.index("--last-fraction") + 1]) if "--last-fraction" in sys.argv else 1
    rows = list(csv.reader(open(path)))
    hdr = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
    H, data = rows[hdr], rows[hdr + 1:]
    ki, mi, vi, ii = H.index("Kernel Name"), H.index("Metric Name"), H.index("Metric Value"), H.index("ID")
    d = collections.OrderedDict()
    for r in data:
        if len(r) > vi:
            d.setdefault((int(r[ii]), r[ki]), {})[r[mi]] = r[vi]
    items = sorted(d.items())
    if "--lib-only" in sys.argv:
        items = [it for it in items if "<unnamed>::" in it[0][1]]
    items = items[len(items) - len(items) // frac:]
    if "--from-last" in sys.argv:
        key = sys.argv[sys.argv.index("--from-last") + 1]
        starts = [i for i, it in enumerate(items) if key in it[0][1]]
        items = items[starts[-1]:]
    if "--last-n" in sys.argv:
        items = items[-int(sys.argv[sys.argv.index("--last-n") + 1]):]
    agg = collections.OrderedDict()
    for (_, k), m in items:
        g = lambda n: float(m.get(n, "0").replace(",", ""))
        name = k.split("(")[0].replace("void ", "").replace("<unnamed>::", "")
        a = agg.setdefault(name, [0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0])
        us = g("gpu__time_duration.sum") / 1e3
        a[0] += 1
        a[1] += us
        a[2] += g("dram__bytes_read.sum") / 1e6
        a[3] += g("dram__bytes_write.sum") / 1e6
        a[4] += g("smsp__inst_executed.sum") / 1e6
        a[5] += g("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active") * us
        a[6] += g("lts__t_bytes.sum") / 1e6
    tot = sum(a[1] for a in agg.values())
    print(f"{title}: {tot / 1e3:.3f} ms over {sum(a[0] for a in agg.values())} launches (ncu: cold caches, serialised; "
          "durations are for shares, not bench values)")
    print(f"{'kernel':46s} {'n':>4s} {'us':>9s} {'share':>6s} {'dram R MB':>10s} {'W MB':>9s} {'dram GB/s':>9s} "
          f"{'L2 MB':>9s} {'warp inst M':>12s} {'tensor pipe %':>13s}")
    for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"{k[:46]:46s} {a[0]:4d} {a[1]:9.1f} {100 * a[1] / tot:5.1f}% {a[2]:10.1f} {a[3]:9.1f} "
              f"{(a[2] + a[3]) / a[1] * 1e3:9.0f} {a[6]:9.1f} {a[4]:12.1f} {a[5] / a[1]:13.1f}")


if __name__ == "__main__":
    main()
