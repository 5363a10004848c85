"""Summarise an ncu --csv metrics log (tools/prof_bwd_ops.py run): per launch duration, DRAM / L2 bytes, shared-memory
bank conflicts, warp instructions.  python tools/summarize_ops_ncu.py gpurun_out/r02u_ops_ncu.csv [--last-half]"""
import csv
import sys


def main():
    rows = list(csv.reader(open(sys.argv[1])))
    hdr = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
    H, data = rows[hdr], rows[hdr + 1:]
    ki, mi, vi, ii = H.index("Kernel Name"), H.index("Metric Name"), H.index("Metric Value"), H.index("ID")
    d = {}
    for r in data:
        if len(r) > vi:
            d.setdefault((int(r[ii]), r[ki][:70]), {})[r[mi]] = r[vi]
    items = sorted(d.items())
    if "--last-half" in sys.argv:
        items = items[len(items) // 2:]
    for (i, k), m in items:
        g = lambda n: float(m.get(n, "0").replace(",", ""))
        print(f"{k:70s} {g('gpu__time_duration.sum') / 1e3:8.1f} us  dram R {g('dram__bytes_read.sum') / 1e6:7.1f} W "
              f"{g('dram__bytes_write.sum') / 1e6:7.1f} MB  L2 {g('lts__t_bytes.sum') / 1e6:8.1f} MB  smem conflicts "
              f"{g('l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum') / 1e6:7.2f} M  warp inst {g('smsp__inst_executed.sum') / 1e6:7.2f} M")


if __name__ == "__main__":
    main()
