"""Per-launch summary of an `ncu --csv --page raw` log: python scripts/ncu_raw_summary.py <csv> [metric ...]"""
import csv
import sys

DEFAULT = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
           "sm__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
           "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size",
           "smsp__inst_executed.sum"]


def main():
    rows = list(csv.reader(open(sys.argv[1])))
    hi = next(i for i, r in enumerate(rows) if r and r[0] == "ID")
    hdr, units = rows[hi], rows[hi + 1]
    want = sys.argv[2:] or DEFAULT
    for r in rows[hi + 2:]:
        if len(r) < len(hdr):
            continue
        print(r[hdr.index("Kernel Name")])
        for w in want:
            if w in hdr:
                print(f"    {w:70s} {r[hdr.index(w)]} {units[hdr.index(w)]}")


main()
