"""Key metrics per kernel from `ncu -i X.ncu-rep --page raw --csv` (first capture of each kernel name, plus the mean
duration over all its captures).  python tools/full_summary.py X_raw.csv "header line" > X_summary.txt"""
import csv
import sys
from collections import OrderedDict

KEYS = """gpu__time_duration.sum sm__throughput.avg.pct_of_peak_sustained_elapsed dram__bytes_read.sum dram__bytes_write.sum
gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed lts__throughput.avg.pct_of_peak_sustained_elapsed
l1tex__throughput.avg.pct_of_peak_sustained_elapsed launch__registers_per_thread launch__grid_size launch__block_size
launch__shared_mem_per_block_dynamic sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active
sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed smsp__issue_active.avg.pct_of_peak_sustained_active
smsp__inst_executed.sum sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active sm__cycles_elapsed.avg smsp__cycles_active.avg
sm__warps_active.avg.pct_of_peak_sustained_active l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum
l1tex__data_pipe_lsu_wavefronts_mem_shared.sum""".split()


def main():
    rows = list(csv.reader(open(sys.argv[1], errors="replace")))
    hdr, units = rows[0], rows[1]
    kn = hdr.index("Kernel Name")
    seen = OrderedDict()
    for r in rows[2:]:
        seen.setdefault(r[kn], []).append(r)
    if len(sys.argv) > 2:
        print(sys.argv[2] + "\n")
    di = hdr.index("gpu__time_duration.sum")
    for name, rs in seen.items():
        r = rs[0]
        mean = sum(float(x[di].replace(",", "")) for x in rs) / len(rs)
        print(f"==== {name} ({len(rs)} captures, mean duration {mean:.3f} {units[di]})")
        for k in KEYS:
            if k in hdr:
                i = hdr.index(k)
                print(f"  {k:<86} {r[i]} {units[i]}")
        st = [(hdr[i], float(r[i].replace(",", "") or 0)) for i in range(len(hdr))
              if "pcsamp_warps_issue_stalled" in hdr[i] and "not_issued" not in hdr[i]]
        for k, v in sorted(st, key=lambda kv: -kv[1])[:8]:
            print(f"  {k:<86} {v:.0f} samples")


if __name__ == "__main__":
    main()
