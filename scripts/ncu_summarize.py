"""`ncu -i x.ncu-rep --page raw --csv` -> the handful of counters the roofline discussion uses, one block per launch."""
import csv
import sys

KEEP = ["Kernel Name", "Grid Size", "Block Size", "gpu__time_duration.sum",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__t_bytes.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "sm__cycles_elapsed.max", "smsp__inst_executed.sum",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum"]
STALL = "smsp__average_warps_issue_stalled_"
rd = csv.reader(l for l in open(sys.argv[1]) if l.startswith('"'))
head = next(rd)
units = next(rd)
print(sys.argv[2] if len(sys.argv) > 2 else "")
for row in rd:
    rec = dict(zip(head, row))
    print("----")
    for k in KEEP:
        if k in rec:
            u = units[head.index(k)]
            print(f"{k:80s} {rec[k]} {u}")
    stalls = []
    for k, v in rec.items():
        if k.startswith(STALL) and k.endswith("_per_issue_active.ratio"):
            try:
                stalls.append((float(v.replace(",", "")), k[len(STALL):-len("_per_issue_active.ratio")]))
            except ValueError:
                pass
    for v, k in sorted(stalls, reverse=True)[:4]:
        print(f"{'stall ' + k + ' (warps per issue-active cycle)':80s} {v:.3f}")
