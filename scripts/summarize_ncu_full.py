"""Summarise an `ncu --set full` report (exported with `ncu -i X.ncu-rep --page raw --csv`) per launch: duration, DRAM
traffic, tensor-pipe activity, issue utilisation, registers, top stall reasons."""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
hdr, units = rows[0], rows[1]


def col(name):
    return hdr.index(name) if name in hdr else None


TENSOR = "TPC.TriageCompute.sm__pipe_tensor_subpipe_hmma_cycles_active_realtime.avg"
stalls = [h for h in hdr if h.startswith("smsp__pcsamp_warps_issue_stalled_") and not h.endswith("_not_issued")]
for r in rows[2:]:
    def g(name, default=""):
        c = col(name)
        return r[c] if c is not None and c < len(r) else default
    name = g("Kernel Name")
    dur = float(g("gpu__time_duration.sum", "0") or 0)
    cyc = float(g("sm__cycles_elapsed.avg", "0") or 0)
    rd, wr = float(g("dram__bytes_read.sum", "0") or 0), float(g("dram__bytes_write.sum", "0") or 0)
    t = g(TENSOR)
    tens = float(t) / 4.0 if t not in ("", "no data") else None   # the counter sums the 4 sub-partitions
    st = sorted(((int(r[hdr.index(h)] or 0), h.replace("smsp__pcsamp_warps_issue_stalled_", "")) for h in stalls), reverse=True)
    tot = sum(s for s, _ in st) or 1
    print("%s\n  grid %s  regs %s  duration %.1f %s  dram read %.1f + write %.1f %s  issue-active %s%%" % (
        name[:110], g("Grid Size"), g("launch__registers_per_thread"), dur, units[col("gpu__time_duration.sum")], rd, wr,
        units[col("dram__bytes_read.sum")], g("smsp__issue_active.avg.pct_of_peak_sustained_active")[:5]))
    if tens is not None and cyc:
        print("  tensor pipe active %.0f of %.0f cycles per SM (%.0f%%)" % (tens, cyc, 100 * tens / cyc))
    print("  stalls: " + ", ".join("%s %.0f%%" % (n, 100.0 * s / tot) for s, n in st[:5]))
