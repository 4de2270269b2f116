"""Where the data-parallel overhead of the step goes: reads the CUPTI timelines `bench.py --trace-out` writes (chrome
trace, one per rank; `.json` or `.json.gz`) and compares the compute stream of a data-parallel rank with the single-GPU
step kernel by kernel.

    python scripts/comm_overlap.py gpurun_out/r02_trace_n1.rank0.json.gz gpurun_out/r02_trace_n2_fill.rank0.json.gz

Prints: the collectives (count, total and mean duration, which compute kernels they ran beside), the stretch of the
compute kernels (same launch order in both traces) by kernel name, and the idle gaps of the compute stream."""
import collections
import gzip
import json
import sys


def load(path):
    op = gzip.open if path.endswith(".gz") else open
    with op(path, "rt") as f:
        d = json.load(f)
    ev = [e for e in d["traceEvents"] if e.get("cat") == "kernel" and "dur" in e]
    ev.sort(key=lambda e: e["ts"])
    return ev


def short(n):
    n = n.split("(")[0]
    for k in ("gemm_kernel", "attn_bwd_kernel", "attn_fwd", "ln_bwd", "ln_fwd", "multi_copy", "nccl", "chunk_mean", "colsum"):
        if k in n:
            return k
    return n[-40:]


def is_comm(e):
    n = e["name"].lower()
    return "nccl" in n


def is_side(e):
    return "multi_copy" in e["name"]


def last_step(ev):
    """Kernels of the last of the two profiled graph replays.  (A replayed graph runs on the executor's own streams, so
    stream ids say nothing: compute kernels are told from the reducer's by name and ordered by start time.)"""
    comp = [e for e in ev if not is_comm(e) and not is_side(e)]
    m2 = comp[len(comp) // 2:]
    t0, t1 = m2[0]["ts"], max(e["ts"] + e["dur"] for e in m2)
    others = [e for e in ev if (is_comm(e) or is_side(e)) and t0 <= e["ts"]]
    return m2, others, t0, t1


def union_busy(evs):
    """Length of the union of the kernels' intervals (kernels of independent graph branches overlap)."""
    tot, cur0, cur1 = 0.0, None, None
    for e in sorted(evs, key=lambda e: e["ts"]):
        a, b = e["ts"], e["ts"] + e["dur"]
        if cur1 is None or a > cur1:
            if cur1 is not None:
                tot += cur1 - cur0
            cur0, cur1 = a, b
        else:
            cur1 = max(cur1, b)
    return tot + (cur1 - cur0 if cur1 is not None else 0.0)


def main():
    base, dp = load(sys.argv[1]), load(sys.argv[2])
    bm, _, b0, b1 = last_step(base)
    dm, do, d0, d1 = last_step(dp)
    print("single GPU : %d compute kernels, span %.2f ms, sum of durations %.2f ms, busy (union) %.2f ms"
          % (len(bm), (b1 - b0) / 1e3, sum(e["dur"] for e in bm) / 1e3, union_busy(bm) / 1e3))
    print("data par.  : %d compute kernels, span %.2f ms, sum of durations %.2f ms, busy (union) %.2f ms"
          % (len(dm), (d1 - d0) / 1e3, sum(e["dur"] for e in dm) / 1e3, union_busy(dm) / 1e3))
    nccl = [e for e in do if is_comm(e)]
    side = [e for e in do if is_side(e)]
    print("collectives: %d, total %.2f ms, mean %.1f us, max %.1f us; gather / scatter kernels: %d, %.2f ms"
          % (len(nccl), sum(e["dur"] for e in nccl) / 1e3, sum(e["dur"] for e in nccl) / max(len(nccl), 1),
             max([e["dur"] for e in nccl] + [0]), len(side), sum(e["dur"] for e in side) / 1e3))
    end_all = max([e["ts"] + e["dur"] for e in nccl + side] + [d1])
    print("after the last compute kernel: %.1f us of reducer work (exposed tail)" % (end_all - d1))
    # align launches by (kernel name, occurrence index inside the step)
    def keyed(m):
        cnt, out = collections.Counter(), {}
        for e in m:
            out[(e["name"], cnt[e["name"]])] = e
            cnt[e["name"]] += 1
        return out
    kb, kd = keyed(bm), keyed(dm)
    common = [k for k in kb if k in kd]
    print("launches matched by (name, occurrence): %d of %d / %d" % (len(common), len(bm), len(dm)))
    stretch = collections.defaultdict(lambda: [0, 0.0, 0.0])
    beside = collections.defaultdict(lambda: [0, 0.0])
    for k in common:
        a, b = kb[k], kd[k]
        n = short(a["name"])
        s = stretch[n]
        s[0] += 1
        s[1] += a["dur"]
        s[2] += b["dur"]
        for c in nccl:                         # overlap of this compute kernel with a collective
            ov = min(b["ts"] + b["dur"], c["ts"] + c["dur"]) - max(b["ts"], c["ts"])
            if ov > 0:
                beside[n][0] += 1
                beside[n][1] += ov
    print("%-28s %6s %10s %10s %9s   %s" % ("kernel", "count", "single ms", "dp ms", "delta ms", "overlap with collectives (count, ms)"))
    for n, (c, a, b) in sorted(stretch.items(), key=lambda kv: -(kv[1][2] - kv[1][1])):
        if abs(b - a) > 5 or beside[n][0]:
            print("%-28s %6d %10.3f %10.3f %+9.3f   %d, %.3f" % (n, c, a / 1e3, b / 1e3, (b - a) / 1e3, beside[n][0], beside[n][1] / 1e3))
    print("total stretch of the matched launches: %+.3f ms" % (sum(b - a for _, a, b in stretch.values()) / 1e3))
    worst = sorted(common, key=lambda k: -(kd[k]["dur"] - kb[k]["dur"]))[:14]
    print("most stretched launches:")
    for k in worst:
        b = kd[k]
        ov = [c for c in nccl if min(b["ts"] + b["dur"], c["ts"] + c["dur"]) - max(b["ts"], c["ts"]) > 0]
        print("  %-16s #%3d %8.1f -> %8.1f us   beside %d collective(s)%s" % (
            short(b["name"]), k[1], kb[k]["dur"], b["dur"], len(ov),
            "".join("  [starts %+.1f us, %.1f us long]" % (c["ts"] - b["ts"], c["dur"]) for c in ov[:2])))
    print("collectives, in order: start (ms into the step), duration, compute kernels running when it starts")
    for c in nccl:
        run = [short(e["name"]) for e in dm if e["ts"] <= c["ts"] < e["ts"] + e["dur"]]
        print("  %8.3f  %7.1f us  %s" % ((c["ts"] - d0) / 1e3, c["dur"], ",".join(run)))


if __name__ == "__main__":
    main()
