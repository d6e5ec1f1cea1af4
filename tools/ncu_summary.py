"""Summarise ncu outputs into profiles/ (text, committed).

  python tools/ncu_summary.py launches <launches.csv>      -> per-kernel time shares
  python tools/ncu_summary.py report <file.ncu-rep> [N]    -> key metrics + top-N source lines by stall samples
"""
import collections
import csv
import io
import subprocess
import sys

KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
        "launch__shared_mem_per_block_dynamic", "launch__grid_size", "launch__block_size",
        "smsp__inst_executed.sum", "sm__cycles_elapsed.max",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct"]


def launches(path):
    rows = list(csv.reader(open(path)))
    hi = [i for i, r in enumerate(rows) if "Kernel Name" in r][0]
    hdr = rows[hi]
    kn, mv, mu = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    agg = collections.defaultdict(list)
    for r in rows[hi + 1:]:
        if len(r) <= mv:
            continue
        v = float(r[mv].replace(",", ""))
        v = v / 1000 if r[mu] == "ns" else (v * 1000 if r[mu] == "ms" else v)
        agg[r[kn][:90]].append(v)
    tot = sum(sum(v) for v in agg.values())
    print(f"# {path}: per-kernel device time (ncu gpu__time_duration.sum, cold-cache, serialised)")
    print(f"{'kernel':92s} {'n':>5s} {'total_us':>10s} {'avg_us':>8s} {'share':>6s}")
    for k, v in sorted(agg.items(), key=lambda kv: -sum(kv[1])):
        print(f"{k:92s} {len(v):5d} {sum(v):10.1f} {sum(v) / len(v):8.1f} {100 * sum(v) / tot:5.1f}%")


def report(path, topn=25):
    raw = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    print(f"# {path}: key metrics per profiled launch")
    kn = hdr.index("Kernel Name")
    for r in rows[2:]:
        print("## " + r[kn][:100])
        for k in KEYS:
            if k in hdr:
                i = hdr.index(k)
                print(f"  {k:72s} {r[i]:>16s} {units[i]}")
    src = subprocess.run(["ncu", "-i", path, "--page", "source", "--csv", "--print-source", "cuda,sass"],
                         capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(src)))
    hidx = [i for i, r in enumerate(rows) if r and r[0] == "Line No"]
    if not hidx:
        return
    hdr = rows[hidx[0]]
    iS, iI = hdr.index("# Samples"), hdr.index("Instructions Executed")
    stall = [(i, h) for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
    agg = {}
    for r in rows[hidx[0] + 1:]:
        if len(r) < len(hdr) or r[0] in ("", "Line No"):
            continue
        try:
            ln, s, ins = int(r[0]), int(r[iS]), int(r[iI])
        except ValueError:
            continue
        a = agg.setdefault(ln, [r[1].strip()[:100], 0, 0, collections.Counter()])
        a[1] += s
        a[2] += ins
        for i, h in stall:
            if r[i] not in ("-", ""):
                a[3][h] += int(r[i])
    tot = sum(a[1] for a in agg.values()) or 1
    toti = sum(a[2] for a in agg.values()) or 1
    allst = collections.Counter()
    for a in agg.values():
        allst.update(a[3])
    print("\n# stall-reason totals over all sampled lines (all launches in the report)")
    for h, c in allst.most_common(8):
        print(f"  {h:24s} {100 * c / sum(allst.values()):5.1f}%")
    print(f"\n# top {topn} source lines by warp-stall samples")
    for ln, a in sorted(agg.items(), key=lambda kv: -kv[1][1])[:topn]:
        top = ", ".join(f"{h}:{c}" for h, c in a[3].most_common(2))
        print(f"  L{ln:4d} {100 * a[1] / tot:5.1f}% samples {100 * a[2] / toti:5.1f}% instr | {a[0]} | {top}")


if __name__ == "__main__":
    if sys.argv[1] == "launches":
        launches(sys.argv[2])
    else:
        report(sys.argv[2], int(sys.argv[3]) if len(sys.argv) > 3 else 25)
