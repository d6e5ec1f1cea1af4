"""Per-kernel instruction / stall-sample share by source line from an ncu report.
usage: python tools/ncu_lines.py <file.ncu-rep> <kernel-substring> [N]"""
import csv, io, subprocess, sys, collections
rep, pat = sys.argv[1], sys.argv[2]
N = int(sys.argv[3]) if len(sys.argv) > 3 else 30
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
agg, hdr, active = {}, None, False
for r in rows:
    if r and r[0] == "Function Name":
        active = pat in r[1]
    elif r and r[0] == "Line No":
        hdr = r
    elif active and hdr and len(r) >= len(hdr) and r[0] != "":
        try:
            ln, s, ins = int(r[0]), int(r[hdr.index("# Samples")]), int(r[hdr.index("Instructions Executed")])
        except ValueError:
            continue
        a = agg.setdefault((ln, r[1].strip()[:105]), [0, 0])
        a[0] += s; a[1] += ins
tot = sum(a[0] for a in agg.values()) or 1
toti = sum(a[1] for a in agg.values()) or 1
print(f"{pat}: {toti} warp-instr, {tot} samples (all profiled launches)")
for (ln, txt), a in sorted(agg.items(), key=lambda kv: -kv[1][1])[:N]:
    print(f"  L{ln:4d} instr {100*a[1]/toti:5.1f}%  samples {100*a[0]/tot:5.1f}% | {txt}")
