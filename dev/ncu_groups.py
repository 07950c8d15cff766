"""Per-instruction view of an .ncu-rep (one kernel, captured with --import-source on): consecutive SASS instructions with the
same execution count AND the same number of active threads are one group -- finer than the plateaus of dev/ncu_summary.py,
which average the lanes over a region and so hide a branch that runs in most iterations at three lanes (how the "weight 1"
add of every dying path was found: 6 instructions, 90 % of the iterations, 2.8 lanes, 2.2 % of all issue slots).

    python dev/ncu_groups.py report.ncu-rep [min share in %, default 0.2] > groups.md
"""
import csv, io, subprocess, sys

rep = sys.argv[1]
floor = float(sys.argv[2]) if len(sys.argv) > 2 else 0.2
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True, check=True).stdout
rows = list(csv.reader(io.StringIO(txt)))
kernel = rows[0][1] if rows and len(rows[0]) > 1 else "?"
hdr = next(r for r in rows if "Instructions Executed" in r)
iex, ith, ipo = hdr.index("Instructions Executed"), hdr.index("Thread Instructions Executed"), hdr.index("Predicated-On Thread Instructions Executed")
ins = []
for r in rows[rows.index(hdr) + 1:]:
    try:
        ins.append((r[0][-4:], r[1], float(r[iex]), float(r[ith]), float(r[ipo])))
    except (ValueError, IndexError):
        pass
total = sum(i[2] for i in ins)
groups = []
for addr, sass, ex, th, po in ins:
    lanes = th / ex if ex else 0.0
    g = groups[-1] if groups else None
    if g and abs(g["ex"] - ex) <= 0.02 * max(ex, 1.0) and abs(g["lanes"] - lanes) < 0.7:
        g["n"] += 1; g["cost"] += ex; g["end"] = addr; g["on"] += po
    else:
        groups.append(dict(start=addr, end=addr, ex=ex, lanes=lanes, n=1, cost=ex, first=sass, on=po))
print(f"# instruction groups of `{kernel[:120]}`\n\nfrom `{rep}`; {total / 1e6:.1f} M warp instructions; groups below {floor} % omitted\n")
print("| SASS | instructions | executions (M) | active lanes | predicated-on lanes | share | first instruction |\n|---|---|---|---|---|---|---|")
for g in groups:
    share = 100 * g["cost"] / total
    if share >= floor:
        print(f"| {g['start']}-{g['end']} | {g['n']} | {g['ex'] / 1e6:.2f} | {g['lanes']:.1f} | {g['on'] / g['cost']:.1f} | {share:.2f} % | `{g['first'].strip()[:70]}` |")
