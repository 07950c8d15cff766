"""ncu launch list (CSV from `ncu --metrics gpu__time_duration.sum --csv`) -> markdown table by kernel.
   python dev/launch_list.py launches.csv out.md "title" """
import csv, sys
rows = [r for r in csv.reader(open(sys.argv[1], errors="ignore")) if len(r) > 14 and r[0].isdigit()]
agg = {}
for r in rows:
    if r[12] != "gpu__time_duration.sum":
        continue
    t = float(r[14].replace(",", "")) * {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "nsecond": 1e-6, "usecond": 1e-3, "msecond": 1.0}.get(r[13], 1e-6)
    a = agg.setdefault(r[4], [0, 0.0])
    a[0] += 1
    a[1] += t
total = sum(v[1] for v in agg.values())
out = [f"# {sys.argv[3]}\n",
       "`ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv` -- per-launch times are serialised and cold-cache; the SHARE "
       "is what must agree with bench.py (the number a run under ncu prints is not a bench value).\n",
       "| kernel | launches | total ms | mean ms | share |", "|---|---|---|---|---|"]
for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    out.append(f"| `{k[:110]}` | {n} | {t:.3f} | {t / n:.3f} | {100 * t / total:.2f}% |")
open(sys.argv[2], "w").write("\n".join(out) + "\n")
print("\n".join(out))
