"""Summarise an .ncu-rep: headline metrics + per-region issue-slot breakdown of the top kernel.
   python dev/ncu_summary.py gpurun_out/prof.ncu-rep [out.md]"""
import csv, io, subprocess, sys

rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, vals = rows[0], rows[1], rows[2]
want = ["gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum", "smsp__thread_inst_executed.sum", "smsp__thread_inst_executed_per_inst_executed.ratio",
        "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
        "smsp__sass_thread_inst_executed_op_ffma_pred_on.sum", "smsp__sass_thread_inst_executed_op_fadd_pred_on.sum",
        "smsp__sass_thread_inst_executed_op_fmul_pred_on.sum", "smsp__warps_eligible.avg.per_cycle_active",
        "smsp__warps_active.avg.per_cycle_active", "dram__bytes_read.sum", "dram__bytes_write.sum", "sm__cycles_elapsed.avg",
        "smsp__sass_average_branch_targets_threads_uniform.pct", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "smsp__inst_executed_op_global_red.sum"]
out = []
out.append(f"# ncu summary of {rep}\n")
out.append(f"kernel: `{vals[hdr.index('Kernel Name')]}`\n")
out.append("| metric | unit | value |\n|---|---|---|")
d = {}
for i, h in enumerate(hdr):
    d[h] = vals[i]
    if h in want:
        out.append(f"| {h} | {units[i]} | {vals[i]} |")
for i, h in enumerate(hdr):
    if "issue_stalled" in h and h.endswith("_per_warp_active.pct"):
        try:
            if float(vals[i]) >= 2.0:
                out.append(f"| {h} | {units[i]} | {vals[i]} |")
        except ValueError:
            pass
try:
    simt = float(d["smsp__thread_inst_executed_per_inst_executed.ratio"]) / 32
    out.append(f"\nSIMT efficiency (threads per warp instruction / 32): **{simt:.3f}**")
    fl = sum(float(d[f"smsp__sass_thread_inst_executed_op_{k}_pred_on.sum"]) * (2 if k == "ffma" else 1) for k in ("ffma", "fadd", "fmul"))
    t = float(d["gpu__time_duration.sum"]) * {"ms": 1e-3, "us": 1e-6, "ns": 1e-9, "s": 1.0, "msecond": 1e-3, "usecond": 1e-6, "nsecond": 1e-9, "second": 1.0}[units[hdr.index("gpu__time_duration.sum")]]
    out.append(f"executed FP32 flop (ffma*2+fadd+fmul, predicated-on threads): {fl:.4g} -> {fl / t / 1e12:.2f} TFLOP/s executed")
except Exception as e:  # noqa
    out.append(f"(derived metrics unavailable: {e})")

src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
h2 = rows[1]
data = rows[2:]
ia, isrc, iex, ith = h2.index("Address"), h2.index("Source"), h2.index("Instructions Executed"), h2.index("Thread Instructions Executed")
ins = [(int(r[ia], 16), int(r[iex]), int(r[ith]), r[isrc].strip()) for r in data]
base = ins[0][0]
tot = sum(i[1] for i in ins)
out.append(f"\n## issue slots by code region (plateaus of equal execution count)\n\ntotal warp instructions {tot:.4g}\n")
out.append("| SASS range | #instr | warp-instr (M) | share | threads/instr | first instruction |\n|---|---|---|---|---|---|")
start = 0
def flush(a, b):
    ex = sum(o[1] for o in ins[a:b]); th = sum(o[2] for o in ins[a:b])
    if ex / tot >= 0.003:
        out.append(f"| {ins[a][0]-base:#06x}-{ins[b-1][0]-base:#06x} | {b-a} | {ex/1e6:.1f} | {100*ex/tot:.1f}% | {th/max(ex,1):.1f} | `{ins[a][3][:44]}` |")
for i in range(1, len(ins)):
    if abs(ins[i][1] - ins[i-1][1]) > 0.15 * max(ins[i][1], ins[i-1][1], 1e6):
        flush(start, i); start = i
flush(start, len(ins))
text = "\n".join(out) + "\n"
if len(sys.argv) > 2:
    open(sys.argv[2], "w").write(text)
print(text)
