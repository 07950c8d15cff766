"""profiles/r2_counters_{C2,C3,C4,C5}.csv (made by dev/r2_ncu_counters.sh on a GPU box) -> profiles/ncu_counters.json,
keyed by the hash of the device sources (bench.py: kernel_build_id) so that bench.py prints them only for this build."""
import csv, json, os, sys, importlib.util
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
spec = importlib.util.spec_from_file_location("bench", os.path.join(ROOT, "bench.py")); b = importlib.util.module_from_spec(spec); spec.loader.exec_module(b)
build = b.kernel_build_id()
out = {}
PEAK_LANES = 148 * 128 * 1.965e9
for cfg, variant, paths in (("C3", "sorted", 1920 * 1080 * 4 * 1024), ("C2", "sorted", 1024 * 768 * 4 * 256), ("C4", "sorted", 3840 * 2160 * 4 * 256),
                            ("C5", "sorted", 1920 * 1080 * 4 * 64)):
    rows = [r for r in csv.reader(open(os.path.join(ROOT, "profiles", f"r2_counters_{cfg}.csv"))) if len(r) > 14 and r[0].isdigit()]
    m = {r[12]: float(r[14].replace(",", "")) for r in rows}
    t = m["gpu__time_duration.sum"] * 1e-9
    ffma, fadd, fmul = (m[f"smsp__sass_thread_inst_executed_op_{k}_pred_on.sum"] for k in ("ffma", "fadd", "fmul"))
    e = out[f"{cfg}_{variant}"] = {
        "kernel_build": build, "source": f"profiles/r2_counters_{cfg}.csv (ncu --metrics, --clock-control none, one launch at the size below)",
        "kernel": rows[0][4], "paths_in_capture": paths, "kernel_ms_under_ncu": t * 1e3,
        "executed_fp32_tflops": (2 * ffma + fadd + fmul) / t / 1e12, "executed_fp32_frac": (2 * ffma + fadd + fmul) / t / 1e12 / 74.45,
        "fp32_lane_utilisation": (ffma + fadd + fmul) / (PEAK_LANES * t),
        "ffma_thread_inst": ffma, "fadd_thread_inst": fadd, "fmul_thread_inst": fmul,
        "warp_inst": m["smsp__inst_executed.sum"], "warp_inst_per_path": m["smsp__inst_executed.sum"] / paths,
        "simt_efficiency": m["smsp__thread_inst_executed.sum"] / m["smsp__inst_executed.sum"] / 32,
        "issue_active_pct": m["smsp__issue_active.avg.pct_of_peak_sustained_active"],
        "pipe_fma_pct": m["sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active"],
        "l1tex_throughput_pct": m["l1tex__throughput.avg.pct_of_peak_sustained_elapsed"],
        "l2_throughput_pct": m["lts__throughput.avg.pct_of_peak_sustained_elapsed"],
        "dram_bytes_per_launch": m["dram__bytes_read.sum"] + m["dram__bytes_write.sum"],
        "dram_bytes_scale_note": "per launch of the capture's size (paths_in_capture); the accumulation buffer (16 B x sub-pixels) is read and "
                                 "written once per launch whatever the sample count",
    }
    print(cfg, f"{t * 1e3:.1f} ms  exec {e['executed_fp32_tflops']:.2f} TF/s frac {e['executed_fp32_frac']:.3f} lanes {e['fp32_lane_utilisation']:.3f} "
               f"simt {e['simt_efficiency']:.3f} issue {e['issue_active_pct']} l1 {e['l1tex_throughput_pct']} l2 {e['l2_throughput_pct']} "
               f"dram {e['dram_bytes_per_launch'] / 1e6:.0f} MB  instr/path {e['warp_inst_per_path']:.1f}")
json.dump(out, open(os.path.join(ROOT, "profiles", "ncu_counters.json"), "w"), indent=1)
