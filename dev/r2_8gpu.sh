#!/bin/bash
# 8-GPU validation: bench (both transports), the host program on all GPUs
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/r2_topo8.txt 2>&1
T="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29621"
timeout 600 $T bench.py --gpus 8 --steps 5 --warmup 3 > gpurun_out/r2_bench_8gpu_peer.json 2> gpurun_out/r2_bench_8gpu_peer.err
timeout 300 $T bench.py --gpus 8 --steps 5 --warmup 3 --transport nccl --no-other-configs --no-cpu-baseline > gpurun_out/r2_bench_8gpu_nccl.json 2> gpurun_out/r2_bench_8gpu_nccl.err
timeout 300 $T bench.py --gpus 8 --steps 3 --warmup 3 --config C4 --no-other-configs --no-cpu-baseline > gpurun_out/r2_bench_C4_8gpu_peer.json 2> gpurun_out/r2_bench_C4_8gpu_peer.err
timeout 300 $T bench.py --gpus 8 --steps 3 --warmup 3 --config C4 --transport nccl --no-other-configs --no-cpu-baseline > gpurun_out/r2_bench_C4_8gpu_nccl.json 2> gpurun_out/r2_bench_C4_8gpu_nccl.err
python - <<'PY'
import json
for f in ("r2_bench_8gpu_peer","r2_bench_8gpu_nccl","r2_bench_C4_8gpu_peer","r2_bench_C4_8gpu_nccl"):
    try:
        l=json.loads(open(f"gpurun_out/{f}.json").read().strip().splitlines()[-1])
        print(f, round(l["value"]), round(l["e2e"]["value"]), round(l["ms_per_step"],3), "kernel", round(l["roofline"]["kernel_ms"],3), "sum+resolve", round(l["sum_resolve_ms"],3), l.get("parity_n",{}).get("status"), l.get("parity_n",{}).get("transport"), {k:(round(v["value"]),round(v["sum_resolve_ms"],3)) for k,v in l.get("other_configs",{}).items()})
    except Exception as e:
        print(f, "FAILED", e); print(open(f"gpurun_out/{f}.err").read()[-800:])
PY
cd cpu-path-tracing_b200
./ptb_main 4096 --size 1920x1080 --devices 0,1,2,3,4,5,6,7 --p6 --out /tmp/a.ppm 2>&1 | tail -4
./ptb_main 4096 --scene dof_glass --size 3840x2160 --devices 0,1,2,3,4,5,6,7 --p6 --out /tmp/c.ppm 2>&1 | tail -4
PTB_GPUS=8 ../oracle/_ref/cpu_path_tracer_b200 4096 2>&1 | tail -2
