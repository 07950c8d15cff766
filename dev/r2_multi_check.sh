#!/bin/bash
# 2-GPU validation of the multi-GPU path behind the C ABI (run with: gpurun --gpus 2 -- bash dev/r2_multi_check.sh)
set -x
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/r2_topo.txt 2>&1
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2_multi_tests.log 2>&1; tail -15 gpurun_out/r2_multi_tests.log
for tr in peer nccl; do
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611 bench.py --gpus 2 --steps 3 --warmup 3 --transport $tr --no-cpu-baseline > gpurun_out/r2_bench_2gpu_$tr.json 2> gpurun_out/r2_bench_2gpu_$tr.err; tail -c 600 gpurun_out/r2_bench_2gpu_$tr.err; python - <<PY
import json
l=json.loads(open("gpurun_out/r2_bench_2gpu_$tr.json").read().strip().splitlines()[-1])
print("$tr", l["value"], l["e2e"]["value"], l["sum_resolve_ms"], l.get("parity_n"), {k:(v["value"],v["sum_resolve_ms"]) for k,v in l.get("other_configs",{}).items()})
PY
done
cd cpu-path-tracing_b200
./ptb_main 4096 --size 1920x1080 --devices 0,1 --p6 --out /tmp/a.ppm 2>&1 | tail -4
./ptb_main 4096 --size 1920x1080 --device 0 --out /tmp/b.ppm 2>&1 | tail -3
./ptb_main 1024 --scene dof_glass --size 3840x2160 --devices 0,1 --p6 --out /tmp/c.ppm 2>&1 | tail -4
./ptb_main 1024 --scene dof_glass --size 3840x2160 --devices 0,1 --out /tmp/d.ppm 2>&1 | tail -4
