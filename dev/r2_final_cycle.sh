#!/bin/bash
# One GPU call after a device-source change: GPU tests, bench-size counters of the new build (-> profiles/ncu_counters.json, which
# bench.py prints only for the build it was made from), the bench line, and the ncu launch list of the same bench command.
mkdir -p gpurun_out
python -m pytest tests -q -m gpu -x > gpurun_out/r2_gpu_tests_1gpu.log 2>&1; tail -2 gpurun_out/r2_gpu_tests_1gpu.log
bash dev/r2_ncu_counters.sh
cp gpurun_out/r2_counters_C?.csv profiles/ && python dev/ncu_counters_json.py && cp profiles/ncu_counters.json gpurun_out/
python bench.py > gpurun_out/r2_bench_final.json 2> gpurun_out/r2_bench_final.err; tail -c 600 gpurun_out/r2_bench_final.json
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r2_bench_reference_final.json 2>/dev/null
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_launches_final.csv python bench.py --steps 2 --warmup 3 --no-other-configs > gpurun_out/r2_launches_final.log 2>&1
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -1
