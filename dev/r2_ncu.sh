#!/bin/bash
# ncu captures of the shipped kernels (run-time compiled sorted megakernel on box_mirror; hierarchy kernel on spheres10k)
set -x
mkdir -p gpurun_out
python dev/prof_render.py box_mirror 960 540 8 2 3 > gpurun_out/r2_prof_bm_plain.log 2>&1 || exit 1
python dev/prof_render.py spheres10k 960 540 8 2 2 > gpurun_out/r2_prof_s10k_plain.log 2>&1 || exit 1
ncu --set full --clock-control none --import-source on -k regex:mega_sorted --launch-skip 2 -c 1 -f -o gpurun_out/r2_sorted_v15_bm python dev/prof_render.py box_mirror 960 540 8 2 3 > gpurun_out/r2_ncu_bm.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:mega_sorted --launch-skip 1 -c 1 -f -o gpurun_out/r2_sorted_v14_s10k python dev/prof_render.py spheres10k 960 540 8 2 2 > gpurun_out/r2_ncu_s10k.log 2>&1
ls -la gpurun_out/*.ncu-rep
tail -2 gpurun_out/r2_prof_bm_plain.log gpurun_out/r2_prof_s10k_plain.log
