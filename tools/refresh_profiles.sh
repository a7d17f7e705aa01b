#!/bin/bash
# tools/refresh_profiles.sh -- run HERE after tools/capture_profiles.sh has run on the GPU box: turns the captures merged back into
# gpurun_out/ into the tracked summaries under profiles/ (raw pages as CSV, key metrics + stall mix + hottest lines, shares by
# function / phase) and copies the text outputs.
set -e
cd "$(dirname "$0")/.."
G=gpurun_out; P=profiles
raw() { ncu -i $G/$1.ncu-rep --page raw --csv > $P/$1_raw.csv 2>/dev/null; }
raw r2_dp_cycle_kernel; python tools/ncu_summary.py $G/r2_dp_cycle_kernel.ncu-rep 40 > $P/r2_dp_cycle_kernel_summary.txt
python tools/ncu_phase.py $G/r2_dp_cycle_kernel.ncu-rep > $P/r2_dp_cycle_kernel_by_function.txt
for c in cfg2 cfg5; do
  raw r2_dp_group_kernel_$c; python tools/ncu_summary.py $G/r2_dp_group_kernel_$c.ncu-rep 40 > $P/r2_dp_group_kernel_${c}_summary.txt
  python tools/ncu_lines.py $G/r2_dp_group_kernel_$c.ncu-rep > $P/r2_dp_group_kernel_${c}_by_phase.txt
done
raw r2_sweep_kernels; python tools/ncu_summary.py $G/r2_sweep_kernels.ncu-rep 40 > $P/r2_sweep_kernels_summary.txt
raw r2_dp_frames_kernel; python tools/ncu_summary.py $G/r2_dp_frames_kernel.ncu-rep 12 > $P/r2_dp_frames_kernel_summary.txt
for f in r2_launches.csv r2_small_kernels.csv r2_new_kernels.csv r2_bench_plain.json r2_bench_sweep.json r2_sweep_probe.txt \
         r2_group_timeline_cfg2.txt r2_scene_timeline.txt r2_kernel_compare.txt r2_closed_loop.json; do cp $G/$f $P/$f; done
python tools/static_evidence.py
ls -la $P | tail -5
