#!/bin/bash
# tools/capture_profiles.sh -- round-2 ncu evidence, run on the GPU box through gpurun; outputs under gpurun_out/ (copied to profiles/).
# Every capture follows a plain run of the same command that exited 0; numbers printed under ncu are never bench values.
set -x
O=gpurun_out
python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-extras > $O/r2_bench_plain.json 2> /dev/null || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/r2_launches.csv \
    python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-extras > /dev/null 2>&1
# config 2, warp-per-scene kernel (the default for N < 32): Decision half + Planning half
python tools/profile_cycle.py 4096 12 > /dev/null || exit 1
ncu --set full --clock-control none --import-source on -k regex:dp_cycle_kernel -s 8 -c 2 -f -o $O/r2_dp_cycle_kernel python tools/profile_cycle.py 4096 12 > /dev/null 2>&1
# config 2 shape through the group kernel, and the config-5 workload (its default kernel)
DP_KERNEL=group ncu --set full --clock-control none --import-source on -k regex:dp_group_kernel -s 8 -c 1 -f -o $O/r2_dp_group_kernel_cfg2 python tools/profile_cycle.py 4096 12 > /dev/null 2>&1
python tools/profile_urban.py 1024 6 200 1 > /dev/null || exit 1
ncu --set full --clock-control none --import-source on -k regex:dp_group_kernel -s 3 -c 1 -f -o $O/r2_dp_group_kernel_cfg5 python tools/profile_urban.py 1024 6 200 1 > /dev/null 2>&1
# config 3: the dense sweep kernel (one cluster launch per call)
python tools/bench_sweep.py 200 > $O/r2_bench_sweep.json || exit 1
ncu --set full --clock-control none --import-source on -k regex:sweep_rows -s 30 -c 2 -f -o $O/r2_sweep_kernels python tools/bench_sweep.py 40 > /dev/null 2>&1
DP_SWEEP_DBG=1 python tools/sweep_probe.py > $O/r2_sweep_probe.txt 2>&1
# the small kernels: operators, reset, map preparation, FMA peak (one line each)
ncu --metrics gpu__time_duration.sum,sm__throughput.avg.pct_of_peak_sustained_elapsed,dram__throughput.avg.pct_of_peak_sustained_elapsed,launch__registers_per_thread,dram__bytes.sum \
    --clock-control none -k regex:'op_|dp_reset|dp_map_prep|sweep_prefix|fma_peak' -c 40 --csv --log-file $O/r2_small_kernels.csv \
    python -m pytest tests/test_gpu_parity.py -m gpu -q -k "operator or nearest or reset" > /dev/null 2>&1
# closed-loop episodes (graph vs direct launches), the output-frame kernel (HBM-bound) and the other new small kernels
python tools/bench_closed_loop.py 25 > $O/r2_closed_loop.json || exit 1
ncu --set full --clock-control none --import-source on -k regex:dp_frames_kernel -s 25 -c 1 -f -o $O/r2_dp_frames_kernel python tools/bench_closed_loop.py 2 > /dev/null 2>&1
ncu --metrics gpu__time_duration.sum,sm__throughput.avg.pct_of_peak_sustained_elapsed,dram__throughput.avg.pct_of_peak_sustained_elapsed,launch__registers_per_thread,dram__bytes.sum \
    --clock-control none -k regex:'dp_world|dp_frames|dp_v2x' -c 240 --csv --log-file $O/r2_new_kernels.csv \
    python -m pytest tests/test_closed_loop.py tests/test_v2x.py -m gpu -q -k "ragged or frames_equal or v2x_equals" > /dev/null 2>&1
python tools/group_timeline.py 4096 12 > $O/r2_group_timeline_cfg2.txt 2>&1
DP_DEBUG_LIB=$PWD/decision-making-and-path-planning_b200/libdmpp_b200_dbg.so python tools/scene_timeline.py 4096 > $O/r2_scene_timeline.txt 2>&1
DP_KERNEL=group python tools/compare_kernels.py 4096 10 highway > $O/r2_kernel_compare.txt 2>&1
DP_KERNEL=warp python tools/compare_kernels.py 4096 10 highway >> $O/r2_kernel_compare.txt 2>&1
DP_KERNEL=group python tools/compare_kernels.py 65536 10 highway >> $O/r2_kernel_compare.txt 2>&1
DP_KERNEL=warp python tools/compare_kernels.py 65536 10 highway >> $O/r2_kernel_compare.txt 2>&1
DP_KERNEL=group python tools/compare_kernels.py 1024 200 junction 40 >> $O/r2_kernel_compare.txt 2>&1
DP_KERNEL=warp python tools/compare_kernels.py 1024 200 junction 40 >> $O/r2_kernel_compare.txt 2>&1
DP_DEBUG_LIB=decision-making-and-path-planning_b200/libdmpp_r1.so DP_KERNEL=warp python tools/compare_kernels.py 4096 10 highway >> $O/r2_kernel_compare.txt 2>&1
DP_DEBUG_LIB=decision-making-and-path-planning_b200/libdmpp_r1.so DP_KERNEL=warp python tools/compare_kernels.py 1024 200 junction 40 >> $O/r2_kernel_compare.txt 2>&1
ls -la $O
