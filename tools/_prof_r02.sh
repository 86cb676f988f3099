#!/bin/bash
# round-2 profiling pass (run under gpurun on one B200); outputs under gpurun_out/
mkdir -p gpurun_out
DISN_NO_GRAPH=1 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/r02_under_ncu.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:point_tc -s 3 -c 1 -f -o gpurun_out/r02_point_tc python tools/kbench.py --rounds 1 --reps 2 "X=0" > gpurun_out/r02_ncu_point.log 2>&1
DISN_NO_GRAPH=1 timeout 900 ncu --set full --clock-control none --import-source on -k regex:conv_tc -s 17 -c 17 -f -o gpurun_out/r02_conv_tc_b8 python bench.py --config 2 --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/r02_ncu_conv.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"mc_|scan_" -s 10 -c 10 -f -o gpurun_out/r02_mc python bench.py --config 1 --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/r02_ncu_mc.log 2>&1
timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r02b_bench_cfg1.json 2>&1
tail -c 900 gpurun_out/r02b_bench_cfg1.json
ls -la gpurun_out/*.ncu-rep
