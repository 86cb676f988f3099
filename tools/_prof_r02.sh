#!/bin/bash
# round-2 profiling pass (run under gpurun on one B200); outputs under gpurun_out/
mkdir -p gpurun_out
DISN_NO_GRAPH=1 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02d_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/r02d_under_ncu.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:point_tc -s 3 -c 1 -f -o gpurun_out/r02d_point_tc python tools/kbench.py --rounds 1 --reps 2 "X=0" > gpurun_out/r02d_ncu_point.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"mc_|scan_" -s 9 -c 9 -f -o gpurun_out/r02d_mc python bench.py --config 1 --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/r02d_ncu_mc.log 2>&1
for c in 1 0 2; do timeout 400 python bench.py --config $c --steps 10 --warmup 3 2>/dev/null | tail -1 > gpurun_out/r02d_bench_cfg$c.json; done
timeout 400 python bench.py --config 4 --steps 3 --warmup 3 2>/dev/null | tail -1 > gpurun_out/r02d_bench_cfg4.json
timeout 600 python bench.py --impl reference --steps 5 2>/dev/null | tail -1 > gpurun_out/r02d_bench_reference.json
python - <<'PY'
import json
for c in (1, 0, 2, 4):
    d = json.load(open("gpurun_out/r02d_bench_cfg%d.json" % c))
    print(c, "%.4g pts/s" % d["value"], "step %.3f ms" % d["ms_per_step"], "kernel %.3f" % d["roofline"]["kernel_ms"], "frac %.3f" % d["roofline"]["frac"],
          "enc %.3f" % d["encoder"]["ms"], "e2e %.4g" % d["e2e"]["value"], d.get("marching_cubes", {}).get("ms"))
PY
