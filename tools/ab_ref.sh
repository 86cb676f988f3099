#!/bin/bash
# A/B the working tree against a git ref on the SAME GPU:   tools/ab_ref.sh <git-ref> [precision] [rounds]
# Run the build part here (no GPU needed), then `gpurun -- 'bash tools/ab_ref.sh --run [precision] [rounds]'`.
if [ "$1" != "--run" ]; then
  REF=${1:?git ref}; rm -rf ab/A && mkdir -p ab/A && git archive "$REF" | tar -x -C ab/A && cp -f MEASURED_PEAKS.json ab/A/ 2>/dev/null
  (cd ab/A && python disn_b200/build.py) && python disn_b200/build.py && echo "built ab/A ($REF) and the working tree; now: gpurun -- 'bash tools/ab_ref.sh --run'"
  exit $?
fi
PREC=${2:-f16f8}; ROUNDS=${3:-2}
run() {
  ( cd "$1" && timeout 200 python bench.py --precision "$PREC" --steps 5 --warmup 3 2>/dev/null | tail -1 > /tmp/ab_line.json )
  python - "$1" <<'PY'
import json, sys
try:
    d = json.load(open('/tmp/ab_line.json'))
    k, mhz = d['roofline']['kernel_ms'], d['clocks']['sm_mhz']
    print("%-8s %s  kernel %.2f ms  %.0f MHz  %.1f Kcyc/tile  value %.4g" % (sys.argv[1], d['config']['precision'], k, mhz, k * mhz / 1792.1, d['value']))
except Exception as e:
    print(sys.argv[1], "bench failed:", e)
PY
}
for i in $(seq "$ROUNDS"); do run ab/A; run .; done
