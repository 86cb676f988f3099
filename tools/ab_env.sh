#!/bin/bash
# A/B two settings of an environment variable on the SAME GPU (chips differ by several % under the power cap, so numbers from
# different gpurun boxes are not comparable):   tools/ab_env.sh DISN_TC_V2 0 1 [precision] [rounds]
# Prints kernel ms, SM clock and kilo-cycles per 128-point tile for alternating runs.
VAR=${1:?env var}; A=${2:?value A}; B=${3:?value B}; PREC=${4:-f16f8}; ROUNDS=${5:-2}
run() {
  env "$VAR=$1" timeout 200 python bench.py --precision "$PREC" --steps 5 --warmup 3 2>/dev/null | tail -1 > /tmp/ab_line.json
  python - "$VAR=$1" <<'PY'
import json, sys
try:
    d = json.load(open('/tmp/ab_line.json'))
    k, mhz = d['roofline']['kernel_ms'], d['clocks']['sm_mhz']
    print("%-16s %s  kernel %.2f ms  %.0f MHz  %.1f Kcyc/tile  value %.4g" % (sys.argv[1], d['config']['precision'], k, mhz, k * mhz / 1792.1, d['value']))
except Exception as e:
    print(sys.argv[1], "bench failed:", e)
PY
}
for i in $(seq "$ROUNDS"); do run "$A"; run "$B"; done
