#!/bin/bash
# A/B timing of experiment builds (tools/build_variants.py) on one GPU: same bench line per variant, kernel time from
# the roofline object.  Usage: tools/ab_variants.sh head a b c ...   Output: gpurun_out/ab_<name>.json
set -u
mkdir -p gpurun_out
for v in "$@"; do
    H2Y_LIB=$PWD/hdr2yuv_b200/_variants/libh2y_$v.so python bench.py --steps 10 --warmup 3 --no-cpu ${BENCH_ARGS:-} > gpurun_out/ab_$v.json 2> gpurun_out/ab_$v.err
    python - "$v" <<'PY'
import json, sys
v = sys.argv[1]
try:
    j = json.loads(open("gpurun_out/ab_%s.json" % v).read().strip().splitlines()[-1])
    r = j["roofline"]
    print("%-6s kernel %.4f ms  frac %.3f  value %.1f %s  prologue %.3f ms" % (v, r.get("kernel_ms_per_launch", 0), r["frac"], j["value"], j["unit"], r.get("prologue_ms_per_step", 0)))
except Exception as e:
    print(v, "FAILED", e)
PY
done
