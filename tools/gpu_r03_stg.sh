#!/bin/bash
# round 2, session 3: staged (cp.async) rows of the forward kernel, A/B on one box + parity tests
set -u
O=gpurun_out/r03
mkdir -p $O
for v in 0 1 0 1; do
  H2Y_NO_STAGED_ROWS=$v python bench.py --steps 20 --warmup 3 --no-cpu > $O/stg_spec_$v.json 2> $O/stg_spec_$v.err
  H2Y_NO_STAGED_ROWS=$v python bench.py --steps 20 --warmup 3 --no-cpu --plan-reuse 0 > $O/stg_classic_$v.json 2> $O/stg_classic_$v.err
  python - $v <<'PY'
import json, sys
v = sys.argv[1]
for n in ("spec", "classic"):
    try:
        j = json.loads(open("gpurun_out/r03/stg_%s_%s.json" % (n, v)).read().strip().splitlines()[-1])
        r = j["roofline"]
        print("no_staged=%s %-8s ms/step %.4f kernel %.4f frac %.3f step_frac %.3f parity %s redone %s" % (v, n, j["ms_per_step"], r["kernel_ms_per_launch"], r["frac"], r.get("step_frac", 0), j.get("parity"), j.get("plan_reuse", {}).get("frames_converted_again")))
    except Exception as e:
        print(v, n, "FAILED", e)
PY
done
timeout 1200 python -m pytest tests/test_plan_reuse_gpu.py tests/test_forward_gpu.py tests/test_baseline_ref_gpu.py -x -q -m gpu > $O/pytest_fwd.log 2>&1; echo "pytest rc=$?"
tail -5 $O/pytest_fwd.log
