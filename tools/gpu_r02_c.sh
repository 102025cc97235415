#!/bin/bash
# round 2, GPU call C: parity tests of the forward path, then the headline bench both ways (single pass / two pass)
set -u
O=gpurun_out/r02
T=${1:-c}
mkdir -p $O
timeout 900 python -m pytest tests/test_plan_reuse_gpu.py tests/test_forward_gpu.py -x -q -m gpu > $O/${T}_pytest.log 2>&1; echo "pytest rc=$?"
tail -3 $O/${T}_pytest.log
timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu > $O/${T}_bench_auto.json 2> $O/${T}_bench_auto.err; echo "bench auto rc=$?"
timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu --plan-reuse 0 > $O/${T}_bench_classic.json 2> $O/${T}_bench_classic.err; echo "bench classic rc=$?"
python - $T <<'PY'
import json,sys
for n in ("auto", "classic"):
    try:
        j = json.loads(open("gpurun_out/r02/%s_bench_%s.json" % (sys.argv[1], n)).read().strip().splitlines()[-1])
        r = j["roofline"]
        print(n, "ms/step %.4f value %.0f kernel %.4f prologue %.4f frac %.3f e2e %.0f redone %s" % (j["ms_per_step"], j["value"], r["kernel_ms_per_launch"], r["prologue_ms_per_step"], r["frac"], j["e2e"]["value"], j.get("plan_reuse", {}).get("frames_converted_again")))
    except Exception as e:
        print(n, "FAILED", e)
PY
