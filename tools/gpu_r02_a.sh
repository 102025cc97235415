#!/bin/bash
# round 2, GPU call A: the plan-reuse tests first (new code), then the whole GPU suite, then the headline bench both ways
set -u
O=gpurun_out/r02; T=${1:-a}
mkdir -p $O
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > $O/${T}_gpu.txt 2>&1
timeout 900 python -m pytest tests/test_plan_reuse_gpu.py -x -q -m gpu > $O/${T}_pytest_plan_reuse.log 2>&1; echo "plan_reuse rc=$?" >> $O/${T}_status.txt
timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu > $O/${T}_bench_auto.json 2> $O/${T}_bench_auto.err; echo "bench auto rc=$?" >> $O/${T}_status.txt
timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu --plan-reuse 0 > $O/${T}_bench_classic.json 2> $O/${T}_bench_classic.err; echo "bench classic rc=$?" >> $O/${T}_status.txt
timeout 1500 python -m pytest tests -x -q -m gpu > $O/${T}_pytest_gpu.log 2>&1; echo "gpu suite rc=$?" >> $O/${T}_status.txt
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $O/${T}_smoke.log 2>&1; echo "smoke rc=$?" >> $O/${T}_status.txt
cat $O/${T}_status.txt
tail -3 $O/${T}_pytest_plan_reuse.log
tail -3 $O/${T}_pytest_gpu.log
tail -2 $O/${T}_smoke.log
python - $T <<'PY'
import json
for n in ("auto", "classic"):
    try:
        j = json.loads(open("gpurun_out/r02/%s_bench_%s.json" % (__import__("sys").argv[1], n)).read().strip().splitlines()[-1])
        r = j["roofline"]
        print(n, "ms/step %.4f value %.0f kernel %.4f prologue %.4f frac %.3f e2e %.0f plan %s" % (j["ms_per_step"], j["value"], r["kernel_ms_per_launch"], r["prologue_ms_per_step"], r["frac"], j["e2e"]["value"], j.get("plan_reuse")))
    except Exception as e:
        print(n, "FAILED", e)
PY
