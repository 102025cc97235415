#!/bin/bash
# TIFF and inverse routes: parity tests, then the three bench workloads (with the reference parity leg)
set -u
O=gpurun_out/r02; T=${1:-t}
mkdir -p $O
timeout 900 python -m pytest tests/test_forward_gpu.py tests/test_inverse_gpu.py -x -q -m gpu -k "tiff or inverse or golden or 1080p" > $O/${T}_pytest.log 2>&1; echo "pytest rc=$?"; tail -2 $O/${T}_pytest.log
for w in tiff1080_bt2020_420 tiff1080_ydzdx_420 inverse4k_b10_2020; do
  timeout 300 python bench.py --steps 20 --warmup 3 --workload $w ${BENCH_EXTRA:-} > $O/${T}_bench_$w.json 2> $O/${T}_bench_$w.err; echo "bench $w rc=$?"
done
python - $T <<'PY'
import json,sys
for w in ("tiff1080_bt2020_420", "tiff1080_ydzdx_420", "inverse4k_b10_2020"):
    try:
        j = json.loads(open("gpurun_out/r02/%s_bench_%s.json" % (sys.argv[1], w)).read().strip().splitlines()[-1])
        r = j["roofline"]
        print(w, "ms/step %.4f kernel %.4f frac %.3f parity %s" % (j["ms_per_step"], r["kernel_ms_per_launch"], r["frac"], j.get("parity")))
    except Exception as e:
        print(w, "FAILED", e)
PY
