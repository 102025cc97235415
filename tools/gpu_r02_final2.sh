#!/bin/bash
# round 2, last session: evidence for the reworked inverse rows kernel on top of gpu_r02_final.sh's set
# (whole GPU suite, smoke, default and inverse bench lines with the reference parity leg, ncu --set full of k_inverse_rows)
set -u
O=gpurun_out/r02/final2
mkdir -p $O
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.limit --format=csv > $O/gpu.txt 2>&1
timeout 1500 python -m pytest tests -q -m gpu > $O/pytest_gpu.log 2>&1; echo "gpu suite rc=$?" | tee -a $O/status.txt; tail -2 $O/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1; echo "smoke rc=$?" | tee -a $O/status.txt; tail -2 $O/smoke.log
timeout 400 python bench.py > $O/bench_default.json 2> $O/bench_default.err; echo "bench default rc=$?" | tee -a $O/status.txt
for w in inverse4k_b10_2020 tiff1080_bt2020_420 tiff1080_ydzdx_420 exr4k_pq12_bt2020_420; do timeout 400 python bench.py --workload $w > $O/bench_$w.json 2> $O/bench_$w.err; echo "bench $w rc=$?" | tee -a $O/status.txt; done
for w in inverse4k_b12_ydzdx inverse4k_b12_2020 inverse4k_b10_709; do timeout 300 python bench.py --workload $w --no-cpu > $O/bench_$w.json 2> $O/bench_$w.err; echo "bench $w rc=$?" | tee -a $O/status.txt; done
ncu --set full --clock-control none --import-source on -k regex:k_inverse_rows -s 3 -c 1 -o $O/ncu_inverse_rows python bench.py --steps 2 --warmup 3 --no-cpu --workload inverse4k_b10_2020 > $O/ncu_inverse_rows.log 2>&1; echo "ncu inverse rc=$?" | tee -a $O/status.txt
ncu --metrics gpu__time_duration.sum --clock-control none -c 100 --csv --log-file $O/launches_inverse.csv python bench.py --steps 2 --warmup 3 --no-cpu --workload inverse4k_b10_2020 > $O/ncu_launches_inverse.log 2>&1; echo "ncu launches rc=$?" | tee -a $O/status.txt
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/r02/final2/bench_*.json")):
    try:
        j = json.loads(open(f).read().strip().splitlines()[-1])
        r = j.get("roofline", {})
        print(f.split("/")[-1], "ms/step %.4f value %.0f kernel %.4f frac %.3f e2e %.0f parity %s" % (j["ms_per_step"], j["value"], r.get("kernel_ms_per_launch", 0), r.get("frac", 0), j["e2e"]["value"], j.get("parity")))
    except Exception as e:
        print(f, "FAILED", e)
PY
