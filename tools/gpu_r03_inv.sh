#!/bin/bash
# round 2, session 3: inverse rows kernel A/B (guard bands per chain, luma L2 prefetch, halo replication by byte permutes)
set -u
O=gpurun_out/r03
mkdir -p $O
export BENCH_ARGS="--workload inverse4k_b10_2020"
bash tools/ab_variants.sh "$@" 2>&1 | tee $O/ab_inverse.txt
timeout 900 python -m pytest tests/test_inverse_gpu.py tests/test_baseline_ref_gpu.py -x -q -m gpu > $O/pytest_inverse.log 2>&1; echo "pytest rc=$?"
tail -5 $O/pytest_inverse.log
