#!/bin/bash
# ncu --set full captures (with per-instruction source counters) of the three dominant kernels
set -u
O=gpurun_out/r03
mkdir -p $O
ncu --set full --clock-control none --import-source on -k regex:k_forward_exr420_rows -s 10 -c 1 -o $O/ncu_rows_spec python bench.py --steps 2 --warmup 3 --no-cpu > $O/ncu_rows_spec.log 2>&1; echo "ncu spec rc=$?"
ncu --set full --clock-control none --import-source on -k regex:k_forward_u16_420_rows -s 3 -c 1 -o $O/ncu_tiff_rows python bench.py --steps 2 --warmup 3 --no-cpu --workload tiff1080_bt2020_420 > $O/ncu_tiff_rows.log 2>&1; echo "ncu tiff rc=$?"
ncu --set full --clock-control none --import-source on -k regex:k_inverse_rows -s 3 -c 1 -o $O/ncu_inverse_rows python bench.py --steps 2 --warmup 3 --no-cpu --workload inverse4k_b10_2020 > $O/ncu_inverse_rows.log 2>&1; echo "ncu inverse rc=$?"
ls -la $O
