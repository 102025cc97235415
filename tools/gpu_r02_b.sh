#!/bin/bash
# round 2, GPU call B: ncu --set full of the single-pass (SPEC) rows kernel and the classic one, same command
set -u
O=gpurun_out/r02
mkdir -p $O
python bench.py --steps 3 --warmup 3 --no-cpu > $O/f_plain.json 2> $O/f_plain.err &&
ncu --set full --clock-control none --import-source on -k regex:k_forward_exr420_rows -s 8 -c 6 -o $O/f_spec python bench.py --steps 3 --warmup 3 --no-cpu > $O/f_ncu.log 2>&1
echo "rc=$?"
tail -5 $O/f_ncu.log
