#!/bin/bash
# launch list of the headline bench command (single pass), per-launch device time
set -u
O=gpurun_out/r02
mkdir -p $O
python bench.py --steps 2 --warmup 3 --no-cpu > $O/h_plain.json 2> $O/h_plain.err &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/h_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu > $O/h_ncu.log 2>&1
echo "rc=$?"
