#!/bin/bash
# round 2, multi-GPU call (gpurun --gpus N): BASELINE configs[4] as stated, the CLI with --devices N, the copy probe
# usage: tools/gpu_r02_multi.sh N
set -u
N=${1:-8}
O=gpurun_out/r02/multi$N
mkdir -p $O
nvidia-smi --query-gpu=index,name,clocks.sm,clocks.max.sm --format=csv > $O/gpus.txt 2>&1
nvidia-smi topo -m > $O/topo.txt 2>&1
nproc > $O/host.txt; free -g >> $O/host.txt; lscpu | head -25 >> $O/host.txt
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"
# configs[4]: 600 frames 3840x2160 EXR -> 12-bit BT.2020 4:2:0 over 8 GPUs = 75 frames per GPU, device-resident and end to end
timeout 600 $TR bench.py --gpus $N --steps 10 --warmup 3 --no-cpu --workload exr4k_pq12_bt2020_420 > $O/bench_pq12.json 2> $O/bench_pq12.err; echo "bench pq12 rc=$?"
timeout 600 $TR bench.py --gpus $N --steps 10 --warmup 3 --no-cpu > $O/bench_default.json 2> $O/bench_default.err; echo "bench default rc=$?"
# the copy probe: threads vs processes, allocation kinds, CPU pinning
nvcc -O2 -o /tmp/h2d_probe tools/probe/h2d_probe.cu -lpthread 2> /dev/null
: > $O/probe.jsonl
for g in 1 $N; do
  for mode in threads procs; do
    for alloc in pinned wc registered; do
      timeout 120 /tmp/h2d_probe $g $mode $alloc 256 24 0 >> $O/probe.jsonl 2>> $O/probe.err
    done
  done
done
for g in 2 4; do
  [ $g -lt $N ] || continue
  for mode in threads procs; do
    timeout 120 /tmp/h2d_probe $g $mode pinned 256 24 0 >> $O/probe.jsonl 2>> $O/probe.err
  done
done
C=$(( $(nproc) / N ))
timeout 120 /tmp/h2d_probe $N threads pinned 256 24 $C >> $O/probe.jsonl 2>> $O/probe.err
timeout 120 /tmp/h2d_probe $N procs pinned 256 24 $C >> $O/probe.jsonl 2>> $O/probe.err
timeout 120 /tmp/h2d_probe $N threads pinned 64 96 0 >> $O/probe.jsonl 2>> $O/probe.err
cat $O/probe.jsonl
# the C++ host on 600 generated frames, --devices 1 and N, destination on tmpfs and on /dev/null
bash tools/gpu_r02_cli600.sh $N 600
timeout 600 python -m pytest tests/test_cli_gpu.py -q -m gpu -k devices_2 > $O/pytest_devices2.log 2>&1; echo "devices_2 test rc=$?"; tail -2 $O/pytest_devices2.log
python - $O <<'PY'
import json, sys
for n in ("pq12", "default"):
    try:
        j = json.loads(open(sys.argv[1] + "/bench_%s.json" % n).read().strip().splitlines()[-1])
        print(n, "gpus", j["n_gpus"], "value %.0f Mpx/s = %.0f fps" % (j["value"], j["frames_per_s"]), "e2e %.0f Mpx/s = %.0f fps" % (j["e2e"]["value"], j["e2e"]["frames_per_s"]), "h2d/gpu", j["e2e"].get("h2d_gbs_per_gpu"))
    except Exception as e:
        print(n, "FAILED", e)
PY
