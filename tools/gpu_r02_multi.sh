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
for g in 1 2 4 $N; do
  [ $g -le $N ] || continue
  for mode in threads procs; do
    for alloc in pinned wc registered; do
      timeout 120 /tmp/h2d_probe $g $mode $alloc 256 24 0 >> $O/probe.jsonl 2>> $O/probe.err
    done
  done
done
C=$(( $(nproc) / N ))
timeout 120 /tmp/h2d_probe $N threads pinned 256 24 $C >> $O/probe.jsonl 2>> $O/probe.err
timeout 120 /tmp/h2d_probe $N procs pinned 256 24 $C >> $O/probe.jsonl 2>> $O/probe.err
timeout 120 /tmp/h2d_probe $N threads pinned 64 96 0 >> $O/probe.jsonl 2>> $O/probe.err
cat $O/probe.jsonl
# the C++ host on 600 generated frames (8 distinct uncompressed 4K half RGB EXR files under 600 names), 12-bit, --devices N
python - <<'PY'
import os, subprocess, sys
sys.path.insert(0, ".")
from hdr2yuv_b200 import build, synth
cli = build.build_cli()
d = "/dev/shm/h2y600"
os.makedirs(d, exist_ok=True)
w, h = 3840, 2160
for i in range(8):
    synth.exr_half_frame_fast(w, h, seed=i, channels=3).tofile(d + "/f.raw")
    subprocess.check_call([cli["h2y_iotool"], "write-exr", d + "/u%d.exr" % i, str(w), str(h), "3", "0", d + "/f.raw"])
for i in range(600):
    p = d + "/shot.%04d.exr" % i
    if not os.path.exists(p):
        os.symlink(d + "/u%d.exr" % (i % 8), p)
PY
for dev in 1 $N; do
  rm -f /dev/shm/h2y600/out.yuv
  timeout 900 hdr2yuv_b200/cli/bin/hdr2yuv --src_filename /dev/shm/h2y600/shot.0000.exr --dst_filename /dev/shm/h2y600/out.yuv \
    --src_transfer_characteristics LINEAR --dst_transfer_characteristics PQ --src_pic_width 3840 --src_pic_height 2160 --src_bit_depth 16 \
    --dst_bit_depth 12 --src_chroma_format_idc 3 --dst_chroma_format_idc 1 --src_matrix_coeffs 0 --dst_matrix_coeffs 9 \
    --src_colour_primaries 1 --dst_colour_primaries 9 --chroma_resampler_type 1 --dst_video_full_range_flag 0 \
    --n_frames 600 --devices $dev > $O/cli_devices$dev.log 2>&1; echo "cli devices $dev rc=$?"
  tail -2 $O/cli_devices$dev.log
  ls -la /dev/shm/h2y600/out.yuv | awk '{print $5}' >> $O/cli_devices$dev.log
  md5sum /dev/shm/h2y600/out.yuv >> $O/cli_devices$dev.log
done
rm -rf /dev/shm/h2y600
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
