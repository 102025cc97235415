#!/bin/bash
# round 2, final evidence call: whole GPU suite, smoke, every bench workload, latency of one 1080p frame, reference arm,
# launch list and ncu --set full captures of the dominant kernels.  Output: gpurun_out/r02/final/
set -u
O=gpurun_out/r02/final
mkdir -p $O
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.limit --format=csv > $O/gpu.txt 2>&1
timeout 1500 python -m pytest tests -q -m gpu > $O/pytest_gpu.log 2>&1; echo "gpu suite rc=$?" | tee -a $O/status.txt; tail -2 $O/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1; echo "smoke rc=$?" | tee -a $O/status.txt; tail -2 $O/smoke.log
timeout 400 python bench.py > $O/bench_default.json 2> $O/bench_default.err; echo "bench default rc=$?" | tee -a $O/status.txt
timeout 400 python bench.py --plan-reuse 0 --no-cpu > $O/bench_two_pass.json 2> $O/bench_two_pass.err; echo "bench two-pass rc=$?" | tee -a $O/status.txt
for w in exr4k_pq12_bt2020_420 tiff1080_bt2020_420 tiff1080_ydzdx_420 tiff1080_ydzdx_444 inverse4k_b10_2020; do
  timeout 400 python bench.py --workload $w > $O/bench_$w.json 2> $O/bench_$w.err; echo "bench $w rc=$?" | tee -a $O/status.txt
done
timeout 400 python bench.py --forward-content natural --no-cpu > $O/bench_natural.json 2> $O/bench_natural.err; echo "bench natural rc=$?" | tee -a $O/status.txt
# configs[0] as written: ONE 1080p frame (device-resident latency and end to end through the API), then through the C++ host
timeout 300 python bench.py --workload tiff1080_bt2020_420 --frames 1 --steps 50 --warmup 5 --no-cpu > $O/bench_tiff1080_single_frame.json 2> $O/bench_tiff1080_single_frame.err; echo "bench single frame rc=$?" | tee -a $O/status.txt
python - <<'PY' > $O/cli_single_frame.log 2>&1
import os, subprocess, sys, time
sys.path.insert(0, ".")
from hdr2yuv_b200 import build, synth
cli = build.build_cli()
d = "/dev/shm/h2y1"
os.makedirs(d, exist_ok=True)
w, h = 1920, 1080
synth.tiff16_frame(w, h, seed=1).tofile(d + "/t.raw")
subprocess.check_call([cli["h2y_iotool"], "write-tiff", d + "/in.tiff", str(w), str(h), "3", d + "/t.raw"])
for rep in range(3):
    if os.path.exists(d + "/out.yuv"):
        os.remove(d + "/out.yuv")
    t = time.time()
    r = subprocess.run([cli["hdr2yuv"], "--src_filename", d + "/in.tiff", "--dst_filename", d + "/out.yuv", "--src_transfer_characteristics", "16",
                        "--dst_transfer_characteristics", "16", "--src_pic_width", str(w), "--src_pic_height", str(h), "--src_bit_depth", "16",
                        "--dst_bit_depth", "10", "--src_chroma_format_idc", "3", "--dst_chroma_format_idc", "1", "--src_matrix_coeffs", "0",
                        "--dst_matrix_coeffs", "9", "--src_colour_primaries", "10", "--dst_colour_primaries", "9", "--chroma_resampler_type", "1"],
                       stdout=subprocess.PIPE, stderr=subprocess.STDOUT)
    print("run %d: rc %d, wall %.3f s (process start, CUDA context, one frame, file write)" % (rep, r.returncode, time.time() - t))
    print(r.stdout.decode().strip().splitlines()[-2:])
PY
cat $O/cli_single_frame.log
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > $O/bench_reference_arm.json 2> $O/bench_reference_arm.err; echo "bench reference rc=$?" | tee -a $O/status.txt
# ---- ncu: launch list of the default command, then full captures of the dominant kernels (one frame group each) ----
ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file $O/launches_default.csv python bench.py --steps 2 --warmup 3 --no-cpu > $O/ncu_launches.log 2>&1; echo "ncu launches rc=$?" | tee -a $O/status.txt
ncu --set full --clock-control none --import-source on -k regex:k_forward_exr420_rows -s 10 -c 2 -o $O/ncu_rows_spec python bench.py --steps 2 --warmup 3 --no-cpu > $O/ncu_rows_spec.log 2>&1; echo "ncu spec rc=$?" | tee -a $O/status.txt
ncu --set full --clock-control none --import-source on -k regex:k_forward_exr420_rows -s 4 -c 1 -o $O/ncu_rows_two_pass python bench.py --steps 2 --warmup 3 --no-cpu --plan-reuse 0 > $O/ncu_rows_two_pass.log 2>&1; echo "ncu two-pass rc=$?" | tee -a $O/status.txt
ncu --set full --clock-control none --import-source on -k regex:k_forward_u16_420_rows -s 3 -c 1 -o $O/ncu_tiff_rows python bench.py --steps 2 --warmup 3 --no-cpu --workload tiff1080_bt2020_420 > $O/ncu_tiff_rows.log 2>&1; echo "ncu tiff rc=$?" | tee -a $O/status.txt
ncu --set full --clock-control none --import-source on -k regex:k_inverse_rows -s 3 -c 1 -o $O/ncu_inverse_rows python bench.py --steps 2 --warmup 3 --no-cpu --workload inverse4k_b10_2020 > $O/ncu_inverse_rows.log 2>&1; echo "ncu inverse rc=$?" | tee -a $O/status.txt
cat $O/status.txt
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/r02/final/bench_*.json")):
    try:
        j = json.loads(open(f).read().strip().splitlines()[-1])
        r = j.get("roofline", {})
        print(f.split("/")[-1], "ms/step %.4f value %.0f kernel %.4f frac %.3f e2e %.0f parity %s" % (j["ms_per_step"], j["value"], r.get("kernel_ms_per_launch", 0), r.get("frac", 0), j["e2e"]["value"], j.get("parity")))
    except Exception as e:
        print(f, "FAILED", e)
PY
