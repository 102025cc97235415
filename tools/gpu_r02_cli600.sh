#!/bin/bash
# the C++ host end to end on 600 generated 4K frames (8 distinct uncompressed half RGB EXR files under 600 names in
# /dev/shm), 12-bit BT.2020 4:2:0: BASELINE configs[4] through cli/bin/hdr2yuv --devices N.  usage: gpu_r02_cli600.sh N [frames]
set -u
N=${1:-1}; F=${2:-600}
O=gpurun_out/r02/multi$N
mkdir -p $O
python - $F <<'PY'
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
for i in range(int(sys.argv[1])):
    p = d + "/shot.%04d.exr" % i
    if not os.path.exists(p):
        os.symlink(d + "/u%d.exr" % (i % 8), p)
PY
for dev in 1 $N; do
  for dst in /dev/shm/h2y600/out.yuv /dev/null; do
  tag=$(basename $dst | tr -d .)
  rm -f /dev/shm/h2y600/out.yuv
  timeout 900 hdr2yuv_b200/cli/bin/hdr2yuv --src_filename /dev/shm/h2y600/shot.0000.exr --dst_filename $dst \
    --src_transfer_characteristics LINEAR --dst_transfer_characteristics PQ --src_pic_width 3840 --src_pic_height 2160 --src_bit_depth 16 \
    --dst_bit_depth 12 --src_chroma_format_idc 3 --dst_chroma_format_idc 1 --src_matrix_coeffs 0 --dst_matrix_coeffs 9 \
    --src_colour_primaries 1 --dst_colour_primaries 9 --chroma_resampler_type 1 --dst_video_full_range_flag 0 \
    --n_frames $F --devices $dev > $O/cli_devices${dev}_$tag.log 2>&1; echo "cli devices $dev -> $dst rc=$?"
  tail -$((dev + 1)) $O/cli_devices${dev}_$tag.log
  [ -f /dev/shm/h2y600/out.yuv ] && md5sum /dev/shm/h2y600/out.yuv | cut -c1-32 >> $O/cli_devices${dev}_$tag.log
  done
  [ $N -eq 1 ] && break
done
rm -rf /dev/shm/h2y600
