"""N>1 host logic on CPU: two gloo ranks partition a frame range exactly like bench.py / the CLI
hosts do, with no data-path collective (frames are independent)."""
import os
import socket
import subprocess
import sys

from hdr2yuv_b200 import sharding

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_partition_is_exact_and_contiguous():
    for n in (0, 1, 7, 60, 600, 601):
        for world in (1, 2, 3, 4, 8):
            ranges = [sharding.frame_range(r, world, n) for r in range(world)]
            assert ranges[0][0] == 0 and ranges[-1][1] == n
            for a, b in zip(ranges, ranges[1:]):
                assert a[1] == b[0]
            sizes = [b - a for a, b in ranges]
            assert max(sizes) - min(sizes) <= 1
    assert sharding.frame_range(3, 8, 600) == (225, 300)       # 75 per GPU
    assert sharding.output_offset(5, 24883200) == 5 * 24883200  # append-compatible .yuv layout (tiff.cpp:440)


_WORKER = r"""
import os, sys
sys.path.insert(0, sys.argv[1])
import torch, torch.distributed as dist
from hdr2yuv_b200 import sharding
dist.init_process_group("gloo", rank=int(os.environ["RANK"]), world_size=int(os.environ["WORLD_SIZE"]))
rank, world = dist.get_rank(), dist.get_world_size()
n = 61
lo, hi = sharding.frame_range(rank, world, n)
owned = torch.zeros(n, dtype=torch.int32)
owned[lo:hi] = 1
# the only cross-rank traffic is bookkeeping: a barrier and the max-over-ranks of the elapsed time
dist.barrier()
ms = sharding.max_over_ranks(10.0 + rank)
dist.all_reduce(owned)      # test-only check that every frame has exactly one owner
assert bool((owned == 1).all()), owned
assert ms == 10.0 + world - 1, ms
total = sharding.sum_over_ranks(hi - lo)
assert total == n, total
dist.destroy_process_group()
print("rank", rank, "ok")
"""


def test_two_rank_gloo_sharding(tmp_path):
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    script = tmp_path / "worker.py"
    script.write_text(_WORKER)
    procs = []
    for rank in range(2):
        env = dict(os.environ, RANK=str(rank), WORLD_SIZE="2", MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
        procs.append(subprocess.Popen([sys.executable, str(script), ROOT], env=env, stdout=subprocess.PIPE,
                                      stderr=subprocess.STDOUT))
    for p in procs:
        out, _ = p.communicate(timeout=180)
        assert p.returncode == 0, out.decode()
