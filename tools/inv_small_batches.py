import sys, time, numpy as np, torch
sys.path.insert(0, '.')
from hdr2yuv_b200 import api, _cabi as cabi
ctx = api.Context(0)
for (w, h) in ((1920, 1080), (3840, 2160), (1280, 720)):
    for n in (1, 2, 4):
        p = cabi.InverseParams(w, h, 10, 2, 1, 0, 0, 0)
        g = torch.Generator(device='cuda'); g.manual_seed(1)
        d_yuv = torch.randint(64, 940, (n, w * h * 3 // 2), device='cuda', generator=g, dtype=torch.int32).to(torch.int16).view(torch.uint8).view(-1)
        d_rgb = torch.zeros(n * w * h * 6, dtype=torch.uint8, device='cuda')
        res = {}
        for kern in ('tile', 'rows', ''):
            ctx.set_option('H2Y_INVERSE_KERNEL', kern if kern else None) if kern else ctx.set_option(None)
            for _ in range(5): ctx.inverse(p, d_yuv, d_rgb, n)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(50): ctx.inverse(p, d_yuv, d_rgb, n)
            e1.record(); torch.cuda.synchronize()
            res[kern or 'auto'] = e0.elapsed_time(e1) / 50 * 1000
        print("%dx%d n=%d  tile %.1f us  rows %.1f us  auto %.1f us" % (w, h, n, res['tile'], res['rows'], res['auto']))
