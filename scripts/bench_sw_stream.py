"""TMA delivery rate of head-slice tiles: round-1 4-D map (16-byte inner boxes, unswizzled core-matrix layout) vs the round-2 3-D map
(128-byte swizzled boxes + 32-byte tail), DiT-XL/2 geometry (rows 32768, 16 heads of 72 in a packed [R, 3456] qkv buffer)."""
import json
import sys

import torch

sys.path.insert(0, ".")
from diffulab_b200 import _lib  # noqa: E402

lib = _lib.load_probes()
B, S, H, hd = 128, 256, 16, 72
rows, ld = B * S, 3 * H * hd
x = torch.randn(rows, ld, device="cuda").bfloat16()
base = x[:, 2 * H * hd:]
st = torch.cuda.current_stream().cuda_stream
clk = torch.cuda.clock_rate() if hasattr(torch.cuda, "clock_rate") else None
for name, fn in (("r1_4d_16B_inner", lambda g, t: lib.dlb_tma_gather_probe(base.data_ptr(), rows, ld, H, hd, g, t, None, st)),
                 ("r2_3d_sw128_sw32", lambda g, t: lib.dlb_attn_sw_stream_probe(base.data_ptr(), rows, ld, H, hd, g, t, st))):
    for grid, tiles in ((148, 64), (148, 256), (296, 128)):
        fn(grid, tiles)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(5):
            fn(grid, tiles)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 5
        nbytes = grid * tiles * 64 * hd * 2
        print(json.dumps({"layout": name, "grid": grid, "tiles_per_cta": tiles, "ms": round(ms, 4), "GB_per_s": round(nbytes / ms / 1e6, 1),
                          "bytes_per_clk_per_sm_at_1.9GHz": round(nbytes / (ms * 1e-3) / 1.9e9 / 148, 1)}))
