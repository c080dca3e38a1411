"""TMA gather probe: layout check + streaming bandwidth of 16-byte-inner 4-D boxes (see csrc/tma_gather_probe.cu)."""
import sys

import torch

sys.path.insert(0, ".")
from diffulab_b200 import _lib  # noqa: E402

lib = _lib.load_probes()
B, S, H, hd = 128, 256, 16, 72
rows, ld = B * S, 3 * H * hd
x = torch.randn(rows, ld, device="cuda").bfloat16()
base = x[:, 2 * H * hd:]  # the V third of a packed qkv projection
cpr_box = (hd + 15) // 16 * 2
dump = torch.zeros(64 * cpr_box * 8, dtype=torch.bfloat16, device="cuda")
st = torch.cuda.current_stream().cuda_stream
rc = lib.dlb_tma_gather_probe(base.data_ptr(), rows, ld, H, hd, 148, 4, dump.data_ptr(), st)
torch.cuda.synchronize()
print("rc", rc, lib.dlb_last_error().decode() if rc else "")
# expected: chunk c, row r at (c*64 + r)*8 elements; CTA 0 -> head 0, rows 0..63
img = dump.view(cpr_box, 64, 8)
exp = torch.zeros_like(img)
for c in range(hd // 8):
    exp[c] = base[0:64, c * 8:(c + 1) * 8]
print("layout ok:", torch.equal(img, exp), "pad chunk zero:", bool((img[hd // 8:] == 0).all()))
for grid, tiles in [(148, 256), (296, 128), (592, 64)]:
    for _ in range(2):
        lib.dlb_tma_gather_probe(base.data_ptr(), rows, ld, H, hd, grid, tiles, None, st)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        lib.dlb_tma_gather_probe(base.data_ptr(), rows, ld, H, hd, grid, tiles, None, st)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 5
    gb = grid * tiles * 64 * hd * 2 / 1e9
    print(f"grid {grid} x {tiles} tiles: {ms:.3f} ms, {gb / ms * 1e3:.0f} GB/s useful, {gb / ms * 1e3 / grid * 1e9 / 1.9e9 / 1e9 * 1e0:.1f} B/clk/CTA")
