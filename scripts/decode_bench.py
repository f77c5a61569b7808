"""Stand-alone streaming kernels against the HBM roofline: a2 decode-all (MdDecodeLevel: 16 B deltas read + 16 B box
written per anchor, anchors regenerated in registers) and a1 anchor grid (16 B written per anchor), at B=64 images of
config-2 level 0 (200x336x3 anchors) so that the working set (2 x 413 MB) exceeds L2.  python scripts/decode_bench.py"""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from minddet_b200 import AnchorGenerator, BoundingBoxDecode, synth

iters = int(sys.argv[1]) if len(sys.argv) > 1 else 20
B, (H, W), stride, A = 64, synth.level_shapes()[0], 4, 3
base = torch.from_numpy(synth.base_anchor_sets()[0]).cuda()
g = torch.Generator(device="cuda").manual_seed(0)
deltas = [torch.randn(B, 4 * A, H, W, device="cuda", generator=g) * 0.2 for _ in range(2)]
dec = BoundingBoxDecode((800, 1344))
gen = AnchorGenerator(4, [8], [0.5, 1.0, 2.0])
peak = 6650.0
pk = os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json")
if os.path.exists(pk):
    peak = json.load(open(pk))["hbm_gbs"]


def timeit(fn):
    for i in range(3):
        fn(i)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(iters):
        fn(i)
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters * 1e3


n = B * H * W * A
t_dec = timeit(lambda i: dec.decode_level(deltas[i & 1], base, stride))
t_anc = timeit(lambda i: gen.grid_anchors((H * 8, W * 8), 4))          # 1600x2688 cells x 3 = 12.9 M anchors, 206 MB written
n_anc = H * 8 * W * 8 * 3
out = {"decode_level": {"us": t_dec, "anchors": n, "algorithmic_bytes": n * 32, "GBps": n * 32 / (t_dec * 1e-6) / 1e9},
       "anchor_grid": {"us": t_anc, "anchors": n_anc, "algorithmic_bytes": n_anc * 16, "GBps": n_anc * 16 / (t_anc * 1e-6) / 1e9}}
for v in out.values():
    v["roofline_frac"] = v["GBps"] / peak
out["peak_GBps"] = peak
print(json.dumps(out))
