"""Mask-head RoIAlign (config 4): 14x14, 128 positive RoIs per image, 8 images, 256 channels: fwd + bwd timing."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from minddet_b200 import SingleRoIExtractor, synth
rng = np.random.default_rng(0)
B, C, R = 8, 256, 1024
feats = [torch.rand(B, C, h, w, device="cuda") for h, w in synth.level_shapes()[:4]]
gts, _, valid = synth.gt_boxes(B, G=32, seed=1)
rois = np.zeros((R, 5), np.float32)
for r in range(R):
    b = r // 128
    g = gts[b, rng.integers(0, max(1, int(valid[b].sum())))]
    rois[r] = [b, *(g + rng.normal(0, 4, 4))]
rois = torch.from_numpy(rois).cuda()
ext = SingleRoIExtractor(14, 2)
dout = torch.rand(R, C, 14, 14, device="cuda")
shapes = [tuple(f.shape) for f in feats]
for name, fn in (("mask fwd", lambda: ext._forward(rois, feats)), ("mask bwd", lambda: ext._backward(rois, dout, shapes))):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        fn()
    e1.record()
    torch.cuda.synchronize()
    print(f"{name}: {e0.elapsed_time(e1) / 10 * 1e3:.1f} us / call")
