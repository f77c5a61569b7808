"""Tiny RoIAlign call for debugging under compute-sanitizer."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from minddet_b200 import SingleRoIExtractor
rng = np.random.default_rng(0)
B, C = 1, 8
feats = [torch.rand(B, C, h, w, device="cuda") for h, w in [(200, 336), (100, 168), (50, 84), (25, 42)]]
rois = torch.tensor([[0, 100, 100, 180, 170], [0, 10, 10, 300, 200]], dtype=torch.float32, device="cuda")
ext = SingleRoIExtractor()
out = ext(rois, *feats)
torch.cuda.synchronize()
print("fwd ok", out.abs().sum().item())
g = ext._backward(rois, torch.ones_like(out), [tuple(f.shape) for f in feats])
torch.cuda.synchronize()
print("bwd ok", [x.sum().item() for x in g])
