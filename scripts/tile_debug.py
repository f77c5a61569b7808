"""Development helper: one tile-stationary backward call on a small problem, compared with the oracle."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import oracle as O
from minddet_b200 import synth
from minddet_b200.ops import SingleRoIExtractor

R = int(sys.argv[1]) if len(sys.argv) > 1 else 160
C = int(sys.argv[2]) if len(sys.argv) > 2 else 64
rng = np.random.default_rng(42)
B = 2
shapes = synth.level_shapes()[:4]
strides = synth.STRIDES[:4]
b = synth.rand_boxes(rng, R, smin=4, smax=900)
rois = np.concatenate([rng.integers(0, B, (R, 1)).astype(np.float32), b], 1).astype(np.float32)
dout = rng.uniform(-1, 1, (R, C, 7, 7)).astype(np.float32)
ext = SingleRoIExtractor()
got = ext._backward(torch.from_numpy(rois).cuda(), torch.from_numpy(dout).cuda(), [(B, C, h, w) for h, w in shapes])
torch.cuda.synchronize()
print("ran")
dref = O.roialign_bwd([(B, C, h, w) for h, w in shapes], strides, rois, dout)
for l in range(4):
    g = got[l].cpu().numpy()
    print(l, "max err", np.abs(g - dref[l]).max(), "of", np.abs(dref[l]).max(), "nan", np.isnan(g).sum())
