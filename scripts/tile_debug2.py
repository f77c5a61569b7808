"""Development helper: single-RoI tile-stationary backward vs the oracle, printing where they differ."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import oracle as O
from minddet_b200 import synth
from minddet_b200.ops import SingleRoIExtractor

C = 32
B = 2
shapes = synth.level_shapes()[:4]
strides = synth.STRIDES[:4]
ext = SingleRoIExtractor()
for name, box in (("dense", [1, 100, 100, 140, 140]), ("wide", [0, 100, 100, 400, 300]), ("two", None)):
    if box is None:
        rois = np.array([[1, 100, 100, 140, 140], [1, 110, 104, 150, 150]], np.float32)
    else:
        rois = np.array([box], np.float32)
    R = rois.shape[0]
    rng = np.random.default_rng(1)
    dout = rng.uniform(-1, 1, (R, C, 7, 7)).astype(np.float32)
    got = ext._backward(torch.from_numpy(rois).cuda(), torch.from_numpy(dout).cuda(), [(B, C, h, w) for h, w in shapes])
    torch.cuda.synchronize()
    dref = O.roialign_bwd([(B, C, h, w) for h, w in shapes], strides, rois, dout)
    for l in range(4):
        g = got[l].cpu().numpy()
        bad = np.argwhere(np.abs(g - dref[l]) > 1e-5)
        print(name, "level", l, "mismatches", len(bad), "nonzero ref", int((dref[l] != 0).sum()), "nonzero got", int((g != 0).sum()))
        for idx in bad[:6]:
            print("   ", tuple(idx), "got", g[tuple(idx)], "ref", dref[l][tuple(idx)])
        if len(bad):
            ch0 = bad[bad[:, 1] == bad[0, 1]]
            print("    channel", bad[0, 1], "rows", sorted(set(ch0[:, 2])), "cols", sorted(set(ch0[:, 3])))
            nz = np.argwhere(dref[l][bad[0, 0], bad[0, 1]] != 0)
            print("    ref footprint rows", nz[:, 0].min(), nz[:, 0].max(), "cols", nz[:, 1].min(), nz[:, 1].max())
