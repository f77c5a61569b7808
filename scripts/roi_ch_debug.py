"""Debug harness for the channel-lane RoIAlign kernels: small controlled cases against the oracle, with error structure
printed (per RoI, per channel, per bin).  python scripts/roi_ch_debug.py [C]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import oracle as O
from minddet_b200 import synth
from minddet_b200.ops import SingleRoIExtractor

C = int(sys.argv[1]) if len(sys.argv) > 1 else 32
B = 2
np.set_printoptions(linewidth=200, precision=4, suppress=True)
shapes = synth.level_shapes()[:4]
strides = synth.STRIDES[:4]
rng = np.random.default_rng(5)
feats = [rng.uniform(-1, 1, (B, C, h, w)).astype(np.float32) for h, w in shapes]
cases = {
    "nq1 tiny": [0, 300, 300, 301, 301],
    "nq2 28px": [0, 40, 40, 67, 67],
    "nq3": [1, 100, 100, 139, 139],
    "nq4": [0, 200, 100, 255, 160],
    "level1": [1, 64, 64, 175, 175],
    "tall": [0, 100, 5, 130, 790],
    "wide 3 chunks": [1, 10, 10, 450, 120],
    "edge tl": [0, -30, -30, 40, 50],
    "edge br": [1, 1300, 760, 1343, 799],
    "outside": [0, -500, -500, -300, -300],
    "bad batch": [-1, 10, 10, 50, 50],
    "whole image": [0, 0, 0, 1343, 799],
}
names = list(cases)
rois = np.array([cases[n] for n in names], np.float32)
ext = SingleRoIExtractor(7, 2, strides, 56)
dv = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()
ft = [dv(f) for f in feats]
out = ext._forward(dv(rois), ft).cpu().numpy()
ref = O.roialign_fwd(feats, strides, rois)
lv = O.roi_levels(rois, 56.0, 4)
print("== forward")
bad = None
for i, n in enumerate(names):
    err = np.abs(out[i] - ref[i])
    tol = 1e-6 + 1e-5 * np.abs(ref[i])
    nb = int((err > tol).sum())
    print(f"{n:16s} level {lv[i]} max|err| {err.max():.3e} bad {nb}/{err.size}")
    if nb and bad is None:
        bad = i
if bad is not None:
    err = np.abs(out[bad] - ref[bad])
    print("first bad RoI:", names[bad], rois[bad])
    print("per-channel max err:", err.reshape(C, -1).max(1))
    cbad = int(err.reshape(C, -1).max(1).argmax())
    print(f"channel {cbad} got:\n", out[bad, cbad], "\nref:\n", ref[bad, cbad])
# backward
dout = rng.uniform(-1, 1, ref.shape).astype(np.float32)
grads = ext._backward(dv(rois), dv(dout), [f.shape for f in feats])
dref = O.roialign_bwd([f.shape for f in feats], strides, rois, dout)
print("== backward (all RoIs)")
for l in range(4):
    e = np.abs(grads[l].cpu().numpy() - dref[l])
    print(f"level {l}: max|err| {e.max():.3e} (scale {np.abs(dref[l]).max():.3f}) bad {(e > 1e-5 * max(1, np.abs(dref[l]).max())).sum()}")
for i, n in enumerate(names):
    g1 = ext._backward(dv(rois[i:i + 1]), dv(dout[i:i + 1]), [f.shape for f in feats])
    d1 = O.roialign_bwd([f.shape for f in feats], strides, rois[i:i + 1], dout[i:i + 1])
    e = max(np.abs(g1[l].cpu().numpy() - d1[l]).max() for l in range(4))
    print(f"   {n:16s} max|err| {e:.3e}")
