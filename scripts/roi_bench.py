"""RoIAlign fwd/bwd alone at config-2 sizes (8 images, 4096 RoIs from the real proposal/assign path).
Used for ncu captures: python scripts/roi_bench.py [iters]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from minddet_b200 import pipeline

iters = int(sys.argv[1]) if len(sys.argv) > 1 else 5
rp = pipeline.RegionPath(seed=0)
dev = pipeline.to_device(pipeline.make_inputs(8, seed=0xD37))
out = rp.forward(dev["cls_scores"], dev["bbox_preds"], dev["feats"], dev["gts"], dev["gt_labels"], dev["gt_valid"])
rois = out["rois"]
shapes = [tuple(f.shape) for f in dev["feats"]]
torch.cuda.synchronize()
for name, fn in (("fwd", lambda: rp.extractor._forward(rois, dev["feats"])),
                 ("bwd", lambda: rp.backward(rois, dev["dout"], shapes))):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    print(f"{name}: {e0.elapsed_time(e1) / iters * 1e3:.1f} us / call")
lv = rp.extractor.map_roi_levels(rois).cpu().numpy()
r = rois.cpu().numpy()
print("rois per level", np.bincount(lv, minlength=4), "median w,h px", np.median(r[:, 3] - r[:, 1]), np.median(r[:, 4] - r[:, 2]))
