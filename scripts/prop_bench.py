"""Proposal + RPN/RCNN assign stages alone at config-2 sizes (8 images); for ncu captures.
python scripts/prop_bench.py [iters]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from minddet_b200 import pipeline

iters = int(sys.argv[1]) if len(sys.argv) > 1 else 5
rp = pipeline.RegionPath(seed=0)
inp = pipeline.make_inputs(8, seed=0xD37)
dev = pipeline.to_device({k: v for k, v in inp.items() if k not in ("feats", "dout")})
anchors, avalid = rp.anchors()


def run():
    props, pmask = rp.proposal(dev["cls_scores"], dev["bbox_preds"])
    rpn = rp.rpn_targets(dev["gts"], dev["gt_valid"], anchors, avalid)
    rcnn = rp.rcnn_targets(dev["gts"], dev["gt_labels"], pmask, props, dev["gt_valid"])
    return props, rpn, rcnn


for _ in range(2):
    run()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(iters):
    run()
e1.record()
torch.cuda.synchronize()
print(f"proposal+assign: {e0.elapsed_time(e1) / iters * 1e3:.1f} us / call")
