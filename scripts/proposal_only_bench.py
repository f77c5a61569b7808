"""MdProposal alone at config-2 sizes (8 images), CUDA events over many calls: python scripts/proposal_only_bench.py [iters]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from minddet_b200 import pipeline

iters = int(sys.argv[1]) if len(sys.argv) > 1 else 50
rp = pipeline.RegionPath(seed=0)
inp = pipeline.make_inputs(8, seed=0xD37)
dev = pipeline.to_device({k: v for k, v in inp.items() if k in ("cls_scores", "bbox_preds")})
for _ in range(5):
    rp.proposal(dev["cls_scores"], dev["bbox_preds"])
torch.cuda.synchronize()
best = 1e9
for rep in range(5):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        rp.proposal(dev["cls_scores"], dev["bbox_preds"])
    e1.record()
    torch.cuda.synchronize()
    best = min(best, e0.elapsed_time(e1) / iters * 1e3)
print(f"{os.environ.get('MD_REGION_LIB', 'default')}: proposal {best:.1f} us / call (best of 5 x {iters})")
