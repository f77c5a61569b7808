"""MdAssignSampleRcnn / MdAssignSample alone at config-2 sizes, CUDA events over many calls."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from minddet_b200 import pipeline

iters = int(sys.argv[1]) if len(sys.argv) > 1 else 50
rp = pipeline.RegionPath(seed=0)
inp = pipeline.make_inputs(8, seed=0xD37)
dev = pipeline.to_device({k: v for k, v in inp.items() if k not in ("feats", "dout")})
anchors, avalid = rp.anchors()
props, pmask = rp.proposal(dev["cls_scores"], dev["bbox_preds"])
fns = {"rcnn": lambda: rp.rcnn_targets(dev["gts"], dev["gt_labels"], pmask, props, dev["gt_valid"]),
       "rpn": lambda: rp.rpn_targets(dev["gts"], dev["gt_valid"], anchors, avalid)}
out = []
for name, fn in fns.items():
    for _ in range(5):
        fn()
    torch.cuda.synchronize()
    best = 1e9
    for rep in range(5):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(iters):
            fn()
        e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1) / iters * 1e3)
    out.append(f"{name} {best:.1f} us")
print(f"{os.path.basename(os.environ.get('MD_REGION_LIB', 'default'))}: " + ", ".join(out))
