"""per-level top-k timing (B=8): python scripts/topk_bench.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from minddet_b200 import TopKPerLevel, synth
op = TopKPerLevel(2000, apply_sigmoid=True)
for (h, w) in synth.level_shapes():
    x = torch.randn(8, 3, h, w, device="cuda") * 2 - 4
    for _ in range(3):
        op(x)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20):
        op(x)
    e1.record()
    torch.cuda.synchronize()
    print(f"level {h}x{w}: {e0.elapsed_time(e1) / 20 * 1e3:.1f} us")
