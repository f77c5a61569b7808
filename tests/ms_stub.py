"""A stand-in for the five names ``minddet_b200/mindspore_ops.py`` takes from MindSpore (``Tensor``, ``nn.Cell``,
``ops.Custom``, ``ops.zeros_like``, ``dtype``), so that file can be EXECUTED on the GPU box without MindSpore.

``ops.Custom(func, out_shape, out_dtype, func_type="aot", bprop=None)`` forwards to ``minddet_b200._aot.Custom``, which
marshals what MindSpore's runtime passes to an aot symbol: (nparam, params, ndims, shapes, dtypes, stream, extra) with
device pointers, outputs pre-allocated from out_shape / out_dtype, the current stream (reference call sites:
centerpoint/det3d_ms/ops/test_custom_pytorch/iou_gpu.py:55-60, ops/nms_cpu.py:10-27).  ``bprop`` follows MindSpore's
convention ``bprop(*inputs, out, dout) -> grads per input`` and is wired through ``torch.autograd``.
Test infrastructure only."""
import sys
import types

import numpy as np
import torch


def install():
    if "mindspore" in sys.modules and getattr(sys.modules["mindspore"], "__md_stub__", False):
        return sys.modules["mindspore"]
    from minddet_b200 import _aot

    ms = types.ModuleType("mindspore")
    ms.__md_stub__ = True

    def Tensor(a, dtype=None):
        t = torch.from_numpy(np.ascontiguousarray(a)) if isinstance(a, np.ndarray) else torch.as_tensor(a)
        if dtype is not None:
            t = t.to(dtype)
        return t.cuda()

    class Cell:
        def __init__(self):
            pass

        def __call__(self, *a, **k):
            return self.construct(*a, **k)

    class Custom:
        def __init__(self, func, out_shape=None, out_dtype=None, func_type="aot", bprop=None):
            assert func_type == "aot"
            self.inner = _aot.Custom(func, out_shape, out_dtype, func_type)
            self.bprop = bprop

        def __call__(self, *inputs):
            if self.bprop is None or not any(t.requires_grad for t in inputs):
                return self.inner(*inputs)
            inner, bprop = self.inner, self.bprop

            class Fn(torch.autograd.Function):
                @staticmethod
                def forward(ctx, *xs):
                    out = inner(*[x.detach() for x in xs])
                    ctx.save_for_backward(*xs, out)
                    return out

                @staticmethod
                def backward(ctx, dout):
                    *xs, out = ctx.saved_tensors
                    grads = bprop(*xs, out, dout.contiguous())
                    return tuple(g if x.requires_grad else None for g, x in zip(grads, xs))

            return Fn.apply(*inputs)

    dtype = types.ModuleType("mindspore.dtype")
    dtype.float32, dtype.int32, dtype.int64, dtype.bool_, dtype.uint8, dtype.float16 = (
        torch.float32, torch.int32, torch.int64, torch.bool, torch.uint8, torch.float16)
    nn = types.ModuleType("mindspore.nn")
    nn.Cell = Cell
    ops = types.ModuleType("mindspore.ops")
    ops.Custom = Custom
    ops.zeros_like = torch.zeros_like
    ms.Tensor, ms.nn, ms.ops, ms.dtype = Tensor, nn, ops, dtype
    sys.modules.update({"mindspore": ms, "mindspore.nn": nn, "mindspore.ops": ops, "mindspore.dtype": dtype})
    return ms
