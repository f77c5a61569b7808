"""ctypes driver that calls libmdregion.so exactly the way MindSpore's ``ops.Custom(func_type="aot")``
runtime does: ``int f(int nparam, void** params, int* ndims, int64_t** shapes, const char** dtypes,
void* stream, void* extra)`` with device pointers, outputs pre-allocated from ``out_shape`` /
``out_dtype`` and the framework's current CUDA stream (reference call sites:
centerpoint/det3d_ms/ops/test_custom_pytorch/iou_gpu.py:17-22,55-60; nms_cpu.py:10-27).

MindSpore is not installable in this environment, so torch supplies device memory and streams
(plumbing only).  There is NO fallback: if the library is missing or a call fails, this raises.
"""
import ctypes
import os

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("MD_REGION_LIB", os.path.join(_HERE, "lib", "libmdregion.so"))   # override: experiments only

_DTYPE_NAMES = {
    torch.float32: "float32", torch.int32: "int32", torch.int64: "int64", torch.uint8: "uint8",
    torch.bool: "bool", torch.float16: "float16", torch.int8: "int8",
}

SYMBOLS = ("MdAnchorGrid", "MdDecodeClip", "MdDecodeLevel", "MdTopKPerLevel", "MdNms", "MdProposal",
           "MdAssignSample", "MdAssignSampleRcnn", "MdRoiLevels", "MdRoiAlignFwd", "MdRoiAlignBwd", "MdRoiAlignBwdAcc",
           "MdRoiAlignFwdExact", "MdRoiAlignBwdExact", "MdRoiAlignBwdPrepare", "MdRoiAlignBwdPlanned",
           # the reference's own GPU symbols (iou3d_nms_kernel.cu:445-601) + the device twin of boxes_iou_nms_cpu
           "BoxesIouBevGpu", "BoxesOverlapBevGpu", "NmsGpu", "NmsNormalGpu", "BoxesIouNmsGpu",
           "MdYoloDecode", "MdYoloNms", "MdMaskTargets", "MdEncode", "MdRcnnPostProcess")

ERRORS = {1: "wrong nparam", 2: "bad dtype/shape", 3: "CUDA error", 4: "unsupported size",
          5: "workspace must grow while the stream is being captured (run the op once eagerly first)"}

_lib = None


def load_library(path=None):
    """dlopen libmdregion.so (built by minddet_b200/build.py).  Raises if it is not there."""
    global _lib
    if _lib is None or path is not None:
        p = path or LIB_PATH
        if not os.path.exists(p):
            raise ImportError(
                f"{p} not found: build it with `python -m minddet_b200.build` (nvcc, sm_100a). "
                "minddet_b200 has no CPU or eager fallback.")
        lib = ctypes.CDLL(p)
        lib.MdVersion.restype = ctypes.c_char_p
        for s in SYMBOLS:
            fn = getattr(lib, s)
            fn.restype = ctypes.c_int
            fn.argtypes = [ctypes.c_int, ctypes.POINTER(ctypes.c_void_p), ctypes.POINTER(ctypes.c_int),
                           ctypes.POINTER(ctypes.POINTER(ctypes.c_int64)), ctypes.POINTER(ctypes.c_char_p),
                           ctypes.c_void_p, ctypes.c_void_p]
        lib.MdRoiAlignPlanBytes.restype = ctypes.c_int64
        lib.MdRoiAlignPlanBytes.argtypes = [ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.POINTER(ctypes.c_int),
                                            ctypes.POINTER(ctypes.c_int)]
        if path is not None:
            return lib
        _lib = lib
    return _lib


class AotError(RuntimeError):
    pass


def call_aot(symbol, inputs, outputs, stream=None, lib=None):
    """Invoke one aot symbol on `inputs + outputs` (torch CUDA tensors, contiguous)."""
    lib = lib or load_library()
    tensors = list(inputs) + list(outputs)
    n = len(tensors)
    for t in tensors:
        if not t.is_cuda:
            raise AotError(f"{symbol}: all tensors must live on the GPU (got {t.device}); there is no CPU path")
        if not t.is_contiguous():
            raise AotError(f"{symbol}: tensors must be contiguous")
    params = (ctypes.c_void_p * n)(*[t.data_ptr() for t in tensors])
    ndims = (ctypes.c_int * n)(*[t.dim() for t in tensors])
    shape_arrays = [(ctypes.c_int64 * max(1, t.dim()))(*t.shape) for t in tensors]
    shapes = (ctypes.POINTER(ctypes.c_int64) * n)(*[ctypes.cast(a, ctypes.POINTER(ctypes.c_int64)) for a in shape_arrays])
    dtypes = (ctypes.c_char_p * n)(*[_DTYPE_NAMES[t.dtype].encode() for t in tensors])
    dev = tensors[0].device
    if any(t.device != dev for t in tensors):
        raise AotError(f"{symbol}: all tensors must live on one device")
    if stream is None:
        stream = torch.cuda.current_stream(dev).cuda_stream
    with torch.cuda.device(dev):          # the library keys its workspace on the CURRENT device (as MindSpore binds it)
        rc = getattr(lib, symbol)(n, params, ndims, shapes, dtypes, ctypes.c_void_p(stream), None)
    if rc != 0:
        raise AotError(f"{symbol} returned {rc} ({ERRORS.get(rc, 'unknown')})")
    return outputs


class Custom:
    """Stand-in for ``mindspore.ops.Custom("<so>:<symbol>", out_shape, out_dtype, func_type="aot")``.

    out_shape: callable(*input_shapes) -> shape or tuple of shapes; out_dtype: dtype, tuple, or callable.
    """

    def __init__(self, func, out_shape, out_dtype, func_type="aot"):
        if func_type != "aot":
            raise ValueError("only func_type='aot' exists here")
        so, _, symbol = func.rpartition(":")
        self.symbol = symbol
        self.lib = load_library(so) if so and os.path.abspath(so) != LIB_PATH else load_library()
        self.out_shape, self.out_dtype = out_shape, out_dtype

    def __call__(self, *inputs):
        shapes = self.out_shape(*[tuple(t.shape) for t in inputs]) if callable(self.out_shape) else self.out_shape
        dts = self.out_dtype(*[t.dtype for t in inputs]) if callable(self.out_dtype) else self.out_dtype
        single = not isinstance(shapes[0], (tuple, list)) if len(shapes) else True
        if single:
            shapes, dts = (shapes,), (dts,)
        elif not isinstance(dts, (tuple, list)):
            dts = (dts,) * len(shapes)
        dev = inputs[0].device
        outs = [torch.empty(tuple(int(d) for d in s), dtype=dt, device=dev) for s, dt in zip(shapes, dts)]
        call_aot(self.symbol, inputs, outs, lib=self.lib)
        return outs[0] if single else tuple(outs)
