"""ctypes/numpy binding of oracle/region_oracle.c (TEST INFRASTRUCTURE ONLY, see __init__)."""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "liboracle.so")
_lib = None

f32p = C.POINTER(C.c_float)
i32p = C.POINTER(C.c_int32)
u8p = C.POINTER(C.c_uint8)

MAX_RATIO = float(np.float32(abs(np.log(0.016))))


def build(force=False):
    """Compile liboracle.so (and oracle/_ref when /root/reference exists)."""
    src = os.path.join(_HERE, "region_oracle.c")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "_build/liboracle.so"], stdout=subprocess.DEVNULL)
    ref_so = os.path.join(_HERE, "_ref", "nms_fast_ref.so")
    ref_so2 = os.path.join(_HERE, "_ref", "iou3d_device_ref.so")   # the reference's __device__ functions on the host (golden generators only)
    if os.path.isdir("/root/reference") and (force or not os.path.exists(ref_so) or not os.path.exists(ref_so2)):
        subprocess.call(["make", "-C", _HERE, "ref"], stdout=subprocess.DEVNULL)


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = C.CDLL(_SO)
        _lib.o_exp.restype = C.c_float
        _lib.o_exp.argtypes = [C.c_float]
        _lib.o_sigmoid.restype = C.c_float
        _lib.o_sigmoid.argtypes = [C.c_float]
        _lib.o_philox_key.restype = C.c_uint32
        _lib.o_philox_key.argtypes = [C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint64]
        _lib.o_nms.restype = C.c_int
        _lib.o_sample.restype = C.c_int64
        _lib.o_sample_step.restype = C.c_int64
        _lib.o_proposal_image.restype = C.c_int
        _lib.o_assign_sample_rpn.restype = C.c_int
        _lib.o_assign_sample_rcnn.restype = C.c_int
    return _lib


def _f(a):
    return np.ascontiguousarray(a, dtype=np.float32)


def _p(a, t):
    return a.ctypes.data_as(t) if a is not None else None


class ProposalCfg(C.Structure):
    _fields_ = [("img_h", C.c_float), ("img_w", C.c_float), ("means", C.c_float * 4),
                ("stds", C.c_float * 4), ("max_ratio", C.c_float), ("nms_thr", C.c_float),
                ("nms_off", C.c_float), ("union_eps", C.c_float), ("nms_inclusive", C.c_int),
                ("apply_sigmoid", C.c_int), ("nms_pre", C.c_int), ("max_num", C.c_int)]


class AssignCfg(C.Structure):
    _fields_ = [("pos_thr", C.c_float), ("neg_thr", C.c_float), ("min_pos_iou", C.c_float),
                ("iou_off", C.c_float), ("mode", C.c_int), ("pos_slots", C.c_int),
                ("neg_slots", C.c_int), ("num_total", C.c_int), ("means", C.c_float * 4),
                ("stds", C.c_float * 4), ("seed", C.c_uint64), ("step", C.c_uint32), ("pad_", C.c_uint32)]


class RegionCfg(C.Structure):
    _fields_ = [("prop", ProposalCfg), ("rpn", AssignCfg), ("rcnn", AssignCfg),
                ("finest_scale", C.c_float), ("roi_P", C.c_int), ("roi_S", C.c_int),
                ("roi_end_mode", C.c_float), ("num_roi_levels", C.c_int), ("do_backward", C.c_int)]


def proposal_cfg(img_h, img_w, nms_pre=2000, max_num=2000, nms_thr=0.7, means=(0, 0, 0, 0),
                 stds=(1, 1, 1, 1), max_ratio=MAX_RATIO, nms_off=0.0, union_eps=1e-8,
                 nms_inclusive=False, apply_sigmoid=True):
    c = ProposalCfg()
    c.img_h, c.img_w = img_h, img_w
    c.means[:] = list(means)
    c.stds[:] = list(stds)
    c.max_ratio, c.nms_thr, c.nms_off, c.union_eps = max_ratio, nms_thr, nms_off, union_eps
    c.nms_inclusive, c.apply_sigmoid, c.nms_pre, c.max_num = int(nms_inclusive), int(apply_sigmoid), nms_pre, max_num
    return c


def assign_cfg(pos_thr, neg_thr, min_pos_iou, pos_slots, neg_slots, num_total, means=(0, 0, 0, 0),
               stds=(1, 1, 1, 1), seed=0, iou_off=1.0, mode=0, step=0):
    c = AssignCfg()
    c.pos_thr, c.neg_thr, c.min_pos_iou, c.iou_off, c.mode = pos_thr, neg_thr, min_pos_iou, iou_off, mode
    c.pos_slots, c.neg_slots, c.num_total, c.seed, c.step = pos_slots, neg_slots, num_total, seed, step
    c.means[:] = list(means)
    c.stds[:] = list(stds)
    return c


# ------------------------------------------------------------------------------------------------
def exp(x):
    x = _f(x)
    y = np.empty_like(x)
    lib().o_exp_array(_p(x, f32p), _p(y, f32p), C.c_int64(x.size))
    return y


def sigmoid(x):
    x = _f(x)
    y = np.empty_like(x)
    lib().o_sigmoid_array(_p(x, f32p), _p(y, f32p), C.c_int64(x.size))
    return y


def philox_key(n, stream, image, seed):
    L = lib()
    return np.array([L.o_philox_key(int(i), stream, image, seed) for i in np.atleast_1d(n)], dtype=np.uint32)


def anchor_grid(base, H, W, stride):
    base = _f(base)
    A = base.shape[0]
    out = np.empty((H * W * A, 4), np.float32)
    lib().o_anchor_grid(_p(base, f32p), A, H, W, C.c_float(stride), _p(out, f32p))
    return out


def decode(anchors, deltas, img_h, img_w, means=(0, 0, 0, 0), stds=(1, 1, 1, 1), max_ratio=MAX_RATIO):
    anchors, deltas = _f(anchors), _f(deltas)
    m, s = _f(means), _f(stds)
    out = np.empty_like(anchors)
    lib().o_decode(_p(anchors, f32p), _p(deltas, f32p), C.c_int64(anchors.shape[0]), _p(m, f32p),
                   _p(s, f32p), C.c_float(max_ratio), C.c_float(img_h), C.c_float(img_w), _p(out, f32p))
    return out


def decode_level_nchw(base, H, W, stride, deltas_nchw, img_h, img_w, means=(0, 0, 0, 0),
                      stds=(1, 1, 1, 1), max_ratio=MAX_RATIO):
    base, d = _f(base), _f(deltas_nchw)
    A = base.shape[0]
    m, s = _f(means), _f(stds)
    out = np.empty((H * W * A, 4), np.float32)
    lib().o_decode_level_nchw(_p(base, f32p), A, H, W, C.c_float(stride), _p(d, f32p), _p(m, f32p),
                              _p(s, f32p), C.c_float(max_ratio), C.c_float(img_h), C.c_float(img_w),
                              _p(out, f32p))
    return out


def encode(props, gts, means=(0, 0, 0, 0), stds=(1, 1, 1, 1)):
    props, gts = _f(props), _f(gts)
    m, s = _f(means), _f(stds)
    out = np.empty_like(props)
    lib().o_encode(_p(props, f32p), _p(gts, f32p), C.c_int64(props.shape[0]), _p(m, f32p), _p(s, f32p), _p(out, f32p))
    return out


def topk(scores, K):
    scores = _f(scores).ravel()
    K = min(K, scores.size)
    v = np.empty(K, np.float32)
    i = np.empty(K, np.int32)
    lib().o_topk(_p(scores, f32p), C.c_int64(scores.size), C.c_int64(K), _p(v, f32p), _p(i, i32p))
    return v, i


def level_scores(logits_nchw, apply_sigmoid=True):
    x = _f(logits_nchw)
    A, H, W = x.shape
    out = np.empty(A * H * W, np.float32)
    lib().o_level_scores(_p(x, f32p), A, H, W, int(apply_sigmoid), _p(out, f32p))
    return out


def nms(boxes, thr, off=0.0, inclusive=False, union_eps=1e-8):
    """boxes (K, >=4) score-sorted.  returns keep mask (K,) uint8."""
    boxes = _f(boxes)
    K, ld = boxes.shape
    keep = np.zeros(K, np.uint8)
    lib().o_nms(_p(boxes, f32p), ld, K, C.c_float(thr), C.c_float(off), int(inclusive),
                C.c_float(union_eps), _p(keep, u8p))
    return keep


def proposal_image(levels, cfg):
    """levels: list of (logits (A,H,W), deltas (4A,H,W), base (A,4), stride)."""
    L = len(levels)
    lg = [_f(x[0]) for x in levels]
    dl = [_f(x[1]) for x in levels]
    bs = [_f(x[2]) for x in levels]
    A = (C.c_int * L)(*[x.shape[0] for x in lg])
    H = (C.c_int * L)(*[x.shape[1] for x in lg])
    W = (C.c_int * L)(*[x.shape[2] for x in lg])
    st = (C.c_float * L)(*[float(x[3]) for x in levels])
    sumK = sum(min(cfg.nms_pre, x.size) for x in lg)
    pl = (f32p * L)(*[_p(x, f32p) for x in lg])
    pd = (f32p * L)(*[_p(x, f32p) for x in dl])
    pb = (f32p * L)(*[_p(x, f32p) for x in bs])
    out = dict(idx=np.empty(sumK, np.int32), score=np.empty(sumK, np.float32),
               box=np.empty((sumK, 4), np.float32), keep=np.empty(sumK, np.uint8),
               props=np.empty((cfg.max_num, 5), np.float32), mask=np.empty(cfg.max_num, np.uint8))
    lib().o_proposal_image(L, pl, pd, pb, A, H, W, st, C.byref(cfg), _p(out["idx"], i32p),
                           _p(out["score"], f32p), _p(out["box"], f32p), _p(out["keep"], u8p),
                           _p(out["props"], f32p), _p(out["mask"], u8p))
    return out


def iou_matrix(boxes, gts, off=1.0):
    boxes, gts = _f(boxes), _f(gts)
    out = np.empty((boxes.shape[0], gts.shape[0]), np.float32)
    lib().o_iou_matrix(_p(boxes, f32p), C.c_int64(boxes.shape[0]), _p(gts, f32p), gts.shape[0],
                       C.c_float(off), _p(out, f32p))
    return out


def _u8(a):
    return None if a is None else np.ascontiguousarray(a, dtype=np.uint8)


def assign(boxes, gts, pos_thr, neg_thr, min_pos_iou, valid=None, gt_valid=None, off=1.0, mode=0):
    boxes, gts = _f(boxes).reshape(-1, 4), _f(gts).reshape(-1, 4)
    valid, gt_valid = _u8(valid), _u8(gt_valid)
    N, G = boxes.shape[0], gts.shape[0]
    a = np.empty(N, np.int32)
    m = np.empty(N, np.float32)
    am = np.empty(N, np.int32)
    lib().o_assign(_p(boxes, f32p), _p(valid, u8p), C.c_int64(N), _p(gts, f32p), _p(gt_valid, u8p), G,
                   C.c_float(pos_thr), C.c_float(neg_thr), C.c_float(min_pos_iou), C.c_float(off),
                   mode, _p(a, i32p), _p(m, f32p), _p(am, i32p))
    return a, m, am


def sample(assigned, want_positive, stream, image, seed, k_slots, step=0):
    assigned = np.ascontiguousarray(assigned, np.int32)
    out = np.empty(k_slots, np.int32)
    c = lib().o_sample_step(_p(assigned, i32p), C.c_int64(assigned.size), int(want_positive), C.c_uint32(stream),
                            C.c_uint32(image), C.c_uint64(seed), C.c_uint32(step), k_slots, _p(out, i32p))
    return out, int(c)


def assign_sample_rpn(anchors, gts, gt_valid, cfg, image, valid=None):
    anchors, gts = _f(anchors), _f(gts).reshape(-1, 4)
    valid, gt_valid = _u8(valid), _u8(gt_valid)
    N, G = anchors.shape[0], gts.shape[0]
    o = dict(assigned=np.empty(N, np.int32), pos_idx=np.empty(cfg.pos_slots, np.int32),
             pos_valid=np.empty(cfg.pos_slots, np.uint8), neg_idx=np.empty(cfg.neg_slots, np.int32),
             neg_valid=np.empty(cfg.neg_slots, np.uint8), pos_gt=np.empty(cfg.pos_slots, np.int32),
             pos_target=np.empty((cfg.pos_slots, 4), np.float32))
    o["num_pos"] = lib().o_assign_sample_rpn(
        _p(anchors, f32p), _p(valid, u8p), C.c_int64(N), _p(gts, f32p), _p(gt_valid, u8p), G, C.byref(cfg),
        C.c_uint32(image), _p(o["assigned"], i32p), _p(o["pos_idx"], i32p), _p(o["pos_valid"], u8p),
        _p(o["neg_idx"], i32p), _p(o["neg_valid"], u8p), _p(o["pos_gt"], i32p), _p(o["pos_target"], f32p))
    return o


def assign_sample_rcnn(props, prop_valid, gts, gt_labels, gt_valid, cfg, image):
    props, gts = _f(props).reshape(-1, 4), _f(gts).reshape(-1, 4)
    prop_valid, gt_valid = _u8(prop_valid), _u8(gt_valid)
    gt_labels = np.ascontiguousarray(gt_labels, np.int32)
    P, G = props.shape[0], gts.shape[0]
    S = cfg.pos_slots + cfg.neg_slots
    o = dict(assigned=np.empty(G + P, np.int32), sel_idx=np.empty(S, np.int32),
             rois=np.empty((S, 4), np.float32), deltas=np.empty((S, 4), np.float32),
             labels=np.empty(S, np.int32), mask=np.empty(S, np.uint8))
    o["num_pos"] = lib().o_assign_sample_rcnn(
        _p(props, f32p), _p(prop_valid, u8p), P, _p(gts, f32p), _p(gt_labels, i32p), _p(gt_valid, u8p), G,
        C.byref(cfg), C.c_uint32(image), _p(o["assigned"], i32p), _p(o["sel_idx"], i32p),
        _p(o["rois"], f32p), _p(o["deltas"], f32p), _p(o["labels"], i32p), _p(o["mask"], u8p))
    return o


def roi_levels(rois5, finest=56.0, num_levels=4):
    rois5 = _f(rois5)
    out = np.empty(rois5.shape[0], np.int32)
    lib().o_roi_levels(_p(rois5, f32p), C.c_int64(rois5.shape[0]), C.c_float(finest), num_levels, _p(out, i32p))
    return out


def _level_args(feats, strides):
    L = len(feats)
    H = (C.c_int * L)(*[x.shape[2] for x in feats])
    W = (C.c_int * L)(*[x.shape[3] for x in feats])
    sc = (C.c_float * L)(*[float(np.float32(1.0) / np.float32(s)) for s in strides])
    return L, H, W, sc


def roialign_fwd(feats, strides, rois5, P=7, S=2, end_mode=0.0, finest=56.0, lvl=None):
    feats = [_f(x) for x in feats]
    rois5 = _f(rois5)
    L, H, W, sc = _level_args(feats, strides)
    B, Cc = feats[0].shape[:2]
    R = rois5.shape[0]
    out = np.empty((R, Cc, P, P), np.float32)
    pf = (f32p * L)(*[_p(x, f32p) for x in feats])
    lv = None if lvl is None else np.ascontiguousarray(lvl, np.int32)
    lib().o_roialign_fwd(L, pf, H, W, sc, B, Cc, _p(rois5, f32p), C.c_int64(R), _p(lv, i32p),
                         C.c_float(finest), P, S, C.c_float(end_mode), _p(out, f32p))
    return out


def roialign_bwd(feat_shapes, strides, rois5, dout, P=7, S=2, end_mode=0.0, finest=56.0, lvl=None):
    dfe = [np.zeros(s, np.float32) for s in feat_shapes]
    rois5, dout = _f(rois5), _f(dout)
    L, H, W, sc = _level_args(dfe, strides)
    B, Cc = dfe[0].shape[:2]
    R = rois5.shape[0]
    pf = (f32p * L)(*[_p(x, f32p) for x in dfe])
    lv = None if lvl is None else np.ascontiguousarray(lvl, np.int32)
    lib().o_roialign_bwd(L, pf, H, W, sc, B, Cc, _p(rois5, f32p), C.c_int64(R), _p(lv, i32p),
                         C.c_float(finest), P, S, C.c_float(end_mode), _p(dout, f32p))
    return dfe


def region_path_batch(logits, deltas, bases, strides, feats, gts, gt_labels, gt_valid, cfg, dout=None, nthreads=1):
    """Whole path for a batch (CPU baseline driver).  logits[l] (B,A,H,W); deltas[l] (B,4A,H,W);
    feats[l] (B,C,H,W) for the RoI levels; gts (B,G,4).  Returns a dict of outputs."""
    L = len(logits)
    logits = [_f(x) for x in logits]
    deltas = [_f(x) for x in deltas]
    bases = [_f(x) for x in bases]
    feats = [_f(x) for x in feats]
    gts = _f(gts)
    gt_labels = np.ascontiguousarray(gt_labels, np.int32)
    gt_valid = _u8(gt_valid)
    B, G = gts.shape[:2]
    A = (C.c_int * L)(*[x.shape[1] for x in logits])
    H = (C.c_int * L)(*[x.shape[2] for x in logits])
    W = (C.c_int * L)(*[x.shape[3] for x in logits])
    st = (C.c_float * L)(*[float(s) for s in strides])
    Cc = feats[0].shape[1]
    Ntot = sum(x[0].size for x in logits)
    S = cfg.rcnn.pos_slots + cfg.rcnn.neg_slots
    M = cfg.prop.max_num
    P = cfg.roi_P
    o = dict(props=np.empty((B, M, 5), np.float32), pmask=np.empty((B, M), np.uint8),
             rpn_assigned=np.empty((B, Ntot), np.int32),
             rpn_pos_idx=np.empty((B, cfg.rpn.pos_slots), np.int32), rpn_pos_valid=np.empty((B, cfg.rpn.pos_slots), np.uint8),
             rpn_neg_idx=np.empty((B, cfg.rpn.neg_slots), np.int32), rpn_neg_valid=np.empty((B, cfg.rpn.neg_slots), np.uint8),
             rois=np.empty((B, S, 5), np.float32), roi_labels=np.empty((B, S), np.int32),
             roi_mask=np.empty((B, S), np.uint8), roi_deltas=np.empty((B, S, 4), np.float32),
             roi_feats=np.empty((B * S, Cc, P, P), np.float32))
    dfe = [np.zeros_like(x) for x in feats] if cfg.do_backward else []
    o["dfeats"] = dfe
    if cfg.do_backward:
        dout = _f(dout)
    nl = len(feats)
    pl = (f32p * L)(*[_p(x, f32p) for x in logits])
    pd = (f32p * L)(*[_p(x, f32p) for x in deltas])
    pb = (f32p * L)(*[_p(x, f32p) for x in bases])
    pf = (f32p * nl)(*[_p(x, f32p) for x in feats])
    pdf = (f32p * nl)(*[_p(x, f32p) for x in dfe]) if cfg.do_backward else None
    lib().o_region_path_batch(
        B, L, pl, pd, pb, A, H, W, st, pf, Cc, _p(gts, f32p), _p(gt_labels, i32p), _p(gt_valid, u8p), G,
        C.byref(cfg), nthreads, _p(o["props"], f32p), _p(o["pmask"], u8p), _p(o["rpn_assigned"], i32p),
        _p(o["rpn_pos_idx"], i32p), _p(o["rpn_pos_valid"], u8p), _p(o["rpn_neg_idx"], i32p), _p(o["rpn_neg_valid"], u8p),
        _p(o["rois"], f32p), _p(o["roi_labels"], i32p), _p(o["roi_mask"], u8p), _p(o["roi_deltas"], f32p),
        _p(o["roi_feats"], f32p), _p(dout, f32p) if cfg.do_backward else None, pdf)
    return o


# ------------------------------------------------------------------------------------------------
# "next" row 1: YOLOv8 post-process (parity unpinned; CONVENTIONS #19-#20)
def yolo_decode(pred, shapes, strides, reg_max=16):
    """pred (4*reg_max+nc, A) -> dets (A,6) [x1,y1,x2,y2,score,label]"""
    pred = _f(pred)
    A = pred.shape[1]
    nc = pred.shape[0] - 4 * reg_max
    L = len(shapes)
    H = (C.c_int * L)(*[s[0] for s in shapes])
    W = (C.c_int * L)(*[s[1] for s in shapes])
    st = (C.c_float * L)(*[float(s) for s in strides])
    assert sum(h * w for h, w in shapes) == A
    out = np.zeros((A, 6), np.float32)
    lib().o_yolo_decode(_p(pred, f32p), reg_max, nc, L, H, W, st, _p(out, f32p))
    return out


def yolo_nms(dets, conf_thr=0.25, nms_pre=2048, iou_thr=0.7, agnostic=False, max_det=300):
    dets = _f(dets)
    out = np.zeros((max_det, 6), np.float32)
    idx = np.zeros(max_det, np.int32)
    L = lib()
    L.o_yolo_nms.restype = C.c_int
    cnt = L.o_yolo_nms(_p(dets, f32p), C.c_int64(dets.shape[0]), C.c_float(conf_thr), int(nms_pre), C.c_float(iou_thr),
                       int(agnostic), int(max_det), _p(out, f32p), _p(idx, i32p))
    return out, idx, cnt


def mask_targets(masks, rois, gt_idx, M=28, S=2):
    """masks (G,H,W) uint8, rois (R,4), gt_idx (R) int32 -> (R,M,M) uint8   (a13, parity unpinned, CONVENTIONS #21)"""
    masks = np.ascontiguousarray(masks, np.uint8)
    rois = _f(rois)
    gt_idx = np.ascontiguousarray(gt_idx, np.int32)
    G, H, W = masks.shape
    R = rois.shape[0]
    out = np.zeros((R, M, M), np.uint8)
    lib().o_mask_targets(_p(masks, u8p), G, H, W, _p(rois, f32p), _p(gt_idx, i32p), C.c_int64(R), M, S, _p(out, u8p))
    return out


def softmax_rows(logits):
    logits = _f(logits)
    out = np.empty_like(logits)
    lib().o_softmax_rows(_p(logits, f32p), C.c_int64(logits.shape[0]), int(logits.shape[1]), _p(out, f32p))
    return out


def rcnn_post(rois, roi_valid, logits, deltas, img_h, img_w, means=(0, 0, 0, 0), stds=(0.1, 0.1, 0.2, 0.2),
              max_ratio=MAX_RATIO, score_thr=0.05, nms_pre=2048, iou_thr=0.5, max_det=100):
    """'next' row 4 (parity unpinned, CONVENTIONS #22) -> (out (max_det,6), keep_idx (max_det), count)"""
    rois, logits, deltas = _f(rois), _f(logits), _f(deltas)
    P, nc1 = logits.shape
    valid = np.ascontiguousarray(roi_valid, np.uint8) if roi_valid is not None else None
    m = (C.c_float * 4)(*[float(x) for x in means])
    s = (C.c_float * 4)(*[float(x) for x in stds])
    out = np.zeros((max_det, 6), np.float32)
    idx = np.zeros(max_det, np.int32)
    L = lib()
    L.o_rcnn_post.restype = C.c_int
    cnt = L.o_rcnn_post(_p(rois, f32p), _p(valid, u8p), _p(logits, f32p), _p(deltas, f32p), C.c_int64(P), int(nc1), m, s,
                        C.c_float(max_ratio), C.c_float(img_h), C.c_float(img_w), C.c_float(score_thr), int(nms_pre),
                        C.c_float(iou_thr), int(max_det), _p(out, f32p), _p(idx, i32p))
    return out, idx, cnt
