"""Seeded synthetic inputs for the region path; numpy only, used by tests,
bench.py and smoke().  Not part of the product path."""
import numpy as np

STRIDES = (4, 8, 16, 32, 64)
IMG_H, IMG_W = 800, 1344          # 800x1333 padded to a multiple of 64


def level_shapes(img_h=IMG_H, img_w=IMG_W, strides=STRIDES):
    return [(-(-img_h // s), -(-img_w // s)) for s in strides]


def base_anchor_sets(strides=STRIDES, scales=(8,), ratios=(0.5, 1.0, 2.0)):
    """mmdet-v1 base anchors, one set per level (base_size == stride). float64 -> round -> fp32."""
    out = []
    for s in strides:
        w = h = float(s)
        x_ctr, y_ctr = 0.5 * (w - 1), 0.5 * (h - 1)
        hr = np.sqrt(np.asarray(ratios, np.float64))
        wr = 1 / hr
        ws = (w * wr[:, None] * np.asarray(scales, np.float64)[None, :]).reshape(-1)
        hs = (h * hr[:, None] * np.asarray(scales, np.float64)[None, :]).reshape(-1)
        b = np.stack([x_ctr - 0.5 * (ws - 1), y_ctr - 0.5 * (hs - 1), x_ctr + 0.5 * (ws - 1), y_ctr + 0.5 * (hs - 1)], -1).round()
        out.append(b.astype(np.float32))
    return out


def rpn_head_outputs(B, shapes, A=3, seed=0xD37):
    """logits ~ N(-4,2) per anchor, dx,dy ~ N(0,0.1), dw,dh ~ N(0,0.2); head layout (B,A,H,W)/(B,4A,H,W)."""
    rng = np.random.default_rng(seed)
    logits, deltas = [], []
    for (h, w) in shapes:
        logits.append(rng.normal(-4.0, 2.0, (B, A, h, w)).astype(np.float32))
        d = rng.normal(0.0, 1.0, (B, A, 4, h, w)).astype(np.float32)
        d[:, :, :2] *= 0.1
        d[:, :, 2:] *= 0.2
        deltas.append(d.reshape(B, 4 * A, h, w))
    return logits, deltas


def gt_boxes(B, G=128, max_valid=32, img_h=IMG_H, img_w=IMG_W, seed=0xD37, num_classes=80):
    """G ~ U{1..max_valid} valid gts per image, centres uniform, sides log-uniform in [16,512], clipped."""
    rng = np.random.default_rng(seed + 1)
    gts = np.zeros((B, G, 4), np.float32)
    labels = np.zeros((B, G), np.int32)
    valid = np.zeros((B, G), np.uint8)
    for b in range(B):
        n = int(rng.integers(1, max_valid + 1))
        c = rng.uniform([0, 0], [img_w, img_h], (n, 2))
        wh = np.exp(rng.uniform(np.log(16), np.log(512), (n, 2)))
        bx = np.concatenate([c - wh / 2, c + wh / 2], 1)
        bx[:, 0::2] = np.clip(bx[:, 0::2], 0, img_w - 1)
        bx[:, 1::2] = np.clip(bx[:, 1::2], 0, img_h - 1)
        gts[b, :n] = bx.astype(np.float32)
        labels[b, :n] = rng.integers(1, num_classes + 1, n)
        valid[b, :n] = 1
    return gts, labels, valid


def features(B, shapes, C=256, seed=0xD37):
    rng = np.random.default_rng(seed + 2)
    return [rng.uniform(-1.0, 1.0, (B, C, h, w)).astype(np.float32) for (h, w) in shapes]


def rand_boxes(rng, n, img_w=float(IMG_W), img_h=float(IMG_H), smin=8.0, smax=400.0, cluster=None, sigma=12.0):
    if cluster is None:
        cx, cy = rng.uniform(0, img_w, n), rng.uniform(0, img_h, n)
    else:
        ctr = rng.uniform([0, 0], [img_w, img_h], (cluster, 2))
        pick = rng.integers(0, cluster, n)
        cx = ctr[pick, 0] + rng.normal(0, sigma, n)
        cy = ctr[pick, 1] + rng.normal(0, sigma, n)
    w = np.exp(rng.uniform(np.log(smin), np.log(smax), n))
    h = np.exp(rng.uniform(np.log(smin), np.log(smax), n))
    b = np.stack([cx - w / 2, cy - h / 2, cx + w / 2, cy + h / 2], 1)
    b[:, 0::2] = np.clip(b[:, 0::2], 0, img_w - 1)
    b[:, 1::2] = np.clip(b[:, 1::2], 0, img_h - 1)
    return b.astype(np.float32)
