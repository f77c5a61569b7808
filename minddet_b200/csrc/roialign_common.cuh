// roialign_common.cuh -- geometry shared by the gather (bit-exact) and TMA (separable) RoIAlign kernels.
// Semantics: oracle/CONVENTIONS.md #14, #15; op order identical to oracle/region_oracle.c
// (roi_level, roi_geometry, sample_coord, make_tap).
#pragma once
#include "common.cuh"
#include "kernels.h"

namespace md {

constexpr int kRoiMaxTaps = 14 * 14 * 4;   // P*P*S*S upper bound handled by the tap table

struct RoiFeat {
    int L, B, C;
    int H[kMaxLv], W[kMaxLv];
    float *feat[kMaxLv];
    const float *cfg;
};

MD_DEVINL int roi_level_of(const float *r /* x1,y1,x2,y2 */, float finest, int num_levels)
{
    const float w = add(sub(r[2], r[0]), 1.0f);
    const float h = add(sub(r[3], r[1]), 1.0f);
    const float s = __fsqrt_rn(mul(w, h));
    const float t = add(div(s, finest), 1e-6f);
    int l = (t >= 2.0f) + (t >= 4.0f) + (t >= 8.0f);
    for (int k = 4; k < num_levels; k++) l += (t >= (float)(1 << k));
    return min(l, num_levels - 1);
}

// one bilinear sample: 4 plane offsets + 4 weights (weights 0 when the sample is out of range)
struct __align__(16) Tap { int o1, o2, o3, o4; float w1, w2, w3, w4; };

MD_DEVINL Tap make_tap(float y, float x, int H, int W)
{
    Tap t; t.o1 = t.o2 = t.o3 = t.o4 = 0; t.w1 = t.w2 = t.w3 = t.w4 = 0.0f;
    if (y < -1.0f || y > (float)H || x < -1.0f || x > (float)W) return t;
    if (y <= 0.0f) y = 0.0f;
    if (x <= 0.0f) x = 0.0f;
    int yl = (int)y, xl = (int)x, yh, xh;
    if (yl >= H - 1) { yh = yl = H - 1; y = (float)yl; } else yh = yl + 1;
    if (xl >= W - 1) { xh = xl = W - 1; x = (float)xl; } else xh = xl + 1;
    const float ly = sub(y, (float)yl), lx = sub(x, (float)xl);
    const float hy = sub(1.0f, ly), hx = sub(1.0f, lx);
    t.o1 = yl * W + xl; t.o2 = yl * W + xh; t.o3 = yh * W + xl; t.o4 = yh * W + xh;
    t.w1 = mul(hy, hx); t.w2 = mul(hy, lx); t.w3 = mul(ly, hx); t.w4 = mul(ly, lx);
    return t;
}

struct RoiGeom { int b, l, H, W; float sw, sh, bw, bh; bool ok; };   // ok: batch index inside [0, B) (CONVENTIONS #23)

MD_DEVINL RoiGeom roi_geometry(const RoiFeat &f, const float *__restrict__ roi, int P)
{
    RoiGeom g;
    float r[4];
#pragma unroll
    for (int k = 0; k < 4; k++) r[k] = __ldg(roi + 1 + k);
    const float bf = __ldg(roi);
    g.ok = bf >= 0.0f && bf < (float)f.B;          // false for NaN too
    g.b = g.ok ? (int)bf : 0;
    g.l = roi_level_of(r, __ldg(f.cfg + 0), f.L);
    g.H = f.H[g.l]; g.W = f.W[g.l];
    const float scale = div(1.0f, __ldg(f.cfg + 4 + g.l));
    const float em = __ldg(f.cfg + 2);
    g.sw = mul(r[0], scale); g.sh = mul(r[1], scale);
    const float ew = mul(add(r[2], em), scale), eh = mul(add(r[3], em), scale);
    const float rw = fmaxf(sub(ew, g.sw), 1.0f), rh = fmaxf(sub(eh, g.sh), 1.0f);
    g.bw = div(rw, (float)P); g.bh = div(rh, (float)P);
    return g;
}
MD_DEVINL float sample_coord(float start, float bin, int p, int i, int S)
{
    const float base = add(start, mul((float)p, bin));
    const float o = div(mul(add((float)i, 0.5f), bin), (float)S);
    return add(base, o);
}


// ---- 1-D sample -> (low index, high index, low weight, high weight, valid); mirrors make_tap ----------
MD_DEVINL bool sample_1d(float v, int extent, int &lo, int &hi, float &wl, float &wh)
{
    if (v < -1.0f || v > (float)extent) return false;
    if (v <= 0.0f) v = 0.0f;
    lo = (int)v;
    if (lo >= extent - 1) { hi = lo = extent - 1; v = (float)lo; } else hi = lo + 1;
    wh = sub(v, (float)lo);
    wl = sub(1.0f, wh);
    return true;
}


// CTA -> (RoI, channel chunk).  Blocks are ordered (segment of `seg` consecutive RoIs, chunk, RoI in segment)
// so that the CTAs resident at any time read the SAME channel planes of (normally) one image: every
// feature byte is then fetched from HBM once and re-used out of L2 by all RoIs that overlap it.
struct WorkItem { int r, chunk; };
MD_DEVINL WorkItem work_item(int bid, int R, int seg, int nchunk)
{
    const int per_seg = seg * nchunk;
    const int sidx = bid / per_seg, base = sidx * seg;
    const int seg_len = min(seg, R - base);
    const int rem = bid - sidx * per_seg;
    WorkItem w;
    w.chunk = rem / seg_len;
    w.r = base + (rem - w.chunk * seg_len);
    return w;
}

}  // namespace md
