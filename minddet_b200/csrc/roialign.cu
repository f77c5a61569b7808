// roialign.cu -- a9..a11: RoI -> level map, FPN RoIAlign forward / backward.  sm_100a.
//
// No reference code exists for this op (the only bilinear gather in the
// tree is centerpoint/det3d_ms/core/utils/center_utils.py:97-131).  Semantics: oracle/CONVENTIONS.md
// #14-16 (Caffe2 / aligned=False RoIAlign, average of S x S samples).  Unlike the upstream graph, which
// runs ROIAlign on all 4 levels for every RoI and selects afterwards, only the mapped level is touched.
//
// Forward arithmetic uses explicitly rounded mul/add in the oracle's order, so the result is
// bit-identical to oracle/region_oracle.c:o_roialign_fwd.  Backward accumulates with float atomics
// (order is not deterministic -> FP tolerance only, as north_star allows).
#include "kernels.h"
#include "common.cuh"

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

__global__ void roi_levels_kernel(const float *__restrict__ rois5, int R, const float *__restrict__ cfg,
                                  int32_t *__restrict__ out)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= R) return;
    float r[4];
#pragma unroll
    for (int k = 0; k < 4; k++) r[k] = __ldg(rois5 + (int64_t)i * 5 + 1 + k);
    out[i] = roi_level_of(r, __ldg(cfg), (int)__ldg(cfg + 1));
}

cudaError_t launch_roi_levels(const float *rois5, int R, const float *cfg, int32_t *out, cudaStream_t s)
{
    if (R == 0) return cudaSuccess;
    roi_levels_kernel<<<(R + 255) / 256, 256, 0, s>>>(rois5, R, cfg, out);
    return cudaGetLastError();
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

struct RoiGeom { int b, l, H, W; float sw, sh, bw, bh; };

MD_DEVINL RoiGeom roi_geometry(const RoiFeat &f, const float *__restrict__ roi, int P)
{
    RoiGeom g;
    float r[4];
#pragma unroll
    for (int k = 0; k < 4; k++) r[k] = __ldg(roi + 1 + k);
    g.b = (int)__ldg(roi);
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

MD_DEVINL void build_taps(const RoiGeom &g, int P, int S, Tap *taps)
{
    const int n = P * P * S * S;
    for (int t = threadIdx.x; t < n; t += blockDim.x) {
        const int ix = t % S, iy = (t / S) % S, bin = t / (S * S);
        const int pw = bin % P, ph = bin / P;
        taps[t] = make_tap(sample_coord(g.sh, g.bh, ph, iy, S), sample_coord(g.sw, g.bw, pw, ix, S), g.H, g.W);
    }
}

// ---- generic gather kernels (any level pitch, any footprint) --------------------------------------------
constexpr int kRoiThreads = 256;

__global__ void __launch_bounds__(kRoiThreads)
roialign_fwd_gather_kernel(const RoiFeat f, const float *__restrict__ rois5, int P, int csplit,
                           float *__restrict__ out)
{
    __shared__ Tap taps[kRoiMaxTaps];
    const int r = blockIdx.x;
    const int S = (int)__ldg(f.cfg + 1);
    const RoiGeom g = roi_geometry(f, rois5 + (int64_t)r * 5, P);
    build_taps(g, P, S, taps);
    __syncthreads();
    const int PP = P * P, SS = S * S;
    const float cnt = (float)SS;
    const int cper = (f.C + csplit - 1) / csplit;
    const int c0 = blockIdx.y * cper, c1 = min(f.C, c0 + cper);
    const int64_t plane = (int64_t)g.H * g.W;
    const float *fb = f.feat[g.l] + (int64_t)g.b * f.C * plane;
    for (int o = c0 * PP + threadIdx.x; o < c1 * PP; o += kRoiThreads) {
        const int c = o / PP, bin = o - c * PP;
        const float *fp = fb + (int64_t)c * plane;
        float sum = 0.0f;
        for (int s = 0; s < SS; s++) {
            const Tap t = taps[bin * SS + s];
            float v = add(mul(t.w1, __ldg(fp + t.o1)), mul(t.w2, __ldg(fp + t.o2)));
            v = add(v, mul(t.w3, __ldg(fp + t.o3)));
            v = add(v, mul(t.w4, __ldg(fp + t.o4)));
            sum = add(sum, v);
        }
        out[(int64_t)r * f.C * PP + o] = div(sum, cnt);
    }
}

__global__ void __launch_bounds__(kRoiThreads)
roialign_bwd_gather_kernel(const RoiFeat f, const float *__restrict__ rois5, int P, int csplit,
                           const float *__restrict__ dout)
{
    __shared__ Tap taps[kRoiMaxTaps];
    const int r = blockIdx.x;
    const int S = (int)__ldg(f.cfg + 1);
    const RoiGeom g = roi_geometry(f, rois5 + (int64_t)r * 5, P);
    build_taps(g, P, S, taps);
    __syncthreads();
    const int PP = P * P, SS = S * S;
    const float cnt = (float)SS;
    const int cper = (f.C + csplit - 1) / csplit;
    const int c0 = blockIdx.y * cper, c1 = min(f.C, c0 + cper);
    const int64_t plane = (int64_t)g.H * g.W;
    float *fb = f.feat[g.l] + (int64_t)g.b * f.C * plane;
    for (int o = c0 * PP + threadIdx.x; o < c1 * PP; o += kRoiThreads) {
        const int c = o / PP, bin = o - c * PP;
        float *fp = fb + (int64_t)c * plane;
        const float gr = div(__ldg(dout + (int64_t)r * f.C * PP + o), cnt);
        for (int s = 0; s < SS; s++) {
            const Tap t = taps[bin * SS + s];
            if (t.w1 != 0.0f) atomicAdd(fp + t.o1, mul(gr, t.w1));
            if (t.w2 != 0.0f) atomicAdd(fp + t.o2, mul(gr, t.w2));
            if (t.w3 != 0.0f) atomicAdd(fp + t.o3, mul(gr, t.w3));
            if (t.w4 != 0.0f) atomicAdd(fp + t.o4, mul(gr, t.w4));
        }
    }
}

static RoiFeat to_roifeat(const FeatSet &fs, const float *cfg)
{
    RoiFeat f{};
    f.L = fs.L; f.B = fs.B; f.C = fs.C; f.cfg = cfg;
    for (int l = 0; l < fs.L; l++) { f.H[l] = fs.H[l]; f.W[l] = fs.W[l]; f.feat[l] = fs.feat[l]; }
    return f;
}

cudaError_t launch_roialign_fwd(const FeatSet &fs, const float *rois5, int R, int P, const float *cfg,
                                float *out, cudaStream_t s)
{
    if (R == 0) return cudaSuccess;
    if (P * P * 4 > kRoiMaxTaps || fs.L > kMaxLv) return cudaErrorInvalidValue;
    const RoiFeat f = to_roifeat(fs, cfg);
    const int csplit = fs.C >= 64 ? 4 : 1;
    roialign_fwd_gather_kernel<<<dim3(R, csplit), kRoiThreads, 0, s>>>(f, rois5, P, csplit, out);
    return cudaGetLastError();
}

cudaError_t launch_roialign_bwd(const FeatSet &fs, const float *rois5, int R, int P, const float *cfg,
                                const float *dout, cudaStream_t s)
{
    if (P * P * 4 > kRoiMaxTaps || fs.L > kMaxLv) return cudaErrorInvalidValue;
    for (int l = 0; l < fs.L; l++) {
        cudaError_t e = cudaMemsetAsync(fs.feat[l], 0, (size_t)fs.B * fs.C * fs.H[l] * fs.W[l] * sizeof(float), s);
        if (e != cudaSuccess) return e;
    }
    if (R == 0) return cudaSuccess;
    const RoiFeat f = to_roifeat(fs, cfg);
    const int csplit = fs.C >= 64 ? 4 : 1;
    roialign_bwd_gather_kernel<<<dim3(R, csplit), kRoiThreads, 0, s>>>(f, rois5, P, csplit, dout);
    return cudaGetLastError();
}

}  // namespace md
