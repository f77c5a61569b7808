// bev.cu -- rotated bird's-eye-view overlap / IoU / NMS behind the reference's OWN aot symbols
// (SURVEY.md section 8(f) row 3): BoxesIouBevGpu, BoxesOverlapBevGpu, NmsGpu, NmsNormalGpu
// (centerpoint/det3d_ms/ops/test_custom_pytorch/iou3d_nms_kernel.cu:445-601, python side iou_gpu.py:14-80)
// plus BoxesIouNmsGpu, the device twin of the CPU op boxes_iou_nms_cpu
// (centerpoint/det3d_ms/ops/iou-bev-nms-org.cpp:237-283, python side nms_cpu.py:10-27).
//
// What is kept: the geometry (edge/edge intersections + contained corners with the 1e-2 margin, angular sort
// about the centroid, fan area; iou-bev-nms-org.cpp:26-224) evaluated in plain fp32 without FMA contraction,
// so that keep indices equal those of the compiled reference file (oracle/_ref) -- and the symbols' I/O.
// What is different from the reference's GPU path:
//   * one warp per (row box, 64 columns): lanes = column boxes, two ballots make the 64-bit mask word
//     (the reference serialises 64 heavy overlaps per thread, :300-344, and computes the lower triangle too);
//   * far-apart pairs are rejected by a circumscribed-circle test (exact: their overlap is 0 in the reference);
//   * the greedy reduce runs on the device (nms_sweep_kernel) instead of cudaMalloc + blocking D2H of the mask
//     + a host loop + H2D (:510-542); everything is enqueued on the caller's stream.
#include "kernels.h"
#include "common.cuh"

namespace md {

struct P2 { float x, y; };
constexpr float kBevEps = 1e-8f;
constexpr int kBevMaxPts = 24;           // 16 edge crossings + 8 corners

MD_DEVINL float cross2(P2 a, P2 b) { return a.x * b.y - a.y * b.x; }
MD_DEVINL float cross3(P2 p1, P2 p2, P2 p0) { return (p1.x - p0.x) * (p2.y - p0.y) - (p2.x - p0.x) * (p1.y - p0.y); }

struct BevBox { float x, y, dx, dy, cs, sn, csn, snn; P2 c[5]; };   // cs/sn of +heading, csn/snn of -heading

MD_DEVINL BevBox load_bev(const float *b)
{
    BevBox r;
    r.x = b[0]; r.y = b[1]; r.dx = b[3]; r.dy = b[4];
    const float ang = b[6];
    r.cs = cosf(ang); r.sn = sinf(ang);
    r.csn = cosf(-ang); r.snn = sinf(-ang);
    const float hx = r.dx / 2, hy = r.dy / 2;
    const float x1 = r.x - hx, y1 = r.y - hy, x2 = r.x + hx, y2 = r.y + hy;
    const float px[4] = { x1, x2, x2, x1 }, py[4] = { y1, y1, y2, y2 };
#pragma unroll
    for (int k = 0; k < 4; k++) {
        r.c[k].x = (px[k] - r.x) * r.cs + (py[k] - r.y) * (-r.sn) + r.x;
        r.c[k].y = (px[k] - r.x) * r.sn + (py[k] - r.y) * r.cs + r.y;
    }
    r.c[4] = r.c[0];
    return r;
}

MD_DEVINL bool corner_inside(const BevBox &b, P2 p)
{
    const float rx = (p.x - b.x) * b.csn + (p.y - b.y) * (-b.snn);
    const float ry = (p.x - b.x) * b.snn + (p.y - b.y) * b.csn;
    return fabsf(rx) < b.dx / 2 + 1e-2f && fabsf(ry) < b.dy / 2 + 1e-2f;
}

// segment (p0,p1) x segment (q0,q1); same decision tree as the reference's intersection()
MD_DEVINL bool seg_cross(P2 p1, P2 p0, P2 q1, P2 q0, P2 &out)
{
    const bool boxes_touch = fminf(p0.x, p1.x) <= fmaxf(q0.x, q1.x) && fminf(q0.x, q1.x) <= fmaxf(p0.x, p1.x) &&
                             fminf(p0.y, p1.y) <= fmaxf(q0.y, q1.y) && fminf(q0.y, q1.y) <= fmaxf(p0.y, p1.y);
    if (!boxes_touch) return false;
    const float s1 = cross3(q0, p1, p0), s2 = cross3(p1, q1, p0), s3 = cross3(p0, q1, q0), s4 = cross3(q1, p1, q0);
    if (!(s1 * s2 > 0 && s3 * s4 > 0)) return false;
    const float s5 = cross3(q1, p1, p0);
    if (fabsf(s5 - s1) > kBevEps) {
        out.x = (s5 * q0.x - s1 * q1.x) / (s5 - s1);
        out.y = (s5 * q0.y - s1 * q1.y) / (s5 - s1);
    } else {
        const float a0 = p0.y - p1.y, b0 = p1.x - p0.x, c0 = p0.x * p1.y - p1.x * p0.y;
        const float a1 = q0.y - q1.y, b1 = q1.x - q0.x, c1 = q0.x * q1.y - q1.x * q0.y;
        const float D = a0 * b1 - a1 * b0;
        out.x = (b0 * c1 - b1 * c0) / D;
        out.y = (a1 * c0 - a0 * c1) / D;
    }
    return true;
}

MD_DEVINL float bev_overlap(const BevBox &a, const BevBox &b)
{
    // far apart (circumscribed circles + slack for the 1e-2 corner margin): no crossing, no contained corner -> 0
    {
        const float ddx = a.x - b.x, ddy = a.y - b.y;
        const float ra = 0.5f * sqrtf(a.dx * a.dx + a.dy * a.dy), rb = 0.5f * sqrtf(b.dx * b.dx + b.dy * b.dy);
        const float reach = ra + rb + 0.1f;
        if (ddx * ddx + ddy * ddy > reach * reach * 1.001f) return 0.0f;
    }
    P2 pts[kBevMaxPts];
    float key[kBevMaxPts];
    int cnt = 0;
    P2 ctr = { 0.0f, 0.0f };
    for (int i = 0; i < 4; i++)
        for (int j = 0; j < 4; j++) {
            P2 x;
            if (seg_cross(a.c[i + 1], a.c[i], b.c[j + 1], b.c[j], x)) {
                ctr.x = ctr.x + x.x; ctr.y = ctr.y + x.y;
                pts[cnt++] = x;
            }
        }
    for (int k = 0; k < 4; k++) {
        if (corner_inside(a, b.c[k])) { ctr.x = ctr.x + b.c[k].x; ctr.y = ctr.y + b.c[k].y; pts[cnt++] = b.c[k]; }
        if (corner_inside(b, a.c[k])) { ctr.x = ctr.x + a.c[k].x; ctr.y = ctr.y + a.c[k].y; pts[cnt++] = a.c[k]; }
    }
    if (cnt < 3) return 0.0f;                   // the reference's fan loop yields 0 for fewer than 3 points
    ctr.x /= cnt; ctr.y /= cnt;
    for (int k = 0; k < cnt; k++) key[k] = atan2f(pts[k].y - ctr.y, pts[k].x - ctr.x);
    // stable ascending sort by angle (same permutation as the reference's bubble sort with a strict '>')
    for (int k = 1; k < cnt; k++) {
        const P2 p = pts[k];
        const float kk = key[k];
        int m = k - 1;
        while (m >= 0 && key[m] > kk) { pts[m + 1] = pts[m]; key[m + 1] = key[m]; m--; }
        pts[m + 1] = p; key[m + 1] = kk;
    }
    float area = 0.0f;
    for (int k = 0; k < cnt - 1; k++) {
        const P2 u = { pts[k].x - pts[0].x, pts[k].y - pts[0].y }, v = { pts[k + 1].x - pts[0].x, pts[k + 1].y - pts[0].y };
        area += cross2(u, v);
    }
    return fabsf(area) * 0.5f;
}

// axis-aligned variant on the same 7-float boxes (iou_normal, iou3d_nms_kernel.cu:347-358)
MD_DEVINL float bev_iou_normal(const float *a, const float *b)
{
    const float left = fmaxf(a[0] - a[3] / 2, b[0] - b[3] / 2), right = fminf(a[0] + a[3] / 2, b[0] + b[3] / 2);
    const float top = fmaxf(a[1] - a[4] / 2, b[1] - b[4] / 2), bottom = fminf(a[1] + a[4] / 2, b[1] + b[4] / 2);
    const float width = fmaxf(right - left, 0.0f), height = fmaxf(bottom - top, 0.0f);
    const float inter = width * height;
    const float sa = a[3] * a[4], sb = b[3] * b[4];
    return inter / fmaxf(sa + sb - inter, kBevEps);
}

// ---- (N,M) overlap / IoU matrices -------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
bev_pair_kernel(const float *__restrict__ boxes_a, int na, const float *__restrict__ boxes_b, int nb, int want_iou,
                float *__restrict__ out)
{
    // block = 8 rows x 32 columns; a row box is shared by the 32 lanes of a warp
    const int col = blockIdx.x * 32 + (threadIdx.x & 31), row = blockIdx.y * 8 + (threadIdx.x >> 5);
    if (row >= na || col >= nb) return;
    const BevBox a = load_bev(boxes_a + (int64_t)row * 7), b = load_bev(boxes_b + (int64_t)col * 7);
    float v = bev_overlap(a, b);
    if (want_iou) v = v / fmaxf(a.dx * a.dy + b.dx * b.dy - v, kBevEps);
    out[(int64_t)row * nb + col] = v;
}

cudaError_t launch_bev_pairs(const float *boxes_a, int na, const float *boxes_b, int nb, int want_iou, float *out, cudaStream_t s)
{
    if (na == 0 || nb == 0) return cudaSuccess;
    bev_pair_kernel<<<dim3((nb + 31) / 32, (na + 7) / 8), 256, 0, s>>>(boxes_a, na, boxes_b, nb, want_iou, out);
    return cudaGetLastError();
}

// ---- suppression bitmask: warp = (row box, one 64-column block), lanes = columns, 2 ballots per word --------
// mode 0: rotated IoU with the eps guard, strict '>'          (NmsGpu)
// mode 1: axis-aligned IoU with the eps guard, strict '>'      (NmsNormalGpu)
// mode 2: rotated, ovr = s / (sa + sb - s) without guard, '>=' (boxes_iou_nms_cpu)
__global__ void __launch_bounds__(256)
bev_mask_kernel(const float *__restrict__ boxes, int n, const float *__restrict__ thr_ptr, int mode, int nbp,
                unsigned long long *__restrict__ mask)
{
    const int nb = (n + 63) >> 6;
    const int warp_global = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    const int row = warp_global / nb, cb = warp_global - row * nb;
    if (row >= n) return;
    unsigned long long word = 0ull;
    if (cb >= (row >> 6)) {                         // upper triangle only
        const float thr = __ldg(thr_ptr);
        const float *pa = boxes + (int64_t)row * 7;
        BevBox a;
        if (mode != 1) a = load_bev(pa);
        const float sa = pa[3] * pa[4];
#pragma unroll
        for (int h = 0; h < 2; h++) {
            const int col = cb * 64 + h * 32 + lane;
            bool sup = false;
            if (col < n && col > row) {
                const float *pb = boxes + (int64_t)col * 7;
                if (mode == 1) {
                    sup = bev_iou_normal(pa, pb) > thr;
                } else {
                    const BevBox b = load_bev(pb);
                    const float s = bev_overlap(a, b);
                    const float sb = pb[3] * pb[4];
                    if (mode == 0) sup = s / fmaxf(sa + sb - s, kBevEps) > thr;
                    else sup = s / (sa + sb - s) >= thr;
                }
            }
            word |= (unsigned long long)__ballot_sync(0xffffffffu, sup) << (32 * h);
        }
    }
    if (lane == 0) mask[(int64_t)row * nbp + cb] = word;
}

// zero-area boxes are removed up front and never suppress anything (iou-bev-nms-org.cpp:250-256, 258-263)
__global__ void bev_zero_area_kernel(const float *__restrict__ boxes, int n, int nbp, unsigned long long *__restrict__ init_removed,
                                     unsigned long long *__restrict__ mask)
{
    const int w = blockIdx.x * blockDim.x + threadIdx.x;      // one thread per 64-box word
    if (w >= nbp) return;
    unsigned long long bits = 0ull;
    for (int k = 0; k < 64; k++) {
        const int i = w * 64 + k;
        if (i < n && boxes[(int64_t)i * 7 + 3] * boxes[(int64_t)i * 7 + 4] == 0.0f) bits |= 1ull << k;
    }
    init_removed[w] = bits;
    // a removed box must not suppress: clear its mask row
    for (int k = 0; k < 64; k++)
        if ((bits >> k) & 1ull)
            for (int j = 0; j < nbp; j++) mask[(int64_t)(w * 64 + k) * nbp + j] = 0ull;
}

__global__ void bev_widen_keep_kernel(const int32_t *__restrict__ keep32, const int32_t *__restrict__ count, int n,
                                      long long *__restrict__ keep64, int32_t *__restrict__ num_out)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) keep64[i] = keep32[i];                       // zero padded beyond count, like the reference's memset
    if (i == 0) num_out[0] = count[0];
}

size_t bev_nms_workspace_bytes(int n)
{
    const int nb = (n + 63) / 64, nbp = (nb + 1) & ~1;
    return (size_t)nb * 64 * nbp * 8 + (size_t)nbp * 8 + (size_t)n * 4 + (size_t)n + 4 + 1024;
}

cudaError_t launch_nms_sweep_single(const unsigned long long *mask, const unsigned long long *init_removed, int n, int nbp,
                                    int32_t *keep_pos, uint8_t *keep_mask, int32_t *count, cudaStream_t s);

// keep: int64 (keep_is_64) or int32; count int32[1]
cudaError_t launch_bev_nms(const float *boxes, int n, const float *thr, int mode, void *ws, void *keep, int keep_is_64,
                           int32_t *num_out, cudaStream_t s)
{
    const int nb = (n + 63) / 64, nbp = (nb + 1) & ~1;
    if (nb > 64) return cudaErrorInvalidValue;               // sweep capacity: 4096 boxes (the reference's older CPU copy: N = 4096)
    unsigned char *p = reinterpret_cast<unsigned char *>(ws);
    unsigned long long *mask = reinterpret_cast<unsigned long long *>(p); p += (size_t)nb * 64 * nbp * 8;
    unsigned long long *init_removed = reinterpret_cast<unsigned long long *>(p); p += (size_t)nbp * 8;
    int32_t *keep32 = reinterpret_cast<int32_t *>(p); p += (size_t)n * 4;
    int32_t *count = reinterpret_cast<int32_t *>(p); p += 4;
    uint8_t *kmask = p;
    if (n == 0) {
        cudaError_t e = cudaMemsetAsync(num_out, 0, 4, s);
        return e;
    }
    const int warps = n * nb;
    bev_mask_kernel<<<(warps + 7) / 8, 256, 0, s>>>(boxes, n, thr, mode, nbp, mask);
    const unsigned long long *init = nullptr;
    if (mode == 2) {
        bev_zero_area_kernel<<<(nbp + 63) / 64, 64, 0, s>>>(boxes, n, nbp, init_removed, mask);
        init = init_removed;
    }
    cudaError_t e = launch_nms_sweep_single(mask, init, n, nbp, keep_is_64 ? keep32 : reinterpret_cast<int32_t *>(keep), kmask,
                                            keep_is_64 ? count : num_out, s);
    if (e != cudaSuccess) return e;
    if (keep_is_64)
        bev_widen_keep_kernel<<<(n + 255) / 256, 256, 0, s>>>(keep32, count, n, reinterpret_cast<long long *>(keep), num_out);
    return cudaGetLastError();
}

}  // namespace md
