// yolo.cu -- "next" row 1: YOLOv8 post-process = DFL decode of every anchor + class-aware
// batched NMS.  No reference code exists (README.md:13 names the model only); semantics: oracle/CONVENTIONS.md
// #19-#20, oracle/region_oracle.c (o_yolo_decode, o_yolo_nms).
//
// Decode is the streaming kernel of this repo: (4*16 + nc) * 4 bytes read and 24 bytes written per anchor
// (algorithmic 4.84 MB + 0.20 MB per 640x640 image), no reuse, so it is judged against the HBM roofline.
// Each thread owns V consecutive anchors (V = 2 by default: 64 registers, 8 CTAs/SM; V = 4 is the 128-bit variant):
// every channel plane is read with one vector streaming load per thread (256-512 contiguous bytes per warp), a
// side's 16 bins live in registers, and the V x 6 results leave as 128-bit stores; the grid is one persistent wave.  All arithmetic that decides a label or a keep index is individually rounded (bit-exact vs the oracle).
// NMS re-uses the region path's machinery: cluster radix select (top nms_pre of the candidates above the confidence
// threshold, sorted), label-aware 64-bit bitmask tiles, on-device sweep.
#include <cstdlib>

#include "kernels.h"
#include "nms.cuh"
#include "select.cuh"

namespace md {

constexpr int kRegMax = 16;
constexpr int kYoloThreads = 128;
constexpr int kYoloMaxLevels = 8;

struct YoloLevels { int n, start[kYoloMaxLevels + 1], W[kYoloMaxLevels]; float stride[kYoloMaxLevels]; };

// cfg_len = floats the caller's cfg tensor really holds: the level count read from cfg[0] is clamped to what fits
MD_DEVINL YoloLevels load_levels(const float *__restrict__ cfg, int cfg_len)
{
    YoloLevels lv;
    lv.n = max(0, min(min((int)__ldg(cfg), kYoloMaxLevels), (cfg_len - 1) / 3));
    int s = 0;
    for (int l = 0; l < lv.n; l++) {
        const int H = (int)__ldg(cfg + 1 + 3 * l);
        lv.W[l] = (int)__ldg(cfg + 2 + 3 * l);
        lv.stride[l] = __ldg(cfg + 3 + 3 * l);
        lv.start[l] = s;
        s += H * lv.W[l];
    }
    lv.start[lv.n] = s;
    return lv;
}

// distance of one side: max-subtracted softmax over the bins, expectation in bins (sequential fp32 accumulation)
MD_DEVINL float dfl_expectation(const float (&x)[kRegMax])
{
    float m = x[0];
#pragma unroll
    for (int i = 1; i < kRegMax; i++) m = fmaxf(m, x[i]);
    float den = 0.0f, num = 0.0f;
#pragma unroll
    for (int i = 0; i < kRegMax; i++) {
        const float e = exact_exp(sub(x[i], m));
        den = add(den, e);
        num = __fmaf_rn(e, (float)i, num);
    }
    return div(num, den);
}

template <int V>   // anchors per thread: 4 (128-bit path) or 1
__global__ void __launch_bounds__(kYoloThreads)
yolo_decode_kernel(const float *__restrict__ pred, int B, int A, int nc, int tiles_per_image, const float *__restrict__ cfg, int cfg_len,
                   float *__restrict__ dets)
{
    __shared__ YoloLevels lv;                 // dynamic level lookup per anchor: keep it out of local memory
    if (threadIdx.x == 0) lv = load_levels(cfg, cfg_len);
    __syncthreads();
    // persistent: the grid is one full wave (148 x resident CTAs); tiles = (image, 128*V anchors) are strided
    for (int tile = blockIdx.x; tile < tiles_per_image * B; tile += gridDim.x) {
    const int b = tile / tiles_per_image;
    const int a0 = ((tile - b * tiles_per_image) * kYoloThreads + threadIdx.x) * V;
    if (a0 >= A) continue;
    const float *base = pred + (int64_t)b * (4 * kRegMax + nc) * A + a0;
    float d[4][V];
#pragma unroll
    for (int s = 0; s < 4; s++) {
        float x[V][kRegMax];
#pragma unroll
        for (int i = 0; i < kRegMax; i++) {
            const float *p = base + (int64_t)(s * kRegMax + i) * A;
            if (V == 4) {
                const float4 v = ldg_stream(reinterpret_cast<const float4 *>(p));
                x[0][i] = v.x; x[1 % V][i] = v.y; x[2 % V][i] = v.z; x[3 % V][i] = v.w;
            } else if (V == 2) {
                const float2 v = __ldcs(reinterpret_cast<const float2 *>(p));
                x[0][i] = v.x; x[1 % V][i] = v.y;
            } else {
                x[0][i] = __ldg(p);
            }
        }
#pragma unroll
        for (int j = 0; j < V; j++) d[s][j] = dfl_expectation(x[j]);
    }
    float best[V];
    int lab[V];
    {
        const float *c = base + (int64_t)(4 * kRegMax) * A;
#pragma unroll
        for (int j = 0; j < V; j++) { best[j] = -3.0e38f; lab[j] = 0; }
#pragma unroll 16
        for (int k = 0; k < nc; k++) {
            float v[V];
            if (V == 4) {
                const float4 t = ldg_stream(reinterpret_cast<const float4 *>(c + (int64_t)k * A));
                v[0] = t.x; v[1 % V] = t.y; v[2 % V] = t.z; v[3 % V] = t.w;
            } else if (V == 2) {
                const float2 t = __ldcs(reinterpret_cast<const float2 *>(c + (int64_t)k * A));
                v[0] = t.x; v[1 % V] = t.y;
            } else {
                v[0] = __ldg(c + (int64_t)k * A);
            }
#pragma unroll
            for (int j = 0; j < V; j++)
                if (k == 0 || v[j] > best[j]) { best[j] = v[j]; lab[j] = k; }     // first maximum wins
        }
    }
    float o[V * 6];
#pragma unroll
    for (int j = 0; j < V; j++) {
        const int a = a0 + j;
        int l = 0;
        while (l + 1 < lv.n && a >= lv.start[l + 1]) l++;
        const int r = a - lv.start[l];
        const int y = r / lv.W[l], x = r - y * lv.W[l];
        const float cx = add((float)x, 0.5f), cy = add((float)y, 0.5f), st = lv.stride[l];
        o[j * 6 + 0] = mul(sub(cx, d[0][j]), st);
        o[j * 6 + 1] = mul(sub(cy, d[1][j]), st);
        o[j * 6 + 2] = mul(add(cx, d[2][j]), st);
        o[j * 6 + 3] = mul(add(cy, d[3][j]), st);
        o[j * 6 + 4] = exact_sigmoid(best[j]);
        o[j * 6 + 5] = (float)lab[j];
    }
    float *dst = dets + ((int64_t)b * A + a0) * 6;
    if (V == 4) {
#pragma unroll
        for (int q = 0; q < 6; q++)
            stg_stream(reinterpret_cast<float4 *>(dst) + q, make_float4(o[4 * q], o[(4 * q + 1) % (V * 6)], o[(4 * q + 2) % (V * 6)], o[(4 * q + 3) % (V * 6)]));
    } else if (V == 2) {
#pragma unroll
        for (int q = 0; q < 3; q++)
            stg_stream(reinterpret_cast<float4 *>(dst) + q, make_float4(o[4 * q], o[(4 * q + 1) % (V * 6)], o[(4 * q + 2) % (V * 6)], o[(4 * q + 3) % (V * 6)]));
    } else {
#pragma unroll
        for (int q = 0; q < 6; q++) dst[q] = o[q];
    }
    }
}

cudaError_t launch_yolo_decode(const float *pred, int B, int C, int A, const float *cfg, int cfg_len, float *dets, cudaStream_t s)
{
    const int nc = C - 4 * kRegMax;
    if (nc < 1) return cudaErrorInvalidValue;
    if (B == 0 || A == 0) return cudaSuccess;
    const bool vec = (A % 4 == 0) && ((reinterpret_cast<uintptr_t>(pred) & 15) == 0) && ((reinterpret_cast<uintptr_t>(dets) & 15) == 0);
    const char *ev = getenv("MD_YOLO_V");
    const int V = !vec ? 1 : (ev ? atoi(ev) : 2);   // measured at config 5: V=2 61 us, V=1 63 us, V=4 66 us (occupancy wins)
    auto kern = V == 4 ? yolo_decode_kernel<4> : (V == 2 ? yolo_decode_kernel<2> : yolo_decode_kernel<1>);
    static int resident[5] = { 0, 0, 0, 0, 0 };         // CTAs per SM per instantiation (occupancy query, once)
    if (!resident[V]) {
        int n = 0;
        cudaError_t e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, kern, kYoloThreads, 0);
        if (e != cudaSuccess) return e;
        resident[V] = n > 0 ? n : 1;
    }
    const int per = V * kYoloThreads;
    const int tiles_per_image = (A + per - 1) / per;
    const long long tiles = (long long)tiles_per_image * B;
    const int grid = (int)(tiles < 148LL * resident[V] ? tiles : 148LL * resident[V]);
    kern<<<grid, kYoloThreads, 0, s>>>(pred, B, A, nc, tiles_per_image, cfg, cfg_len, dets);
    return cudaGetLastError();
}

// ---- class-aware NMS -------------------------------------------------------------------------------------------
struct YoloSrc {
    const float *dets; int A, nms_pre; const float *cfg;      // cfg: conf_thr, iou_thr, agnostic
    struct Ctx { const float *base; float conf; };
    __device__ int segment_of(int i, int it) const { return it ? -1 : i; }
    __device__ Ctx prepare(int seg) const { return Ctx{ dets + (int64_t)seg * A * 6, __ldg(cfg) }; }
    __device__ bool active(const Ctx &) const { return true; }
    __device__ int length(const Ctx &) const { return A; }
    __device__ int want(const Ctx &) const { return nms_pre; }
    __device__ uint32_t index_of(const Ctx &, int m) const { return (uint32_t)m; }
    __device__ bool load(const Ctx &c, int m, uint32_t &key) const
    {
        const float sc = __ldg(c.base + (int64_t)m * 6 + 4);
        if (!(sc > c.conf)) return false;
        key = score_key(sc);
        return true;
    }
};
struct YoloSink {
    const float *dets; int A, nms_pre;
    float4 *ws_boxes; int32_t *ws_labels; int32_t *cand_idx; int32_t *selected;
    __device__ void emit(int seg, int rank, unsigned long long comp) const
    {
        const int32_t a = (int32_t)(~(uint32_t)comp);
        const float *p = dets + ((int64_t)seg * A + a) * 6;
        const int64_t o = (int64_t)seg * nms_pre + rank;
        ws_boxes[o] = make_float4(__ldg(p), __ldg(p + 1), __ldg(p + 2), __ldg(p + 3));
        ws_labels[o] = (int32_t)__ldg(p + 5);
        cand_idx[o] = a;
    }
    __device__ void pad(int seg, int rank) const
    {
        const int64_t o = (int64_t)seg * nms_pre + rank;
        ws_boxes[o] = make_float4(0, 0, 0, 0);        // zero boxes have IoU 0 with everything: they never suppress
        ws_labels[o] = -1;
        cand_idx[o] = -1;
    }
    __device__ void finish(int seg, int sel, int) const { selected[seg] = sel; }
};

__global__ void yolo_gather_kernel(const float *__restrict__ dets, int A, int nms_pre, int max_det,
                                   const int32_t *__restrict__ cand_idx, const int32_t *__restrict__ selected,
                                   const int32_t *__restrict__ keep_pos, const int32_t *__restrict__ count,
                                   float *__restrict__ out, int32_t *__restrict__ keep_idx, int32_t *__restrict__ num_out)
{
    const int b = blockIdx.x;
    const int sel = selected[b];
    const int32_t *kp = keep_pos + (int64_t)b * nms_pre;
    // kept real candidates are a prefix of keep_pos (ascending positions; padding sits at positions >= sel)
    int lo = 0, hi = count[b];
    while (lo < hi) { const int mid = (lo + hi) >> 1; if (kp[mid] < sel) lo = mid + 1; else hi = mid; }
    const int n = min(lo, max_det);
    for (int i = threadIdx.x; i < max_det; i += blockDim.x) {
        float *o = out + ((int64_t)b * max_det + i) * 6;
        int32_t a = -1;
        if (i < n) {
            a = cand_idx[(int64_t)b * nms_pre + kp[i]];
            const float *p = dets + ((int64_t)b * A + a) * 6;
#pragma unroll
            for (int k = 0; k < 6; k++) o[k] = __ldg(p + k);
        } else {
#pragma unroll
            for (int k = 0; k < 6; k++) o[k] = 0.0f;
        }
        keep_idx[(int64_t)b * max_det + i] = a;
    }
    if (threadIdx.x == 0) num_out[b] = n;
}

struct YoloWs { float4 *boxes; int32_t *labels, *selected, *keep_pos, *count; uint8_t *keep_mask; float *nms_cfg; unsigned long long *mask; size_t total; };
static inline size_t al256(size_t x) { return (x + 255) & ~(size_t)255; }
static YoloWs carve_yolo_ws(void *ws, int B, int nms_pre)
{
    YoloWs w;
    unsigned char *p = reinterpret_cast<unsigned char *>(ws);
    size_t o = 0;
    w.boxes = reinterpret_cast<float4 *>(p + o); o += al256((size_t)B * nms_pre * 16);
    w.labels = reinterpret_cast<int32_t *>(p + o); o += al256((size_t)B * nms_pre * 4);
    w.keep_pos = reinterpret_cast<int32_t *>(p + o); o += al256((size_t)B * nms_pre * 4);
    w.keep_mask = reinterpret_cast<uint8_t *>(p + o); o += al256((size_t)B * nms_pre);
    w.selected = reinterpret_cast<int32_t *>(p + o); o += al256((size_t)B * 4);
    w.count = reinterpret_cast<int32_t *>(p + o); o += al256((size_t)B * 4);
    w.nms_cfg = reinterpret_cast<float *>(p + o); o += 256;
    w.mask = reinterpret_cast<unsigned long long *>(p + o); o += nms_workspace_bytes(B, nms_pre);
    w.total = o;
    return w;
}
size_t yolo_nms_workspace_bytes(int B, int nms_pre) { return carve_yolo_ws(nullptr, B, nms_pre).total; }

// MD_CFG_NMS block for run_nms from the yolo cfg: thr = cfg[1], offset 0, strict, eps 1e-8
__global__ void yolo_nms_cfg_kernel(const float *__restrict__ cfg, float *__restrict__ out)
{
    out[0] = cfg[1]; out[1] = 0.0f; out[2] = 0.0f; out[3] = 1e-8f;
}

cudaError_t launch_yolo_nms(const float *dets, int B, int A, const float *cfg, void *ws, int nms_pre, int max_det,
                            float *out, int32_t *keep_idx, int32_t *num_out, int32_t *cand_idx, cudaStream_t s)
{
    if (nms_pre > kSelMaxK || A >= (1 << kSelMaxIndexBits)) return cudaErrorInvalidValue;
    if (B == 0) return cudaSuccess;
    const YoloWs w = carve_yolo_ws(ws, B, nms_pre);
    yolo_nms_cfg_kernel<<<1, 1, 0, s>>>(cfg, w.nms_cfg);
    cudaError_t e = launch_select_sorted(YoloSrc{ dets, A, nms_pre, cfg },
                                         YoloSink{ dets, A, nms_pre, w.boxes, w.labels, cand_idx, w.selected }, B, A, s);
    if (e != cudaSuccess) return e;
    NmsSegs sg{};
    sg.boxes = reinterpret_cast<const float *>(w.boxes); sg.ld = 4; sg.seg_stride = nms_pre; sg.L = 1; sg.K[0] = nms_pre;
    const int nb = (nms_pre + 63) / 64;
    sg.nbp = (nb + 1) & ~1; sg.rows_pad = nb * 64;
    sg.labels = w.labels; sg.agnostic = cfg + 2;
    sg.dyn_k = w.selected;                      // candidates really selected per image: the padding rows cost nothing
    e = run_nms(sg, B, nms_pre, w.nms_cfg, w.mask, w.keep_pos, nms_pre, w.keep_mask, nms_pre, w.count, s);
    if (e != cudaSuccess) return e;
    yolo_gather_kernel<<<B, 128, 0, s>>>(dets, A, nms_pre, max_det, cand_idx, w.selected, w.keep_pos, w.count, out, keep_idx, num_out);
    return cudaGetLastError();
}

}  // namespace md
