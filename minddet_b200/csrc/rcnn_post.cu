// rcnn_post.cu -- "next" row 4: the step after RoIAlign at inference.  Softmax over the class logits,
// per-class delta decode of the RoIs, score threshold, class-aware NMS, top max_det -- plus the stand-alone
// BoundingBoxEncode.  No reference code exists; structural template of the reference's own head post-process:
// pointpillars/src/predict.py:43-98, centerpoint/det3d_ms/models/bbox_heads/center_head.py:398-463
// (score mask -> TopK -> gather -> NMS -> gather keep -> min(count, post_max)).  Semantics: oracle/CONVENTIONS.md #22.
#include "kernels.h"
#include "nms.cuh"
#include "select.cuh"

namespace md {

// one warp per RoI row: p_j = exp(x_j - max) / sum, sequential fp32 sum in class order (lane 0), like the oracle
__global__ void __launch_bounds__(256)
softmax_rows_kernel(const float *__restrict__ logits, int64_t rows, int nc1, float *__restrict__ probs)
{
    const int64_t r = (int64_t)blockIdx.x * 8 + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (r >= rows) return;
    const float *x = logits + r * nc1;
    float *o = probs + r * nc1;
    float m = -3.0e38f;
    for (int j = lane; j < nc1; j += 32) m = fmaxf(m, __ldg(x + j));
#pragma unroll
    for (int s = 16; s > 0; s >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, s));
    for (int j = lane; j < nc1; j += 32) o[j] = exact_exp(sub(__ldg(x + j), m));
    __syncwarp();
    float den = 0.0f;
    if (lane == 0)
        for (int j = 0; j < nc1; j++) den = add(den, o[j]);          // the oracle's summation order
    den = __shfl_sync(0xffffffffu, den, 0);
    for (int j = lane; j < nc1; j += 32) o[j] = div(o[j], den);
}

struct RcnnSrc {
    const float *probs; const uint8_t *roi_valid; int P, nc1, nms_pre; const float *cfg;      // cfg[11] = score_thr
    struct Ctx { const float *base; const uint8_t *valid; float thr; };
    __device__ int segment_of(int i, int it) const { return it ? -1 : i; }
    __device__ Ctx prepare(int seg) const
    {
        return Ctx{ probs + (int64_t)seg * P * nc1, roi_valid ? roi_valid + (int64_t)seg * P : nullptr, __ldg(cfg + 11) };
    }
    __device__ bool active(const Ctx &) const { return true; }
    __device__ int length(const Ctx &) const { return P * nc1; }
    __device__ int want(const Ctx &) const { return nms_pre; }
    __device__ uint32_t index_of(const Ctx &, int m) const { return (uint32_t)m; }
    __device__ bool load(const Ctx &c, int m, uint32_t &key) const
    {
        const int r = m / nc1, cls = m - r * nc1;
        if (cls == 0 || (c.valid && !c.valid[r])) return false;
        const float p = __ldg(c.base + m);
        if (!(p > c.thr)) return false;
        key = score_key(p);
        return true;
    }
};
struct RcnnSink {
    const float *rois; int roi_ld; const float *deltas; int P, nc1, nms_pre; const float *cfg;
    float4 *ws_boxes; float *ws_scores; int32_t *ws_labels; int32_t *cand_idx; int32_t *selected;
    __device__ void emit(int seg, int rank, unsigned long long comp) const
    {
        const int32_t id = (int32_t)(~(uint32_t)comp);
        const int r = id / nc1, cls = id - r * nc1;
        const float *pr = rois + ((int64_t)seg * P + r) * roi_ld + (roi_ld == 5 ? 1 : 0);
        const float4 a = make_float4(__ldg(pr), __ldg(pr + 1), __ldg(pr + 2), __ldg(pr + 3));
        const float4 d = __ldg(reinterpret_cast<const float4 *>(deltas) + ((int64_t)seg * P + r) * nc1 + cls);
        const int64_t o = (int64_t)seg * nms_pre + rank;
        ws_boxes[o] = decode_box(a, d, load_decode_cfg(cfg));
        ws_scores[o] = key_score((uint32_t)(comp >> 32));
        ws_labels[o] = cls;
        cand_idx[o] = id;
    }
    __device__ void pad(int seg, int rank) const
    {
        const int64_t o = (int64_t)seg * nms_pre + rank;
        ws_boxes[o] = make_float4(0, 0, 0, 0);
        ws_scores[o] = 0.0f;
        ws_labels[o] = -1;
        cand_idx[o] = -1;
    }
    __device__ void finish(int seg, int sel, int) const { selected[seg] = sel; }
};

__global__ void rcnn_gather_kernel(int nms_pre, int max_det, const float4 *__restrict__ ws_boxes, const float *__restrict__ ws_scores,
                                   const int32_t *__restrict__ ws_labels, const int32_t *__restrict__ cand_idx,
                                   const int32_t *__restrict__ selected, const int32_t *__restrict__ keep_pos,
                                   const int32_t *__restrict__ count, float *__restrict__ out, int32_t *__restrict__ keep_idx,
                                   int32_t *__restrict__ num_out)
{
    const int b = blockIdx.x;
    const int sel = selected[b];
    const int32_t *kp = keep_pos + (int64_t)b * nms_pre;
    int lo = 0, hi = count[b];
    while (lo < hi) { const int mid = (lo + hi) >> 1; if (kp[mid] < sel) lo = mid + 1; else hi = mid; }
    const int n = min(lo, max_det);
    for (int i = threadIdx.x; i < max_det; i += blockDim.x) {
        float *o = out + ((int64_t)b * max_det + i) * 6;
        int32_t id = -1;
        float4 bx = make_float4(0, 0, 0, 0);
        float sc = 0.0f, lb = 0.0f;
        if (i < n) {
            const int64_t p = (int64_t)b * nms_pre + kp[i];
            bx = ws_boxes[p]; sc = ws_scores[p]; lb = (float)ws_labels[p]; id = cand_idx[p];
        }
        o[0] = bx.x; o[1] = bx.y; o[2] = bx.z; o[3] = bx.w; o[4] = sc; o[5] = lb;
        keep_idx[(int64_t)b * max_det + i] = id;
    }
    if (threadIdx.x == 0) num_out[b] = n;
}

__global__ void rcnn_nms_cfg_kernel(const float *__restrict__ cfg, float *__restrict__ out)
{
    out[0] = cfg[12]; out[1] = 0.0f; out[2] = 0.0f; out[3] = 1e-8f;      // iou_thr, offset 0, strict, eps
}

struct RcnnWs { float *probs; float4 *boxes; float *scores; int32_t *labels, *selected, *keep_pos, *count; uint8_t *keep_mask; float *nms_cfg; unsigned long long *mask; size_t total; };
static inline size_t al256r(size_t x) { return (x + 255) & ~(size_t)255; }
static RcnnWs carve_rcnn_ws(void *ws, int B, int P, int nc1, int nms_pre)
{
    RcnnWs w;
    unsigned char *p = reinterpret_cast<unsigned char *>(ws);
    size_t o = 0;
    w.probs = reinterpret_cast<float *>(p + o); o += al256r((size_t)B * P * nc1 * 4);
    w.boxes = reinterpret_cast<float4 *>(p + o); o += al256r((size_t)B * nms_pre * 16);
    w.scores = reinterpret_cast<float *>(p + o); o += al256r((size_t)B * nms_pre * 4);
    w.labels = reinterpret_cast<int32_t *>(p + o); o += al256r((size_t)B * nms_pre * 4);
    w.keep_pos = reinterpret_cast<int32_t *>(p + o); o += al256r((size_t)B * nms_pre * 4);
    w.keep_mask = reinterpret_cast<uint8_t *>(p + o); o += al256r((size_t)B * nms_pre);
    w.selected = reinterpret_cast<int32_t *>(p + o); o += al256r((size_t)B * 4);
    w.count = reinterpret_cast<int32_t *>(p + o); o += al256r((size_t)B * 4);
    w.nms_cfg = reinterpret_cast<float *>(p + o); o += 256;
    w.mask = reinterpret_cast<unsigned long long *>(p + o); o += nms_workspace_bytes(B, nms_pre);
    w.total = o;
    return w;
}
size_t rcnn_post_workspace_bytes(int B, int P, int nc1, int nms_pre) { return carve_rcnn_ws(nullptr, B, P, nc1, nms_pre).total; }

// cfg: MD_CFG_DECODE (0..10) | score_thr (11) | iou_thr (12)
cudaError_t launch_rcnn_post(const float *rois, int roi_ld, const uint8_t *roi_valid, const float *logits, const float *deltas,
                             int B, int P, int nc1, const float *cfg, void *ws, int nms_pre, int max_det,
                             float *out, int32_t *keep_idx, int32_t *num_out, int32_t *cand_idx, cudaStream_t s)
{
    if (nms_pre > kSelMaxK || (int64_t)P * nc1 >= (1 << kSelMaxIndexBits)) return cudaErrorInvalidValue;
    if (B == 0) return cudaSuccess;
    const RcnnWs w = carve_rcnn_ws(ws, B, P, nc1, nms_pre);
    const int64_t rows = (int64_t)B * P;
    if (rows > 0) softmax_rows_kernel<<<(unsigned)((rows + 7) / 8), 256, 0, s>>>(logits, rows, nc1, w.probs);
    rcnn_nms_cfg_kernel<<<1, 1, 0, s>>>(cfg, w.nms_cfg);
    cudaError_t e = launch_select_sorted(RcnnSrc{ w.probs, roi_valid, P, nc1, nms_pre, cfg },
                                         RcnnSink{ rois, roi_ld, deltas, P, nc1, nms_pre, cfg, w.boxes, w.scores, w.labels, cand_idx, w.selected },
                                         B, P * nc1, s);
    if (e != cudaSuccess) return e;
    NmsSegs sg{};
    sg.boxes = reinterpret_cast<const float *>(w.boxes); sg.ld = 4; sg.seg_stride = nms_pre; sg.L = 1; sg.K[0] = nms_pre;
    const int nb = (nms_pre + 63) / 64;
    sg.nbp = (nb + 1) & ~1; sg.rows_pad = nb * 64;
    sg.labels = w.labels; sg.agnostic = nullptr;
    sg.dyn_k = w.selected;                      // candidates really selected per image: the padding rows cost nothing
    e = run_nms(sg, B, nms_pre, w.nms_cfg, w.mask, w.keep_pos, nms_pre, w.keep_mask, nms_pre, w.count, s);
    if (e != cudaSuccess) return e;
    rcnn_gather_kernel<<<B, 128, 0, s>>>(nms_pre, max_det, w.boxes, w.scores, w.labels, cand_idx, w.selected, w.keep_pos, w.count,
                                         out, keep_idx, num_out);
    return cudaGetLastError();
}

// ---- BoundingBoxEncode on rows (legacy +1 bbox2delta; logf -> FP tolerance) -----------------------------------------
__global__ void __launch_bounds__(256)
encode_rows_kernel(const float4 *__restrict__ props, const float4 *__restrict__ gts, int64_t K, const float *__restrict__ cfg,
                   float4 *__restrict__ out)
{
    float mean[4], stdv[4];
#pragma unroll
    for (int i = 0; i < 4; i++) { mean[i] = __ldg(cfg + i); stdv[i] = __ldg(cfg + 4 + i); }
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < K; i += (int64_t)gridDim.x * blockDim.x)
        stg_stream(out + i, encode_box(ldg_stream(props + i), ldg_stream(gts + i), mean, stdv));
}
cudaError_t launch_encode_rows(const float *props, const float *gts, int64_t K, const float *cfg, float *out, cudaStream_t s)
{
    if (K == 0) return cudaSuccess;
    int blocks = (int)((K + 255) / 256);
    if (blocks > 148 * 8) blocks = 148 * 8;
    encode_rows_kernel<<<blocks, 256, 0, s>>>(reinterpret_cast<const float4 *>(props), reinterpret_cast<const float4 *>(gts), K, cfg,
                                              reinterpret_cast<float4 *>(out));
    return cudaGetLastError();
}

}  // namespace md
