// proposal.cu -- a1..a6 of the region path: anchors, delta decode + clip, per-level top-k,
// bitmask NMS with an on-device sweep, cross-level merge.  sm_100a.
//
// Reference lineage (what is being replaced / followed):
//   * NMS bitmask tiles: nms_normal_kernel, centerpoint/det3d_ms/ops/test_custom_pytorch/
//     iou3d_nms_kernel.cu:361-405 -- same 64x64 tile -> one u64 word layout, but only the upper
//     triangle is launched (the reference's early-out is commented out, :376) ...
//   * ... and the host-side serial reduce (:571-593: cudaMalloc, blocking D2H of N*ceil(N/64)*8 bytes,
//     CPU sweep, H2D) is replaced by nms_sweep_kernel: the greedy sweep runs on the device, one CTA per
//     (image, level) -- a resolver warp walking the 64-box chunks with a warp-wide OR fixed point, helper warps
//     streaming the mask in and folding the kept rows -- zero host round trips.
//   * launch_proposal runs two lanes side by side (finest level | the other levels: top-k -> mask -> sweep each) on the
//     caller's stream and the workspace's helper stream, joined in front of the cross-level merge.
//   * graph order top-k -> gather -> NMS -> gather-keep: center_head.py:435-459.
// Semantics: oracle/CONVENTIONS.md #1-8, #17; oracle/region_oracle.c (o_topk, o_nms, o_proposal_image).
#include <cstdlib>

#include "kernels.h"
#include "nms.cuh"
#include "select.cuh"

namespace md {

// =====================================================================================================
// a1: anchor grid -- pure 128-bit stores, 16 B / anchor
// =====================================================================================================
// one grid row (blockIdx.y) per feature-map row; 32-bit index math (one division by A per anchor)
__global__ void __launch_bounds__(256)
anchor_grid_kernel(const float *__restrict__ base, int A, int H, int W, const float *__restrict__ cfg,
                   float4 *__restrict__ out)
{
    const float stride = __ldg(cfg);
    const int row_len = W * A;
    for (int h = blockIdx.y; h < H; h += gridDim.y) {
        const float sy = mul((float)h, stride);
        float4 *orow = out + (int64_t)h * row_len;
        for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < row_len; i += gridDim.x * blockDim.x) {
            const int w = i / A, a = i - w * A;
            const float sx = mul((float)w, stride);
            const float4 b = __ldg(reinterpret_cast<const float4 *>(base) + a);
            stg_stream(orow + i, make_float4(add(b.x, sx), add(b.y, sy), add(b.z, sx), add(b.w, sy)));
        }
    }
}

cudaError_t launch_anchor_grid(const float *base, int A, int H, int W, const float *cfg, float *out, cudaStream_t s)
{
    if ((int64_t)H * W * A == 0) return cudaSuccess;
    const int row_len = W * A;
    const int gx = (row_len + 255) / 256 < 8 ? (row_len + 255) / 256 : 8;
    anchor_grid_kernel<<<dim3(gx, H < 65535 ? H : 65535), 256, 0, s>>>(base, A, H, W, cfg, reinterpret_cast<float4 *>(out));
    return cudaGetLastError();
}

// =====================================================================================================
// a2: decode on gathered rows (K,4)
// =====================================================================================================
__global__ void __launch_bounds__(256)
decode_rows_kernel(const float4 *__restrict__ anchors, const float4 *__restrict__ deltas, int64_t K,
                   const float *__restrict__ cfg, float4 *__restrict__ out)
{
    const DecodeCfg c = load_decode_cfg(cfg);
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < K; i += (int64_t)gridDim.x * blockDim.x)
        stg_stream(out + i, decode_box(ldg_stream(anchors + i), ldg_stream(deltas + i), c));
}

cudaError_t launch_decode_rows(const float *anchors, const float *deltas, int64_t K, const float *cfg, float *out, cudaStream_t s)
{
    if (K == 0) return cudaSuccess;
    int blocks = (int)((K + 255) / 256);
    if (blocks > 148 * 8) blocks = 148 * 8;
    decode_rows_kernel<<<blocks, 256, 0, s>>>(reinterpret_cast<const float4 *>(anchors),
                                              reinterpret_cast<const float4 *>(deltas), K, cfg,
                                              reinterpret_cast<float4 *>(out));
    return cudaGetLastError();
}

// =====================================================================================================
// a2 (decode-all form): RPN head layout (B,4A,H,W) -> (B,HW*A,4); anchors regenerated in registers.
// Algorithmic bytes: 16 (deltas) + 16 (boxes) per anchor; no anchor read.
// Each thread owns 4 consecutive cells: 4A 128-bit loads (one per delta plane), results are staged in
// shared memory so that the (cell, anchor)-interleaved output leaves as fully coalesced 128-bit stores.
// =====================================================================================================
constexpr int kDecThreads = 128;
constexpr int kDecCells = kDecThreads * 4;

// AU > 0: the anchor count is a compile-time constant and ALL 4*AU delta planes of a thread's four cells are requested
// before the first box is decoded (AU*64 bytes in flight per thread instead of 64: the kernel is a pure stream, and with
// one anchor at a time its loads queued behind its own exps).  AU == 0: any A, one anchor at a time.
template <bool VEC, int AU>
__global__ void __launch_bounds__(kDecThreads)
decode_level_kernel(const float *__restrict__ deltas, const float *__restrict__ base, int A, int H, int W,
                    const float *__restrict__ cfg, float4 *__restrict__ out)
{
    extern __shared__ float4 stage[];   // [kDecCells * A]
    const DecodeCfg c = load_decode_cfg(cfg);
    const float stride = __ldg(cfg + 11);
    const int HW = H * W;
    const int b = blockIdx.y;
    const int tiles = (HW + kDecCells - 1) / kDecCells;
    constexpr int AB = AU > 0 ? AU : 1;      // anchors per batch of loads
    for (int tile = blockIdx.x; tile < tiles; tile += gridDim.x) {
        const int p0 = tile * kDecCells + threadIdx.x * 4;
        float sx[4], sy[4];
#pragma unroll
        for (int j = 0; j < 4; j++) {
            const int p = p0 + j;
            const int w = p % W, h = p / W;
            sx[j] = mul((float)w, stride); sy[j] = mul((float)h, stride);
        }
        for (int a0 = 0; a0 < A; a0 += AB) {
            float d[AB][4][4];   // [anchor][coord][cell]
#pragma unroll
            for (int i = 0; i < AB; i++)
#pragma unroll
                for (int k = 0; k < 4; k++) {
                    const float *plane = deltas + ((int64_t)b * 4 * A + (a0 + i) * 4 + k) * HW;
                    if (VEC) {
                        float4 v = make_float4(0, 0, 0, 0);
                        if (p0 < HW) v = ldg_stream(reinterpret_cast<const float4 *>(plane + p0));
                        d[i][k][0] = v.x; d[i][k][1] = v.y; d[i][k][2] = v.z; d[i][k][3] = v.w;
                    } else {
#pragma unroll
                        for (int j = 0; j < 4; j++) d[i][k][j] = (p0 + j < HW) ? __ldg(plane + p0 + j) : 0.0f;
                    }
                }
#pragma unroll
            for (int i = 0; i < AB; i++) {
                const float4 bs = __ldg(reinterpret_cast<const float4 *>(base) + a0 + i);
#pragma unroll
                for (int j = 0; j < 4; j++) {
                    const float4 anc = make_float4(add(bs.x, sx[j]), add(bs.y, sy[j]), add(bs.z, sx[j]), add(bs.w, sy[j]));
                    stage[(threadIdx.x * 4 + j) * A + a0 + i] =
                        decode_box(anc, make_float4(d[i][0][j], d[i][1][j], d[i][2][j], d[i][3][j]), c);
                }
            }
        }
        __syncthreads();
        const int cells = min(kDecCells, HW - tile * kDecCells);
        float4 *dst = out + ((int64_t)b * HW + (int64_t)tile * kDecCells) * A;
        for (int i = threadIdx.x; i < cells * A; i += kDecThreads) stg_stream(dst + i, stage[i]);
        __syncthreads();
    }
}

cudaError_t launch_decode_level(const float *deltas, const float *base, int B, int A, int H, int W,
                                const float *cfg, float *out, cudaStream_t s)
{
    const int HW = H * W;
    if (B == 0 || HW == 0) return cudaSuccess;
    const size_t smem = (size_t)kDecCells * A * sizeof(float4);
    if (smem > 200 * 1024) return cudaErrorInvalidValue;
    const bool vec = (HW % 4 == 0) && ((reinterpret_cast<uintptr_t>(deltas) & 15) == 0);
    const int tiles = (HW + kDecCells - 1) / kDecCells;
    const int gx = tiles;                                  // one tile per CTA: several waves hide the load -> sync -> store phases
    auto kern = vec ? (A == 3 ? decode_level_kernel<true, 3> : decode_level_kernel<true, 0>) : decode_level_kernel<false, 0>;
    if (smem > 48 * 1024) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
    }
    kern<<<dim3(gx, B), kDecThreads, smem, s>>>(deltas, base, A, H, W, cfg, reinterpret_cast<float4 *>(out));
    return cudaGetLastError();
}

// =====================================================================================================
// a3: top-k sources / sinks for select_sorted_kernel
// =====================================================================================================
struct TopkSrc {            // one level, B segments
    const float *scores; int A, HW, N, K; const float *cfg_sigmoid;
    struct Ctx { const float *base; bool sigmoid; };
    __device__ int segment_of(int i, int it) const { return it ? -1 : i; }
    __device__ bool active(const Ctx &) const { return true; }
    __device__ Ctx prepare(int seg) const { return Ctx{ scores + (int64_t)seg * N, __ldg(cfg_sigmoid) != 0.0f }; }
    __device__ int length(const Ctx &) const { return N; }
    __device__ int want(const Ctx &) const { return K; }
    __device__ uint32_t index_of(const Ctx &, int m) const
    {
        if (A == 0) return (uint32_t)m;
        const int a = m / HW, p = m - a * HW;
        return (uint32_t)(p * A + a);
    }
    __device__ const uint32_t *raw_ptr(const Ctx &c, int m) const { return reinterpret_cast<const uint32_t *>(c.base + m); }
    __device__ bool from_raw(const Ctx &c, int, uint32_t raw, uint32_t &key) const
    {
        float x = __uint_as_float(raw);
        if (c.sigmoid) x = exact_sigmoid(x);
        key = score_key(x);
        return true;
    }
    __device__ bool load(const Ctx &c, int m, uint32_t &key) const { return from_raw(c, m, __ldg(raw_ptr(c, m)), key); }
};
struct TopkSink {
    float *values; int32_t *indices; int K;
    __device__ void emit(int seg, int rank, unsigned long long comp) const
    {
        values[(int64_t)seg * K + rank] = key_score((uint32_t)(comp >> 32));
        indices[(int64_t)seg * K + rank] = (int32_t)(~(uint32_t)comp);
    }
    __device__ void pad(int seg, int rank) const
    {
        values[(int64_t)seg * K + rank] = 0.0f;
        indices[(int64_t)seg * K + rank] = -1;
    }
    __device__ void finish(int, int, int) const {}
};

cudaError_t launch_topk(const float *scores, int B, int A, int HW, int K, const float *cfg_sigmoid,
                        float *values, int32_t *indices, cudaStream_t s)
{
    TopkSrc src{ scores, A, HW, (A == 0 ? HW : A * HW), K, cfg_sigmoid };
    TopkSink sink{ values, indices, K };
    return launch_select_sorted(src, sink, B, src.N, s);
}

// ---- Proposal: all levels in one launch; the sink gathers deltas, regenerates the anchor, decodes ----
struct PropLevels {
    int L, B, nms_pre;
    int A[kMaxLv], W[kMaxLv], HW[kMaxLv], K[kMaxLv];
    const float *scores[kMaxLv], *deltas[kMaxLv], *base[kMaxLv];
    const float *cfg;
};
struct PropSrc {
    PropLevels p;
    int l0, nl;                 // this launch serves levels [l0, l0 + nl)
    struct Ctx { const float *base; int A, HW, N; bool sigmoid; };
    // One cluster per (image, level), level-major: the clusters of the finest (longest) level start first.  About 15
    // clusters of eight 139 KB CTAs are resident at a time; the hardware hands the freed slots to the next clusters.
    // (Measured: giving each image two clusters that work through {0, L-1} and {1 .. L-2} in turn -- 16 clusters -- is
    // slower, 101 vs 66 us: the 16th cluster does not fit and starts when the first level-0 segment ends.)
    __device__ int segment_of(int i, int it) const
    {
        if (it) return -1;
        const int l = i / p.B, b = i - l * p.B;
        return b * p.L + l0 + l;
    }
    __device__ Ctx prepare(int seg) const
    {
        const int l = seg % p.L, b = seg / p.L;
        const int N = p.A[l] * p.HW[l];
        return Ctx{ p.scores[l] + (int64_t)b * N, p.A[l], p.HW[l], N, __ldg(p.cfg + 15) != 0.0f };
    }
    __device__ bool active(const Ctx &) const { return true; }
    __device__ int length(const Ctx &c) const { return c.N; }
    __device__ int want(const Ctx &) const { return p.nms_pre; }   // pads up to nms_pre; selection caps at N
    __device__ uint32_t index_of(const Ctx &c, int m) const
    {
        const int a = m / c.HW, q = m - a * c.HW;
        return (uint32_t)(q * c.A + a);
    }
    __device__ const uint32_t *raw_ptr(const Ctx &c, int m) const { return reinterpret_cast<const uint32_t *>(c.base + m); }
    __device__ bool from_raw(const Ctx &c, int, uint32_t raw, uint32_t &key) const
    {
        float x = __uint_as_float(raw);
        if (c.sigmoid) x = exact_sigmoid(x);
        key = score_key(x);
        return true;
    }
    __device__ bool load(const Ctx &c, int m, uint32_t &key) const { return from_raw(c, m, __ldg(raw_ptr(c, m)), key); }
};
struct PropSink {
    PropLevels p;
    float4 *ws_boxes; float *ws_scores; int32_t *topk_idx;
    struct Pre { float4 dl, bs; float sx, sy; };
    // the four strided delta loads are the long pole of emit: start them before the ranking
    __device__ Pre prefetch(int seg, unsigned long long comp) const
    {
        const int l = seg % p.L, b = seg / p.L;
        const int A = p.A[l], W = p.W[l], HW = p.HW[l];
        const uint32_t n = ~(uint32_t)comp;
        const int a = n % A, q = n / A;
        const int w = q % W, h = q / W;
        const float stride = __ldg(p.cfg + 16 + l);
        Pre r;
        r.sx = mul((float)w, stride); r.sy = mul((float)h, stride);
        r.bs = __ldg(reinterpret_cast<const float4 *>(p.base[l]) + a);
        const float *d = p.deltas[l] + ((int64_t)b * 4 * A + a * 4) * HW + q;
        r.dl = make_float4(__ldg(d), __ldg(d + HW), __ldg(d + 2 * (int64_t)HW), __ldg(d + 3 * (int64_t)HW));
        return r;
    }
    __device__ void emit_pre(int seg, int rank, unsigned long long comp, const Pre &r) const
    {
        const float4 anc = make_float4(add(r.bs.x, r.sx), add(r.bs.y, r.sy), add(r.bs.z, r.sx), add(r.bs.w, r.sy));
        const DecodeCfg c = load_decode_cfg(p.cfg);
        const int64_t o = (int64_t)seg * p.nms_pre + rank;
        ws_boxes[o] = decode_box(anc, r.dl, c);
        ws_scores[o] = key_score((uint32_t)(comp >> 32));
        topk_idx[o] = (int32_t)(~(uint32_t)comp);
    }
    __device__ void emit(int seg, int rank, unsigned long long comp) const { emit_pre(seg, rank, comp, prefetch(seg, comp)); }
    __device__ void pad(int seg, int rank) const
    {
        const int64_t o = (int64_t)seg * p.nms_pre + rank;
        ws_boxes[o] = make_float4(0, 0, 0, 0);
        ws_scores[o] = 0.0f;
        topk_idx[o] = -1;
    }
    __device__ void finish(int, int, int) const {}
};

// =====================================================================================================
// a4: NMS.  (1) bitmask tiles (upper triangle), (2) on-device greedy sweep.
// =====================================================================================================
struct NmsCfg { float thr, off, eps; bool inclusive; };
MD_DEVINL NmsCfg load_nms_cfg(const float *__restrict__ cfg)
{
    NmsCfg c;
    c.thr = __ldg(cfg + 0); c.off = __ldg(cfg + 1); c.inclusive = __ldg(cfg + 2) != 0.0f; c.eps = __ldg(cfg + 3);
    return c;
}
struct BoxA { float x1, y1, x2, y2, area; };

// Decision `inter / max(union, eps) > thr` (>= when inclusive), identical to the oracle's rounded expression.
// Hot path without the division and without branches: e = inter - thr*union differs from its real value by
// < 2e-7*union, and the correctly rounded quotient from the real one by < 6e-8, so outside a 2e-6*union band the
// sign of e decides exactly as the division would.  Pairs inside the band (and NaN / non-positive unions) are
// flagged `unsure` and re-evaluated with the exact expression after the loop.
struct NmsVote { bool sup, unsure; };
MD_DEVINL NmsVote nms_vote(const BoxA &a, const BoxA &b, const NmsCfg &c)
{
    const float left = fmaxf(a.x1, b.x1), right = fminf(a.x2, b.x2);
    const float top = fmaxf(a.y1, b.y1), bottom = fminf(a.y2, b.y2);
    const float w = fmaxf(add(sub(right, left), c.off), 0.0f);
    const float h = fmaxf(add(sub(bottom, top), c.off), 0.0f);
    const float inter = mul(w, h);
    float uni = sub(add(a.area, b.area), inter);
    uni = fmaxf(uni, c.eps);                                      // eps == 0 leaves positive unions untouched
    const float e = sub(inter, mul(c.thr, uni)), band = mul(2e-6f, uni);
    NmsVote v;
    v.sup = (uni > 0.0f) & (e > band);
    v.unsure = !((uni > 0.0f) & ((e > band) | (e < -band)));
    return v;
}
MD_DEVINL bool nms_suppresses(const BoxA &a, const BoxA &b, const NmsCfg &c, bool zero_cond)
{
    const float left = fmaxf(a.x1, b.x1), right = fminf(a.x2, b.x2);
    const float top = fmaxf(a.y1, b.y1), bottom = fminf(a.y2, b.y2);
    const float w = fmaxf(add(sub(right, left), c.off), 0.0f);
    const float h = fmaxf(add(sub(bottom, top), c.off), 0.0f);
    const float inter = mul(w, h);
    if (c.eps > 0.0f && inter == 0.0f) return zero_cond;   // 0 / max(u,eps) == 0 exactly
    float uni = sub(add(a.area, b.area), inter);
    if (c.eps > 0.0f) uni = fmaxf(uni, c.eps);
    const float iou = div(inter, uni);
    return c.inclusive ? (iou >= c.thr) : (iou > c.thr);
}

MD_DEVINL int nms_segment(const NmsSegs &sg, int z) { return sg.nl ? (z / sg.nl) * sg.L + sg.l0 + z % sg.nl : z; }

// grid: (tile, 1, segment) over the upper triangle of 64 x 64 tiles; 64 threads: thread r owns row box r of the tile.
// (kMaskGroup > 1: grid (row block, group of column blocks, segment) and a block walks its group with double-buffered
// column boxes -- kept for reference, measured slower.)
constexpr int kMaskGroup = 1;   // measured: 4 tiles per block 67 us, 1 tile per block 58 us (short blocks hide the box loads)

template <bool LABELS>
__global__ void __launch_bounds__(64)
nms_mask_kernel(const NmsSegs sg, const float *__restrict__ cfg, unsigned long long *__restrict__ mask)
{
    pdl_entry();
    const int seg = nms_segment(sg, blockIdx.z);
    const int K = sg.dyn_k ? min(sg.K[seg % sg.L], max(__ldg(sg.dyn_k + seg), 0)) : sg.K[seg % sg.L];
    const int nb = (K + 63) >> 6;
    int i, j_begin, j_end;
    if (kMaskGroup == 1) {
        // triangular index t -> (row block i, column block j), j >= i: no empty blocks in the grid
        const int t = blockIdx.x;
        if (t >= nb * (nb + 1) / 2) return;
        i = (int)((2.0f * nb + 1.0f - sqrtf((2.0f * nb + 1.0f) * (2.0f * nb + 1.0f) - 8.0f * t)) * 0.5f);
        while (i > 0 && i * (2 * nb - i + 1) / 2 > t) i--;
        while ((i + 1) * (2 * nb - i) / 2 <= t) i++;
        j_begin = i + (t - i * (2 * nb - i + 1) / 2);
        j_end = j_begin + 1;
    } else {
        i = blockIdx.x;
        j_begin = max(i, (int)blockIdx.y * kMaskGroup); j_end = min(nb, ((int)blockIdx.y + 1) * kMaskGroup);
        if (i >= nb || j_begin >= j_end) return;
    }

    if (LABELS && sg.labels_sorted && kMaskGroup == 1 && j_begin > i) {
        // rows grouped by label: every row label <= the last row's <= the first column's <= every column label (in the
        // permutation's order), so the tile holds a same-label pair only if those two are equal
        const int32_t *lab = sg.labels + (int64_t)seg * sg.seg_stride;
        if (__ldg(lab + j_begin * 64) != __ldg(lab + min(K, i * 64 + 64) - 1)) {
            const int r = i * 64 + threadIdx.x;
            if (r < K) mask[((int64_t)seg * sg.rows_pad + r) * sg.nbp + j_begin] = 0ull;
            return;
        }
    }
    const NmsCfg c = load_nms_cfg(cfg);
    const bool zero_cond = c.inclusive ? (0.0f >= c.thr) : (0.0f > c.thr);
    const float *boxes = sg.boxes + (int64_t)seg * sg.seg_stride * sg.ld;
    __shared__ float4 cbox[2][64];         // 128-bit + 32-bit broadcast loads per column (a 5-float struct costs 5 LDS)
    __shared__ float carea[2][64], cthr[2][64];  // area, thr * area
    __shared__ int32_t col_label[2][64];
    const int tid = threadIdx.x;
    const int32_t *labels = (LABELS && sg.labels && !(sg.agnostic && __ldg(sg.agnostic) != 0.0f)) ? sg.labels + (int64_t)seg * sg.seg_stride : nullptr;
    // "plain" box: non-negative width and height and an area the eps clamp cannot touch -- what the fast test below
    // assumes of both boxes of a pair (NaNs fail every comparison and so are not plain either)
    const float tiny = mul(4.0f, c.eps);
    const int ridx = i * 64 + tid;
    const bool row_ok = ridx < K;
    BoxA a = { 0.0f, 0.0f, 0.0f, 0.0f, 0.0f };
    bool row_plain = true;
    if (row_ok) {
        const float *p = boxes + (int64_t)ridx * sg.ld;
        a.x1 = __ldg(p); a.y1 = __ldg(p + 1); a.x2 = __ldg(p + 2); a.y2 = __ldg(p + 3);
        a.area = mul(add(sub(a.x2, a.x1), c.off), add(sub(a.y2, a.y1), c.off));
        row_plain = (a.x2 >= a.x1) & (a.y2 >= a.y1) & (a.area >= tiny) & (a.area > 0.0f);
    }
    const bool use_labels = LABELS && labels;
    const int32_t la = (use_labels && row_ok) ? labels[ridx] : 0;
    // class-aware NMS: the labels of this warp's 32 row boxes as a 64-bucket set.  A column whose label falls in no bucket
    // cannot be suppressed by any row of the warp, and the test is warp-uniform, so the whole 17-instruction vote of that
    // column is skipped (80 random classes: two columns in three).  Exact: the real label comparison stays in the vote.
    uint32_t rs_lo = 0u, rs_hi = 0u;
    if (LABELS) {
        const int bkt = la & 63;
        rs_lo = __reduce_or_sync(0xffffffffu, (use_labels && row_ok && bkt < 32) ? 1u << bkt : 0u);
        rs_hi = __reduce_or_sync(0xffffffffu, (use_labels && row_ok && bkt >= 32) ? 1u << (bkt - 32) : 0u);
    }
    // fast test only with the plain IoU of the detection configs (no legacy +1, a threshold away from 0)
    const bool fast_cfg = c.off == 0.0f && c.thr >= 0.05f && c.thr <= 1.0f;
    const float k1 = add(1.0f, c.thr), cb = div(3e-6f, c.thr), ta = mul(c.thr, a.area);

    for (int j = j_begin; j < j_end; j++) {
        const int buf = j & 1;
        bool plain = row_plain;
        {
            const int cidx = j * 64 + tid;
            float4 b = make_float4(0, 0, 0, 0);
            float area = 0.0f;
            if (LABELS) col_label[buf][tid] = (labels && cidx < K) ? labels[cidx] : 0;
            if (cidx < K) {
                const float *p = boxes + (int64_t)cidx * sg.ld;
                b = make_float4(__ldg(p), __ldg(p + 1), __ldg(p + 2), __ldg(p + 3));
                area = mul(add(sub(b.z, b.x), c.off), add(sub(b.w, b.y), c.off));
                plain = plain & (b.z >= b.x) & (b.w >= b.y) & (area >= tiny) & (area > 0.0f);
            }
            cbox[buf][tid] = b;
            carea[buf][tid] = area;
            cthr[buf][tid] = mul(c.thr, area);
        }
        // one barrier per tile: tile j + 2 rewrites this buffer only after the barrier of tile j + 1, which every thread
        // reaches after it is done with tile j
        const bool fast = !__syncthreads_or(!plain) && fast_cfg;
        if (!row_ok) continue;
        const int ncol = min(64, K - j * 64);
        const int start = (i == j) ? tid + 1 : 0;
        // all 64 columns, fully unrolled (constant bit positions), branch-free; columns outside [start, ncol) are masked
        // off below; the rare pairs the division-free test cannot decide are redone exactly afterwards
        uint32_t half[2] = { 0u, 0u }, unsure[2] = { 0u, 0u };
        if (fast) {
            // iou > thr  <=>  e = inter * (1 + thr) - thr * (area_a + area_b) > 0 in real arithmetic.  Evaluated in fp32,
            // e is off by < 1e-6 * (area_a + area_b), and the oracle's rounded quotient moves the boundary by
            // < 2e-7 * union, so outside the band |e| <= 3e-6 * (area_a + area_b) the sign of e decides exactly as the
            // division does.  17 instructions per pair (the general test below: 32).  A half with any pair inside the
            // band is redone exactly as a whole.
#pragma unroll
            for (int hh = 0; hh < 2; hh++) {
                if (hh * 32 >= ncol || hh * 32 + 32 <= start) continue;
                uint32_t acc = 0u;
                bool close = false;
#pragma unroll
                for (int kk = 0; kk < 32; kk++) {
                    const int k = hh * 32 + kk;
                    if (LABELS && use_labels) {
                        const int cbk = col_label[buf][k] & 63;
                        if (!(((cbk < 32 ? rs_lo : rs_hi) >> (cbk & 31)) & 1u)) continue;
                    }
                    const float4 bx = cbox[buf][k];
                    const float w = fmaxf(sub(fminf(a.x2, bx.z), fmaxf(a.x1, bx.x)), 0.0f);
                    const float h = fmaxf(sub(fminf(a.y2, bx.w), fmaxf(a.y1, bx.y)), 0.0f);
                    const float sum = add(ta, cthr[buf][k]);
                    const float e = __fmaf_rn(mul(w, h), k1, -sum), band = mul(sum, cb);
                    bool sup = e > band;
                    if (LABELS) sup = sup && (!use_labels || col_label[buf][k] == la);
                    if (sup) acc |= 1u << kk;
                    close = close || !(fabsf(e) > band);
                }
                half[hh] = acc;
                unsure[hh] = close ? 0xFFFFFFFFu : 0u;
            }
        } else {
#pragma unroll
            for (int hh = 0; hh < 2; hh++) {
                if (hh * 32 >= ncol || hh * 32 + 32 <= start) continue;
                uint32_t acc = 0u, uns = 0u;
#pragma unroll
                for (int kk = 0; kk < 32; kk++) {
                    const int k = hh * 32 + kk;
                    const float4 bx = cbox[buf][k];
                    const BoxA b = { bx.x, bx.y, bx.z, bx.w, carea[buf][k] };
                    const NmsVote v = nms_vote(a, b, c);
                    bool sup = v.sup;
                    if (LABELS) sup = sup && (!use_labels || col_label[buf][k] == la);
                    acc |= (sup ? 1u : 0u) << kk;
                    uns |= (v.unsure ? 1u : 0u) << kk;
                }
                half[hh] = acc;
                unsure[hh] = uns;
            }
        }
#pragma unroll
        for (int hh = 0; hh < 2; hh++) {
            uint32_t u = unsure[hh];
            while (u) {
                const int kk = __ffs(u) - 1, k = hh * 32 + kk;
                u &= u - 1;
                const float4 bx = cbox[buf][k];
                const BoxA b = { bx.x, bx.y, bx.z, bx.w, carea[buf][k] };
                bool sup = nms_suppresses(a, b, c, zero_cond);
                if (LABELS) sup = sup && (!use_labels || col_label[buf][k] == la);
                half[hh] = (half[hh] & ~(1u << kk)) | ((sup ? 1u : 0u) << kk);
            }
        }
        unsigned long long bits = ((unsigned long long)half[1] << 32) | half[0];
        const unsigned long long lo_mask = start >= 64 ? 0ull : (~0ull << start);
        const unsigned long long hi_mask = ncol >= 64 ? ~0ull : ((1ull << ncol) - 1ull);
        bits &= lo_mask & hi_mask;
        mask[((int64_t)seg * sg.rows_pad + ridx) * sg.nbp + j] = bits;
    }
}

MD_DEVINL void cp_async16(void *smem, const void *gmem)
{
    const uint32_t sa = (uint32_t)__cvta_generic_to_shared(smem);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" :: "r"(sa), "l"(gmem) : "memory");
}
MD_DEVINL void cp_async4(void *smem, const void *gmem)
{
    const uint32_t sa = (uint32_t)__cvta_generic_to_shared(smem);
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" :: "r"(sa), "l"(gmem) : "memory");
}
MD_DEVINL void cp_async8(void *smem, const void *gmem)
{
    const uint32_t sa = (uint32_t)__cvta_generic_to_shared(smem);
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" :: "r"(sa), "l"(gmem) : "memory");
}
MD_DEVINL void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N> MD_DEVINL void cp_async_wait() { asm volatile("cp.async.wait_group %0;" :: "n"(N) : "memory"); }

// One CTA per segment: warp 0 (the resolver) walks the 64-box chunks in score order, warps 1..8 (the helpers) stream
// the chunks' bitmask rows into a kSweepStages-deep shared-memory ring with cp.async and fold them into `removed`.
//   resolver, chunk c:  kept <- alive & ~OR_{b in kept} diag[b]   iterated to its (unique) fixed point == greedy NMS
//                       (the OR across the warp is redux.or), then the kept rows' word c + 1 - the only word chunk c + 1
//                       cannot start without - is ORed into removed[c + 1] from a packed copy of that column;
//   helpers, chunk c-1: the kept rows' words c + 1 .. nb - 1 are ORed into removed[] one iteration later, off the
//                       critical path (8 rows per warp, lane j = word j).
// One __syncthreads per chunk; keep_mask / keep_pos are written by all threads after the walk from kept_all[].
constexpr int kSweepHelpers = 8;                            // helper warps
constexpr int kSweepThreads = 32 * (1 + kSweepHelpers);
// Two instantiations: <32 words, 6 stages> for K <= 2048 (a chunk is issued 4 iterations, ~1 us, before the resolver needs
// its diagonal) and <64 words, 4 stages> for K <= 4096 (the ring is bounded by shared memory: 4 x 32 KB).
constexpr int kSweepMaxNbAny = 64;
template <int kSweepMaxNb, int kSweepStages> constexpr size_t sweep_smem_bytes()
{
    return (size_t)kSweepStages * 64 * kSweepMaxNb * sizeof(unsigned long long);
}

template <int kSweepMaxNb, int kSweepStages>
__global__ void __launch_bounds__(kSweepThreads)
nms_sweep_kernel(const NmsSegs sg, const unsigned long long *__restrict__ mask,
                 const unsigned long long *__restrict__ init_removed /* nullable: (nseg, nbp) boxes dead on entry */,
                 int32_t *__restrict__ keep_pos, int keep_stride, uint8_t *__restrict__ keep_mask,
                 int mask_stride, int32_t *__restrict__ count,
                 unsigned long long *__restrict__ kept_bits /* non-null: only the kept bitmask (nseg, nbp) is written */)
{
    pdl_entry();
    extern __shared__ __align__(16) unsigned long long stage_raw[];
    unsigned long long (*stage)[64 * kSweepMaxNb] = reinterpret_cast<unsigned long long (*)[64 * kSweepMaxNb]>(stage_raw);
    __shared__ unsigned long long diag[kSweepStages][64];    // word c     of the rows of chunk c (the 64 x 64 diagonal block)
    __shared__ unsigned long long next[kSweepStages][64];    // word c + 1 of the rows of chunk c
    __shared__ unsigned long long removed[kSweepMaxNb];      // word j of "suppressed by an earlier kept box"
    __shared__ unsigned long long kept_all[kSweepMaxNb];
    __shared__ int kept_before[kSweepMaxNb + 1];
    __shared__ float score_sh[64 * kSweepMaxNb];
    const int seg = nms_segment(sg, blockIdx.x);
    const int K = sg.dyn_k ? min(sg.K[seg % sg.L], max(__ldg(sg.dyn_k + seg), 0)) : sg.K[seg % sg.L];
    const int nb = (K + 63) >> 6, nbp = sg.nbp;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int htid = tid - 32, hw = warp - 1;                // helper thread / warp index
    const unsigned long long *m = mask + (int64_t)seg * sg.rows_pad * nbp;
    int32_t *kp = keep_pos + (int64_t)seg * keep_stride;
    uint8_t *km = keep_mask + (int64_t)seg * mask_stride;

    // helpers only.  Rows are contiguous in the workspace and in the stage (same pitch), so a chunk is ONE flat copy of
    // 64 * nbp words in 16-byte pieces; the two columns the resolver reads are copied once more, packed (in the stage
    // they sit nbp words apart, i.e. in one bank)
    auto issue = [&](int c) {
        if (c < nb) {
            unsigned long long *dst = stage[c % kSweepStages];
            const unsigned long long *src = m + (int64_t)c * 64 * nbp;
            const int total = 32 * nbp;                        // 16-byte pieces
            for (int q = htid; q < total; q += 32 * kSweepHelpers) cp_async16(dst + 2 * q, src + 2 * q);
            if (htid < 64) cp_async8(&diag[c % kSweepStages][htid], src + (int64_t)htid * nbp + c);
            else if (htid < 128 && c + 1 < nb) cp_async8(&next[c % kSweepStages][htid - 64], src + (int64_t)(htid - 64) * nbp + c + 1);
        }
        cp_async_commit();                                      // empty groups keep the wait count uniform
    };

    if (tid < kSweepMaxNb) removed[tid] = (init_removed && tid < nb) ? init_removed[(int64_t)seg * nbp + tid] : 0ull;
    if (warp > 0) {
        // scores of the segment (only when the kept keys are wanted): they ride in the first cp.async group, so the
        // epilogue finds them in shared memory instead of paying an L2 round trip per kept box
        if (sg.kept_keys) {
            const float *sc = sg.scores + (int64_t)seg * sg.seg_stride;
            for (int q = htid; q < K; q += 32 * kSweepHelpers) cp_async4(&score_sh[q], sc + q);
        }
        for (int c = 0; c < kSweepStages - 2; c++) issue(c);
    }
    constexpr int kRowsPerWarp = 64 / kSweepHelpers;
    for (int c = 0; c < nb; c++) {
        if (warp > 0) cp_async_wait<kSweepStages - 3>();        // chunk c has landed (this thread's pieces)
        __syncthreads();    // ... all pieces; removed[c] is final; kept_all[c - 1] is visible; stage (c - 2) % S is free
        if (warp == 0) {
            const int nrows = min(64, K - c * 64);
            const unsigned long long vmask = nrows == 64 ? ~0ull : ((1ull << nrows) - 1ull);
            const unsigned long long alive = ~removed[c] & vmask;
            const unsigned long long *dg = diag[c % kSweepStages], *nx = next[c % kSweepStages];
            const unsigned long long d_lo = dg[lane], d_hi = dg[lane + 32];
            unsigned long long n_lo = 0ull, n_hi = 0ull;
            if (c + 1 < nb) { n_lo = nx[lane]; n_hi = nx[lane + 32]; }
            unsigned long long kept = alive;
            for (int it = 0; it < 64; it++) {
                unsigned long long sup = (((kept >> lane) & 1ull) ? d_lo : 0ull) |
                                         (((kept >> (lane + 32)) & 1ull) ? d_hi : 0ull);
                const uint32_t lo = __reduce_or_sync(0xffffffffu, (uint32_t)sup);
                const uint32_t hi = __reduce_or_sync(0xffffffffu, (uint32_t)(sup >> 32));
                const unsigned long long nk = alive & ~(((unsigned long long)hi << 32) | lo);
                if (nk == kept) break;
                kept = nk;
            }
            if (c + 1 < nb) {
                unsigned long long sup = (((kept >> lane) & 1ull) ? n_lo : 0ull) |
                                         (((kept >> (lane + 32)) & 1ull) ? n_hi : 0ull);
                const uint32_t lo = __reduce_or_sync(0xffffffffu, (uint32_t)sup);
                const uint32_t hi = __reduce_or_sync(0xffffffffu, (uint32_t)(sup >> 32));
                const unsigned long long v = ((unsigned long long)hi << 32) | lo;
                if (lane == 0 && v) atomicOr(&removed[c + 1], v);    // helpers OR chunk c - 1's rows into the same word
            }
            if (lane == 0) kept_all[c] = kept;
        } else {
            issue(c + kSweepStages - 2);
            // rows of chunk c - 1, words c + 1 ..: `kept` is warp-uniform, so a row that was not kept costs one predicate
            if (c >= 1) {
                const unsigned long long kept = kept_all[c - 1] >> (hw * kRowsPerWarp);
#pragma unroll
                for (int w0 = 0; w0 < kSweepMaxNb; w0 += 32) {
                    const int w = w0 + lane;                         // lane j = words j, j + 32
                    if (w > c && w < nb) {
                        const unsigned long long *col = stage[(c - 1) % kSweepStages] + (hw * kRowsPerWarp) * nbp + w;
                        unsigned long long acc = 0ull;
#pragma unroll
                        for (int b = 0; b < kRowsPerWarp; b++)
                            if ((kept >> b) & 1ull) acc |= col[b * nbp];
                        if (acc) atomicOr(&removed[w], acc);
                    }
                }
            }
        }
    }
    cp_async_wait<0>();
    __syncthreads();
    if (kept_bits) {                                             // label-major run: nms_unpermute_kernel writes the outputs
        for (int w = tid; w < nbp; w += kSweepThreads) kept_bits[(int64_t)seg * nbp + w] = w < nb ? kept_all[w] : 0ull;
        return;
    }
    // outputs: kept_before[c] = boxes kept in chunks < c, then every thread writes its boxes
    if (warp == 0) {
        int base = 0;
#pragma unroll
        for (int w0 = 0; w0 < kSweepMaxNb; w0 += 32) {
            const int n = w0 + lane < nb ? __popcll(kept_all[w0 + lane]) : 0;
            int incl = n;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                const int t = __shfl_up_sync(0xffffffffu, incl, d);
                if (lane >= d) incl += t;
            }
            kept_before[w0 + lane] = base + incl - n;
            base += __shfl_sync(0xffffffffu, incl, 31);
        }
        if (lane == 0) kept_before[kSweepMaxNb] = base;
    }
    __syncthreads();
    const int nkept = kept_before[kSweepMaxNb];
    for (int i = tid; i < mask_stride; i += kSweepThreads) {
        bool k = false;
        if (i < K) {
            const unsigned long long w = kept_all[i >> 6];
            const int b = i & 63;
            k = (w >> b) & 1ull;
            if (k) {
                const int r = kept_before[i >> 6] + __popcll(w & ((1ull << b) - 1ull));
                kp[r] = i;
                if (sg.kept_keys) sg.kept_keys[(int64_t)seg * keep_stride + r] = score_key(score_sh[i]);
            }
        }
        km[i] = k;
    }
    for (int i = nkept + tid; i < keep_stride; i += kSweepThreads) kp[i] = 0;
    if (tid == 0) count[seg] = nkept;
}

// ---- class-aware NMS: label-major permutation ------------------------------------------------------------------
// Boxes of different labels never suppress each other, so greedy NMS is independent per label.  The candidates of a
// segment (score order) are stably re-ordered by label: the suppression matrix becomes block diagonal, the mask kernel
// skips every tile whose label ranges do not meet (nms_mask_kernel), the sweep runs unchanged on the permuted order, and
// nms_unpermute_kernel maps the kept flags back to score order and writes the sweep's usual outputs.  A dense crowd of
// 2048 candidates over 80 classes needs ~1/5 of the tiles; an agnostic run (device flag) keeps the identity order.
constexpr int kPermThreads = 1024;
constexpr int kPermMaxK = 4096;

__global__ void __launch_bounds__(kPermThreads)
nms_label_perm_kernel(const NmsSegs sg, int Kp, int32_t *__restrict__ perm, float4 *__restrict__ boxes_p, int32_t *__restrict__ labels_p)
{
    __shared__ unsigned long long keys[kPermMaxK];
    const int seg = blockIdx.x, tid = threadIdx.x;
    const int K = sg.dyn_k ? min(sg.K[seg % sg.L], max(__ldg(sg.dyn_k + seg), 0)) : sg.K[seg % sg.L];
    const bool agn = sg.agnostic && __ldg(sg.agnostic) != 0.0f;
    const int32_t *lab = sg.labels + (int64_t)seg * sg.seg_stride;
    int n2 = 64;
    while (n2 < K) n2 <<= 1;
    for (int i = tid; i < n2; i += kPermThreads)
        keys[i] = i < K ? ((unsigned long long)(agn ? 0u : ((uint32_t)lab[i] ^ 0x80000000u)) << 32) | (uint32_t)i : ~0ull;
    __syncthreads();
    // bitonic sort, one compare-exchange per thread and step (pair t -> elements i, i + j); steps whose partners sit
    // inside one warp's 64 elements need no block barrier
    for (int k = 2; k <= n2; k <<= 1)
        for (int j = k >> 1; j > 0; j >>= 1) {
            for (int t = tid; t < (n2 >> 1); t += kPermThreads) {
                const int i = 2 * t - (t & (j - 1)), l = i + j;
                const unsigned long long a = keys[i], b = keys[l];
                if ((a > b) == ((i & k) == 0)) { keys[i] = b; keys[l] = a; }
            }
            if (j > 32 || j == 1) __syncthreads(); else __syncwarp();   // j == 1 ends a phase: the next one starts across warps
        }
    const float *boxes = sg.boxes + (int64_t)seg * sg.seg_stride * sg.ld;
    for (int i = tid; i < K; i += kPermThreads) {
        const int pos = (int)(uint32_t)keys[i];
        const float *p = boxes + (int64_t)pos * sg.ld;
        perm[(int64_t)seg * Kp + i] = pos;
        boxes_p[(int64_t)seg * Kp + i] = make_float4(__ldg(p), __ldg(p + 1), __ldg(p + 2), __ldg(p + 3));
        labels_p[(int64_t)seg * Kp + i] = agn ? 0 : lab[pos];
    }
}

// kept flags of the permuted order -> score order; outputs exactly as nms_sweep_kernel's epilogue writes them
__global__ void __launch_bounds__(kPermThreads)
nms_unpermute_kernel(const NmsSegs sg, int Kp, const int32_t *__restrict__ perm, const unsigned long long *__restrict__ kept_bits,
                     int32_t *__restrict__ keep_pos, int keep_stride, uint8_t *__restrict__ keep_mask, int mask_stride,
                     int32_t *__restrict__ count)
{
    __shared__ uint32_t words[kPermMaxK / 32];
    __shared__ int before[kPermMaxK / 32 + 1];
    const int seg = blockIdx.x, tid = threadIdx.x, lane = tid & 31;
    const int K = sg.dyn_k ? min(sg.K[seg % sg.L], max(__ldg(sg.dyn_k + seg), 0)) : sg.K[seg % sg.L];
    if (tid < kPermMaxK / 32) words[tid] = 0u;
    __syncthreads();
    for (int p = tid; p < K; p += kPermThreads)
        if ((kept_bits[(int64_t)seg * sg.nbp + (p >> 6)] >> (p & 63)) & 1ull) {
            const int pos = perm[(int64_t)seg * Kp + p];
            atomicOr(&words[pos >> 5], 1u << (pos & 31));
        }
    __syncthreads();
    if (tid < 32) {
        int base = 0;
        for (int w0 = 0; w0 < kPermMaxK / 32; w0 += 32) {
            const int n = __popc(words[w0 + lane]);
            int incl = n;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                const int t = __shfl_up_sync(0xffffffffu, incl, d);
                if (lane >= d) incl += t;
            }
            before[w0 + lane] = base + incl - n;
            base += __shfl_sync(0xffffffffu, incl, 31);
        }
        if (lane == 0) before[kPermMaxK / 32] = base;
    }
    __syncthreads();
    const int nkept = before[kPermMaxK / 32];
    int32_t *kp = keep_pos + (int64_t)seg * keep_stride;
    uint8_t *km = keep_mask + (int64_t)seg * mask_stride;
    for (int i = tid; i < mask_stride; i += kPermThreads) {
        bool k = false;
        if (i < K) {
            const uint32_t w = words[i >> 5];
            k = (w >> (i & 31)) & 1u;
            if (k) kp[before[i >> 5] + __popc(w & ((1u << (i & 31)) - 1u))] = i;
        }
        km[i] = k;
    }
    for (int i = nkept + tid; i < keep_stride; i += kPermThreads) kp[i] = 0;
    if (tid == 0) count[seg] = nkept;
}

static size_t nms_mask_bytes(int nseg, int Kmax)
{
    const int nb = (Kmax + 63) / 64, nbp = (nb + 1) & ~1;
    return (((size_t)nseg * nb * 64 * nbp * sizeof(unsigned long long)) + 255) & ~(size_t)255;
}
// mask tiles + the label-major scratch (permutation, permuted boxes and labels, kept bitmask)
size_t nms_workspace_bytes(int nseg, int Kmax)
{
    const int nb = (Kmax + 63) / 64, nbp = (nb + 1) & ~1;
    const size_t rows = (size_t)nseg * nb * 64;
    return nms_mask_bytes(nseg, Kmax) + rows * (4 + 16 + 4) + (size_t)nseg * nbp * 8 + 4 * 256;
}

cudaError_t run_nms(const NmsSegs &sg, int nseg, int Kmax, const float *cfg, unsigned long long *mask,
                    int32_t *keep_pos, int keep_stride, uint8_t *keep_mask, int mask_stride,
                    int32_t *count, cudaStream_t s)
{
    if (nseg == 0) return cudaSuccess;
    const int nb = (Kmax + 63) / 64;
    if (nb > kSweepMaxNbAny) return cudaErrorInvalidValue;
    const int tiles = nb * (nb + 1) / 2;
    if (sg.labels && kMaskGroup == 1 && Kmax <= kPermMaxK && tiles > 0) {
        // class-aware: label-major permutation -> block-diagonal mask -> sweep -> back to score order
        const int Kp = nb * 64;
        unsigned char *x = reinterpret_cast<unsigned char *>(mask) + nms_mask_bytes(nseg, Kmax);
        float4 *boxes_p = reinterpret_cast<float4 *>(x); x += (((size_t)nseg * Kp * 16) + 255) & ~(size_t)255;
        int32_t *perm = reinterpret_cast<int32_t *>(x); x += (((size_t)nseg * Kp * 4) + 255) & ~(size_t)255;
        int32_t *labels_p = reinterpret_cast<int32_t *>(x); x += (((size_t)nseg * Kp * 4) + 255) & ~(size_t)255;
        unsigned long long *kept = reinterpret_cast<unsigned long long *>(x);
        nms_label_perm_kernel<<<nseg, kPermThreads, 0, s>>>(sg, Kp, perm, boxes_p, labels_p);
        NmsSegs sp = sg;
        sp.boxes = reinterpret_cast<const float *>(boxes_p); sp.ld = 4; sp.seg_stride = Kp;
        sp.labels = labels_p; sp.agnostic = nullptr; sp.labels_sorted = 1;
        nms_mask_kernel<true><<<dim3(tiles, 1, nseg), 64, 0, s>>>(sp, cfg, mask);     // behind the permutation kernel: plain launch
        if (nb <= 32) {
            cudaError_t e = cudaFuncSetAttribute(nms_sweep_kernel<32, 6>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sweep_smem_bytes<32, 6>());
            if (e != cudaSuccess) return e;
            nms_sweep_kernel<32, 6><<<nseg, kSweepThreads, sweep_smem_bytes<32, 6>(), s>>>(sp, mask, nullptr, keep_pos, keep_stride, keep_mask, mask_stride, count, kept);
        } else {
            cudaError_t e = cudaFuncSetAttribute(nms_sweep_kernel<64, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sweep_smem_bytes<64, 4>());
            if (e != cudaSuccess) return e;
            nms_sweep_kernel<64, 4><<<nseg, kSweepThreads, sweep_smem_bytes<64, 4>(), s>>>(sp, mask, nullptr, keep_pos, keep_stride, keep_mask, mask_stride, count, kept);
        }
        nms_unpermute_kernel<<<nseg, kPermThreads, 0, s>>>(sg, Kp, perm, kept, keep_pos, keep_stride, keep_mask, mask_stride, count);
        return cudaGetLastError();
    }
    if (tiles > 0) {
        const dim3 grid = kMaskGroup == 1 ? dim3(tiles, 1, nseg) : dim3(nb, (nb + kMaskGroup - 1) / kMaskGroup, nseg);
        cudaError_t e = sg.labels ? launch_pdl(nms_mask_kernel<true>, grid, dim3(64), 0, s, sg, cfg, mask)
                                  : launch_pdl(nms_mask_kernel<false>, grid, dim3(64), 0, s, sg, cfg, mask);
        if (e != cudaSuccess) return e;
    }
    if (nb <= 32) {
        // the attribute is per DEVICE and the value is a constant: set it on every launch (no process-wide "done" flag)
        cudaError_t e = cudaFuncSetAttribute(nms_sweep_kernel<32, 6>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sweep_smem_bytes<32, 6>());
        if (e != cudaSuccess) return e;
        return launch_pdl(nms_sweep_kernel<32, 6>, dim3(nseg), dim3(kSweepThreads), sweep_smem_bytes<32, 6>(), s, sg, mask, (const unsigned long long *)nullptr,
                          keep_pos, keep_stride, keep_mask, mask_stride, count, (unsigned long long *)nullptr);
    } else {
        cudaError_t e = cudaFuncSetAttribute(nms_sweep_kernel<64, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sweep_smem_bytes<64, 4>());
        if (e != cudaSuccess) return e;
        return launch_pdl(nms_sweep_kernel<64, 4>, dim3(nseg), dim3(kSweepThreads), sweep_smem_bytes<64, 4>(), s, sg, mask, (const unsigned long long *)nullptr,
                          keep_pos, keep_stride, keep_mask, mask_stride, count, (unsigned long long *)nullptr);
    }
    return cudaGetLastError();
}

cudaError_t launch_nms_sweep_single(const unsigned long long *mask, const unsigned long long *init_removed, int n, int nbp,
                                    int32_t *keep_pos, uint8_t *keep_mask, int32_t *count, cudaStream_t s)
{
    NmsSegs sg{};
    sg.L = 1; sg.K[0] = n; sg.nbp = nbp; sg.rows_pad = ((n + 63) / 64) * 64;
    const int nb = (n + 63) / 64;
    if (nb > kSweepMaxNbAny) return cudaErrorInvalidValue;
    if (nb <= 32) {
        cudaError_t e = cudaFuncSetAttribute(nms_sweep_kernel<32, 6>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sweep_smem_bytes<32, 6>());
        if (e != cudaSuccess) return e;
        nms_sweep_kernel<32, 6><<<1, kSweepThreads, sweep_smem_bytes<32, 6>(), s>>>(sg, mask, init_removed, keep_pos, n, keep_mask, n, count, nullptr);
    } else {
        cudaError_t e = cudaFuncSetAttribute(nms_sweep_kernel<64, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sweep_smem_bytes<64, 4>());
        if (e != cudaSuccess) return e;
        nms_sweep_kernel<64, 4><<<1, kSweepThreads, sweep_smem_bytes<64, 4>(), s>>>(sg, mask, init_removed, keep_pos, n, keep_mask, n, count, nullptr);
    }
    return cudaGetLastError();
}

cudaError_t launch_nms(const float *boxes, int ld, int B, int K, const float *cfg, void *ws,
                       int32_t *keep_idx, uint8_t *mask, int32_t *count, cudaStream_t s)
{
    NmsSegs sg{};
    sg.boxes = boxes; sg.ld = ld; sg.seg_stride = K; sg.L = 1; sg.K[0] = K;
    const int nb = (K + 63) / 64;
    sg.nbp = (nb + 1) & ~1; sg.rows_pad = nb * 64;
    return run_nms(sg, B, K, cfg, reinterpret_cast<unsigned long long *>(ws), keep_idx, K, mask, K, count, s);
}

// =====================================================================================================
// a5: cross-level merge.  Kept boxes come first ordered by (score desc, concat index asc); suppressed
// boxes follow in concat order (they all carry the -65536 sentinel upstream).  Every level's kept list
// is already sorted, so a box's final rank is a sum of binary searches -- no second global top-k.
// Precondition: scores > -65536.
// =====================================================================================================
constexpr int kMergeThreads = 512;
constexpr int kMergeSplit = 32;     // measured 8 / 16 / 32 / 64 / 128: MdProposal 155.1 / 150.5 / 146.9 / 149.4 / 153.0 us

// suppressed boxes: rank = total_kept + (#suppressed valid boxes before me in concat order)
// (second half of merge_levels_kernel: same grid, disjoint output ranks)
__device__ void merge_suppressed(int L, int nms_pre, int max_num,
                                 const float4 *__restrict__ ws_boxes, const float *__restrict__ ws_scores,
                                 const uint8_t *__restrict__ keep_mask, const int32_t *__restrict__ keep_pos,
                                 const int32_t *__restrict__ count, const NmsSegs &sg,
                                 float *__restrict__ props, uint8_t *__restrict__ pmask)
{
    __shared__ int cnt[kMaxLv], cumk[kMaxLv + 1], cumv[kMaxLv + 1];
    const int b = blockIdx.x, tid = threadIdx.x;
    if (tid == 0) {
        int c = 0, v = 0;
        for (int l = 0; l < L; l++) { cnt[l] = count[b * L + l]; cumk[l] = c; c += cnt[l]; cumv[l] = v; v += sg.K[l]; }
        cumk[L] = c; cumv[L] = v;
    }
    __syncthreads();
    const int total_kept = cumk[L], total_valid = cumv[L];
    for (int e = blockIdx.y * kMergeThreads + tid; e < L * nms_pre && total_kept < max_num; e += kMergeSplit * kMergeThreads) {
        const int l = e / nms_pre, i = e - l * nms_pre;
        if (i >= sg.K[l]) continue;
        const int64_t seg = (int64_t)b * L + l;
        if (keep_mask[seg * nms_pre + i]) continue;
        int lo = 0, hi = cnt[l];
        const int32_t *kp = keep_pos + seg * nms_pre;
        while (lo < hi) { const int mid = (lo + hi) >> 1; if (kp[mid] < i) lo = mid + 1; else hi = mid; }
        const int concat = cumv[l] + i;
        const int kept_before = cumk[l] + lo;
        const int rank = total_kept + (concat - kept_before);
        if (rank < max_num) {
            const float4 bx = ws_boxes[seg * nms_pre + i];
            float *o = props + ((int64_t)b * max_num + rank) * 5;
            o[0] = bx.x; o[1] = bx.y; o[2] = bx.z; o[3] = bx.w; o[4] = ws_scores[seg * nms_pre + i];
            pmask[(int64_t)b * max_num + rank] = 0;
        }
    }
    // zero padding when fewer than max_num boxes exist at all
    for (int r = total_valid + blockIdx.y * kMergeThreads + tid; r < max_num; r += kMergeSplit * kMergeThreads) {
        float *o = props + ((int64_t)b * max_num + r) * 5;
        o[0] = o[1] = o[2] = o[3] = o[4] = 0.0f;
        pmask[(int64_t)b * max_num + r] = 0;
    }
}

__global__ void __launch_bounds__(kMergeThreads)
merge_levels_kernel(int L, int nms_pre, int max_num, const float4 *__restrict__ ws_boxes,
                    const uint32_t *__restrict__ kept_keys,
                    const int32_t *__restrict__ keep_pos, const int32_t *__restrict__ count,
                    const float *__restrict__ ws_scores, const uint8_t *__restrict__ keep_mask, const NmsSegs sg,
                    float *__restrict__ props, uint8_t *__restrict__ pmask)
{
    extern __shared__ uint32_t kkeys[];   // [L][nms_pre] keys of the kept boxes, descending per level
    __shared__ int cnt[kMaxLv];
    const int b = blockIdx.x;
    const int tid = threadIdx.x;
    if (tid < L) cnt[tid] = count[b * L + tid];
    __syncthreads();
    // keys of the kept boxes, written in kept order by the NMS sweep (gathering them here through keep_pos -> score,
    // once per y-block, was 80 MB of L2 sector traffic and most of this kernel's 25 us)
    if ((nms_pre & 3) == 0) {
        // 16-byte cp.async pieces, all in flight at once (a load-then-store loop waited one L2 round trip per piece)
        const int q = nms_pre >> 2;
        for (int i = tid; i < L * q; i += kMergeThreads) {
            const int l = i / q, p4 = (i - l * q) * 4;
            if (p4 < cnt[l]) cp_async16(kkeys + l * nms_pre + p4, kept_keys + ((int64_t)b * L + l) * nms_pre + p4);
        }
        cp_async_commit();
        cp_async_wait<0>();
    } else {
        for (int i = tid; i < L * nms_pre; i += kMergeThreads) {
            const int l = i / nms_pre, t = i - l * nms_pre;
            if (t < cnt[l]) kkeys[i] = kept_keys[((int64_t)b * L + l) * nms_pre + t];
        }
    }
    __syncthreads();
    // one kept box per thread: entry t of level l's kept list.  Its rank inside its level is t itself; the other levels
    // contribute a branch-free lower bound each, all in lockstep (the searches are independent)
    int top_step = 1;
    while (top_step * 2 <= nms_pre) top_step <<= 1;
    for (int e = blockIdx.y * kMergeThreads + tid; e < L * nms_pre; e += kMergeSplit * kMergeThreads) {
        const int l = e / nms_pre, t = e - l * nms_pre;
        if (t >= cnt[l]) continue;
        const int64_t seg = (int64_t)b * L + l;
        const int i = keep_pos[seg * nms_pre + t];               // position in the level's pre-NMS list
        const float4 bx = ws_boxes[seg * nms_pre + i];
        const uint32_t key = kkeys[e];
        int pos[kMaxLv];
#pragma unroll
        for (int l2 = 0; l2 < kMaxLv; l2++) pos[l2] = 0;
        for (int st = top_step; st > 0; st >>= 1) {
#pragma unroll
            for (int l2 = 0; l2 < kMaxLv; l2++) {
                if (l2 >= L) break;
                // #kept in level l2 that sort before me: key2 > key, or == when l2 < l
                const int probe = pos[l2] + st;
                if (probe <= cnt[l2]) {
                    const uint32_t k2 = kkeys[l2 * nms_pre + probe - 1];
                    if ((l2 < l) ? (k2 >= key) : (k2 > key)) pos[l2] = probe;
                }
            }
        }
        int rank = t;
#pragma unroll
        for (int l2 = 0; l2 < kMaxLv; l2++)
            if (l2 < L && l2 != l) rank += pos[l2];
        if (rank < max_num) {
            float *o = props + ((int64_t)b * max_num + rank) * 5;
            o[0] = bx.x; o[1] = bx.y; o[2] = bx.z; o[3] = bx.w; o[4] = key_score(key);
            pmask[(int64_t)b * max_num + rank] = 1;
        }
    }
    merge_suppressed(L, nms_pre, max_num, ws_boxes, ws_scores, keep_mask, keep_pos, count, sg, props, pmask);
}

// workspace layout (all per segment = b*L + l, nms_pre rows each)
struct PropWs {
    float4 *boxes; float *scores; int32_t *keep_pos; uint32_t *kept_keys; int32_t *count; unsigned long long *mask;
};
static inline size_t align256(size_t x) { return (x + 255) & ~(size_t)255; }
static PropWs carve_prop_ws(void *ws, int nseg, int nms_pre, size_t *total)
{
    PropWs w;
    size_t o = 0;
    unsigned char *p = reinterpret_cast<unsigned char *>(ws);
    w.boxes = reinterpret_cast<float4 *>(p + o); o += align256((size_t)nseg * nms_pre * sizeof(float4));
    w.scores = reinterpret_cast<float *>(p + o); o += align256((size_t)nseg * nms_pre * sizeof(float));
    w.keep_pos = reinterpret_cast<int32_t *>(p + o); o += align256((size_t)nseg * nms_pre * sizeof(int32_t));
    w.kept_keys = reinterpret_cast<uint32_t *>(p + o); o += align256((size_t)nseg * nms_pre * sizeof(uint32_t));
    w.count = reinterpret_cast<int32_t *>(p + o); o += align256((size_t)nseg * sizeof(int32_t));
    w.mask = reinterpret_cast<unsigned long long *>(p + o); o += nms_workspace_bytes(nseg, nms_pre);
    if (total) *total = o;
    return w;
}
size_t proposal_workspace_bytes(int B, int L, int nms_pre)
{
    size_t t = 0;
    carve_prop_ws(nullptr, B * L, nms_pre, &t);
    return t;
}

cudaError_t launch_proposal(const LevelSet &lv, int B, int nms_pre, int max_num, const float *cfg, void *ws,
                            float *props, uint8_t *pmask, int32_t *topk_idx, uint8_t *keep, cudaStream_t s, const SideLane *side)
{
    if (lv.L < 1 || lv.L > kMaxLv || nms_pre > kSelMaxK) return cudaErrorInvalidValue;
    const int L = lv.L, nseg = B * L;
    if (nseg == 0) return cudaSuccess;
    PropWs w = carve_prop_ws(ws, nseg, nms_pre, nullptr);
    PropLevels pl{};
    pl.L = L; pl.B = B; pl.nms_pre = nms_pre; pl.cfg = cfg;
    int maxN = 0, Kmax = 0;
    NmsSegs sg{};
    for (int l = 0; l < L; l++) {
        pl.A[l] = lv.A[l]; pl.W[l] = lv.W[l]; pl.HW[l] = lv.H[l] * lv.W[l];
        const int N = lv.A[l] * lv.H[l] * lv.W[l];
        if (N >= (1 << kSelMaxIndexBits)) return cudaErrorInvalidValue;
        pl.K[l] = N < nms_pre ? N : nms_pre;
        pl.scores[l] = lv.scores[l]; pl.deltas[l] = lv.deltas[l]; pl.base[l] = lv.base[l];
        sg.K[l] = pl.K[l];
        if (N > maxN) maxN = N;
        if (pl.K[l] > Kmax) Kmax = pl.K[l];
    }
    sg.boxes = reinterpret_cast<const float *>(w.boxes); sg.ld = 4; sg.seg_stride = nms_pre; sg.L = L;
    sg.scores = w.scores; sg.kept_keys = w.kept_keys;
    const int nb = (nms_pre + 63) / 64;
    sg.nbp = (nb + 1) & ~1; sg.rows_pad = nb * 64;
    cudaError_t e;
    if (side && L >= 2) {
        // Two lanes side by side, joined in front of the merge: the finest level on the caller's stream, the other
        // levels on the workspace's helper stream, each lane top-k -> NMS mask -> NMS sweep.
        //  * top-k: the finest level's clusters need 139 KB of shared memory per CTA (one CTA per SM, ~15 clusters
        //    resident); sized by the second level the others need 60 KB, so all of them are resident at once instead of
        //    queueing behind the big ones for three rounds;
        //  * the other levels finish their top-k first, and their (issue-bound) mask kernel then fills the SMs the
        //    (latency-bound) level-0 top-k leaves mostly idle.
        // Measured inside the step (bench.py): 1.008 -> 0.991 ms provided nothing DRAM-heavy runs beside this chain (with
        // the gradient zero-fill started at the top of the step the lanes made it 12 us longer).  Also measured: the
        // caller's lane taking the second level's NMS as well (it finishes its top-k later but has less NMS work):
        // MdProposal 147 -> 158 us, step +13 us -- two issue-bound mask kernels side by side only slow each other.
        int maxN1 = 0;
        for (int l = 1; l < L; l++) maxN1 = max(maxN1, lv.A[l] * lv.H[l] * lv.W[l]);
        e = cudaEventRecord(side->fork, s);
        if (e == cudaSuccess) e = cudaStreamWaitEvent(side->stream, side->fork, 0);
        // the NMS mask + sweep per lane too (MD_PROP_NMS_LANES=0: one NMS launch for all levels after the join)
        static const bool nms_lanes = !(getenv("MD_PROP_NMS_LANES") && getenv("MD_PROP_NMS_LANES")[0] == '0');
        int Kmax1 = 0;
        for (int l = 1; l < L; l++) Kmax1 = max(Kmax1, pl.K[l]);
        NmsSegs sg1 = sg, sg0 = sg;
        sg1.l0 = 1; sg1.nl = L - 1;
        sg0.l0 = 0; sg0.nl = 1;
        if (e == cudaSuccess) e = launch_select_sorted(PropSrc{ pl, 1, L - 1 }, PropSink{ pl, w.boxes, w.scores, topk_idx }, B * (L - 1), maxN1, side->stream);
        if (e == cudaSuccess && nms_lanes) e = run_nms(sg1, B * (L - 1), Kmax1, cfg + 11, w.mask, w.keep_pos, nms_pre, keep, nms_pre, w.count, side->stream);
        if (e == cudaSuccess) e = cudaEventRecord(side->join, side->stream);
        if (e == cudaSuccess) e = launch_select_sorted(PropSrc{ pl, 0, 1 }, PropSink{ pl, w.boxes, w.scores, topk_idx }, B, lv.A[0] * lv.H[0] * lv.W[0], s);
        if (e == cudaSuccess && nms_lanes) e = run_nms(sg0, B, pl.K[0], cfg + 11, w.mask, w.keep_pos, nms_pre, keep, nms_pre, w.count, s);
        if (e == cudaSuccess) e = cudaStreamWaitEvent(s, side->join, 0);
        if (e == cudaSuccess && !nms_lanes) e = run_nms(sg, nseg, Kmax, cfg + 11, w.mask, w.keep_pos, nms_pre, keep, nms_pre, w.count, s);
    } else {
        e = launch_select_sorted(PropSrc{ pl, 0, L }, PropSink{ pl, w.boxes, w.scores, topk_idx }, nseg, maxN, s);
        if (e == cudaSuccess) e = run_nms(sg, nseg, Kmax, cfg + 11, w.mask, w.keep_pos, nms_pre, keep, nms_pre, w.count, s);
    }
    if (e != cudaSuccess) return e;
    const size_t smem = (size_t)L * nms_pre * sizeof(uint32_t);
    if (smem > 48 * 1024) {      // per device, constant value: set on every launch that needs it
        e = cudaFuncSetAttribute(merge_levels_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
        if (e != cudaSuccess) return e;
    }
    merge_levels_kernel<<<dim3(B, kMergeSplit), kMergeThreads, smem, s>>>(
        L, nms_pre, max_num, w.boxes, w.kept_keys, w.keep_pos, w.count, w.scores, keep, sg, props, pmask);
    return cudaGetLastError();
}

}  // namespace md
