// assign.cu -- a7/a8: MaxIoU assignment + deterministic (Philox) sampling.  sm_100a.
//
// Algorithm family of the reference: create_target_np, pointpillars/src/core/target_assigner.py:84-134
// (per-anchor argmax :90-93, per-gt max with every tie force-matched :94-103, thresholds :104-108) with
// the IoU of pointpillars/src/core/box_np_ops.py:639-679 (iou_jit, eps=1).  The reference materialises the
// (N x G) overlap matrix on the host per sample; here it is NEVER materialised: every thread streams the
// (<=G) valid gts its warp can touch from shared memory and builds the per-gt maxima and its own (max IoU,
// arg-max) in one pass (kernel 1, 8 bytes per box kept); kernel 2 labels from those 8 bytes and walks the gts
// again only in the few warps that hold a box good enough to be some gt's best (the force-match rule).
// Sampling replaces npr.choice (:116-128) by "k smallest Philox keys" (oracle/CONVENTIONS.md #13) on
// the cluster radix-select of select.cuh.
#include "kernels.h"
#include "select.cuh"
#include "launch.cuh"

namespace md {

constexpr int kAsThreads = 256;
constexpr int kAsMaxG = 1024;

struct AsIn {
    const float *boxes; int ld; int64_t image_stride;   // rows of ld floats; stride 0 = shared by batch
    const uint8_t *box_valid; int64_t valid_stride;      // may be null
    const float *gts; const uint8_t *gt_valid; int G;
    int N;
    const float *cfg;
};

struct GtS { float4 box; float area; int j; };

MD_DEVINL float4 load_box(const float *p, int ld)
{
    if (ld == 4) return __ldg(reinterpret_cast<const float4 *>(p));
    return make_float4(__ldg(p), __ldg(p + 1), __ldg(p + 2), __ldg(p + 3));
}

// compact the valid gts of image b into shared memory (original order kept): ordered block-wide compaction with
// warp ballots (a serial walk by thread 0 cost ~3 us at the head of every CTA)
MD_DEVINL int stage_gts(const AsIn &in, int b, float off, GtS *sg, int *s_count)
{
    __shared__ int warp_tot[kAsThreads / 32];
    __shared__ int base_s;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (threadIdx.x == 0) base_s = 0;
    __syncthreads();
    for (int j0 = 0; j0 < in.G; j0 += kAsThreads) {
        const int j = j0 + threadIdx.x;
        const bool v = j < in.G && (!in.gt_valid || in.gt_valid[(int64_t)b * in.G + j]);
        const uint32_t m = __ballot_sync(0xffffffffu, v);
        if (lane == 0) warp_tot[warp] = __popc(m);
        __syncthreads();
        int before = base_s;
        for (int w = 0; w < warp; w++) before += warp_tot[w];
        if (v) {
            GtS g;
            g.box = __ldg(reinterpret_cast<const float4 *>(in.gts) + (int64_t)b * in.G + j);
            g.area = area_legacy(g.box, off);
            g.j = j;
            sg[before + __popc(m & ((1u << lane) - 1u))] = g;
        }
        __syncthreads();
        if (threadIdx.x == 0) {
            int t = 0;
            for (int w = 0; w < kAsThreads / 32; w++) t += warp_tot[w];
            base_s += t;
        }
        __syncthreads();
    }
    if (threadIdx.x == 0) *s_count = base_s;
    __syncthreads();
    return *s_count;
}

// Warp-level cull: the 32 consecutive boxes of a warp are (for anchors) ~11 neighbouring grid cells, so almost
// no gt can overlap any of them.  Lane k tests gt (base + k) against the bounding box of the warp's boxes and the
// ballot is the ordered hit list (gt order decides arg-max ties and which force-match wins).  Skipping the
// others is exact: their IoU with every box of the warp is 0 (iw or ih <= 0); a 1-pixel margin keeps the test
// conservative under the fp32 rounding of the real iw / ih expressions.
struct WarpBox { float x1, y1, x2, y2; };
MD_DEVINL WarpBox warp_bbox(bool have, float4 a)
{
    WarpBox w = { have ? a.x : 3.0e38f, have ? a.y : 3.0e38f, have ? a.z : -3.0e38f, have ? a.w : -3.0e38f };
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        w.x1 = fminf(w.x1, __shfl_xor_sync(0xffffffffu, w.x1, o)); w.y1 = fminf(w.y1, __shfl_xor_sync(0xffffffffu, w.y1, o));
        w.x2 = fmaxf(w.x2, __shfl_xor_sync(0xffffffffu, w.x2, o)); w.y2 = fmaxf(w.y2, __shfl_xor_sync(0xffffffffu, w.y2, o));
    }
    return w;
}
MD_DEVINL uint32_t warp_hits(const GtS *sg, int ng, int base, const WarpBox &w, float off)
{
    const int k = base + (threadIdx.x & 31);
    bool hit = false;
    if (k < ng) {
        const float4 g = sg[k].box;
        hit = !(g.z + off < w.x1 - 1.0f || w.x2 + off < g.x - 1.0f || g.w + off < w.y1 - 1.0f || w.y2 + off < g.y - 1.0f);
    }
    return __ballot_sync(0xffffffffu, hit);
}

__global__ void __launch_bounds__(kAsThreads)
assign_gtmax_kernel(const AsIn in, uint32_t *__restrict__ gmax /* (B,G) float bits, zeroed */,
                    float2 *__restrict__ best /* (B,N): every box's (max IoU, arg-max gt as int bits) for the label pass */,
                    int32_t *__restrict__ head_assigned /* nullable: RCNN flavour, the gts-as-proposals head of `assigned` */,
                    int64_t head_stride, int32_t *__restrict__ head_cand)
{
    pdl_entry();
    extern __shared__ unsigned char smem_raw[];
    GtS *sg = reinterpret_cast<GtS *>(smem_raw);
    uint32_t *smax = reinterpret_cast<uint32_t *>(sg + in.G);
    __shared__ int s_count;
    const int b = blockIdx.y;
    const float off = __ldg(in.cfg + 3);
    for (int j = threadIdx.x; j < in.G; j += kAsThreads) smax[j] = 0u;
    if (head_assigned && blockIdx.x == 0) {
        // gts-as-proposals head of the RCNN candidate list: a valid gt is assigned to itself and is a positive candidate
        // of its own image (it used to be a kernel of its own)
        int nv = 0;
        for (int j = threadIdx.x; j < in.G; j += kAsThreads) {
            const bool v = !in.gt_valid || in.gt_valid[(int64_t)b * in.G + j];
            head_assigned[(int64_t)b * head_stride + j] = v ? j + 1 : -1;
            nv += v;
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) nv += __shfl_xor_sync(0xffffffffu, nv, o);
        if ((threadIdx.x & 31) == 0 && nv) atomicAdd(head_cand + b * 2, nv);
    }
    const int ng = stage_gts(in, b, off, sg, &s_count);
    const float *boxes = in.boxes + (int64_t)b * in.image_stride;
    for (int n0 = blockIdx.x * kAsThreads; n0 < in.N; n0 += gridDim.x * kAsThreads) {
        const int n = n0 + threadIdx.x;
        const bool have = n < in.N && !(in.box_valid && !in.box_valid[(int64_t)b * in.valid_stride + n]);
        const float4 a = have ? load_box(boxes + (int64_t)n * in.ld, in.ld) : make_float4(0, 0, 0, 0);
        const WarpBox wb = warp_bbox(have, a);
        float m = 0.0f;
        int am = 0;
        for (int base = 0; base < ng; base += 32) {
            uint32_t hits = warp_hits(sg, ng, base, wb, off);
            while (hits) {                                       // ascending gt order: the first of equal maxima wins
                const int k = base + __ffs(hits) - 1;
                hits &= hits - 1;
                const float o = have ? iou_legacy(a, sg[k].box, sg[k].area, off) : 0.0f;
                if (o > m) { m = o; am = sg[k].j; }
                if (o > 0.0f) {
                    const uint32_t ob = __float_as_uint(o);      // o > 0: uint order == float order
                    if (ob > smax[k]) atomicMax(&smax[k], ob);
                }
            }
        }
        if (n < in.N) best[(int64_t)b * in.N + n] = make_float2(m, __int_as_float(am));
    }
    __syncthreads();
    for (int k = threadIdx.x; k < ng; k += kAsThreads)
        if (smax[k]) atomicMax(&gmax[(int64_t)b * in.G + sg[k].j], smax[k]);
}

__global__ void __launch_bounds__(kAsThreads)
assign_label_kernel(const AsIn in, const uint32_t *__restrict__ gmax, const float2 *__restrict__ best,
                    int32_t *__restrict__ assigned, int64_t assigned_stride, int assigned_offset,
                    int32_t *__restrict__ cand_count /* (B,2) pos|neg */)
{
    pdl_entry();
    extern __shared__ unsigned char smem_raw[];
    GtS *sg = reinterpret_cast<GtS *>(smem_raw);
    float *smax = reinterpret_cast<float *>(sg + in.G);
    __shared__ int s_count;
    const int b = blockIdx.y;
    const float pos_thr = __ldg(in.cfg + 0), neg_thr = __ldg(in.cfg + 1), min_pos = __ldg(in.cfg + 2);
    const float off = __ldg(in.cfg + 3);
    const int mode = (int)__ldg(in.cfg + 4);
    const int ng = stage_gts(in, b, off, sg, &s_count);
    __shared__ uint32_t s_gmin;
    if (threadIdx.x == 0) s_gmin = 0x7F800000u;          // +inf
    __syncthreads();
    for (int k = threadIdx.x; k < ng; k += kAsThreads) {
        const float gm = __uint_as_float(gmax[(int64_t)b * in.G + sg[k].j]);
        smax[k] = gm;
        // smallest per-gt maximum that can force a match: a box whose own best IoU is below it cannot be any gt's best
        if (gm > 0.0f && (mode != 0 || gm >= min_pos)) atomicMin(&s_gmin, __float_as_uint(gm));   // gm > 0: uint order == float order
    }
    __syncthreads();
    const float gmin = __uint_as_float(s_gmin);
    const float *boxes = in.boxes + (int64_t)b * in.image_stride;
    int32_t *out = assigned + (int64_t)b * assigned_stride + assigned_offset;
    int npos = 0, nneg = 0;                              // per-thread candidate counts (samplers need the totals)
    for (int n0 = blockIdx.x * kAsThreads; n0 < in.N; n0 += gridDim.x * kAsThreads) {
        const int n = n0 + threadIdx.x;
        const bool have = n < in.N && (!in.box_valid || in.box_valid[(int64_t)b * in.valid_stride + n]);
        // (max IoU, arg-max) come from the first pass; the IoUs are only re-evaluated -- to find WHICH gt forces the
        // match -- by the few warps that hold a box that can be some gt's best (12 bytes per box instead of a second
        // walk over the gts: the RPN label pass went from 34 to ~8 us)
        float m = 0.0f;
        int am = 0, force = 0;
        if (n < in.N) { const float2 bm = __ldg(best + (int64_t)b * in.N + n); m = bm.x; am = __float_as_int(bm.y); }
        if (__any_sync(0xffffffffu, have && m >= gmin)) {
            const float4 a = have ? load_box(boxes + (int64_t)n * in.ld, in.ld) : make_float4(0, 0, 0, 0);
            const WarpBox wb = warp_bbox(have, a);
            for (int base = 0; base < ng; base += 32) {
                uint32_t hits = warp_hits(sg, ng, base, wb, off);
                while (hits) {                               // ascending gt order, uniform across the warp
                    const int k = base + __ffs(hits) - 1;
                    hits &= hits - 1;
                    const float o = iou_legacy(a, sg[k].box, sg[k].area, off);
                    const float gm = smax[k];
                    if (o == gm && gm > 0.0f && (mode != 0 || gm >= min_pos)) force = sg[k].j + 1;
                }
            }
        }
        if (n >= in.N) continue;
        int32_t as = -1;
        if (have) {
            if (mode == 0) {
                if (m >= 0.0f && m < neg_thr) as = 0;
                if (m >= pos_thr) as = am + 1;
                if (force) as = force;
            } else {
                if (force || m >= pos_thr) as = am + 1;
                else if (m < neg_thr) as = 0;
            }
        }
        out[n] = as;
        npos += as > 0; nneg += as == 0;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) { npos += __shfl_xor_sync(0xffffffffu, npos, o); nneg += __shfl_xor_sync(0xffffffffu, nneg, o); }
    if ((threadIdx.x & 31) == 0) {
        if (npos) atomicAdd(cand_count + b * 2, npos);
        if (nneg) atomicAdd(cand_count + b * 2 + 1, nneg);
    }
}

// ---- fast path of the samplers ------------------------------------------------------------------------------
// "k smallest (Philox key, index)" over ~N candidates does not need a radix select over all of them: with the
// candidate totals known (counted by the label kernel), ONE fully parallel pass keeps the candidates whose key is
// below a threshold chosen so that ~kListCap/2 survive, and the cluster select then runs on that short list.
// Exact whenever the list holds at least min(k, #candidates) entries and did not overflow (then the k smallest
// overall are all in it); otherwise the segment is flagged and the full-scan select below handles it.
constexpr int kListCap = 8192;
struct SampleLists { unsigned long long *items; int32_t *count; const int32_t *cand_count; };   // per (image, kind)

// Keys above the threshold are dropped.  Aim for ~2 * want survivors (want is > 8 sigma below that mean, and a list that
// still ends up short is detected and redone by the full scan): the list then fits the select's direct path (<= 1024
// entries: one sort, no radix passes) instead of carrying up to kListCap entries through four passes.
MD_DEVINL uint32_t sample_threshold(int ncand, int want)
{
    int target = 2 * want;
    target = target < 256 ? 256 : (target > kListCap / 2 ? kListCap / 2 : target);
    if (ncand <= target) return 0xFFFFFFFFu;
    return (uint32_t)((((unsigned long long)target) << 32) / (unsigned long long)ncand);
}
MD_DEVINL bool list_is_exact(const SampleLists &L, int seg, int want)
{
    const int cnt = L.count[seg], nc = L.cand_count[seg];
    return cnt <= kListCap && cnt >= min(want, nc);
}

__global__ void __launch_bounds__(256)
sample_prefilter_kernel(const int32_t *__restrict__ assigned, int N, int Sp, int Sn, uint32_t stream_base,
                        const int32_t *__restrict__ seed, int has_step, const float *__restrict__ cfg, const SampleLists L)
{
    pdl_entry();
    const int b = blockIdx.y, lane = threadIdx.x & 31;
    if (__ldg(cfg + 14) != 0.0f) {                       // MD_AS_FORCE_FULL: mark both lists overflowed -> full-scan select
        if (blockIdx.x == 0 && threadIdx.x < 2) L.count[b * 2 + threadIdx.x] = kListCap + 1;
        return;
    }
    const uint32_t s0 = (uint32_t)seed[0], s1 = (uint32_t)seed[1], step = has_step ? (uint32_t)seed[2] : 0u;
    const uint32_t thr_pos = sample_threshold(L.cand_count[b * 2], Sp), thr_neg = sample_threshold(L.cand_count[b * 2 + 1], Sn);
    const int32_t *a = assigned + (int64_t)b * N;
    for (int n0 = blockIdx.x * 256; n0 < N; n0 += gridDim.x * 256) {
        const int n = n0 + threadIdx.x;
        int kind = -1;
        uint32_t r = 0u;
        if (n < N) {
            const int32_t v = __ldg(a + n);
            kind = v > 0 ? 0 : (v == 0 ? 1 : -1);
            if (kind >= 0) {
                r = philox_key((uint32_t)n, stream_base + kind, (uint32_t)b, s0, s1, step);
                if (r > (kind ? thr_neg : thr_pos)) kind = -1;
            }
        }
#pragma unroll
        for (int k = 0; k < 2; k++) {
            const uint32_t m = __ballot_sync(0xffffffffu, kind == k);
            if (m) {
                int pos = 0;
                const int leader = __ffs(m) - 1;
                if (lane == leader) pos = atomicAdd(L.count + b * 2 + k, __popc(m));
                pos = __shfl_sync(0xffffffffu, pos, leader) + __popc(m & ((1u << lane) - 1u));
                if (kind == k && pos < kListCap)
                    L.items[(int64_t)(b * 2 + k) * kListCap + pos] = ((unsigned long long)(~r) << 32) | (uint32_t)(~(uint32_t)n);
            }
        }
    }
}

struct ListSrc {
    SampleLists L; int Sp, Sn; int nimg;
    struct Ctx { const unsigned long long *items; int len, want; bool ok; };
    __device__ int segment_of(int i, int it) const { return it ? -1 : (i < nimg ? 2 * i + 1 : 2 * (i - nimg)); }
    __device__ Ctx prepare(int seg) const
    {
        const int want = (seg & 1) ? Sn : Sp;
        return Ctx{ L.items + (int64_t)seg * kListCap, min(L.count[seg], kListCap), want, list_is_exact(L, seg, want) };
    }
    __device__ bool active(const Ctx &c) const { return c.ok; }
    __device__ int length(const Ctx &c) const { return c.len; }
    __device__ int want(const Ctx &c) const { return c.want; }
    __device__ uint32_t index_of(const Ctx &c, int m) const { return ~(uint32_t)c.items[m]; }
    __device__ bool load(const Ctx &c, int m, uint32_t &key) const { key = (uint32_t)(c.items[m] >> 32); return true; }
};

// ---- sampling on the cluster radix-select -------------------------------------------------------------
struct SampleSrc {
    const int32_t *assigned; int N; int Sp, Sn; uint32_t stream_base; const int32_t *seed; int has_step; int nimg; SampleLists L;
    struct Ctx { const int32_t *base; int kind; uint32_t image, seed_lo, seed_hi, step; bool on; };
    // negatives (odd segments, ~all anchors are candidates) first, positives after
    __device__ int segment_of(int i, int it) const { return it ? -1 : (i < nimg ? 2 * i + 1 : 2 * (i - nimg)); }
    __device__ Ctx prepare(int seg) const
    {
        const int b = seg >> 1;
        return Ctx{ assigned + (int64_t)b * N, seg & 1, (uint32_t)b, (uint32_t)seed[0], (uint32_t)seed[1],
                    has_step ? (uint32_t)seed[2] : 0u, !list_is_exact(L, seg, (seg & 1) ? Sn : Sp) };
    }
    __device__ bool active(const Ctx &c) const { return c.on; }      // full scan only for segments the list path declined
    __device__ int length(const Ctx &) const { return N; }
    __device__ int want(const Ctx &c) const { return c.kind ? Sn : Sp; }
    __device__ uint32_t index_of(const Ctx &, int m) const { return (uint32_t)m; }
    __device__ bool load(const Ctx &c, int m, uint32_t &key) const
    {
        const int32_t a = __ldg(c.base + m);
        const bool cand = c.kind ? (a == 0) : (a > 0);
        if (!cand) return false;
        key = ~philox_key((uint32_t)m, stream_base + c.kind, c.image, c.seed_lo, c.seed_hi, c.step);
        return true;
    }
};
struct SampleSink {
    int32_t *pos_idx, *neg_idx; int Sp, Sn; int64_t pos_stride, neg_stride; int32_t *cand_count; /* (B,2) */
    __device__ void emit(int seg, int rank, unsigned long long comp) const
    {
        const int b = seg >> 1;
        const int32_t idx = (int32_t)(~(uint32_t)comp);
        if (seg & 1) neg_idx[(int64_t)b * neg_stride + rank] = idx;
        else pos_idx[(int64_t)b * pos_stride + rank] = idx;
    }
    __device__ void pad(int seg, int rank) const
    {
        const int b = seg >> 1;
        if (seg & 1) neg_idx[(int64_t)b * neg_stride + rank] = 0;
        else pos_idx[(int64_t)b * pos_stride + rank] = 0;
    }
    __device__ void finish(int, int, int) const {}
};

// ---- finalisers -------------------------------------------------------------------------------------
__global__ void rpn_finalize_kernel(const AsIn in, int Sp, int Sn, const int32_t *__restrict__ cand_count,
                                    const int32_t *__restrict__ assigned, int32_t *__restrict__ pos_idx,
                                    uint8_t *__restrict__ pos_valid, int32_t *__restrict__ neg_idx,
                                    uint8_t *__restrict__ neg_valid, int32_t *__restrict__ pos_gt,
                                    float4 *__restrict__ pos_target, int32_t *__restrict__ num_pos_out, int32_t *step)
{
    pdl_entry();
    const int b = blockIdx.x;
    if (step && b == 0 && threadIdx.x == 0) *step += 1;          // the samplers of this call are done: next call, next draw
    const int P = cand_count[b * 2], Q = cand_count[b * 2 + 1];
    const int num_total = (int)__ldg(in.cfg + 5);
    const int num_pos = min(P, Sp);
    const int nneg = max(0, min(min(Q, Sn), num_total - num_pos));
    float mean[4], stdv[4];
#pragma unroll
    for (int i = 0; i < 4; i++) { mean[i] = __ldg(in.cfg + 6 + i); stdv[i] = __ldg(in.cfg + 10 + i); }
    const float *boxes = in.boxes + (int64_t)b * in.image_stride;
    for (int i = threadIdx.x; i < Sp; i += blockDim.x) {
        const bool v = i < num_pos;
        const int64_t o = (int64_t)b * Sp + i;
        pos_valid[o] = v;
        int32_t g = 0;
        float4 t = make_float4(0, 0, 0, 0);
        if (v) {
            const int32_t n = pos_idx[o];
            g = assigned[(int64_t)b * in.N + n] - 1;
            t = encode_box(load_box(boxes + (int64_t)n * in.ld, in.ld),
                           __ldg(reinterpret_cast<const float4 *>(in.gts) + (int64_t)b * in.G + g), mean, stdv);
        } else {
            pos_idx[o] = 0;
        }
        pos_gt[o] = g;
        pos_target[o] = t;
    }
    for (int i = threadIdx.x; i < Sn; i += blockDim.x) {
        const bool v = i < nneg;
        const int64_t o = (int64_t)b * Sn + i;
        neg_valid[o] = v;
        if (!v) neg_idx[o] = 0;
    }
    if (threadIdx.x == 0) num_pos_out[b] = num_pos;
}

__global__ void rcnn_finalize_kernel(const float *__restrict__ props5, int P_, const float *__restrict__ gts,
                                     const int32_t *__restrict__ gt_labels, int G, const float *__restrict__ cfg,
                                     int Sp, int Sn, const int32_t *__restrict__ cand_count,
                                     const int32_t *__restrict__ assigned, int32_t *__restrict__ sel_idx,
                                     float *__restrict__ rois5, float4 *__restrict__ deltas,
                                     int32_t *__restrict__ labels, uint8_t *__restrict__ mask,
                                     int32_t *__restrict__ pos_gt, int32_t *__restrict__ num_pos_out, int32_t *step)
{
    pdl_entry();
    if (step && blockIdx.x == 0 && threadIdx.x == 0) *step += 1;   // the samplers of this call are done: next call, next draw
    const int b = blockIdx.x;
    const int S = Sp + Sn, N = G + P_;
    const int Pc = cand_count[b * 2], Qc = cand_count[b * 2 + 1];
    const int num_total = (int)__ldg(cfg + 5);
    const int num_pos = min(Pc, Sp);
    const int nneg = max(0, min(min(Qc, Sn), num_total - num_pos));
    float mean[4], stdv[4];
#pragma unroll
    for (int i = 0; i < 4; i++) { mean[i] = __ldg(cfg + 6 + i); stdv[i] = __ldg(cfg + 10 + i); }
    for (int i = threadIdx.x; i < S; i += blockDim.x) {
        const bool is_pos = i < Sp;
        const bool v = is_pos ? (i < num_pos) : ((i - Sp) < nneg);
        const int64_t o = (int64_t)b * S + i;
        int32_t n = v ? sel_idx[o] : 0;
        sel_idx[o] = n;
        float4 box;
        if (n < G) box = __ldg(reinterpret_cast<const float4 *>(gts) + (int64_t)b * G + n);
        else box = load_box(props5 + ((int64_t)b * P_ + (n - G)) * 5, 5);
        float4 d = make_float4(0, 0, 0, 0);
        int32_t lab = 0, g = 0;
        if (v && is_pos) {
            g = assigned[(int64_t)b * N + n] - 1;
            d = encode_box(box, __ldg(reinterpret_cast<const float4 *>(gts) + (int64_t)b * G + g), mean, stdv);
            lab = gt_labels[(int64_t)b * G + g];
        }
        float *r = rois5 + o * 5;
        r[0] = (float)b; r[1] = box.x; r[2] = box.y; r[3] = box.z; r[4] = box.w;
        deltas[o] = d;
        labels[o] = lab;
        mask[o] = v;
        if (is_pos) pos_gt[(int64_t)b * Sp + i] = g;
    }
    if (threadIdx.x == 0) num_pos_out[b] = num_pos;
}

// workspace: gmax (B,G) u32 | cand_count (B,2) i32 | list_count (B,2) i32 | lists (B,2,kListCap) u64
struct AssignWs { uint32_t *gmax; int32_t *cand, *list_count; unsigned long long *items; float2 *best; size_t total; };
static AssignWs carve_assign_ws(void *ws, int B, int G, int N)
{
    AssignWs w;
    unsigned char *p = reinterpret_cast<unsigned char *>(ws);
    size_t o = 0;
    w.gmax = reinterpret_cast<uint32_t *>(p + o); o += ((size_t)B * G * 4 + 255) & ~(size_t)255;
    w.cand = reinterpret_cast<int32_t *>(p + o); w.list_count = w.cand + B * 2; o += ((size_t)B * 4 * 4 + 255) & ~(size_t)255;
    w.items = reinterpret_cast<unsigned long long *>(p + o); o += (size_t)B * 2 * kListCap * 8;
    w.best = reinterpret_cast<float2 *>(p + o); o += ((size_t)B * N * sizeof(float2) + 255) & ~(size_t)255;   // (max IoU, arg-max) per box
    w.total = o + 256;
    return w;
}
size_t assign_workspace_bytes(int B, int G, int N) { return carve_assign_ws(nullptr, B, G, N).total; }

static cudaError_t run_assign(const AsIn &in, int B, const AssignWs &w, int32_t *assigned, int64_t assigned_stride,
                              int assigned_offset, bool zero_counts, cudaStream_t s, bool gt_head = false)
{
    if (in.G > kAsMaxG) return cudaErrorInvalidValue;
    cudaError_t e = cudaMemsetAsync(w.gmax, 0, (size_t)B * in.G * 4, s);
    if (e != cudaSuccess) return e;
    if (zero_counts) {
        e = cudaMemsetAsync(w.cand, 0, (size_t)B * 4 * 4, s);
        if (e != cudaSuccess) return e;
    }
    int gx = (in.N + kAsThreads - 1) / kAsThreads;
    const int cap = (148 * 8 + B - 1) / B;
    if (gx > cap) gx = cap;
    if (gx < 1) gx = 1;
    const size_t smem = (size_t)in.G * (sizeof(GtS) + 4) + 16;
    e = launch_pdl(assign_gtmax_kernel, dim3(gx, B), dim3(kAsThreads), smem, s, in, w.gmax, w.best, gt_head ? assigned : (int32_t *)nullptr, assigned_stride, w.cand);
    if (e != cudaSuccess) return e;
    return launch_pdl(assign_label_kernel, dim3(gx, B), dim3(kAsThreads), smem, s, in, (const uint32_t *)w.gmax, (const float2 *)w.best, assigned, assigned_stride, assigned_offset, w.cand);
}

// list fast path, then the full-scan select for the segments it declined (normally none: both kernels of the
// second launch return at once)
static cudaError_t run_samplers(const int32_t *assigned, int B, int N, int Sp, int Sn, uint32_t stream_base,
                                const int32_t *seed, int has_step, const float *cfg, const AssignWs &w, const SampleSink &sink, cudaStream_t s)
{
    SampleLists L{ w.items, w.list_count, w.cand };
    int gx = (N + 255) / 256;
    const int cap = (148 * 8 + B - 1) / B;
    if (gx > cap) gx = cap;
    cudaError_t e = launch_pdl(sample_prefilter_kernel, dim3(gx, B), dim3(256), 0, s, assigned, N, Sp, Sn, stream_base, seed, has_step, cfg, L);
    if (e != cudaSuccess) return e;
    e = launch_select_sorted(ListSrc{ L, Sp, Sn, B }, sink, 2 * B, kListCap, s);
    if (e != cudaSuccess) return e;
    return launch_select_sorted(SampleSrc{ assigned, N, Sp, Sn, stream_base, seed, has_step, B, L }, sink, 2 * B, N, s);
}

cudaError_t launch_assign_sample_rpn(const float *boxes, int boxes_per_image, const uint8_t *box_valid,
                                     int B, int N, const float *gts, const uint8_t *gt_valid, int G,
                                     const float *cfg, const int32_t *seed, int seed_len, void *ws, int Sp, int Sn,
                                     int32_t *assigned, int32_t *pos_idx, uint8_t *pos_valid, int32_t *neg_idx,
                                     uint8_t *neg_valid, int32_t *pos_gt, float *pos_target, int32_t *num_pos,
                                     cudaStream_t s)
{
    if (B == 0) return cudaSuccess;
    if (N >= (1 << kSelMaxIndexBits) || Sp > kSelMaxK || Sn > kSelMaxK) return cudaErrorInvalidValue;
    const AssignWs w = carve_assign_ws(ws, B, G, N);
    AsIn in{ boxes, 4, boxes_per_image ? (int64_t)N * 4 : 0, box_valid, boxes_per_image ? (int64_t)N : 0, gts, gt_valid, G, N, cfg };
    cudaError_t e = run_assign(in, B, w, assigned, N, 0, true, s);
    if (e != cudaSuccess) return e;
    SampleSink sink{ pos_idx, neg_idx, Sp, Sn, Sp, Sn, w.cand };
    e = run_samplers(assigned, B, N, Sp, Sn, 0u, seed, seed_len >= 3, cfg, w, sink, s);
    if (e != cudaSuccess) return e;
    return launch_pdl(rpn_finalize_kernel, dim3(B), dim3(256), 0, s, in, Sp, Sn, w.cand, assigned, pos_idx, pos_valid, neg_idx, neg_valid,
                      pos_gt, reinterpret_cast<float4 *>(pos_target), num_pos,
                      seed_len >= 3 ? const_cast<int32_t *>(seed) + 2 : (int32_t *)nullptr);
}

cudaError_t launch_assign_sample_rcnn(const float *props5, const uint8_t *prop_mask, int B, int P,
                                      const float *gts, const int32_t *gt_labels, const uint8_t *gt_valid, int G,
                                      const float *cfg, const int32_t *seed, int seed_len, void *ws, int Sp, int Sn,
                                      float *rois5, float *deltas, int32_t *labels, uint8_t *mask,
                                      int32_t *assigned, int32_t *sel_idx, int32_t *pos_gt, int32_t *num_pos,
                                      cudaStream_t s)
{
    if (B == 0) return cudaSuccess;
    const int N = G + P, S = Sp + Sn;
    if (N >= (1 << kSelMaxIndexBits) || Sp > kSelMaxK || Sn > kSelMaxK) return cudaErrorInvalidValue;
    const AssignWs w = carve_assign_ws(ws, B, G, P);
    AsIn in{ props5, 5, (int64_t)P * 5, prop_mask, (int64_t)P, gts, gt_valid, G, P, cfg };
    cudaError_t e = cudaMemsetAsync(w.cand, 0, (size_t)B * 4 * 4, s);
    if (e != cudaSuccess) return e;
    e = run_assign(in, B, w, assigned, N, G, false, s, G > 0);
    if (e != cudaSuccess) return e;
    SampleSink sink{ sel_idx, sel_idx + Sp, Sp, Sn, S, S, w.cand };
    e = run_samplers(assigned, B, N, Sp, Sn, 2u, seed, seed_len >= 3, cfg, w, sink, s);
    if (e != cudaSuccess) return e;
    return launch_pdl(rcnn_finalize_kernel, dim3(B), dim3(256), 0, s, props5, P, gts, gt_labels, G, cfg, Sp, Sn, w.cand, assigned, sel_idx,
                      rois5, reinterpret_cast<float4 *>(deltas), labels, mask, pos_gt, num_pos,
                      seed_len >= 3 ? const_cast<int32_t *>(seed) + 2 : (int32_t *)nullptr);
}

}  // namespace md
