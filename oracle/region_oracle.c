/*
 * region_oracle.c -- CPU ORACLE for the Faster R-CNN region path.  TEST INFRASTRUCTURE ONLY.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's CPU legs may load this library; the product
 * (minddet_b200/) never does.  Every function restates, in strict fp32 (compile with
 * -ffp-contract=off, no fast-math), the algorithm recorded in SURVEY.md section 8(a) under the
 * conventions frozen in oracle/CONVENTIONS.md.
 *
 * PARITY STATUS: the reference checkout (/root/reference) contains no Faster R-CNN code, so
 *   - o_nms           is pinned (bit-exact, tests/golden) to pointpillars/src/core/nms.py:85-112
 *                     (nms_jit, mode offset=0/inclusive) and nms.py:7-41 (apply_nms, offset=1/strict);
 *                     the default mode (offset 0, strict >, union guard) is pinned bit-exact on lattice
 *                     boxes to iou_normal, centerpoint/det3d_ms/ops/test_custom_pytorch/
 *                     iou3d_nms_kernel.cu:347-358, cut out of that file and run on the host
 *                     (oracle/ref_cu_device_harness.cpp; sweep restated from :361-405, :526-536).
 *   - o_iou_pair_p1   is pinned (<=1e-6) to pointpillars/src/core/box_np_ops.py:639-679 (iou_jit, eps=1);
 *                     with offset 0 (<=2e-6) to iou_jit(eps=0) and eval_utils.py:118-165 (image_box_overlap).
 *   - o_assign mode 1 is pinned (bit-exact) to pointpillars/src/core/target_assigner.py:84-134; mode 0
 *                     reproduces the same reference-run fixtures and may differ from mode 1 only on
 *                     anchors that tie one gt's best IoU while their argmax is another gt (tested).
 *   - o_topk          values + indices on unique scores are pinned (bit-exact) to pointpillars/src/core/nms.py:66-83
 *                     (topk_); the tie order is a decision.
 *   - anchors grid order is pinned to pointpillars/src/core/box_np_ops.py:453-523.
 *   - RoIAlign 4-tap bilinear weights (interior points) are pinned (2e-5) to bilinear_interpolate_torch,
 *                     centerpoint/det3d_ms/core/utils/center_utils.py:97-131.
 *   - decode, top-k tie order, sampling, RoI level map, RoIAlign bin geometry / edge rules / backward:
 *     PARITY UNPINNED (no reference code exists; cross-checked against torchvision where conventions
 *     coincide).
 */
#include <math.h>
#include <pthread.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define O_API __attribute__((visibility("default")))

/* ------------------------------------------------------------------------------------------ */
/* exact math (CONVENTIONS #6)                                                                 */
/* ------------------------------------------------------------------------------------------ */
static inline float f_from_bits(uint32_t u) { float f; memcpy(&f, &u, 4); return f; }
static inline uint32_t bits_from_f(float f) { uint32_t u; memcpy(&u, &f, 4); return u; }

O_API float o_exp(float x)
{
    /* fused multiply-adds are single roundings of the exact product-sum (fmaf here, fma.rn.f32 on the device), so
     * the result is still bit-reproducible; round-to-nearest via the 1.5*2^23 magic constant keeps the whole
     * function on plain add/mul/fma units (no float<->int conversion instructions on the device). */
    if (x > 88.0f) x = 88.0f;
    if (x < -87.0f) x = -87.0f;
    float t = x * 1.44269504088896341f;
    float z = t + 12582912.0f;
    float kf = z - 12582912.0f;
    float r = fmaf(kf, -0.693359375f, x);
    r = fmaf(kf, 2.12194440e-4f, r);
    float p = 1.9875691500E-4f;
    p = fmaf(p, r, 1.3981999507E-3f);
    p = fmaf(p, r, 8.3334519073E-3f);
    p = fmaf(p, r, 4.1665795894E-2f);
    p = fmaf(p, r, 1.6666665459E-1f);
    p = fmaf(p, r, 5.0000001201E-1f);
    float r2 = r * r;
    p = fmaf(p, r2, r);
    p = p + 1.0f;
    int32_t k = (int32_t)bits_from_f(z) - 0x4B400000;
    float scale = f_from_bits((uint32_t)(k + 127) << 23);
    return p * scale;
}

O_API float o_sigmoid(float x)
{
    float e = o_exp(-x);
    return 1.0f / (1.0f + e);
}

O_API void o_exp_array(const float *x, float *y, int64_t n)
{
    for (int64_t i = 0; i < n; i++) y[i] = o_exp(x[i]);
}
O_API void o_sigmoid_array(const float *x, float *y, int64_t n)
{
    for (int64_t i = 0; i < n; i++) y[i] = o_sigmoid(x[i]);
}

/* monotone uint32 image of an fp32 value (CONVENTIONS #4) */
static inline uint32_t score_key(float f)
{
    uint32_t b = bits_from_f(f);
    return (b & 0x80000000u) ? ~b : (b | 0x80000000u);
}

/* ------------------------------------------------------------------------------------------ */
/* Philox-4x32-10 (CONVENTIONS #13)                                                            */
/* ------------------------------------------------------------------------------------------ */
static inline void philox_round(uint32_t c[4], uint32_t k0, uint32_t k1)
{
    uint64_t p0 = (uint64_t)0xD2511F53u * c[0];
    uint64_t p1 = (uint64_t)0xCD9E8D57u * c[2];
    uint32_t n0 = (uint32_t)(p1 >> 32) ^ c[1] ^ k0;
    uint32_t n1 = (uint32_t)p1;
    uint32_t n2 = (uint32_t)(p0 >> 32) ^ c[3] ^ k1;
    uint32_t n3 = (uint32_t)p0;
    c[0] = n0; c[1] = n1; c[2] = n2; c[3] = n3;
}
O_API uint32_t o_philox_key_step(uint32_t n, uint32_t stream, uint32_t image, uint32_t step, uint64_t seed)
{
    uint32_t c[4] = { n, stream, image, step };                 /* CONVENTIONS #24 */
    uint32_t k0 = (uint32_t)seed, k1 = (uint32_t)(seed >> 32);
    for (int r = 0; r < 10; r++) {
        philox_round(c, k0, k1);
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
    return c[0];
}
O_API uint32_t o_philox_key(uint32_t n, uint32_t stream, uint32_t image, uint64_t seed)
{
    return o_philox_key_step(n, stream, image, 0u, seed);
}

/* ------------------------------------------------------------------------------------------ */
/* u64 select + sort helpers                                                                   */
/* ------------------------------------------------------------------------------------------ */
static int cmp_u64_desc(const void *a, const void *b)
{
    uint64_t x = *(const uint64_t *)a, y = *(const uint64_t *)b;
    return (x < y) - (x > y);
}
static int cmp_u64_asc(const void *a, const void *b)
{
    uint64_t x = *(const uint64_t *)a, y = *(const uint64_t *)b;
    return (x > y) - (x < y);
}
/* rearrange v[0..n) so that the k largest values occupy v[0..k) (unordered). keys are unique. */
static void select_largest(uint64_t *v, int64_t n, int64_t k)
{
    int64_t lo = 0, hi = n - 1;
    if (k <= 0 || k >= n) return;
    while (lo < hi) {
        uint64_t a = v[lo], b = v[lo + (hi - lo) / 2], c = v[hi];
        uint64_t piv = (a > b) ? ((b > c) ? b : (a > c ? c : a)) : ((a > c) ? a : (b > c ? c : b));
        int64_t i = lo, j = hi;
        while (i <= j) {
            while (v[i] > piv) i++;
            while (v[j] < piv) j--;
            if (i <= j) { uint64_t t = v[i]; v[i] = v[j]; v[j] = t; i++; j--; }
        }
        if (k - 1 <= j) hi = j;
        else if (k - 1 >= i) lo = i;
        else break;
    }
}

/* ------------------------------------------------------------------------------------------ */
/* a1: anchors (CONVENTIONS #7). base anchors come from the host (float64 -> round -> fp32).  */
/* ------------------------------------------------------------------------------------------ */
O_API void o_anchor_grid(const float *base, int A, int H, int W, float stride, float *out)
{
    for (int h = 0; h < H; h++)
        for (int w = 0; w < W; w++) {
            float sx = (float)w * stride, sy = (float)h * stride;
            for (int a = 0; a < A; a++) {
                float *o = out + (((int64_t)h * W + w) * A + a) * 4;
                o[0] = base[a * 4 + 0] + sx;
                o[1] = base[a * 4 + 1] + sy;
                o[2] = base[a * 4 + 2] + sx;
                o[3] = base[a * 4 + 3] + sy;
            }
        }
}

/* ------------------------------------------------------------------------------------------ */
/* a2: delta decode + clip (CONVENTIONS #8)                                                    */
/* ------------------------------------------------------------------------------------------ */
static inline void decode_one(const float *a, const float *d, const float *means, const float *stds,
                              float max_ratio, float img_h, float img_w, float *o)
{
    float dx = d[0] * stds[0]; dx = dx + means[0];
    float dy = d[1] * stds[1]; dy = dy + means[1];
    float dw = d[2] * stds[2]; dw = dw + means[2];
    float dh = d[3] * stds[3]; dh = dh + means[3];
    dw = fminf(fmaxf(dw, -max_ratio), max_ratio);
    dh = fminf(fmaxf(dh, -max_ratio), max_ratio);
    float pw = a[2] - a[0]; pw = pw + 1.0f;
    float ph = a[3] - a[1]; ph = ph + 1.0f;
    float px = a[0] + a[2]; px = px * 0.5f;
    float py = a[1] + a[3]; py = py * 0.5f;
    float gw = pw * o_exp(dw);
    float gh = ph * o_exp(dh);
    float gx = pw * dx; gx = px + gx;
    float gy = ph * dy; gy = py + gy;
    float hw = gw * 0.5f, hh = gh * 0.5f;
    float x1 = gx - hw; x1 = x1 + 0.5f;
    float y1 = gy - hh; y1 = y1 + 0.5f;
    float x2 = gx + hw; x2 = x2 - 0.5f;
    float y2 = gy + hh; y2 = y2 - 0.5f;
    float mw = img_w - 1.0f, mh = img_h - 1.0f;
    o[0] = fminf(fmaxf(x1, 0.0f), mw);
    o[1] = fminf(fmaxf(y1, 0.0f), mh);
    o[2] = fminf(fmaxf(x2, 0.0f), mw);
    o[3] = fminf(fmaxf(y2, 0.0f), mh);
}

O_API void o_decode(const float *anchors, const float *deltas, int64_t K, const float *means,
                    const float *stds, float max_ratio, float img_h, float img_w, float *out)
{
    for (int64_t i = 0; i < K; i++)
        decode_one(anchors + i * 4, deltas + i * 4, means, stds, max_ratio, img_h, img_w, out + i * 4);
}

/* decode-all form straight from the NCHW head output (4A,H,W) with anchors regenerated */
O_API void o_decode_level_nchw(const float *base, int A, int H, int W, float stride,
                               const float *deltas_nchw, const float *means, const float *stds,
                               float max_ratio, float img_h, float img_w, float *out)
{
    int64_t HW = (int64_t)H * W;
    for (int h = 0; h < H; h++)
        for (int w = 0; w < W; w++)
            for (int a = 0; a < A; a++) {
                float anc[4], d[4];
                float sx = (float)w * stride, sy = (float)h * stride;
                anc[0] = base[a * 4 + 0] + sx; anc[1] = base[a * 4 + 1] + sy;
                anc[2] = base[a * 4 + 2] + sx; anc[3] = base[a * 4 + 3] + sy;
                for (int c = 0; c < 4; c++) d[c] = deltas_nchw[(int64_t)(a * 4 + c) * HW + (int64_t)h * W + w];
                decode_one(anc, d, means, stds, max_ratio, img_h, img_w,
                           out + (((int64_t)h * W + w) * A + a) * 4);
            }
}

/* delta encode (CONVENTIONS #9), FP tolerance */
O_API void o_encode(const float *props, const float *gts, int64_t K, const float *means,
                    const float *stds, float *out)
{
    for (int64_t i = 0; i < K; i++) {
        const float *p = props + i * 4, *g = gts + i * 4;
        float px = (p[0] + p[2]) * 0.5f, py = (p[1] + p[3]) * 0.5f;
        float pw = (p[2] - p[0]) + 1.0f, ph = (p[3] - p[1]) + 1.0f;
        float gx = (g[0] + g[2]) * 0.5f, gy = (g[1] + g[3]) * 0.5f;
        float gw = (g[2] - g[0]) + 1.0f, gh = (g[3] - g[1]) + 1.0f;
        float dx = (gx - px) / pw, dy = (gy - py) / ph;
        float dw = logf(gw / pw), dh = logf(gh / ph);
        out[i * 4 + 0] = (dx - means[0]) / stds[0];
        out[i * 4 + 1] = (dy - means[1]) / stds[1];
        out[i * 4 + 2] = (dw - means[2]) / stds[2];
        out[i * 4 + 3] = (dh - means[3]) / stds[3];
    }
}

/* ------------------------------------------------------------------------------------------ */
/* a3: top-k (CONVENTIONS #4)                                                                  */
/* ------------------------------------------------------------------------------------------ */
O_API void o_topk(const float *scores, int64_t N, int64_t K, float *out_val, int32_t *out_idx)
{
    uint64_t *v = (uint64_t *)malloc(sizeof(uint64_t) * (size_t)(N > 0 ? N : 1));
    for (int64_t i = 0; i < N; i++)
        v[i] = ((uint64_t)score_key(scores[i]) << 32) | (uint64_t)(0xFFFFFFFFu - (uint32_t)i);
    if (K > N) K = N;
    select_largest(v, N, K);
    qsort(v, (size_t)K, sizeof(uint64_t), cmp_u64_desc);
    for (int64_t i = 0; i < K; i++) {
        uint32_t idx = 0xFFFFFFFFu - (uint32_t)(v[i] & 0xFFFFFFFFu);
        out_idx[i] = (int32_t)idx;
        out_val[i] = scores[idx];
    }
    free(v);
}

/* RPN head scores for one (image, level): logits NCHW (A,H,W) -> activation -> flat n=(h*W+w)*A+a */
O_API void o_level_scores(const float *logits_nchw, int A, int H, int W, int apply_sigmoid, float *out)
{
    int64_t HW = (int64_t)H * W;
    for (int a = 0; a < A; a++)
        for (int64_t p = 0; p < HW; p++) {
            float x = logits_nchw[a * HW + p];
            out[p * A + a] = apply_sigmoid ? o_sigmoid(x) : x;
        }
}

/* ------------------------------------------------------------------------------------------ */
/* a4: greedy NMS on score-sorted boxes (CONVENTIONS #1-3)                                     */
/* ------------------------------------------------------------------------------------------ */
static inline float nms_iou(const float *a, const float *b, float off, float union_eps)
{
    float left = fmaxf(a[0], b[0]), right = fminf(a[2], b[2]);
    float top = fmaxf(a[1], b[1]), bottom = fminf(a[3], b[3]);
    float w = right - left; w = w + off; w = fmaxf(w, 0.0f);
    float h = bottom - top; h = h + off; h = fmaxf(h, 0.0f);
    float inter = w * h;
    float aw = a[2] - a[0]; aw = aw + off;
    float ah = a[3] - a[1]; ah = ah + off;
    float bw = b[2] - b[0]; bw = bw + off;
    float bh = b[3] - b[1]; bh = bh + off;
    float sa = aw * ah, sb = bw * bh;
    float uni = sa + sb; uni = uni - inter;
    if (union_eps > 0.0f) uni = fmaxf(uni, union_eps);
    return inter / uni;
}

/* boxes: K rows of `ld` floats (x1,y1,x2,y2,...). keep[K] = 1 kept / 0 suppressed. returns #kept */
O_API int o_nms(const float *boxes, int ld, int K, float thr, float off, int inclusive,
                float union_eps, uint8_t *keep)
{
    uint8_t *sup = (uint8_t *)calloc((size_t)(K > 0 ? K : 1), 1);
    int cnt = 0;
    for (int i = 0; i < K; i++) {
        keep[i] = 0;
        if (sup[i]) continue;
        keep[i] = 1; cnt++;
        const float *a = boxes + (int64_t)i * ld;
        for (int j = i + 1; j < K; j++) {
            if (sup[j]) continue;
            float v = nms_iou(a, boxes + (int64_t)j * ld, off, union_eps);
            if (inclusive ? (v >= thr) : (v > thr)) sup[j] = 1;
        }
    }
    free(sup);
    return cnt;
}

/* ------------------------------------------------------------------------------------------ */
/* a5/a6: Proposal for one image (levels -> topk -> gather -> decode -> nms -> merge)          */
/* ------------------------------------------------------------------------------------------ */
typedef struct {
    float img_h, img_w;
    float means[4], stds[4];
    float max_ratio;
    float nms_thr, nms_off, union_eps;
    int nms_inclusive;
    int apply_sigmoid;
    int nms_pre;   /* per level */
    int max_num;
} OProposalCfg;

/*
 * logits[l]: (A,H_l,W_l), deltas[l]: (4A,H_l,W_l), base[l]: (A,4).
 * Per-stage debug outputs (may be NULL), laid out level after level with K_l = min(nms_pre, N_l):
 *   st_idx (sumK) int32, st_score (sumK), st_box (sumK,4), st_keep (sumK) u8
 * Final: props (max_num,5), pmask (max_num) u8.   returns sumK
 */
O_API int o_proposal_image(int L, const float *const *logits, const float *const *deltas,
                           const float *const *base, const int *A_, const int *H_, const int *W_,
                           const float *stride_, const OProposalCfg *cfg,
                           int32_t *st_idx, float *st_score, float *st_box, uint8_t *st_keep,
                           float *props, uint8_t *pmask)
{
    int sumK = 0;
    for (int l = 0; l < L; l++) {
        int64_t N = (int64_t)A_[l] * H_[l] * W_[l];
        sumK += (int)(N < cfg->nms_pre ? N : cfg->nms_pre);
    }
    int32_t *idx = (int32_t *)malloc(sizeof(int32_t) * (size_t)sumK);
    float *sc = (float *)malloc(sizeof(float) * (size_t)sumK);
    float *bx = (float *)malloc(sizeof(float) * (size_t)sumK * 5);
    uint8_t *kp = (uint8_t *)malloc((size_t)sumK);
    int off = 0;
    for (int l = 0; l < L; l++) {
        int A = A_[l], H = H_[l], W = W_[l];
        int64_t N = (int64_t)A * H * W, HW = (int64_t)H * W;
        int K = (int)(N < cfg->nms_pre ? N : cfg->nms_pre);
        float *flat = (float *)malloc(sizeof(float) * (size_t)N);
        o_level_scores(logits[l], A, H, W, cfg->apply_sigmoid, flat);
        o_topk(flat, N, K, sc + off, idx + off);
        free(flat);
        for (int i = 0; i < K; i++) {
            int32_t n = idx[off + i];
            int a = n % A; int64_t p = n / A; int h = (int)(p / W), w = (int)(p % W);
            float anc[4], d[4];
            float sx = (float)w * stride_[l], sy = (float)h * stride_[l];
            anc[0] = base[l][a * 4 + 0] + sx; anc[1] = base[l][a * 4 + 1] + sy;
            anc[2] = base[l][a * 4 + 2] + sx; anc[3] = base[l][a * 4 + 3] + sy;
            for (int c = 0; c < 4; c++) d[c] = deltas[l][(int64_t)(a * 4 + c) * HW + p];
            float *o = bx + (int64_t)(off + i) * 5;
            decode_one(anc, d, cfg->means, cfg->stds, cfg->max_ratio, cfg->img_h, cfg->img_w, o);
            o[4] = sc[off + i];
        }
        o_nms(bx + (int64_t)off * 5, 5, K, cfg->nms_thr, cfg->nms_off, cfg->nms_inclusive,
              cfg->union_eps, kp + off);
        off += K;
    }
    /* merge (CONVENTIONS #17) */
    float *ms = (float *)malloc(sizeof(float) * (size_t)sumK);
    for (int i = 0; i < sumK; i++) ms[i] = kp[i] ? sc[i] : -65536.0f;
    int M = cfg->max_num < sumK ? cfg->max_num : sumK;
    float *tv = (float *)malloc(sizeof(float) * (size_t)(M > 0 ? M : 1));
    int32_t *ti = (int32_t *)malloc(sizeof(int32_t) * (size_t)(M > 0 ? M : 1));
    o_topk(ms, sumK, M, tv, ti);
    for (int i = 0; i < cfg->max_num; i++) {
        if (i < M) {
            memcpy(props + (int64_t)i * 5, bx + (int64_t)ti[i] * 5, 5 * sizeof(float));
            pmask[i] = kp[ti[i]];
        } else {
            memset(props + (int64_t)i * 5, 0, 5 * sizeof(float));
            pmask[i] = 0;
        }
    }
    if (st_idx) memcpy(st_idx, idx, sizeof(int32_t) * (size_t)sumK);
    if (st_score) memcpy(st_score, sc, sizeof(float) * (size_t)sumK);
    if (st_keep) memcpy(st_keep, kp, (size_t)sumK);
    if (st_box)
        for (int i = 0; i < sumK; i++) memcpy(st_box + (int64_t)i * 4, bx + (int64_t)i * 5, 16);
    free(idx); free(sc); free(bx); free(kp); free(ms); free(tv); free(ti);
    return sumK;
}

/* ------------------------------------------------------------------------------------------ */
/* a7: IoU with the legacy +off convention (CONVENTIONS #10)                                   */
/* ------------------------------------------------------------------------------------------ */
static inline float iou_pair(const float *a, const float *g, float garea, float off)
{
    float iw = fminf(a[2], g[2]) - fmaxf(a[0], g[0]); iw = iw + off;
    if (!(iw > 0.0f)) return 0.0f;
    float ih = fminf(a[3], g[3]) - fmaxf(a[1], g[1]); ih = ih + off;
    if (!(ih > 0.0f)) return 0.0f;
    float aw = a[2] - a[0]; aw = aw + off;
    float ah = a[3] - a[1]; ah = ah + off;
    float aarea = aw * ah;
    float inter = iw * ih;
    float ua = aarea + garea; ua = ua - inter;
    return inter / ua;
}
static inline float box_area(const float *g, float off)
{
    float w = g[2] - g[0]; w = w + off;
    float h = g[3] - g[1]; h = h + off;
    return w * h;
}
O_API void o_iou_matrix(const float *boxes, int64_t N, const float *gts, int G, float off, float *out)
{
    for (int j = 0; j < G; j++) {
        float ga = box_area(gts + j * 4, off);
        for (int64_t n = 0; n < N; n++) out[n * G + j] = iou_pair(boxes + n * 4, gts + j * 4, ga, off);
    }
}

/* ------------------------------------------------------------------------------------------ */
/* a8: MaxIoU assignment (CONVENTIONS #11, #12)                                                */
/* ------------------------------------------------------------------------------------------ */
O_API void o_assign(const float *boxes, const uint8_t *valid, int64_t N, const float *gts,
                    const uint8_t *gt_valid, int G, float pos_thr, float neg_thr, float min_pos_iou,
                    float off, int mode, int32_t *assigned, float *max_iou_out, int32_t *argmax_out)
{
    float *gmax = (float *)calloc((size_t)(G > 0 ? G : 1), sizeof(float));
    float *garea = (float *)calloc((size_t)(G > 0 ? G : 1), sizeof(float));
    float *amax = (float *)malloc(sizeof(float) * (size_t)(N > 0 ? N : 1));
    int32_t *aarg = (int32_t *)malloc(sizeof(int32_t) * (size_t)(N > 0 ? N : 1));
    for (int j = 0; j < G; j++) garea[j] = box_area(gts + j * 4, off);
    for (int64_t n = 0; n < N; n++) {
        float m = 0.0f; int32_t am = 0;
        int v = !valid || valid[n];
        if (v)
            for (int j = 0; j < G; j++) {
                if (gt_valid && !gt_valid[j]) continue;
                float o = iou_pair(boxes + n * 4, gts + j * 4, garea[j], off);
                if (o > m) { m = o; am = j; }
                if (o > gmax[j]) gmax[j] = o;
            }
        amax[n] = m; aarg[n] = am;
    }
    for (int64_t n = 0; n < N; n++) {
        int v = !valid || valid[n];
        int32_t as = -1;
        if (v) {
            float m = amax[n];
            if (mode == 0) {
                if (m >= 0.0f && m < neg_thr) as = 0;
                if (m >= pos_thr) as = aarg[n] + 1;
                for (int j = 0; j < G; j++) {
                    if (gt_valid && !gt_valid[j]) continue;
                    if (!(gmax[j] >= min_pos_iou) || !(gmax[j] > 0.0f)) continue;
                    float o = iou_pair(boxes + n * 4, gts + j * 4, garea[j], off);
                    if (o == gmax[j]) as = j + 1;
                }
            } else {
                int force = 0;
                for (int j = 0; j < G && !force; j++) {
                    if (gt_valid && !gt_valid[j]) continue;
                    if (!(gmax[j] > 0.0f)) continue;
                    float o = iou_pair(boxes + n * 4, gts + j * 4, garea[j], off);
                    if (o == gmax[j]) force = 1;
                }
                if (force || m >= pos_thr) as = aarg[n] + 1;
                else if (m < neg_thr) as = 0;
            }
        }
        assigned[n] = as;
    }
    if (max_iou_out) memcpy(max_iou_out, amax, sizeof(float) * (size_t)N);
    if (argmax_out) memcpy(argmax_out, aarg, sizeof(int32_t) * (size_t)N);
    free(gmax); free(garea); free(amax); free(aarg);
}

/* sampling (CONVENTIONS #13): the k candidates with smallest (rkey, n), ascending. returns #cands */
O_API int64_t o_sample_step(const int32_t *assigned, int64_t N, int want_positive, uint32_t stream,
                            uint32_t image, uint64_t seed, uint32_t step, int k_slots, int32_t *out_idx)
{
    uint64_t *v = (uint64_t *)malloc(sizeof(uint64_t) * (size_t)(N > 0 ? N : 1));
    int64_t c = 0;
    for (int64_t n = 0; n < N; n++) {
        int is = want_positive ? (assigned[n] > 0) : (assigned[n] == 0);
        if (!is) continue;
        uint32_t rk = o_philox_key_step((uint32_t)n, stream, image, step, seed);
        /* select_largest picks the largest: invert so that smallest (rkey,n) is largest */
        v[c++] = ~(((uint64_t)rk << 32) | (uint64_t)(uint32_t)n);
    }
    int64_t k = c < k_slots ? c : k_slots;
    select_largest(v, c, k);
    qsort(v, (size_t)k, sizeof(uint64_t), cmp_u64_desc);
    for (int i = 0; i < k_slots; i++) out_idx[i] = (i < k) ? (int32_t)(uint32_t)(~v[i] & 0xFFFFFFFFu) : 0;
    free(v);
    return c;
}

O_API int64_t o_sample(const int32_t *assigned, int64_t N, int want_positive, uint32_t stream,
                       uint32_t image, uint64_t seed, int k_slots, int32_t *out_idx)
{
    return o_sample_step(assigned, N, want_positive, stream, image, seed, 0u, k_slots, out_idx);
}

typedef struct {
    float pos_thr, neg_thr, min_pos_iou, iou_off;
    int mode;
    int pos_slots, neg_slots, num_total;
    float means[4], stds[4];
    uint64_t seed;
    uint32_t step;      /* CONVENTIONS #24: per-call counter, 4th Philox counter word */
    uint32_t pad_;
} OAssignCfg;

/*
 * RPN flavour (BboxAssignSample): compact outputs.
 *   assigned (N) i32; pos_idx/pos_valid (pos_slots); neg_idx/neg_valid (neg_slots);
 *   pos_gt (pos_slots) i32 (assigned-1, 0 when invalid); pos_target (pos_slots,4) encode(anchor, gt)
 * returns num_pos
 */
O_API int o_assign_sample_rpn(const float *anchors, const uint8_t *valid, int64_t N, const float *gts,
                              const uint8_t *gt_valid, int G, const OAssignCfg *cfg, uint32_t image,
                              int32_t *assigned, int32_t *pos_idx, uint8_t *pos_valid,
                              int32_t *neg_idx, uint8_t *neg_valid, int32_t *pos_gt, float *pos_target)
{
    o_assign(anchors, valid, N, gts, gt_valid, G, cfg->pos_thr, cfg->neg_thr, cfg->min_pos_iou,
             cfg->iou_off, cfg->mode, assigned, NULL, NULL);
    int64_t P = o_sample_step(assigned, N, 1, 0u, image, cfg->seed, cfg->step, cfg->pos_slots, pos_idx);
    int64_t Q = o_sample_step(assigned, N, 0, 1u, image, cfg->seed, cfg->step, cfg->neg_slots, neg_idx);
    int num_pos = (int)(P < cfg->pos_slots ? P : cfg->pos_slots);
    int64_t nneg = cfg->num_total - num_pos;
    if (nneg > Q) nneg = Q;
    if (nneg > cfg->neg_slots) nneg = cfg->neg_slots;
    if (nneg < 0) nneg = 0;
    for (int i = 0; i < cfg->pos_slots; i++) {
        pos_valid[i] = i < num_pos;
        if (!pos_valid[i]) pos_idx[i] = 0;
        int32_t g = pos_valid[i] ? assigned[pos_idx[i]] - 1 : 0;
        pos_gt[i] = g;
        if (pos_valid[i]) o_encode(anchors + (int64_t)pos_idx[i] * 4, gts + (int64_t)g * 4, 1, cfg->means, cfg->stds, pos_target + i * 4);
        else memset(pos_target + i * 4, 0, 16);
    }
    for (int i = 0; i < cfg->neg_slots; i++) {
        neg_valid[i] = i < nneg;
        if (!neg_valid[i]) neg_idx[i] = 0;
    }
    return num_pos;
}

/*
 * RCNN flavour (BboxAssignSampleForRcnn): candidates = concat(gts, proposals); gts-as-proposals are
 * pre-assigned to themselves (or -1 when invalid).  S = pos_slots + neg_slots.
 *   rois (S,4), deltas (S,4), labels (S) i32, mask (S) u8, assigned (G+P) i32, sel_idx (S) i32
 */
O_API int o_assign_sample_rcnn(const float *props, const uint8_t *prop_valid, int P_, const float *gts,
                               const int32_t *gt_labels, const uint8_t *gt_valid, int G,
                               const OAssignCfg *cfg, uint32_t image, int32_t *assigned,
                               int32_t *sel_idx, float *rois, float *deltas, int32_t *labels, uint8_t *mask)
{
    int64_t N = (int64_t)G + P_;
    float *all = (float *)malloc(sizeof(float) * (size_t)N * 4);
    memcpy(all, gts, sizeof(float) * (size_t)G * 4);
    memcpy(all + (size_t)G * 4, props, sizeof(float) * (size_t)P_ * 4);
    for (int j = 0; j < G; j++) assigned[j] = (!gt_valid || gt_valid[j]) ? j + 1 : -1;
    o_assign(props, prop_valid, P_, gts, gt_valid, G, cfg->pos_thr, cfg->neg_thr, cfg->min_pos_iou,
             cfg->iou_off, cfg->mode, assigned + G, NULL, NULL);
    int S = cfg->pos_slots + cfg->neg_slots;
    int64_t Pc = o_sample_step(assigned, N, 1, 2u, image, cfg->seed, cfg->step, cfg->pos_slots, sel_idx);
    int64_t Qc = o_sample_step(assigned, N, 0, 3u, image, cfg->seed, cfg->step, cfg->neg_slots, sel_idx + cfg->pos_slots);
    int num_pos = (int)(Pc < cfg->pos_slots ? Pc : cfg->pos_slots);
    int64_t nneg = cfg->num_total - num_pos;
    if (nneg > Qc) nneg = Qc;
    if (nneg > cfg->neg_slots) nneg = cfg->neg_slots;
    if (nneg < 0) nneg = 0;
    for (int i = 0; i < S; i++) {
        int is_pos = i < cfg->pos_slots;
        int v = is_pos ? (i < num_pos) : ((i - cfg->pos_slots) < nneg);
        if (!v) sel_idx[i] = 0;
        mask[i] = (uint8_t)v;
        memcpy(rois + i * 4, all + (int64_t)sel_idx[i] * 4, 16);
        memset(deltas + i * 4, 0, 16);
        labels[i] = 0;
        if (v && is_pos) {
            int32_t g = assigned[sel_idx[i]] - 1;
            o_encode(rois + i * 4, gts + (int64_t)g * 4, 1, cfg->means, cfg->stds, deltas + i * 4);
            labels[i] = gt_labels[g];
        }
    }
    free(all);
    return num_pos;
}

/* ------------------------------------------------------------------------------------------ */
/* a9: RoI -> pyramid level (CONVENTIONS #14)                                                  */
/* ------------------------------------------------------------------------------------------ */
static inline int roi_level(const float *r, float finest, int num_levels)
{
    float w = r[2] - r[0]; w = w + 1.0f;
    float h = r[3] - r[1]; h = h + 1.0f;
    float s = sqrtf(w * h);
    float t = s / finest; t = t + 1e-6f;
    int l = (t >= 2.0f) + (t >= 4.0f) + (t >= 8.0f);
    for (int k = 4; k < num_levels; k++) l += (t >= (float)(1 << k));
    if (l > num_levels - 1) l = num_levels - 1;
    return l;
}
O_API void o_roi_levels(const float *rois5, int64_t R, float finest, int num_levels, int32_t *out)
{
    for (int64_t i = 0; i < R; i++) out[i] = roi_level(rois5 + i * 5 + 1, finest, num_levels);
}

/* ------------------------------------------------------------------------------------------ */
/* a10/a11: RoIAlign forward / backward (CONVENTIONS #15, #16)                                 */
/* ------------------------------------------------------------------------------------------ */
typedef struct { int yl, xl, yh, xh; float w1, w2, w3, w4; int ok; } Tap;

static inline Tap make_tap(float y, float x, int H, int W)
{
    Tap t; memset(&t, 0, sizeof(t));
    if (y < -1.0f || y > (float)H || x < -1.0f || x > (float)W) return t;
    t.ok = 1;
    if (y <= 0.0f) y = 0.0f;
    if (x <= 0.0f) x = 0.0f;
    int yl = (int)y, xl = (int)x, yh, xh;
    if (yl >= H - 1) { yh = yl = H - 1; y = (float)yl; } else yh = yl + 1;
    if (xl >= W - 1) { xh = xl = W - 1; x = (float)xl; } else xh = xl + 1;
    float ly = y - (float)yl, lx = x - (float)xl;
    float hy = 1.0f - ly, hx = 1.0f - lx;
    t.yl = yl; t.xl = xl; t.yh = yh; t.xh = xh;
    t.w1 = hy * hx; t.w2 = hy * lx; t.w3 = ly * hx; t.w4 = ly * lx;
    return t;
}

static inline void roi_geometry(const float *r, float scale, float end_mode, int P,
                                float *sw, float *sh, float *bw, float *bh)
{
    float start_w = r[0] * scale, start_h = r[1] * scale;
    float end_w = r[2] + end_mode; end_w = end_w * scale;
    float end_h = r[3] + end_mode; end_h = end_h * scale;
    float rw = fmaxf(end_w - start_w, 1.0f), rh = fmaxf(end_h - start_h, 1.0f);
    *sw = start_w; *sh = start_h; *bw = rw / (float)P; *bh = rh / (float)P;
}
static inline float sample_coord(float start, float bin, int p, int i, int S)
{
    float base = (float)p * bin; base = start + base;
    float o = ((float)i + 0.5f) * bin; o = o / (float)S;
    return base + o;
}

/*
 * feats[l]: (B,C,H_l,W_l) NCHW; rois5 (R,5) = [batch,x1,y1,x2,y2]; lvl (R) (pass NULL to map here)
 * out (R,C,P,P)
 */
O_API void o_roialign_fwd(int L, const float *const *feats, const int *H_, const int *W_,
                          const float *scale_, int B, int C, const float *rois5, int64_t R,
                          const int32_t *lvl_in, float finest, int P, int S, float end_mode, float *out)
{
    Tap *taps = (Tap *)malloc(sizeof(Tap) * (size_t)P * P * S * S);
    for (int64_t r = 0; r < R; r++) {
        const float *roi = rois5 + r * 5;
        if (!(roi[0] >= 0.0f && roi[0] < (float)B)) {   /* CONVENTIONS #23: batch index out of range (or NaN) -> zeros */
            memset(out + (int64_t)r * C * P * P, 0, sizeof(float) * (size_t)C * P * P);
            continue;
        }
        int b = (int)roi[0];
        int l = lvl_in ? lvl_in[r] : roi_level(roi + 1, finest, L);
        int H = H_[l], W = W_[l];
        float sw, sh, bw, bh;
        roi_geometry(roi + 1, scale_[l], end_mode, P, &sw, &sh, &bw, &bh);
        for (int ph = 0; ph < P; ph++)
            for (int pw = 0; pw < P; pw++)
                for (int iy = 0; iy < S; iy++)
                    for (int ix = 0; ix < S; ix++)
                        taps[((ph * P + pw) * S + iy) * S + ix] =
                            make_tap(sample_coord(sh, bh, ph, iy, S), sample_coord(sw, bw, pw, ix, S), H, W);
        float cnt = (float)(S * S);
        for (int c = 0; c < C; c++) {
            const float *f = feats[l] + ((int64_t)b * C + c) * H * W;
            for (int bin = 0; bin < P * P; bin++) {
                float sum = 0.0f;
                for (int s = 0; s < S * S; s++) {
                    Tap *t = &taps[bin * S * S + s];
                    float val = 0.0f;
                    if (t->ok) {
                        float a1 = t->w1 * f[t->yl * W + t->xl];
                        float a2 = t->w2 * f[t->yl * W + t->xh];
                        float a3 = t->w3 * f[t->yh * W + t->xl];
                        float a4 = t->w4 * f[t->yh * W + t->xh];
                        val = a1 + a2; val = val + a3; val = val + a4;
                    }
                    sum = sum + val;
                }
                out[((int64_t)r * C + c) * P * P + bin] = sum / cnt;
            }
        }
    }
    free(taps);
}

/* dfeats[l] must be zero-initialised by the caller */
O_API void o_roialign_bwd(int L, float *const *dfeats, const int *H_, const int *W_,
                          const float *scale_, int B, int C, const float *rois5, int64_t R,
                          const int32_t *lvl_in, float finest, int P, int S, float end_mode,
                          const float *dout)
{
    Tap *taps = (Tap *)malloc(sizeof(Tap) * (size_t)P * P * S * S);
    for (int64_t r = 0; r < R; r++) {
        const float *roi = rois5 + r * 5;
        if (!(roi[0] >= 0.0f && roi[0] < (float)B)) continue;   /* CONVENTIONS #23: no gradient */
        int b = (int)roi[0];
        int l = lvl_in ? lvl_in[r] : roi_level(roi + 1, finest, L);
        int H = H_[l], W = W_[l];
        float sw, sh, bw, bh;
        roi_geometry(roi + 1, scale_[l], end_mode, P, &sw, &sh, &bw, &bh);
        for (int ph = 0; ph < P; ph++)
            for (int pw = 0; pw < P; pw++)
                for (int iy = 0; iy < S; iy++)
                    for (int ix = 0; ix < S; ix++)
                        taps[((ph * P + pw) * S + iy) * S + ix] =
                            make_tap(sample_coord(sh, bh, ph, iy, S), sample_coord(sw, bw, pw, ix, S), H, W);
        float cnt = (float)(S * S);
        for (int c = 0; c < C; c++) {
            float *f = dfeats[l] + ((int64_t)b * C + c) * H * W;
            for (int bin = 0; bin < P * P; bin++) {
                float g = dout[((int64_t)r * C + c) * P * P + bin] / cnt;
                for (int s = 0; s < S * S; s++) {
                    Tap *t = &taps[bin * S * S + s];
                    if (!t->ok) continue;
                    f[t->yl * W + t->xl] += g * t->w1;
                    f[t->yl * W + t->xh] += g * t->w2;
                    f[t->yh * W + t->xl] += g * t->w3;
                    f[t->yh * W + t->xh] += g * t->w4;
                }
            }
        }
    }
    free(taps);
}

/* ------------------------------------------------------------------------------------------ */
/* whole region path over a batch (CPU baseline driver; pthreads over images)                  */
/* ------------------------------------------------------------------------------------------ */
typedef struct {
    OProposalCfg prop;
    OAssignCfg rpn, rcnn;
    float finest_scale;
    int roi_P, roi_S;
    float roi_end_mode;
    int num_roi_levels;
    int do_backward;
} ORegionCfg;

typedef struct {
    int B, L; const float *const *logits; const float *const *deltas; const float *const *base;
    const int *A_, *H_, *W_; const float *stride_; const float *const *feats; int C;
    const float *gts; const int32_t *gt_labels; const uint8_t *gt_valid; int G;
    const ORegionCfg *cfg;
    float *props; uint8_t *pmask; int32_t *rpn_assigned;
    int32_t *rpn_pos_idx; uint8_t *rpn_pos_valid; int32_t *rpn_neg_idx; uint8_t *rpn_neg_valid;
    float *rois5; int32_t *roi_labels; uint8_t *roi_mask; float *roi_deltas;
    float *roi_feats; const float *dout; float *const *dfeats;
    int64_t Ntot; float scale[16];
    int next; /* atomic work counter */
} RegionCtx;

static void region_image(RegionCtx *x, int b)
{
    const ORegionCfg *cfg = x->cfg;
    int L = x->L, G = x->G, C = x->C, B = x->B;
    int64_t Ntot = x->Ntot;
    int S = cfg->rcnn.pos_slots + cfg->rcnn.neg_slots;
    int M = cfg->prop.max_num;
    const float *lg[16] = { 0 }, *dl[16] = { 0 };
    for (int l = 0; l < L; l++) {
        int64_t n = (int64_t)x->A_[l] * x->H_[l] * x->W_[l];
        lg[l] = x->logits[l] + (int64_t)b * n;
        dl[l] = x->deltas[l] + (int64_t)b * n * 4;
    }
    o_proposal_image(L, lg, dl, x->base, x->A_, x->H_, x->W_, x->stride_, &cfg->prop, NULL, NULL, NULL, NULL,
                     x->props + (int64_t)b * M * 5, x->pmask + (int64_t)b * M);
    /* RPN targets */
    float *anchors = (float *)malloc(sizeof(float) * (size_t)Ntot * 4);
    int64_t o = 0;
    for (int l = 0; l < L; l++) {
        o_anchor_grid(x->base[l], x->A_[l], x->H_[l], x->W_[l], x->stride_[l], anchors + o * 4);
        o += (int64_t)x->A_[l] * x->H_[l] * x->W_[l];
    }
    int32_t *pos_gt = (int32_t *)malloc(sizeof(int32_t) * (size_t)cfg->rpn.pos_slots);
    float *pos_t = (float *)malloc(sizeof(float) * (size_t)cfg->rpn.pos_slots * 4);
    o_assign_sample_rpn(anchors, NULL, Ntot, x->gts + (int64_t)b * G * 4, x->gt_valid + (int64_t)b * G, G,
                        &cfg->rpn, (uint32_t)b, x->rpn_assigned + (int64_t)b * Ntot,
                        x->rpn_pos_idx + (int64_t)b * cfg->rpn.pos_slots, x->rpn_pos_valid + (int64_t)b * cfg->rpn.pos_slots,
                        x->rpn_neg_idx + (int64_t)b * cfg->rpn.neg_slots, x->rpn_neg_valid + (int64_t)b * cfg->rpn.neg_slots,
                        pos_gt, pos_t);
    free(anchors); free(pos_gt); free(pos_t);
    /* RCNN targets */
    float *p4 = (float *)malloc(sizeof(float) * (size_t)M * 4);
    for (int i = 0; i < M; i++) memcpy(p4 + i * 4, x->props + ((int64_t)b * M + i) * 5, 16);
    int32_t *as2 = (int32_t *)malloc(sizeof(int32_t) * (size_t)(G + M));
    int32_t *sel = (int32_t *)malloc(sizeof(int32_t) * (size_t)S);
    float *r4 = (float *)malloc(sizeof(float) * (size_t)S * 4);
    o_assign_sample_rcnn(p4, x->pmask + (int64_t)b * M, M, x->gts + (int64_t)b * G * 4,
                         x->gt_labels + (int64_t)b * G, x->gt_valid + (int64_t)b * G, G, &cfg->rcnn,
                         (uint32_t)b, as2, sel, r4, x->roi_deltas + (int64_t)b * S * 4,
                         x->roi_labels + (int64_t)b * S, x->roi_mask + (int64_t)b * S);
    for (int i = 0; i < S; i++) {
        float *r = x->rois5 + ((int64_t)b * S + i) * 5;
        r[0] = (float)b; memcpy(r + 1, r4 + i * 4, 16);
    }
    free(p4); free(as2); free(sel); free(r4);
    /* RoIAlign */
    int64_t fsz = (int64_t)S * C * cfg->roi_P * cfg->roi_P;
    o_roialign_fwd(cfg->num_roi_levels, x->feats, x->H_, x->W_, x->scale, B, C, x->rois5 + (int64_t)b * S * 5, S,
                   NULL, cfg->finest_scale, cfg->roi_P, cfg->roi_S, cfg->roi_end_mode,
                   x->roi_feats + (int64_t)b * fsz);
    if (cfg->do_backward)
        o_roialign_bwd(cfg->num_roi_levels, x->dfeats, x->H_, x->W_, x->scale, B, C, x->rois5 + (int64_t)b * S * 5, S,
                       NULL, cfg->finest_scale, cfg->roi_P, cfg->roi_S, cfg->roi_end_mode,
                       x->dout + (int64_t)b * fsz);
}

static void *region_worker(void *arg)
{
    RegionCtx *x = (RegionCtx *)arg;
    for (;;) {
        int b = __atomic_fetch_add(&x->next, 1, __ATOMIC_RELAXED);
        if (b >= x->B) break;
        region_image(x, b);   /* images are independent: each scatters only into its own batch slice */
    }
    return NULL;
}

/*
 * Per level l: logits[l] (B,A,H,W), deltas[l] (B,4A,H,W), feats[l] (B,C,H,W) for l < num_roi_levels.
 * gts (B,G,4), gt_labels (B,G), gt_valid (B,G).
 * Outputs: props (B,max_num,5), pmask (B,max_num), rpn_assigned (B,Ntot), rois (B,S,5) [batch,x1..y2],
 *          roi_labels (B,S), roi_mask (B,S), roi_feats (B*S,C,P,P), dfeats[l] (B,C,H,W) (+= , pre-zeroed)
 * dout (B*S,C,P,P) upstream gradient for backward.  nthreads pthreads, one image per task.
 */
O_API void o_region_path_batch(int B, int L, const float *const *logits, const float *const *deltas,
                               const float *const *base, const int *A_, const int *H_, const int *W_,
                               const float *stride_, const float *const *feats, int C,
                               const float *gts, const int32_t *gt_labels, const uint8_t *gt_valid, int G,
                               const ORegionCfg *cfg, int nthreads,
                               float *props, uint8_t *pmask, int32_t *rpn_assigned,
                               int32_t *rpn_pos_idx, uint8_t *rpn_pos_valid, int32_t *rpn_neg_idx, uint8_t *rpn_neg_valid,
                               float *rois5, int32_t *roi_labels, uint8_t *roi_mask, float *roi_deltas,
                               float *roi_feats, const float *dout, float *const *dfeats)
{
    RegionCtx x;
    memset(&x, 0, sizeof(x));
    x.B = B; x.L = L; x.logits = logits; x.deltas = deltas; x.base = base; x.A_ = A_; x.H_ = H_; x.W_ = W_;
    x.stride_ = stride_; x.feats = feats; x.C = C; x.gts = gts; x.gt_labels = gt_labels; x.gt_valid = gt_valid;
    x.G = G; x.cfg = cfg; x.props = props; x.pmask = pmask; x.rpn_assigned = rpn_assigned;
    x.rpn_pos_idx = rpn_pos_idx; x.rpn_pos_valid = rpn_pos_valid; x.rpn_neg_idx = rpn_neg_idx; x.rpn_neg_valid = rpn_neg_valid;
    x.rois5 = rois5; x.roi_labels = roi_labels; x.roi_mask = roi_mask; x.roi_deltas = roi_deltas;
    x.roi_feats = roi_feats; x.dout = dout; x.dfeats = dfeats;
    for (int l = 0; l < L; l++) { x.Ntot += (int64_t)A_[l] * H_[l] * W_[l]; x.scale[l] = 1.0f / stride_[l]; }
    if (nthreads < 1) nthreads = 1;
    if (nthreads > 256) nthreads = 256;
    pthread_t th[256];
    for (int t = 1; t < nthreads; t++) pthread_create(&th[t], NULL, region_worker, &x);
    region_worker(&x);
    for (int t = 1; t < nthreads; t++) pthread_join(th[t], NULL);
}

/* ------------------------------------------------------------------------------------------ */
/* a14 ("next" row 1): YOLOv8 post-process.  No reference code exists (README.md:13 names the   */
/* model only): PARITY UNPINNED, conventions #19-#20 of oracle/CONVENTIONS.md.                 */
/* ------------------------------------------------------------------------------------------ */
/* pred: (4*reg_max + nc, A) channel-major, A = sum_l H_l*W_l anchors, level after level, row-major, x fastest.
 * dets: (A,6) = x1,y1,x2,y2,score,label.  DFL: max-subtracted softmax over reg_max bins, expectation in bins,
 * sequential fp32 accumulation; box = (cell centre -/+ distance) * stride; score = sigmoid(max logit), first max wins. */
O_API void o_yolo_decode(const float *pred, int reg_max, int nc, int L, const int *H_, const int *W_,
                         const float *stride_, float *dets)
{
    int64_t A = 0;
    for (int l = 0; l < L; l++) A += (int64_t)H_[l] * W_[l];
    int64_t a0 = 0;
    for (int l = 0; l < L; l++) {
        for (int y = 0; y < H_[l]; y++)
            for (int x = 0; x < W_[l]; x++) {
                const int64_t a = a0 + (int64_t)y * W_[l] + x;
                float d[4];
                for (int s = 0; s < 4; s++) {
                    const float *p = pred + (int64_t)(s * reg_max) * A + a;
                    float m = p[0];
                    for (int i = 1; i < reg_max; i++) m = fmaxf(m, p[(int64_t)i * A]);
                    float den = 0.0f, num = 0.0f;
                    for (int i = 0; i < reg_max; i++) {
                        float t = p[(int64_t)i * A] - m;
                        float e = o_exp(t);
                        den = den + e;
                        num = fmaf(e, (float)i, num);
                    }
                    d[s] = num / den;
                }
                const float cx = (float)x + 0.5f, cy = (float)y + 0.5f, st = stride_[l];
                float *o = dets + a * 6;
                float t;
                t = cx - d[0]; o[0] = t * st;
                t = cy - d[1]; o[1] = t * st;
                t = cx + d[2]; o[2] = t * st;
                t = cy + d[3]; o[3] = t * st;
                const float *c = pred + (int64_t)(4 * reg_max) * A + a;
                float best = c[0];
                int lab = 0;
                for (int k = 1; k < nc; k++) {
                    float v = c[(int64_t)k * A];
                    if (v > best) { best = v; lab = k; }
                }
                o[4] = o_sigmoid(best);
                o[5] = (float)lab;
            }
        a0 += (int64_t)H_[l] * W_[l];
    }
}

/* Class-aware NMS of one image: candidates score > conf_thr; the nms_pre best by (score desc, index asc); greedy
 * with nms_iou(off 0, eps 1e-8, strict >) between boxes of the same label (any label when agnostic); the first
 * max_det kept rows in order.  out (max_det,6) zero padded, keep_idx (max_det) = anchor index or -1. returns count. */
O_API int o_yolo_nms(const float *dets, int64_t A, float conf_thr, int nms_pre, float iou_thr, int agnostic,
                     int max_det, float *out, int32_t *keep_idx)
{
    uint64_t *v = (uint64_t *)malloc((size_t)(A > 0 ? A : 1) * sizeof(uint64_t));
    int64_t n = 0;
    for (int64_t a = 0; a < A; a++)
        if (dets[a * 6 + 4] > conf_thr) v[n++] = ((uint64_t)score_key(dets[a * 6 + 4]) << 32) | (uint32_t)~(uint32_t)a;
    qsort(v, (size_t)n, sizeof(uint64_t), cmp_u64_desc);
    int64_t K = n < nms_pre ? n : nms_pre;
    uint8_t *sup = (uint8_t *)calloc((size_t)(K > 0 ? K : 1), 1);
    int cnt = 0;
    for (int i = 0; i < max_det; i++) {
        for (int k = 0; k < 6; k++) out[i * 6 + k] = 0.0f;
        keep_idx[i] = -1;
    }
    for (int64_t i = 0; i < K; i++) {
        if (sup[i]) continue;
        const int64_t ai = (int64_t)(uint32_t)~(uint32_t)v[i];
        const float *bi = dets + ai * 6;
        if (cnt < max_det) {
            for (int k = 0; k < 6; k++) out[cnt * 6 + k] = bi[k];
            keep_idx[cnt] = (int32_t)ai;
        }
        cnt++;
        for (int64_t j = i + 1; j < K; j++) {
            if (sup[j]) continue;
            const float *bj = dets + (int64_t)(uint32_t)~(uint32_t)v[j] * 6;
            if (!agnostic && bi[5] != bj[5]) continue;
            if (nms_iou(bi, bj, 0.0f, 1e-8f) > iou_thr) sup[j] = 1;
        }
    }
    free(sup);
    free(v);
    return cnt < max_det ? cnt : max_det;
}

/* ------------------------------------------------------------------------------------------ */
/* a13 ("next" row 2): Mask R-CNN mask targets.  No reference code (README.md:7): PARITY        */
/* UNPINNED, convention #21.                                                                   */
/* ------------------------------------------------------------------------------------------ */
/* masks (G,H,W) uint8 {0,1}; rois (R,4) image coordinates; gt_idx (R) row of `masks` per RoI (<0: empty target);
 * out (R,M,M) uint8.  RoIAlign (aligned=False, scale 1, S x S samples averaged, same tap rules as a10) of the gt's
 * mask plane, binarised with >= 0.5. */
O_API void o_mask_targets(const uint8_t *masks, int G, int H, int W, const float *rois, const int32_t *gt_idx,
                          int64_t R, int M, int S, uint8_t *out)
{
    for (int64_t r = 0; r < R; r++) {
        uint8_t *o = out + r * M * M;
        const int g = gt_idx[r];
        if (g < 0 || g >= G) { memset(o, 0, (size_t)M * M); continue; }
        const uint8_t *plane = masks + (int64_t)g * H * W;
        float sw, sh, bw, bh;
        roi_geometry(rois + r * 4, 1.0f, 0.0f, M, &sw, &sh, &bw, &bh);
        for (int ph = 0; ph < M; ph++)
            for (int pw = 0; pw < M; pw++) {
                float sum = 0.0f;
                for (int iy = 0; iy < S; iy++)
                    for (int ix = 0; ix < S; ix++) {
                        Tap t = make_tap(sample_coord(sh, bh, ph, iy, S), sample_coord(sw, bw, pw, ix, S), H, W);
                        if (!t.ok) continue;
                        float v = t.w1 * (float)plane[(int64_t)t.yl * W + t.xl];
                        float u = t.w2 * (float)plane[(int64_t)t.yl * W + t.xh]; v = v + u;
                        u = t.w3 * (float)plane[(int64_t)t.yh * W + t.xl]; v = v + u;
                        u = t.w4 * (float)plane[(int64_t)t.yh * W + t.xh]; v = v + u;
                        sum = sum + v;
                    }
                float avg = sum / (float)(S * S);
                o[ph * M + pw] = avg >= 0.5f ? 1 : 0;
            }
    }
}

/* ------------------------------------------------------------------------------------------ */
/* "next" row 4: RCNN-head post-process (per-class decode, score threshold, class-aware NMS,    */
/* top max_det).  No reference code: PARITY UNPINNED, convention #22.                           */
/* ------------------------------------------------------------------------------------------ */
/* probs (P, nc1) softmax over nc1 = num_classes + 1 logits, background = class 0:
 *   m = max_j x_j; e_j = o_exp(x_j - m); den = sum_j e_j (sequential); p_j = e_j / den */
O_API void o_softmax_rows(const float *logits, int64_t P, int nc1, float *probs)
{
    for (int64_t r = 0; r < P; r++) {
        const float *x = logits + r * nc1;
        float *o = probs + r * nc1;
        float m = x[0];
        for (int j = 1; j < nc1; j++) m = fmaxf(m, x[j]);
        float den = 0.0f;
        for (int j = 0; j < nc1; j++) { float t = x[j] - m; o[j] = o_exp(t); den = den + o[j]; }
        for (int j = 0; j < nc1; j++) o[j] = o[j] / den;
    }
}

/* rois (P,4), roi_valid (P), logits (P,nc1), deltas (P,nc1*4).  Candidates: (r, c>=1) with roi_valid and
 * prob > score_thr; the nms_pre best by (prob desc, r*nc1+c asc); box = decode_one(roi, deltas[r][c]); greedy NMS
 * (offset 0, strict >, eps 1e-8) between equal labels; first max_det kept.  out (max_det,6) = box, prob, label;
 * keep_idx (max_det) = r*nc1+c or -1.  returns count. */
O_API int o_rcnn_post(const float *rois, const uint8_t *roi_valid, const float *logits, const float *deltas,
                      int64_t P, int nc1, const float *means, const float *stds, float max_ratio, float img_h,
                      float img_w, float score_thr, int nms_pre, float iou_thr, int max_det, float *out, int32_t *keep_idx)
{
    float *probs = (float *)malloc((size_t)P * nc1 * sizeof(float));
    o_softmax_rows(logits, P, nc1, probs);
    uint64_t *v = (uint64_t *)malloc((size_t)(P * nc1 > 0 ? P * nc1 : 1) * sizeof(uint64_t));
    int64_t n = 0;
    for (int64_t r = 0; r < P; r++) {
        if (roi_valid && !roi_valid[r]) continue;
        for (int c = 1; c < nc1; c++) {
            float p = probs[r * nc1 + c];
            if (p > score_thr) v[n++] = ((uint64_t)score_key(p) << 32) | (uint32_t)~(uint32_t)(r * nc1 + c);
        }
    }
    qsort(v, (size_t)n, sizeof(uint64_t), cmp_u64_desc);
    int64_t K = n < nms_pre ? n : nms_pre;
    float *box = (float *)malloc((size_t)(K > 0 ? K : 1) * 4 * sizeof(float));
    int32_t *lab = (int32_t *)malloc((size_t)(K > 0 ? K : 1) * sizeof(int32_t));
    for (int64_t i = 0; i < K; i++) {
        const int64_t id = (int64_t)(uint32_t)~(uint32_t)v[i];
        const int64_t r = id / nc1;
        const int c = (int)(id - r * nc1);
        decode_one(rois + r * 4, deltas + (r * nc1 + c) * 4, means, stds, max_ratio, img_h, img_w, box + i * 4);
        lab[i] = c;
    }
    uint8_t *sup = (uint8_t *)calloc((size_t)(K > 0 ? K : 1), 1);
    int cnt = 0;
    for (int i = 0; i < max_det; i++) { for (int k = 0; k < 6; k++) out[i * 6 + k] = 0.0f; keep_idx[i] = -1; }
    for (int64_t i = 0; i < K; i++) {
        if (sup[i]) continue;
        if (cnt < max_det) {
            const int64_t id = (int64_t)(uint32_t)~(uint32_t)v[i];
            for (int k = 0; k < 4; k++) out[cnt * 6 + k] = box[i * 4 + k];
            out[cnt * 6 + 4] = probs[id];
            out[cnt * 6 + 5] = (float)lab[i];
            keep_idx[cnt] = (int32_t)id;
        }
        cnt++;
        for (int64_t j = i + 1; j < K; j++) {
            if (sup[j] || lab[j] != lab[i]) continue;
            if (nms_iou(box + i * 4, box + j * 4, 0.0f, 1e-8f) > iou_thr) sup[j] = 1;
        }
    }
    free(sup); free(lab); free(box); free(v); free(probs);
    return cnt < max_det ? cnt : max_det;
}
