// kernels.h -- host-side launchers of the region-path kernels (internal; the public boundary is
// include/md_region_aot.h).  Every launcher enqueues on `stream` and never synchronises.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace md {

constexpr int kMaxLv = 8;

struct LevelSet {           // per pyramid level, host-known geometry + device pointers
    int L;
    int A[kMaxLv], H[kMaxLv], W[kMaxLv];
    const float *scores[kMaxLv];   // (B,A,H,W)
    const float *deltas[kMaxLv];   // (B,4A,H,W)
    const float *base[kMaxLv];     // (A,4)
};

// a1 / a2
cudaError_t launch_anchor_grid(const float *base, int A, int H, int W, const float *cfg, float *out, cudaStream_t s);
cudaError_t launch_decode_rows(const float *anchors, const float *deltas, int64_t K, const float *cfg, float *out, cudaStream_t s);
cudaError_t launch_decode_level(const float *deltas, const float *base, int B, int A, int H, int W,
                                const float *cfg, float *out, cudaStream_t s);

// a3: scores (B,A,H,W) head layout (A>0) or flat (B,N) (A==0 -> index == memory order)
cudaError_t launch_topk(const float *scores, int B, int A, int HW, int K, const float *cfg_sigmoid,
                        float *values, int32_t *indices, cudaStream_t s);

// a4: boxes (B,K,ld) score-sorted
size_t nms_workspace_bytes(int nseg, int Kmax);
cudaError_t launch_nms(const float *boxes, int ld, int B, int K, const float *cfg, void *ws,
                       int32_t *keep_idx, uint8_t *mask, int32_t *count, cudaStream_t s);

// a3..a6
size_t proposal_workspace_bytes(int B, int L, int nms_pre);
// A helper stream with its fork / join events (owned by the per-stream workspace): lets one entry point run two of its
// kernels side by side.  Fork-join through events, so it is also legal while the caller's stream is being captured.
struct SideLane { cudaStream_t stream; cudaEvent_t fork, join; };
cudaError_t launch_proposal(const LevelSet &lv, int B, int nms_pre, int max_num, const float *cfg, void *ws,
                            float *props, uint8_t *pmask, int32_t *topk_idx, uint8_t *keep, cudaStream_t s, const SideLane *side = nullptr);

// a7/a8.  seed = {seed_lo, seed_hi[, step]}: with seed_len >= 3 the Philox counter carries seed[2] and the op increments it
// on the device when its samplers are done (a replayed CUDA graph draws a fresh sample every replay)
size_t assign_workspace_bytes(int B, int G, int N);   // N = boxes per image
cudaError_t launch_assign_sample_rpn(const float *boxes, int boxes_per_image, const uint8_t *box_valid,
                                     int B, int N, const float *gts, const uint8_t *gt_valid, int G,
                                     const float *cfg, const int32_t *seed, int seed_len, void *ws, int Sp, int Sn,
                                     int32_t *assigned, int32_t *pos_idx, uint8_t *pos_valid, int32_t *neg_idx,
                                     uint8_t *neg_valid, int32_t *pos_gt, float *pos_target, int32_t *num_pos,
                                     cudaStream_t s);
cudaError_t launch_assign_sample_rcnn(const float *props5, const uint8_t *prop_mask, int B, int P,
                                      const float *gts, const int32_t *gt_labels, const uint8_t *gt_valid, int G,
                                      const float *cfg, const int32_t *seed, int seed_len, void *ws, int Sp, int Sn,
                                      float *rois5, float *deltas, int32_t *labels, uint8_t *mask,
                                      int32_t *assigned, int32_t *sel_idx, int32_t *pos_gt, int32_t *num_pos,
                                      cudaStream_t s);

// "next" row 3: rotated BEV overlap / IoU / NMS behind the reference's own symbols (bev.cu)
cudaError_t launch_bev_pairs(const float *boxes_a, int na, const float *boxes_b, int nb, int want_iou, float *out, cudaStream_t s);
size_t bev_nms_workspace_bytes(int n);
// mode 0 rotated '>' (NmsGpu), 1 axis-aligned '>' (NmsNormalGpu), 2 rotated '>=' + zero-area pre-removal (boxes_iou_nms_cpu)
cudaError_t launch_bev_nms(const float *boxes, int n, const float *thr, int mode, void *ws, void *keep, int keep_is_64,
                           int32_t *num_out, cudaStream_t s);

// "next" row 1: YOLOv8 post-process (yolo.cu); reg_max = 16
cudaError_t launch_yolo_decode(const float *pred, int B, int C, int A, const float *cfg, int cfg_len, float *dets, cudaStream_t s);
size_t yolo_nms_workspace_bytes(int B, int nms_pre);
cudaError_t launch_yolo_nms(const float *dets, int B, int A, const float *cfg, void *ws, int nms_pre, int max_det,
                            float *out, int32_t *keep_idx, int32_t *num_out, int32_t *cand_idx, cudaStream_t s);

// "next" row 4: RCNN-head post-process + BoundingBoxEncode (rcnn_post.cu)
size_t rcnn_post_workspace_bytes(int B, int P, int nc1, int nms_pre);
cudaError_t launch_rcnn_post(const float *rois, int roi_ld, const uint8_t *roi_valid, const float *logits, const float *deltas,
                             int B, int P, int nc1, const float *cfg, void *ws, int nms_pre, int max_det,
                             float *out, int32_t *keep_idx, int32_t *num_out, int32_t *cand_idx, cudaStream_t s);
cudaError_t launch_encode_rows(const float *props, const float *gts, int64_t K, const float *cfg, float *out, cudaStream_t s);

// a9..a11
struct FeatSet {
    int L, B, C;
    int H[kMaxLv], W[kMaxLv];
    float *feat[kMaxLv];    // (B,C,H,W)  (const for fwd, written by bwd)
};
// a13: masks (B,G,H,W) u8, rois5 (R,5) [batch,x1,y1,x2,y2], gt_idx (R) -> out (R,M,M) u8; cfg = {sample_num}
cudaError_t launch_mask_targets(const uint8_t *masks, int B, int G, int H, int W, const float *rois5, const int32_t *gt_idx,
                                int R, int M, const float *cfg, uint8_t *out, cudaStream_t s);
cudaError_t launch_roi_levels(const float *rois5, int R, const float *cfg, int32_t *out, cudaStream_t s);
size_t roialign_workspace_bytes(int R);
size_t roialign_bwd_workspace_bytes(const FeatSet &fs, int R);   // + the tile-stationary backward's plans and lists
// mode (cfg slot MD_ROI_MODE): 0 = TMA separable kernels (+ gather for RoIs they decline), 1 = gather only
// ctl: the (device, stream) control block (zero-initialised ints that persist between calls; the channel-lane kernels keep
// their work tickets in ctl[MD_CTL_ROI_FWD ..] / ctl[MD_CTL_ROI_BWD ..] and re-arm them before they exit)
enum { MD_CTL_ROI_FWD = 0, MD_CTL_ROI_BWD = 8, MD_CTL_INTS = 1024 };
cudaError_t launch_roialign_fwd(const FeatSet &fs, const float *rois5, int R, int P, const float *cfg,
                                float *out, void *ws, int *ctl, int mode, cudaStream_t s);
cudaError_t launch_roialign_bwd(const FeatSet &fs, const float *rois5, int R, int P, const float *cfg,
                                const float *dout, void *ws, int *ctl, int mode, cudaStream_t s, bool accumulate = false);

// two-op form of the tile-stationary backward (7x7, 2 samples, C % 32 == 0): `plan` = roialign_plan_bytes() bytes owned by the caller;
// prepare reads the RoIs only (fs: level shapes), planned = the backward proper (every dX byte written once)
size_t roialign_plan_bytes(const FeatSet &fs, int R);
cudaError_t launch_roialign_bwd_prepare(const FeatSet &fs, const float *rois5, int R, int P, const float *cfg, void *plan, cudaStream_t s);
cudaError_t launch_roialign_bwd_planned(const FeatSet &fs, const float *rois5, int R, int P, const float *cfg, const float *dout, void *plan,
                                        cudaStream_t s);

}  // namespace md
