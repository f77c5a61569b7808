/*
 * md_region_aot.h -- the drop-in boundary of libmdregion.so.
 *
 * Every entry point has the MindSpore `ops.Custom(func_type="aot")` signature, i.e. exactly what the
 * reference binds for its own custom ops:
 *   - GPU exemplar:  minddet/models/centerpoint/det3d_ms/ops/test_custom_pytorch/iou3d_nms_kernel.cu:445-446
 *                    (`extern "C" int NmsNormalGpu(int nparam, void** params, int* ndims,
 *                      int64_t** shapes, const char** dtypes, void* stream, void* extra)`)
 *   - CPU exemplar:  minddet/models/centerpoint/det3d_ms/ops/iou-bev-nms-org.cpp:237
 *   - Python side:   minddet/models/centerpoint/det3d_ms/ops/test_custom_pytorch/iou_gpu.py:55-60,
 *                    minddet/models/centerpoint/det3d_ms/ops/nms_cpu.py:10-27
 *
 * Contract:
 *   params[0 .. n_in-1] are inputs, params[n_in .. nparam-1] outputs, all DEVICE pointers owned by the
 *   caller (outputs pre-sized from out_shape/out_dtype); ndims[i]/shapes[i][d]/dtypes[i] describe
 *   param i; `stream` is the cudaStream_t the framework executes on -- all work is enqueued on it and
 *   never synchronised; scratch comes from a per-(device,stream) cached workspace inside the library.
 *   Float attributes travel as a trailing 1-D float32 `cfg` INPUT TENSOR that is read on the device
 *   (the reference passes its NMS threshold the same way, iou_gpu.py:57, iou3d_nms_kernel.cu:500);
 *   integer attributes travel in the output shapes.  No dynamic output shapes: fixed size + mask/count.
 * Return value: 0 success; 1 wrong nparam; 2 bad dtype/shape; 3 CUDA error; 4 unsupported size;
 *   5 the op needs to allocate (first call on this stream, or a bigger shape than any before) while the stream is being
 *   captured into a CUDA graph -- run it once eagerly (or in graph mode's warm-up) first.
 *   (the reference returns 1 on wrong nparam, iou-bev-nms-org.cpp:238, and 0 on success, :282.)
 * Workspace lifetime: scratch blocks are never freed or moved while the library is loaded (growth retires the old block
 *   but keeps it mapped), so a captured CUDA graph stays valid whatever is called afterwards.  The device used is the one
 *   current on the calling thread (MindSpore binds one device per process): make the tensors' device current.
 * There is NO CPU fallback: without a CUDA device every entry returns 3.
 *
 * cfg slot tables: the MD_CFG_* enums below.
 */
#ifndef MD_REGION_AOT_H_
#define MD_REGION_AOT_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ---- cfg tensor slots (float32, read on the device) ------------------------------------------ */
enum { /* MD_CFG_DECODE: 11 floats */
    MD_DEC_IMG_H = 0, MD_DEC_IMG_W = 1, MD_DEC_MEAN0 = 2, MD_DEC_STD0 = 6, MD_DEC_MAX_RATIO = 10,
    MD_DEC_LEN = 11, MD_DEC_STRIDE = 11 /* MdDecodeLevel only */
};
enum { /* MD_CFG_NMS: 4 floats */
    MD_NMS_THR = 0, MD_NMS_OFFSET = 1, MD_NMS_INCLUSIVE = 2, MD_NMS_UNION_EPS = 3, MD_NMS_LEN = 4
};
enum { /* MD_CFG_PROPOSAL: MD_CFG_DECODE (0..10) + these + one stride per level */
    MD_PROP_NMS_THR = 11, MD_PROP_NMS_OFFSET = 12, MD_PROP_NMS_INCLUSIVE = 13, MD_PROP_UNION_EPS = 14,
    MD_PROP_APPLY_SIGMOID = 15, MD_PROP_STRIDE0 = 16
};
enum { /* MD_CFG_ASSIGN: 16 floats */
    MD_AS_POS_THR = 0, MD_AS_NEG_THR = 1, MD_AS_MIN_POS_IOU = 2, MD_AS_IOU_OFFSET = 3, MD_AS_MODE = 4,
    MD_AS_NUM_TOTAL = 5, MD_AS_MEAN0 = 6, MD_AS_STD0 = 10,
    MD_AS_FORCE_FULL = 14, /* != 0: skip the samplers' short-list fast path (test hook; results are identical) */
    MD_AS_LEN = 16
};
enum { /* MD_CFG_ROI: 4 floats + one stride per level */
    MD_ROI_FINEST = 0, MD_ROI_SAMPLE_NUM = 1, MD_ROI_END_MODE = 2, MD_ROI_STRIDE0 = 4
};

#if defined(__GNUC__)
#define MD_API __attribute__((visibility("default")))
#else
#define MD_API
#endif

#define MD_AOT_ARGS int nparam, void **params, int *ndims, int64_t **shapes, const char **dtypes, void *stream, void *extra

/* a1  AnchorGenerator.grid_anchors
 *   in : base (A,4) f32 | cfg f32[>=1] = {stride}
 *   out: anchors (H,W,A,4) f32                       (n = (h*W+w)*A+a) */
MD_API int MdAnchorGrid(MD_AOT_ARGS);

/* a2  BoundingBoxDecode on gathered rows
 *   in : anchors (K,4) f32 | deltas (K,4) f32 | cfg f32[11] = MD_CFG_DECODE
 *   out: boxes (K,4) f32 */
MD_API int MdDecodeClip(MD_AOT_ARGS);

/* a2  decode-all form, straight from the RPN head layout, anchors regenerated in registers
 *   in : deltas (B,4A,H,W) f32 | base (A,4) f32 | cfg f32[12] = MD_CFG_DECODE + {stride}
 *   out: boxes (B,H*W*A,4) f32 */
MD_API int MdDecodeLevel(MD_AOT_ARGS);

/* a3  TopK(sorted=True) per (image, level)
 *   in : scores (B,A,H,W) f32 [head layout, index n=(h*W+w)*A+a]  or (B,N) f32 [flat]
 *        | cfg f32[1] = {apply_sigmoid}
 *   out: values (B,K) f32 | indices (B,K) int32 */
MD_API int MdTopKPerLevel(MD_AOT_ARGS);

/* a4  NMSWithMask on score-sorted boxes (reference I/O shape: NmsNormalGpu, keep[N] + count[1])
 *   in : boxes (K,C>=4) or (B,K,C>=4) f32 [x1,y1,x2,y2,...] | cfg f32[4] = MD_CFG_NMS
 *   out: keep_idx (B,K) int32 (kept positions, ascending, zero padded) | mask (B,K) uint8/bool
 *        | count (B) int32 */
MD_API int MdNms(MD_AOT_ARGS);

/* a3..a6  Proposal (top-k -> gather -> decode -> NMS per level -> cross-level merge), L levels
 *   in : scores_0..scores_{L-1} (B,A,H_l,W_l) | deltas_0..deltas_{L-1} (B,4A,H_l,W_l)
 *        | base_0..base_{L-1} (A,4) | cfg f32[16+L] = MD_CFG_PROPOSAL
 *   out: proposals (B,max_num,5) f32 | mask (B,max_num) uint8/bool
 *        | topk_idx (B,L,nms_pre) int32 (-1 padded) | keep (B,L,nms_pre) uint8/bool
 *   (nparam = 3L+1+4; nms_pre and max_num are read from the output shapes) */
MD_API int MdProposal(MD_AOT_ARGS);

/* Sampling randomness (a8, both flavours): seed = {seed_lo, seed_hi} or {seed_lo, seed_hi, step}.  The sample is the
 * k candidates with the smallest Philox-4x32-10 keys, counter (candidate, stream, image, step), key (seed_lo, seed_hi).
 * With the 3-word form the op INCREMENTS seed[2] on the device when its samplers are done, so the next call -- or the
 * next replay of a captured CUDA graph -- draws a fresh sample (the reference draws fresh npr.choice samples per call,
 * pointpillars/src/core/target_assigner.py:116-128).  Keep that tensor alive between calls (a Parameter / buffer); the
 * 2-word form means step 0 for ever (reproducible single calls).  oracle/CONVENTIONS.md #13, #24. */
/* a7/a8  BboxAssignSample (RPN flavour)
 *   in : boxes (N,4) [shared by the batch] or (B,N,4) f32 | box_valid (N)/(B,N) uint8/bool
 *        | gts (B,G,4) f32 | gt_valid (B,G) uint8/bool | cfg f32[16] = MD_CFG_ASSIGN | seed int32[2 or 3]
 *   out: assigned (B,N) int32 | pos_idx (B,Sp) int32 | pos_valid (B,Sp) uint8 | neg_idx (B,Sn) int32
 *        | neg_valid (B,Sn) uint8 | pos_gt (B,Sp) int32 | pos_target (B,Sp,4) f32 | num_pos (B) int32 */
MD_API int MdAssignSample(MD_AOT_ARGS);

/* a8  BboxAssignSampleForRcnn (gts are prepended to the proposals as candidates)
 *   in : proposals (B,P,5) f32 | prop_mask (B,P) uint8/bool | gts (B,G,4) f32 | gt_labels (B,G) int32
 *        | gt_valid (B,G) uint8/bool | cfg f32[16] = MD_CFG_ASSIGN | seed int32[2 or 3]
 *   out: rois (B,S,5) f32 [batch,x1,y1,x2,y2] | deltas (B,S,4) f32 | labels (B,S) int32
 *        | mask (B,S) uint8 | assigned (B,G+P) int32 | sel_idx (B,S) int32 | pos_gt (B,Sp) int32
 *        | num_pos (B) int32          (S = Sp+Sn; Sp is read from pos_gt's shape) */
MD_API int MdAssignSampleRcnn(MD_AOT_ARGS);

/* a9   RoI -> pyramid level
 *   in : rois (R,5) f32 | cfg f32[2] = {finest_scale, num_levels}
 *   out: levels (R) int32 */
MD_API int MdRoiLevels(MD_AOT_ARGS);

/* a9+a10  SingleRoIExtractor forward (level map + RoIAlign on the mapped level only)
 *   in : rois (R,5) f32 | feat_0..feat_{L-1} (B,C,H_l,W_l) f32 | cfg f32[4+L] = MD_CFG_ROI
 *   out: roi_feats (R,C,P,P) f32            (nparam = L+2+1)
 *   Default path: TMA-staged tiles + separable bilinear operators (rtol 1e-5 / atol 1e-6 vs the oracle). */
MD_API int MdRoiAlignFwd(MD_AOT_ARGS);

/* a11  ROIAlignGrad (bprop of MdRoiAlignFwd); every output byte is written (nothing has to be initialised by the caller).
 * 7x7 / 2 samples / C % 32 == 0: the tile-stationary kernel of roialign_tile.cu writes every dX byte exactly once (no
 * zero-fill; RoIs it declines are added by the gather kernel afterwards); otherwise zero-fill + scatter-add.
 *   in : rois (R,5) f32 | dout (R,C,P,P) f32 | cfg f32[4+L]
 *   out: dfeat_0..dfeat_{L-1} (B,C,H_l,W_l) f32     (nparam = 3+L) */
MD_API int MdRoiAlignBwd(MD_AOT_ARGS);

/* a11, accumulating form: dfeat_l += ROIAlignGrad(dout) into tensors the caller owns and has initialised (zeros for a
 * plain bprop; an existing gradient for gradient accumulation).  The zero-fill is 731 MB of pure DRAM writes per step
 * at config 2; as its own node it has no producer, so a graph executor can run it beside the latency-bound Proposal
 * chain instead of in front of the RoIAlign backward (bench.py with MD_BENCH_BWD=acc does; since the tile-stationary
 * MdRoiAlignBwd needs no zero-fill at all, the default step uses that one).
 *   in : rois (R,5) f32 | dout (R,C,P,P) f32 | cfg f32[4+L] | acc_0..acc_{L-1} (B,C,H_l,W_l) f32 (read-modify-write)
 *   out: done (1) int32 = 0      (nparam = 3+L+1) */
MD_API int MdRoiAlignBwdAcc(MD_AOT_ARGS);

/* a11, two-op form (7x7, 2 samples, C % 32 == 0 only; anything else returns 4): everything the tile-stationary backward
 * derives from the RoIs -- per-RoI separable plans, per-tile RoI lists, work items -- goes into a tensor the framework owns, so
 * that a graph can run MdRoiAlignBwdPrepare beside the forward (it needs the RoIs and the level shapes only) and the backward
 * proper starts with its tile kernel.  Plays the role of a "saved for backward" tensor; MdRoiAlignBwd is the same work in one call.
 *   MdRoiAlignPlanBytes(R, B, C, L, H[L], W[L]) -> bytes of the plan tensor (host function; MD_TILE_CHUNK must not change between
 *     the query and the calls)
 *   MdRoiAlignBwdPrepare: in rois (R,5) f32 | feat_0..feat_{L-1} (B,C,H_l,W_l) f32 (shapes only, not read) | cfg f32[4+L]
 *                         out plan int32[>= bytes / 4]                                       (nparam = 3+L)
 *   MdRoiAlignBwdPlanned: in rois (R,5) f32 | dout (R,C,7,7) f32 | cfg f32[4+L] | plan int32[..]
 *                         out dfeat_0..dfeat_{L-1} (B,C,H_l,W_l) f32, every byte written      (nparam = 4+L) */
MD_API int MdRoiAlignBwdPrepare(MD_AOT_ARGS);
MD_API int MdRoiAlignBwdPlanned(MD_AOT_ARGS);
MD_API int64_t MdRoiAlignPlanBytes(int R, int B, int C, int L, const int *H, const int *W);

/* Bit-exact variants: same I/O, gather kernels only (forward bit-identical to the oracle's op order;
 * used for levels/footprints the TMA path declines, and selectable by symbol because attributes cannot
 * be read on the host from a device cfg tensor). */
MD_API int MdRoiAlignFwdExact(MD_AOT_ARGS);
MD_API int MdRoiAlignBwdExact(MD_AOT_ARGS);

/* ---- "next" row 3: the reference's OWN GPU symbols, same names and parameter lists
 * (minddet/models/centerpoint/det3d_ms/ops/test_custom_pytorch/iou3d_nms_kernel.cu:445, :470, :493, :548; python side
 * iou_gpu.py:14-80), so `ops.Custom("<this .so>:NmsGpu", ...)` replaces `iou_nms.so:NmsGpu` without touching the cell.
 * Boxes are (N,7) f32 [x, y, z, dx, dy, dz, heading]; N <= 4096 for the NMS symbols.
 *   BoxesIouBevGpu / BoxesOverlapBevGpu : in boxes_a (N,7) | boxes_b (M,7)        out ans (N,M) f32
 *   NmsGpu (rotated) / NmsNormalGpu (axis-aligned): in boxes (N,7) score-sorted | thresh f32[1]
 *                                          out keep (N) int64 zero padded | num_to_keep int32[1]    (IoU > thresh suppresses)
 *   BoxesIouNmsGpu: device twin of the CPU op boxes_iou_nms_cpu (iou-bev-nms-org.cpp:237-283; nms_cpu.py:10-27):
 *                   in boxes (N,7) | thresh f32[1]   out keep (N) int32 | count int32[1]
 *                   (IoU >= thresh suppresses, zero-area boxes are dropped first; N is read from the shape instead of
 *                   the reference's hard-coded 1000) */
MD_API int BoxesIouBevGpu(MD_AOT_ARGS);
MD_API int BoxesOverlapBevGpu(MD_AOT_ARGS);
MD_API int NmsGpu(MD_AOT_ARGS);
MD_API int NmsNormalGpu(MD_AOT_ARGS);
MD_API int BoxesIouNmsGpu(MD_AOT_ARGS);

/* ---- "next" row 4: BoundingBoxEncode + RCNN-head post-process (oracle/CONVENTIONS.md #9, #22)
 *   MdEncode: in proposals (K,4) | gts (K,4) | cfg f32[8] = {means[4], stds[4]}      out deltas (K,4) f32
 *   MdRcnnPostProcess: in rois (B,P,4|5) f32 | roi_valid (B,P) uint8/bool | cls_logits (B,P,nc+1) f32 (class 0 = background)
 *                         | bbox_deltas (B,P,(nc+1)*4) f32 | cfg f32[13] = MD_CFG_DECODE (0..10) + {score_thr, iou_thr}
 *                      out dets (B,max_det,6) f32 [x1,y1,x2,y2,score,label] | keep_idx (B,max_det) int32 = roi*(nc+1)+class, -1 padded
 *                         | count (B) int32 | cand_idx (B,nms_pre) int32 (score-sorted NMS candidates, -1 padded; nms_pre <= 2048) */
MD_API int MdEncode(MD_AOT_ARGS);
MD_API int MdRcnnPostProcess(MD_AOT_ARGS);

/* ---- "next" row 2: Mask R-CNN mask targets (oracle/CONVENTIONS.md #21).  The 14x14 mask
 * RoIAlign itself is MdRoiAlignFwd / MdRoiAlignBwd with a (R,C,14,14) output.
 *   MdMaskTargets: in gt_masks (B,G,H,W) uint8/bool | rois (R,5) f32 [batch,x1,y1,x2,y2] | gt_idx (R) int32 (<0: none)
 *                     | cfg f32[1] = {sample_num}
 *                  out targets (R,M,M) uint8/bool        (M is read from the output shape; 28 upstream) */
MD_API int MdMaskTargets(MD_AOT_ARGS);

/* ---- "next" row 1: YOLOv8 post-process; reg_max = 16 (oracle/CONVENTIONS.md #19-#20)
 *   MdYoloDecode: in pred (B, 64+nc, A) f32 | cfg f32[1+3L] = {L, then per level H, W, stride}
 *                 out dets (B, A, 6) f32 [x1,y1,x2,y2,score,label]
 *   MdYoloNms   : in dets (B, A, 6) | cfg f32[3] = {conf_thr, iou_thr, agnostic}
 *                 out out (B,max_det,6) f32 zero padded | keep_idx (B,max_det) int32 (-1 padded) | count (B) int32
 *                     | cand_idx (B,nms_pre) int32: the score-sorted candidates that entered NMS (-1 padded)
 *                 (nms_pre <= 2048 and max_det are read from the output shapes) */
MD_API int MdYoloDecode(MD_AOT_ARGS);
MD_API int MdYoloNms(MD_AOT_ARGS);

/* library info: returns a static string "libmdregion <version> sm_100a" */
MD_API const char *MdVersion(void);

#ifdef __cplusplus
}
#endif
#endif /* MD_REGION_AOT_H_ */
