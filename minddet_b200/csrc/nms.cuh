// nms.cuh -- segment descriptor + launcher of the bitmask NMS (proposal.cu), shared with yolo.cu
#pragma once
#include "kernels.h"

namespace md {

struct NmsSegs {                 // segment s -> boxes + K
    const float *boxes; int ld;  // rows of `ld` floats, segment stride = seg_stride rows
    int seg_stride;              // rows between consecutive segments
    int L;                       // K depends on (s % L)
    int K[kMaxLv];
    int nbp;                     // mask row pitch in u64 words (even)
    int rows_pad;                // mask rows per segment (multiple of 64)
    const int32_t *labels;       // optional (nseg, seg_stride): only boxes of equal label suppress each other ...
    const float *agnostic;       // ... unless *agnostic != 0 (device flag); both null for class-agnostic NMS
    int l0, nl;                  // optional level subset: launch index z serves segment (z / nl) * L + l0 + z % nl (nl == 0: z)
    const float *scores;         // optional (nseg, seg_stride) with kept_keys: the sweep also writes the monotone keys of
    uint32_t *kept_keys;         // the kept boxes' scores, in kept order, (nseg, keep_stride) -- the cross-level merge's input
    int labels_sorted;           // the rows of a segment are grouped by label (run_nms's label-major permutation): a tile whose row and
                                 // column label ranges do not meet is all zeros and is not evaluated
    const int32_t *dyn_k;        // optional (nseg): rows really present in a segment (device side); rows beyond it are padding and
                                 // neither the mask tiles nor the sweep touch them (a post-process with few candidates pays for those only)
};

// cfg: MD_CFG_NMS (thr, offset, inclusive, union_eps) on the device; keep_pos/keep_mask strides in elements
cudaError_t run_nms(const NmsSegs &sg, int nseg, int Kmax, const float *cfg, unsigned long long *mask,
                    int32_t *keep_pos, int keep_stride, uint8_t *keep_mask, int mask_stride,
                    int32_t *count, cudaStream_t s);

}  // namespace md
