// TEST INFRASTRUCTURE (oracle/): host harness around the REFERENCE's own __device__ functions.
//
// The reference's GPU NMS / IoU arithmetic lives in a CUDA file,
//   /root/reference/minddet/models/centerpoint/det3d_ms/ops/test_custom_pytorch/iou3d_nms_kernel.cu
//   (`const float EPS` :44, `struct Point` and the rotated-box helpers :45-133, `box_overlap` :135-256, `iou_bev` :258-265,
//    `iou_normal` :347-358; kernels `nms_kernel` :300-344, `nms_normal_kernel` :361-405, host reduce :526-536),
// which needs libtorch and a GPU to run as a whole.  The __device__ functions themselves are plain C++ float arithmetic,
// so oracle/Makefile cuts them out of the file WHERE IT LIES into oracle/_ref/iou3d_device_extract.inc (git-ignored, never
// committed, deleted again once the .so is linked) and this harness compiles them for the host with `__device__` defined
// away.  What runs below is therefore the reference's own expression tree; only the pair loops and the greedy sweep around
// it (the kernels' bit mask + the host reduce: box j is removed iff a kept i < j has iou(i, j) > thr, strict) are restated
// here, because a __global__ kernel cannot run on the host.
// Built with -ffp-contract=off (nvcc's default FMA contraction is the one thing the host run cannot reproduce): the
// iou_normal fixtures sit on a 0.25-pixel lattice where every intermediate except the final division is exact, so
// contraction could not change a bit of them; the rotated functions go through sinf / cosf / atan2f (glibc here, CUDA's on
// the device) and are compared within a tolerance anyway.
#include <math.h>

#include <algorithm>
#include <cstdint>
#include <cstdio>
#include <vector>

using std::max;   // the CUDA built-ins min / max the extracted code calls
using std::min;

#define __device__
#include "_ref/iou3d_device_extract.inc"
#undef __device__

#define REF_API extern "C" __attribute__((visibility("default")))

// (n,7) x (m,7) [x, y, z, dx, dy, dz, heading] -> (n,m); which: 0 = iou_normal, 1 = iou_bev, 2 = box_overlap
REF_API void ref_cu_pair_matrix(int which, const float *a, int n, const float *b, int m, float *out)
{
    for (int i = 0; i < n; i++)
        for (int j = 0; j < m; j++) {
            const float *pa = a + i * 7, *pb = b + j * 7;
            out[(int64_t)i * m + j] = which == 0 ? iou_normal(pa, pb) : which == 1 ? iou_bev(pa, pb) : box_overlap(pa, pb);
        }
}

// boxes (n,7) already sorted by score; keep[0..count) = indices kept, in order (NmsGpu / NmsNormalGpu output convention);
// rotated != 0: iou_bev (nms_kernel), else iou_normal (nms_normal_kernel)
REF_API void ref_cu_nms(int rotated, const float *boxes, int n, float thr, int64_t *keep, int *count)
{
    std::vector<unsigned char> removed(n, 0);
    int k = 0;
    for (int i = 0; i < n; i++) {
        if (removed[i]) continue;
        keep[k++] = i;
        for (int j = i + 1; j < n; j++) {
            const float v = rotated ? iou_bev(boxes + i * 7, boxes + j * 7) : iou_normal(boxes + i * 7, boxes + j * 7);
            if (v > thr) removed[j] = 1;
        }
    }
    *count = k;
}

REF_API float ref_cu_eps(void) { return EPS; }
