// TEST INFRASTRUCTURE (oracle/): host harness around the REFERENCE's own `iou_normal`.
//
// The reference's axis-aligned NMS arithmetic lives in a CUDA file,
//   /root/reference/minddet/models/centerpoint/det3d_ms/ops/test_custom_pytorch/iou3d_nms_kernel.cu
//   (`const float EPS` :42, `__device__ inline float iou_normal` :347-358, `nms_normal_kernel` :361-405,
//    host reduce of NmsNormalGpu :526-536),
// which needs libtorch and a GPU to run as a whole.  `iou_normal` itself is plain C float arithmetic, so
// oracle/Makefile cuts that one function and the EPS constant out of the file WHERE IT LIES into
// oracle/_ref/iou_normal_extract.inc (git-ignored, never committed, deleted again once the .so is linked) and this harness compiles them for the host with
// `__device__` defined away.  What runs below is therefore the reference's own expression tree; only the pair loop and
// the greedy sweep around it (the kernel's bit mask + the host reduce: box j is removed iff a kept i < j has
// iou_normal(i, j) > thr, strict) are restated here, because a __global__ kernel cannot run on the host.
// Built with -ffp-contract=off; the fixtures made from it sit on a 0.25-pixel lattice where every intermediate except the
// final division is exact, so nvcc's FMA contraction could not change a bit of them either.
#include <cmath>
#include <cstdint>
#include <vector>

#define __device__
#include "_ref/iou_normal_extract.inc"

extern "C" {

// (n,7) x (m,7) [x, y, z, dx, dy, dz, heading] -> (n,m) iou_normal
__attribute__((visibility("default"))) void ref_iou_normal_matrix(const float *a, int n, const float *b, int m, float *out)
{
    for (int i = 0; i < n; i++)
        for (int j = 0; j < m; j++) out[(int64_t)i * m + j] = iou_normal(a + i * 7, b + j * 7);
}

// boxes (n,7) already sorted by score; keep[0..count) = indices kept, in order (NmsNormalGpu's output convention)
__attribute__((visibility("default"))) void ref_nms_normal(const float *boxes, int n, float thr, int64_t *keep, int *count)
{
    std::vector<unsigned char> removed(n, 0);
    int k = 0;
    for (int i = 0; i < n; i++) {
        if (removed[i]) continue;
        keep[k++] = i;
        for (int j = i + 1; j < n; j++)
            if (iou_normal(boxes + i * 7, boxes + j * 7) > thr) removed[j] = 1;
    }
    *count = k;
}

float ref_iou_normal_eps(void) __attribute__((visibility("default")));
float ref_iou_normal_eps(void) { return EPS; }
}
