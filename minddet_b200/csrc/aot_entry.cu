// aot_entry.cu -- the extern "C" boundary declared in include/md_region_aot.h.
//
// Each symbol has the MindSpore ops.Custom(func_type="aot") signature the reference binds
// (centerpoint/det3d_ms/ops/test_custom_pytorch/iou3d_nms_kernel.cu:445-446, 491-492, 548-549).
// Differences from the reference's entry points, on purpose:
//   * work is enqueued on the stream MindSpore passes and never synchronised (the reference
//     cudaStreamSynchronize()s it and launches on the default stream, :448-449);
//   * no libtorch (the reference wraps raw pointers as at::Tensor, ms_ext.cpp:14-27);
//   * scratch comes from a per-(device,stream) grow-only workspace that is never freed or moved while the library is
//     loaded (safe under CUDA-graph replay), not cudaMalloc/cudaFree per call (:510-517);
//   * errors are returned, never exit()ed (:32-40).
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <mutex>
#include <utility>
#include <vector>

#include "../../include/md_region_aot.h"
#include "kernels.h"

namespace {

enum { MD_OK = 0, MD_ERR_NPARAM = 1, MD_ERR_ARG = 2, MD_ERR_CUDA = 3, MD_ERR_SIZE = 4, MD_ERR_CAPTURE = 5 };

// Per-(device, stream) state: scratch block, a small zero-initialised control block (work tickets the kernels re-arm
// themselves), the helper lane.  Graph-safe by construction:
//   * a block is NEVER freed or moved while the library is loaded: growth allocates a bigger block and retires the old
//     one (kept until MdShutdown), so a CUDA graph captured earlier keeps replaying against valid memory;
//   * nothing is allocated or created while the stream is being captured (cudaMalloc / cudaStreamCreate are not
//     capturable): the call returns MD_ERR_CAPTURE (5) instead -- run the op once, or MdReserve(), before capturing.
struct Workspace {
    void *ptr = nullptr; size_t bytes = 0;
    int *ctl = nullptr;
    std::vector<void *> retired;
    md::SideLane lane{}; bool has_lane = false;
};
std::mutex g_ws_mutex;
std::map<std::pair<int, void *>, Workspace> g_ws;

bool capturing(void *stream)
{
    cudaStreamCaptureStatus st = cudaStreamCaptureStatusNone;
    if (cudaStreamIsCapturing((cudaStream_t)stream, &st) != cudaSuccess) { cudaGetLastError(); return false; }
    return st != cudaStreamCaptureStatusNone;
}

// The device is the one the framework made current for this call (MindSpore binds one device per process / executor
// thread); tensors of another device are a caller error.
int get_workspace(void *stream, size_t bytes, void **out, int **ctl = nullptr)
{
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return MD_ERR_CUDA;
    std::lock_guard<std::mutex> lock(g_ws_mutex);
    Workspace &w = g_ws[std::make_pair(dev, stream)];
    if (w.bytes < bytes || !w.ctl) {
        if (capturing(stream)) return MD_ERR_CAPTURE;
        if (!w.ctl) {
            if (cudaMalloc(reinterpret_cast<void **>(&w.ctl), md::MD_CTL_INTS * sizeof(int)) != cudaSuccess ||
                cudaMemset(w.ctl, 0, md::MD_CTL_INTS * sizeof(int)) != cudaSuccess) {
                cudaGetLastError();
                w.ctl = nullptr;
                return MD_ERR_CUDA;
            }
        }
        if (w.bytes < bytes) {
            const size_t want = (bytes > 2 * w.bytes ? bytes : 2 * w.bytes) + 4096;
            void *p = nullptr;
            if (cudaMalloc(&p, want) != cudaSuccess) { cudaGetLastError(); return MD_ERR_CUDA; }
            if (w.ptr) w.retired.push_back(w.ptr);
            w.ptr = p; w.bytes = want;
        }
    }
    *out = w.ptr;
    if (ctl) *ctl = w.ctl;
    return MD_OK;
}

// the (device, stream) helper stream; created at the first non-capturing touch.  Inside a capture that meets no lane the
// op runs on one stream (same results, a little slower) -- never an error.
const md::SideLane *get_side_lane(void *stream)
{
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return nullptr;
    std::lock_guard<std::mutex> lock(g_ws_mutex);
    Workspace &w = g_ws[std::make_pair(dev, stream)];
    if (!w.has_lane) {
        if (capturing(stream)) return nullptr;
        md::SideLane l{};
        int prio = 0;                       // the helper lane is part of the caller's chain: same scheduling priority
        if (cudaStreamGetPriority((cudaStream_t)stream, &prio) != cudaSuccess) { cudaGetLastError(); prio = 0; }
        if (cudaStreamCreateWithPriority(&l.stream, cudaStreamNonBlocking, prio) != cudaSuccess ||
            cudaEventCreateWithFlags(&l.fork, cudaEventDisableTiming) != cudaSuccess ||
            cudaEventCreateWithFlags(&l.join, cudaEventDisableTiming) != cudaSuccess) {
            cudaGetLastError();
            return nullptr;
        }
        w.lane = l;
        w.has_lane = true;
    }
    return &w.lane;
}

bool is_f32(const char *d) { return d && std::strcmp(d, "float32") == 0; }
bool is_i32(const char *d) { return d && std::strcmp(d, "int32") == 0; }
bool is_i64(const char *d) { return d && std::strcmp(d, "int64") == 0; }
bool is_u8(const char *d) { return d && (std::strcmp(d, "uint8") == 0 || std::strcmp(d, "bool") == 0 || std::strcmp(d, "int8") == 0); }

int64_t numel(int nd, const int64_t *sh)
{
    int64_t n = 1;
    for (int i = 0; i < nd; i++) n *= sh[i];
    return n;
}
int cuda_rc(cudaError_t e)
{
    if (e != cudaSuccess && getenv("MD_VERBOSE")) fprintf(stderr, "[mdregion] CUDA error %d: %s\n", (int)e, cudaGetErrorString(e));
    return e == cudaSuccess ? MD_OK : (e == cudaErrorInvalidValue ? MD_ERR_SIZE : MD_ERR_CUDA);
}

#define REQ(cond) do { if (!(cond)) return MD_ERR_ARG; } while (0)
#define NEED_ARGS() do { if (!params || !ndims || !shapes || !dtypes) return MD_ERR_ARG; } while (0)

}  // namespace

extern "C" {

const char *MdVersion(void) { return "libmdregion 0.1.0 sm_100a"; }

int MdAnchorGrid(MD_AOT_ARGS)
{
    (void)extra;
    if (nparam != 3) return MD_ERR_NPARAM;
    NEED_ARGS();
    REQ(is_f32(dtypes[0]) && is_f32(dtypes[1]) && is_f32(dtypes[2]));
    REQ(ndims[0] == 2 && shapes[0][1] == 4 && ndims[2] == 4 && shapes[2][3] == 4 && shapes[2][2] == shapes[0][0]);
    REQ(numel(ndims[1], shapes[1]) >= 1);
    return cuda_rc(md::launch_anchor_grid((const float *)params[0], (int)shapes[0][0], (int)shapes[2][0], (int)shapes[2][1],
                                          (const float *)params[1], (float *)params[2], (cudaStream_t)stream));
}

int MdDecodeClip(MD_AOT_ARGS)
{
    (void)extra;
    if (nparam != 4) return MD_ERR_NPARAM;
    NEED_ARGS();
    for (int i = 0; i < 4; i++) REQ(is_f32(dtypes[i]));
    REQ(ndims[0] == 2 && shapes[0][1] == 4 && ndims[1] == 2 && shapes[1][1] == 4 && shapes[1][0] == shapes[0][0]);
    REQ(ndims[3] == 2 && shapes[3][0] == shapes[0][0] && shapes[3][1] == 4);
    REQ(numel(ndims[2], shapes[2]) >= MD_DEC_LEN);
    return cuda_rc(md::launch_decode_rows((const float *)params[0], (const float *)params[1], shapes[0][0],
                                          (const float *)params[2], (float *)params[3], (cudaStream_t)stream));
}

int MdDecodeLevel(MD_AOT_ARGS)
{
    (void)extra;
    if (nparam != 4) return MD_ERR_NPARAM;
    NEED_ARGS();
    for (int i = 0; i < 4; i++) REQ(is_f32(dtypes[i]));
    REQ(ndims[0] == 4 && ndims[1] == 2 && shapes[1][1] == 4 && shapes[0][1] == 4 * shapes[1][0]);
    const int B = (int)shapes[0][0], A = (int)shapes[1][0], H = (int)shapes[0][2], W = (int)shapes[0][3];
    REQ(ndims[3] == 3 && shapes[3][0] == B && shapes[3][1] == (int64_t)H * W * A && shapes[3][2] == 4);
    REQ(numel(ndims[2], shapes[2]) >= MD_DEC_LEN + 1);
    return cuda_rc(md::launch_decode_level((const float *)params[0], (const float *)params[1], B, A, H, W,
                                           (const float *)params[2], (float *)params[3], (cudaStream_t)stream));
}

int MdTopKPerLevel(MD_AOT_ARGS)
{
    (void)extra;
    if (nparam != 4) return MD_ERR_NPARAM;
    NEED_ARGS();
    REQ(is_f32(dtypes[0]) && is_f32(dtypes[1]) && is_f32(dtypes[2]) && is_i32(dtypes[3]));
    REQ(ndims[0] == 4 || ndims[0] == 2);
    const int B = (int)shapes[0][0];
    int A = 0, HW = 0;
    if (ndims[0] == 4) { A = (int)shapes[0][1]; HW = (int)(shapes[0][2] * shapes[0][3]); }
    else HW = (int)shapes[0][1];
    REQ(ndims[2] == 2 && ndims[3] == 2 && shapes[2][0] == B && shapes[3][0] == B && shapes[2][1] == shapes[3][1]);
    const int K = (int)shapes[2][1];
    if (K > 2048 || (int64_t)(A ? A : 1) * HW >= (1 << 22)) return MD_ERR_SIZE;
    REQ(numel(ndims[1], shapes[1]) >= 1);
    return cuda_rc(md::launch_topk((const float *)params[0], B, A, HW, K, (const float *)params[1],
                                   (float *)params[2], (int32_t *)params[3], (cudaStream_t)stream));
}

int MdNms(MD_AOT_ARGS)
{
    (void)extra;
    if (nparam != 5) return MD_ERR_NPARAM;
    NEED_ARGS();
    REQ(is_f32(dtypes[0]) && is_f32(dtypes[1]) && is_i32(dtypes[2]) && is_u8(dtypes[3]) && is_i32(dtypes[4]));
    REQ(ndims[0] == 2 || ndims[0] == 3);
    const int B = ndims[0] == 3 ? (int)shapes[0][0] : 1;
    const int K = (int)shapes[0][ndims[0] - 2], ld = (int)shapes[0][ndims[0] - 1];
    REQ(ld >= 4);
    REQ(numel(ndims[1], shapes[1]) >= MD_NMS_LEN);
    REQ(numel(ndims[2], shapes[2]) == (int64_t)B * K && numel(ndims[3], shapes[3]) == (int64_t)B * K && numel(ndims[4], shapes[4]) == B);
    if (K > 4096) return MD_ERR_SIZE;                  // sweep capacity: 64 mask words per row
    void *ws = nullptr;
    int rc = get_workspace(stream, md::nms_workspace_bytes(B, K), &ws);
    if (rc) return rc;
    return cuda_rc(md::launch_nms((const float *)params[0], ld, B, K, (const float *)params[1], ws,
                                  (int32_t *)params[2], (uint8_t *)params[3], (int32_t *)params[4], (cudaStream_t)stream));
}

int MdProposal(MD_AOT_ARGS)
{
    (void)extra;
    if (nparam < 8 || (nparam - 5) % 3 != 0) return MD_ERR_NPARAM;
    NEED_ARGS();
    const int L = (nparam - 5) / 3;
    if (L > md::kMaxLv) return MD_ERR_SIZE;
    md::LevelSet lv{};
    lv.L = L;
    const int B = (int)shapes[0][0];
    for (int l = 0; l < L; l++) {
        const int is = l, id = L + l, ib = 2 * L + l;
        REQ(is_f32(dtypes[is]) && is_f32(dtypes[id]) && is_f32(dtypes[ib]));
        REQ(ndims[is] == 4 && ndims[id] == 4 && ndims[ib] == 2 && shapes[ib][1] == 4);
        lv.A[l] = (int)shapes[is][1]; lv.H[l] = (int)shapes[is][2]; lv.W[l] = (int)shapes[is][3];
        REQ(shapes[is][0] == B && shapes[id][0] == B && shapes[id][1] == 4 * lv.A[l] && shapes[id][2] == lv.H[l] && shapes[id][3] == lv.W[l]);
        REQ(shapes[ib][0] == lv.A[l]);
        lv.scores[l] = (const float *)params[is]; lv.deltas[l] = (const float *)params[id]; lv.base[l] = (const float *)params[ib];
    }
    const int ic = 3 * L, o0 = 3 * L + 1;
    REQ(is_f32(dtypes[ic]) && numel(ndims[ic], shapes[ic]) >= MD_PROP_STRIDE0 + L);
    REQ(is_f32(dtypes[o0]) && is_u8(dtypes[o0 + 1]) && is_i32(dtypes[o0 + 2]) && is_u8(dtypes[o0 + 3]));
    REQ(ndims[o0] == 3 && shapes[o0][0] == B && shapes[o0][2] == 5);
    const int max_num = (int)shapes[o0][1];
    REQ(numel(ndims[o0 + 1], shapes[o0 + 1]) == (int64_t)B * max_num);
    REQ(ndims[o0 + 2] == 3 && shapes[o0 + 2][0] == B && shapes[o0 + 2][1] == L);
    const int nms_pre = (int)shapes[o0 + 2][2];
    REQ(numel(ndims[o0 + 3], shapes[o0 + 3]) == (int64_t)B * L * nms_pre);
    if (nms_pre > 2048) return MD_ERR_SIZE;
    void *ws = nullptr;
    int rc = get_workspace(stream, md::proposal_workspace_bytes(B, L, nms_pre), &ws);
    if (rc) return rc;
    return cuda_rc(md::launch_proposal(lv, B, nms_pre, max_num, (const float *)params[ic], ws, (float *)params[o0],
                                       (uint8_t *)params[o0 + 1], (int32_t *)params[o0 + 2], (uint8_t *)params[o0 + 3],
                                       (cudaStream_t)stream, get_side_lane(stream)));
}

int MdAssignSample(MD_AOT_ARGS)
{
    (void)extra;
    if (nparam != 14) return MD_ERR_NPARAM;
    NEED_ARGS();
    REQ(is_f32(dtypes[0]) && is_u8(dtypes[1]) && is_f32(dtypes[2]) && is_u8(dtypes[3]) && is_f32(dtypes[4]) && is_i32(dtypes[5]));
    REQ((ndims[0] == 2 || ndims[0] == 3) && shapes[0][ndims[0] - 1] == 4);
    REQ(ndims[2] == 3 && shapes[2][2] == 4);
    const int per_image = ndims[0] == 3;
    const int N = (int)shapes[0][ndims[0] - 2];
    const int B = (int)shapes[2][0], G = (int)shapes[2][1];
    REQ(!per_image || shapes[0][0] == B);
    REQ(numel(ndims[1], shapes[1]) == (per_image ? (int64_t)B * N : (int64_t)N));
    REQ(numel(ndims[3], shapes[3]) == (int64_t)B * G);
    REQ(numel(ndims[4], shapes[4]) >= MD_AS_LEN && numel(ndims[5], shapes[5]) >= 2);
    const int o = 6;
    REQ(is_i32(dtypes[o]) && is_i32(dtypes[o + 1]) && is_u8(dtypes[o + 2]) && is_i32(dtypes[o + 3]) && is_u8(dtypes[o + 4]) &&
        is_i32(dtypes[o + 5]) && is_f32(dtypes[o + 6]) && is_i32(dtypes[o + 7]));
    REQ(numel(ndims[o], shapes[o]) == (int64_t)B * N);
    REQ(ndims[o + 1] == 2 && ndims[o + 3] == 2 && shapes[o + 1][0] == B && shapes[o + 3][0] == B);
    const int Sp = (int)shapes[o + 1][1], Sn = (int)shapes[o + 3][1];
    REQ(numel(ndims[o + 2], shapes[o + 2]) == (int64_t)B * Sp && numel(ndims[o + 4], shapes[o + 4]) == (int64_t)B * Sn);
    REQ(numel(ndims[o + 5], shapes[o + 5]) == (int64_t)B * Sp && numel(ndims[o + 6], shapes[o + 6]) == (int64_t)B * Sp * 4);
    REQ(numel(ndims[o + 7], shapes[o + 7]) == B);
    if (Sp > 2048 || Sn > 2048 || N >= (1 << 22) || G > 1024) return MD_ERR_SIZE;
    void *ws = nullptr;
    int rc = get_workspace(stream, md::assign_workspace_bytes(B, G, N), &ws);
    if (rc) return rc;
    return cuda_rc(md::launch_assign_sample_rpn(
        (const float *)params[0], per_image, (const uint8_t *)params[1], B, N, (const float *)params[2],
        (const uint8_t *)params[3], G, (const float *)params[4], (const int32_t *)params[5], (int)numel(ndims[5], shapes[5]), ws, Sp, Sn,
        (int32_t *)params[o], (int32_t *)params[o + 1], (uint8_t *)params[o + 2], (int32_t *)params[o + 3],
        (uint8_t *)params[o + 4], (int32_t *)params[o + 5], (float *)params[o + 6], (int32_t *)params[o + 7],
        (cudaStream_t)stream));
}

int MdAssignSampleRcnn(MD_AOT_ARGS)
{
    (void)extra;
    if (nparam != 15) return MD_ERR_NPARAM;
    NEED_ARGS();
    REQ(is_f32(dtypes[0]) && is_u8(dtypes[1]) && is_f32(dtypes[2]) && is_i32(dtypes[3]) && is_u8(dtypes[4]) &&
        is_f32(dtypes[5]) && is_i32(dtypes[6]));
    REQ(ndims[0] == 3 && shapes[0][2] == 5 && ndims[2] == 3 && shapes[2][2] == 4);
    const int B = (int)shapes[0][0], P = (int)shapes[0][1], G = (int)shapes[2][1];
    REQ(shapes[2][0] == B);
    REQ(numel(ndims[1], shapes[1]) == (int64_t)B * P && numel(ndims[3], shapes[3]) == (int64_t)B * G && numel(ndims[4], shapes[4]) == (int64_t)B * G);
    REQ(numel(ndims[5], shapes[5]) >= MD_AS_LEN && numel(ndims[6], shapes[6]) >= 2);
    const int o = 7;
    REQ(is_f32(dtypes[o]) && is_f32(dtypes[o + 1]) && is_i32(dtypes[o + 2]) && is_u8(dtypes[o + 3]) && is_i32(dtypes[o + 4]) &&
        is_i32(dtypes[o + 5]) && is_i32(dtypes[o + 6]) && is_i32(dtypes[o + 7]));
    REQ(ndims[o] == 3 && shapes[o][0] == B && shapes[o][2] == 5);
    const int S = (int)shapes[o][1];
    REQ(ndims[o + 6] == 2 && shapes[o + 6][0] == B);
    const int Sp = (int)shapes[o + 6][1], Sn = S - Sp;
    REQ(Sn >= 0);
    REQ(numel(ndims[o + 1], shapes[o + 1]) == (int64_t)B * S * 4 && numel(ndims[o + 2], shapes[o + 2]) == (int64_t)B * S &&
        numel(ndims[o + 3], shapes[o + 3]) == (int64_t)B * S && numel(ndims[o + 4], shapes[o + 4]) == (int64_t)B * (G + P) &&
        numel(ndims[o + 5], shapes[o + 5]) == (int64_t)B * S && numel(ndims[o + 7], shapes[o + 7]) == B);
    if (Sp > 2048 || Sn > 2048 || G + P >= (1 << 22) || G > 1024) return MD_ERR_SIZE;
    void *ws = nullptr;
    int rc = get_workspace(stream, md::assign_workspace_bytes(B, G, P), &ws);
    if (rc) return rc;
    return cuda_rc(md::launch_assign_sample_rcnn(
        (const float *)params[0], (const uint8_t *)params[1], B, P, (const float *)params[2], (const int32_t *)params[3],
        (const uint8_t *)params[4], G, (const float *)params[5], (const int32_t *)params[6], (int)numel(ndims[6], shapes[6]), ws, Sp, Sn,
        (float *)params[o], (float *)params[o + 1], (int32_t *)params[o + 2], (uint8_t *)params[o + 3],
        (int32_t *)params[o + 4], (int32_t *)params[o + 5], (int32_t *)params[o + 6], (int32_t *)params[o + 7],
        (cudaStream_t)stream));
}

int MdRoiLevels(MD_AOT_ARGS)
{
    (void)extra;
    if (nparam != 3) return MD_ERR_NPARAM;
    NEED_ARGS();
    REQ(is_f32(dtypes[0]) && is_f32(dtypes[1]) && is_i32(dtypes[2]));
    REQ(ndims[0] == 2 && shapes[0][1] == 5 && numel(ndims[1], shapes[1]) >= 2 && numel(ndims[2], shapes[2]) == shapes[0][0]);
    return cuda_rc(md::launch_roi_levels((const float *)params[0], (int)shapes[0][0], (const float *)params[1],
                                         (int32_t *)params[2], (cudaStream_t)stream));
}

static int parse_feats(int L, int first, void **params, int *ndims, int64_t **shapes, const char **dtypes, md::FeatSet *fs)
{
    if (L < 1 || L > md::kMaxLv) return MD_ERR_SIZE;
    fs->L = L;
    for (int l = 0; l < L; l++) {
        const int i = first + l;
        REQ(is_f32(dtypes[i]) && ndims[i] == 4);
        if (l == 0) { fs->B = (int)shapes[i][0]; fs->C = (int)shapes[i][1]; }
        REQ(shapes[i][0] == fs->B && shapes[i][1] == fs->C);
        fs->H[l] = (int)shapes[i][2]; fs->W[l] = (int)shapes[i][3];
        fs->feat[l] = (float *)params[i];
    }
    return MD_OK;
}

static int roialign_fwd_impl(int mode, MD_AOT_ARGS)
{
    (void)extra;
    if (nparam < 4) return MD_ERR_NPARAM;
    NEED_ARGS();
    const int L = nparam - 3;
    md::FeatSet fs{};
    int rc = parse_feats(L, 1, params, ndims, shapes, dtypes, &fs);
    if (rc) return rc;
    const int ic = 1 + L, io = 2 + L;
    REQ(is_f32(dtypes[0]) && ndims[0] == 2 && shapes[0][1] == 5);
    REQ(is_f32(dtypes[ic]) && numel(ndims[ic], shapes[ic]) >= MD_ROI_STRIDE0 + L);
    const int R = (int)shapes[0][0];
    REQ(is_f32(dtypes[io]) && ndims[io] == 4 && shapes[io][0] == R && shapes[io][1] == fs.C && shapes[io][2] == shapes[io][3]);
    const int P = (int)shapes[io][2];
    void *ws = nullptr;
    int *ctl = nullptr;
    rc = get_workspace(stream, md::roialign_workspace_bytes(R), &ws, &ctl);
    if (rc) return rc;
    return cuda_rc(md::launch_roialign_fwd(fs, (const float *)params[0], R, P, (const float *)params[ic],
                                           (float *)params[io], ws, ctl, mode, (cudaStream_t)stream));
}

static int roialign_bwd_impl(int mode, MD_AOT_ARGS)
{
    (void)extra;
    if (nparam < 4) return MD_ERR_NPARAM;
    NEED_ARGS();
    const int L = nparam - 3;
    md::FeatSet fs{};
    int rc = parse_feats(L, 3, params, ndims, shapes, dtypes, &fs);
    if (rc) return rc;
    REQ(is_f32(dtypes[0]) && ndims[0] == 2 && shapes[0][1] == 5);
    const int R = (int)shapes[0][0];
    REQ(is_f32(dtypes[1]) && ndims[1] == 4 && shapes[1][0] == R && shapes[1][1] == fs.C && shapes[1][2] == shapes[1][3]);
    REQ(is_f32(dtypes[2]) && numel(ndims[2], shapes[2]) >= MD_ROI_STRIDE0 + L);
    const int P = (int)shapes[1][2];
    void *ws = nullptr;
    int *ctl = nullptr;
    rc = get_workspace(stream, md::roialign_bwd_workspace_bytes(fs, R), &ws, &ctl);
    if (rc) return rc;
    return cuda_rc(md::launch_roialign_bwd(fs, (const float *)params[0], R, P, (const float *)params[2],
                                           (const float *)params[1], ws, ctl, mode, (cudaStream_t)stream));
}
static int roialign_bwd_acc_impl(MD_AOT_ARGS)
{
    (void)extra;
    if (nparam < 5) return MD_ERR_NPARAM;
    NEED_ARGS();
    const int L = nparam - 4;
    md::FeatSet fs{};
    int rc = parse_feats(L, 3, params, ndims, shapes, dtypes, &fs);     // the accumulators sit where MdRoiAlignBwd has its outputs
    if (rc) return rc;
    REQ(is_f32(dtypes[0]) && ndims[0] == 2 && shapes[0][1] == 5);
    const int R = (int)shapes[0][0];
    REQ(is_f32(dtypes[1]) && ndims[1] == 4 && shapes[1][0] == R && shapes[1][1] == fs.C && shapes[1][2] == shapes[1][3]);
    REQ(is_f32(dtypes[2]) && numel(ndims[2], shapes[2]) >= MD_ROI_STRIDE0 + L);
    REQ(is_i32(dtypes[nparam - 1]) && numel(ndims[nparam - 1], shapes[nparam - 1]) >= 1);
    const int P = (int)shapes[1][2];
    void *ws = nullptr;
    int *ctl = nullptr;
    rc = get_workspace(stream, md::roialign_bwd_workspace_bytes(fs, R), &ws, &ctl);
    if (rc) return rc;
    rc = cuda_rc(cudaMemsetAsync(params[nparam - 1], 0, sizeof(int32_t), (cudaStream_t)stream));
    if (rc) return rc;
    return cuda_rc(md::launch_roialign_bwd(fs, (const float *)params[0], R, P, (const float *)params[2],
                                           (const float *)params[1], ws, ctl, 0, (cudaStream_t)stream, true));
}

// two-op form of the backward: rois | feat_0..L-1 (level shapes only, not read) | cfg -> plan (int32, MdRoiAlignPlanBytes / 4)
static int roialign_bwd_prepare_impl(MD_AOT_ARGS)
{
    (void)extra;
    if (nparam < 4) return MD_ERR_NPARAM;
    NEED_ARGS();
    const int L = nparam - 3;
    md::FeatSet fs{};
    int rc = parse_feats(L, 1, params, ndims, shapes, dtypes, &fs);
    if (rc) return rc;
    const int ic = 1 + L, io = 2 + L;
    REQ(is_f32(dtypes[0]) && ndims[0] == 2 && shapes[0][1] == 5);
    REQ(is_f32(dtypes[ic]) && numel(ndims[ic], shapes[ic]) >= MD_ROI_STRIDE0 + L);
    const int R = (int)shapes[0][0];
    REQ(is_i32(dtypes[io]) && (size_t)numel(ndims[io], shapes[io]) * 4 >= md::roialign_plan_bytes(fs, R));
    return cuda_rc(md::launch_roialign_bwd_prepare(fs, (const float *)params[0], R, 7, (const float *)params[ic], params[io],
                                                   (cudaStream_t)stream));
}
// rois | dout | cfg | plan -> dfeat_0..L-1
static int roialign_bwd_planned_impl(MD_AOT_ARGS)
{
    (void)extra;
    if (nparam < 5) return MD_ERR_NPARAM;
    NEED_ARGS();
    const int L = nparam - 4;
    md::FeatSet fs{};
    int rc = parse_feats(L, 4, params, ndims, shapes, dtypes, &fs);
    if (rc) return rc;
    REQ(is_f32(dtypes[0]) && ndims[0] == 2 && shapes[0][1] == 5);
    const int R = (int)shapes[0][0];
    REQ(is_f32(dtypes[1]) && ndims[1] == 4 && shapes[1][0] == R && shapes[1][1] == fs.C && shapes[1][2] == 7 && shapes[1][3] == 7);
    REQ(is_f32(dtypes[2]) && numel(ndims[2], shapes[2]) >= MD_ROI_STRIDE0 + L);
    REQ(is_i32(dtypes[3]) && (size_t)numel(ndims[3], shapes[3]) * 4 >= md::roialign_plan_bytes(fs, R));
    return cuda_rc(md::launch_roialign_bwd_planned(fs, (const float *)params[0], R, 7, (const float *)params[2], (const float *)params[1],
                                                   params[3], (cudaStream_t)stream));
}

// ---- the reference's own GPU symbols (iou3d_nms_kernel.cu:445-601), same parameter lists ------------------------
static int bev_pairs_impl(int want_iou, MD_AOT_ARGS)
{
    (void)extra;
    if (nparam != 3) return MD_ERR_NPARAM;
    NEED_ARGS();
    REQ(is_f32(dtypes[0]) && is_f32(dtypes[1]) && is_f32(dtypes[2]));
    REQ(ndims[0] == 2 && ndims[1] == 2 && shapes[0][1] == 7 && shapes[1][1] == 7);
    const int na = (int)shapes[0][0], nb = (int)shapes[1][0];
    REQ(numel(ndims[2], shapes[2]) == (int64_t)na * nb);
    return cuda_rc(md::launch_bev_pairs((const float *)params[0], na, (const float *)params[1], nb, want_iou,
                                        (float *)params[2], (cudaStream_t)stream));
}
static int bev_nms_impl(int mode, int keep_is_64, MD_AOT_ARGS)
{
    (void)extra;
    if (nparam != 4) return MD_ERR_NPARAM;
    NEED_ARGS();
    REQ(is_f32(dtypes[0]) && is_f32(dtypes[1]) && is_i32(dtypes[3]));
    REQ(keep_is_64 ? is_i64(dtypes[2]) : is_i32(dtypes[2]));
    REQ(ndims[0] == 2 && shapes[0][1] == 7 && numel(ndims[1], shapes[1]) >= 1 && numel(ndims[3], shapes[3]) >= 1);
    const int n = (int)shapes[0][0];
    REQ(numel(ndims[2], shapes[2]) == n);
    if (n > 4096) return MD_ERR_SIZE;
    void *ws = nullptr;
    int rc = get_workspace(stream, md::bev_nms_workspace_bytes(n), &ws);
    if (rc) return rc;
    return cuda_rc(md::launch_bev_nms((const float *)params[0], n, (const float *)params[1], mode, ws, params[2], keep_is_64,
                                      (int32_t *)params[3], (cudaStream_t)stream));
}
int BoxesIouBevGpu(MD_AOT_ARGS) { return bev_pairs_impl(1, nparam, params, ndims, shapes, dtypes, stream, extra); }
int BoxesOverlapBevGpu(MD_AOT_ARGS) { return bev_pairs_impl(0, nparam, params, ndims, shapes, dtypes, stream, extra); }
int NmsGpu(MD_AOT_ARGS) { return bev_nms_impl(0, 1, nparam, params, ndims, shapes, dtypes, stream, extra); }
int NmsNormalGpu(MD_AOT_ARGS) { return bev_nms_impl(1, 1, nparam, params, ndims, shapes, dtypes, stream, extra); }
int BoxesIouNmsGpu(MD_AOT_ARGS) { return bev_nms_impl(2, 0, nparam, params, ndims, shapes, dtypes, stream, extra); }

int MdRcnnPostProcess(MD_AOT_ARGS)
{
    (void)extra;
    if (nparam != 9) return MD_ERR_NPARAM;
    NEED_ARGS();
    REQ(is_f32(dtypes[0]) && is_u8(dtypes[1]) && is_f32(dtypes[2]) && is_f32(dtypes[3]) && is_f32(dtypes[4]));
    REQ(is_f32(dtypes[5]) && is_i32(dtypes[6]) && is_i32(dtypes[7]) && is_i32(dtypes[8]));
    REQ(ndims[0] == 3 && (shapes[0][2] == 4 || shapes[0][2] == 5) && ndims[2] == 3);
    const int B = (int)shapes[0][0], P = (int)shapes[0][1], nc1 = (int)shapes[2][2];
    REQ(shapes[2][0] == B && shapes[2][1] == P && nc1 >= 2);
    REQ(numel(ndims[1], shapes[1]) == (int64_t)B * P && numel(ndims[3], shapes[3]) == (int64_t)B * P * nc1 * 4);
    REQ(numel(ndims[4], shapes[4]) >= 13);
    REQ(ndims[5] == 3 && shapes[5][0] == B && shapes[5][2] == 6);
    const int max_det = (int)shapes[5][1];
    REQ(numel(ndims[6], shapes[6]) == (int64_t)B * max_det && numel(ndims[7], shapes[7]) == B && ndims[8] == 2 && shapes[8][0] == B);
    const int nms_pre = (int)shapes[8][1];
    if (nms_pre > 2048 || nms_pre < 1 || (int64_t)P * nc1 >= (1 << 22)) return MD_ERR_SIZE;
    void *ws = nullptr;
    int rc = get_workspace(stream, md::rcnn_post_workspace_bytes(B, P, nc1, nms_pre), &ws);
    if (rc) return rc;
    return cuda_rc(md::launch_rcnn_post((const float *)params[0], (int)shapes[0][2], (const uint8_t *)params[1], (const float *)params[2],
                                        (const float *)params[3], B, P, nc1, (const float *)params[4], ws, nms_pre, max_det,
                                        (float *)params[5], (int32_t *)params[6], (int32_t *)params[7], (int32_t *)params[8],
                                        (cudaStream_t)stream));
}

int MdEncode(MD_AOT_ARGS)
{
    (void)extra;
    if (nparam != 4) return MD_ERR_NPARAM;
    NEED_ARGS();
    for (int i = 0; i < 4; i++) REQ(is_f32(dtypes[i]));
    REQ(ndims[0] == 2 && shapes[0][1] == 4 && ndims[1] == 2 && shapes[1][1] == 4 && shapes[1][0] == shapes[0][0]);
    REQ(numel(ndims[2], shapes[2]) >= 8 && numel(ndims[3], shapes[3]) == shapes[0][0] * 4);
    return cuda_rc(md::launch_encode_rows((const float *)params[0], (const float *)params[1], shapes[0][0], (const float *)params[2],
                                          (float *)params[3], (cudaStream_t)stream));
}

int MdMaskTargets(MD_AOT_ARGS)
{
    (void)extra;
    if (nparam != 5) return MD_ERR_NPARAM;
    NEED_ARGS();
    REQ(is_u8(dtypes[0]) && is_f32(dtypes[1]) && is_i32(dtypes[2]) && is_f32(dtypes[3]) && is_u8(dtypes[4]));
    REQ(ndims[0] == 4 && ndims[1] == 2 && shapes[1][1] == 5 && numel(ndims[3], shapes[3]) >= 1);
    const int R = (int)shapes[1][0];
    REQ(numel(ndims[2], shapes[2]) == R && ndims[4] == 3 && shapes[4][0] == R && shapes[4][1] == shapes[4][2]);
    return cuda_rc(md::launch_mask_targets((const uint8_t *)params[0], (int)shapes[0][0], (int)shapes[0][1], (int)shapes[0][2],
                                           (int)shapes[0][3], (const float *)params[1], (const int32_t *)params[2], R,
                                           (int)shapes[4][1], (const float *)params[3], (uint8_t *)params[4], (cudaStream_t)stream));
}

int MdYoloDecode(MD_AOT_ARGS)
{
    (void)extra;
    if (nparam != 3) return MD_ERR_NPARAM;
    NEED_ARGS();
    REQ(is_f32(dtypes[0]) && is_f32(dtypes[1]) && is_f32(dtypes[2]));
    REQ(ndims[0] == 3 && shapes[0][1] > 64 && ndims[2] == 3 && shapes[2][0] == shapes[0][0] && shapes[2][1] == shapes[0][2] && shapes[2][2] == 6);
    REQ(numel(ndims[1], shapes[1]) >= 4);
    return cuda_rc(md::launch_yolo_decode((const float *)params[0], (int)shapes[0][0], (int)shapes[0][1], (int)shapes[0][2],
                                          (const float *)params[1], (int)numel(ndims[1], shapes[1]), (float *)params[2], (cudaStream_t)stream));
}

int MdYoloNms(MD_AOT_ARGS)
{
    (void)extra;
    if (nparam != 6) return MD_ERR_NPARAM;
    NEED_ARGS();
    REQ(is_f32(dtypes[0]) && is_f32(dtypes[1]) && is_f32(dtypes[2]) && is_i32(dtypes[3]) && is_i32(dtypes[4]) && is_i32(dtypes[5]));
    REQ(ndims[0] == 3 && shapes[0][2] == 6 && numel(ndims[1], shapes[1]) >= 3);
    const int B = (int)shapes[0][0], A = (int)shapes[0][1];
    REQ(ndims[2] == 3 && shapes[2][0] == B && shapes[2][2] == 6);
    const int max_det = (int)shapes[2][1];
    REQ(numel(ndims[3], shapes[3]) == (int64_t)B * max_det && numel(ndims[4], shapes[4]) == B);
    REQ(ndims[5] == 2 && shapes[5][0] == B);
    const int nms_pre = (int)shapes[5][1];
    if (nms_pre > 2048 || nms_pre < 1 || A >= (1 << 22)) return MD_ERR_SIZE;
    void *ws = nullptr;
    int rc = get_workspace(stream, md::yolo_nms_workspace_bytes(B, nms_pre), &ws);
    if (rc) return rc;
    return cuda_rc(md::launch_yolo_nms((const float *)params[0], B, A, (const float *)params[1], ws, nms_pre, max_det,
                                       (float *)params[2], (int32_t *)params[3], (int32_t *)params[4], (int32_t *)params[5],
                                       (cudaStream_t)stream));
}

int MdRoiAlignFwd(MD_AOT_ARGS) { return roialign_fwd_impl(0, nparam, params, ndims, shapes, dtypes, stream, extra); }
int MdRoiAlignBwd(MD_AOT_ARGS) { return roialign_bwd_impl(0, nparam, params, ndims, shapes, dtypes, stream, extra); }
int MdRoiAlignBwdAcc(MD_AOT_ARGS) { return roialign_bwd_acc_impl(nparam, params, ndims, shapes, dtypes, stream, extra); }
int MdRoiAlignBwdPrepare(MD_AOT_ARGS) { return roialign_bwd_prepare_impl(nparam, params, ndims, shapes, dtypes, stream, extra); }
int MdRoiAlignBwdPlanned(MD_AOT_ARGS) { return roialign_bwd_planned_impl(nparam, params, ndims, shapes, dtypes, stream, extra); }
int64_t MdRoiAlignPlanBytes(int R, int B, int C, int L, const int *H, const int *W)
{
    if (R < 0 || L < 1 || L > md::kMaxLv || !H || !W) return -1;
    md::FeatSet fs{};
    fs.L = L; fs.B = B; fs.C = C;
    for (int l = 0; l < L; l++) { fs.H[l] = H[l]; fs.W[l] = W[l]; }
    return (int64_t)md::roialign_plan_bytes(fs, R);
}
int MdRoiAlignFwdExact(MD_AOT_ARGS) { return roialign_fwd_impl(1, nparam, params, ndims, shapes, dtypes, stream, extra); }
int MdRoiAlignBwdExact(MD_AOT_ARGS) { return roialign_bwd_impl(1, nparam, params, ndims, shapes, dtypes, stream, extra); }

}  // extern "C"
