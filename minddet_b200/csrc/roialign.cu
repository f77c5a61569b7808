// roialign.cu -- a9..a11: RoI -> level map, FPN RoIAlign forward / backward.  sm_100a.
//
// No reference code exists for this op (the only bilinear gather in the
// tree is centerpoint/det3d_ms/core/utils/center_utils.py:97-131).  Semantics: oracle/CONVENTIONS.md
// #14-16 (Caffe2 / aligned=False RoIAlign, average of S x S samples).  Unlike the upstream graph, which
// runs ROIAlign on all 4 levels for every RoI and selects afterwards, only the mapped level is touched.
//
// Forward arithmetic uses explicitly rounded mul/add in the oracle's order, so the result is
// bit-identical to oracle/region_oracle.c:o_roialign_fwd.  Backward accumulates with float atomics
// (order is not deterministic -> FP tolerance only, as north_star allows).
#include <cstdlib>

#include "kernels.h"
#include "roialign_common.cuh"
#include "launch.cuh"

namespace md {

__global__ void roi_levels_kernel(const float *__restrict__ rois5, int R, const float *__restrict__ cfg,
                                  int32_t *__restrict__ out)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= R) return;
    float r[4];
#pragma unroll
    for (int k = 0; k < 4; k++) r[k] = __ldg(rois5 + (int64_t)i * 5 + 1 + k);
    out[i] = roi_level_of(r, __ldg(cfg), (int)__ldg(cfg + 1));
}

cudaError_t launch_roi_levels(const float *rois5, int R, const float *cfg, int32_t *out, cudaStream_t s)
{
    if (R == 0) return cudaSuccess;
    roi_levels_kernel<<<(R + 255) / 256, 256, 0, s>>>(rois5, R, cfg, out);
    return cudaGetLastError();
}

MD_DEVINL void build_taps(const RoiGeom &g, int P, int S, Tap *taps)
{
    const int n = P * P * S * S;
    for (int t = threadIdx.x; t < n; t += blockDim.x) {
        const int ix = t % S, iy = (t / S) % S, bin = t / (S * S);
        const int pw = bin % P, ph = bin / P;
        taps[t] = make_tap(sample_coord(g.sh, g.bh, ph, iy, S), sample_coord(g.sw, g.bw, pw, ix, S), g.H, g.W);
    }
}

// ---- generic gather kernels (any level pitch, any footprint) --------------------------------------------
constexpr int kRoiThreads = 256;

__global__ void __launch_bounds__(kRoiThreads)
roialign_fwd_gather_kernel(const RoiFeat f, const float *__restrict__ rois5, int R, int P, int csplit,
                           float *__restrict__ out, const int32_t *__restrict__ only_flagged)
{
    pdl_entry();
    __shared__ Tap taps[kRoiMaxTaps];
    // grid.x <= R: a block walks RoIs blockIdx.x, blockIdx.x + gridDim.x, ... (as the fallback behind the TMA kernel the
    // grid is a few blocks per SM and nearly every RoI is skipped after one flag load)
    for (int r = blockIdx.x; r < R; r += gridDim.x) {
        if (only_flagged && !only_flagged[r]) continue;   // handled by the TMA kernel (uniform over the block)
        const int S = (int)__ldg(f.cfg + 1);
        const RoiGeom g = roi_geometry(f, rois5 + (int64_t)r * 5, P);
        const int PP = P * P, SS = S * S;
        const int cper = (f.C + csplit - 1) / csplit;
        const int c0 = blockIdx.y * cper, c1 = min(f.C, c0 + cper);
        if (!g.ok || S < 1 || PP * SS > kRoiMaxTaps) {     // bad batch index / sample count the tap table cannot hold: zeros
            for (int o = c0 * PP + threadIdx.x; o < c1 * PP; o += kRoiThreads) out[(int64_t)r * f.C * PP + o] = 0.0f;
            continue;
        }
        build_taps(g, P, S, taps);
        __syncthreads();
        const float cnt = (float)SS;
        const int64_t plane = (int64_t)g.H * g.W;
        const float *fb = f.feat[g.l] + (int64_t)g.b * f.C * plane;
        for (int o = c0 * PP + threadIdx.x; o < c1 * PP; o += kRoiThreads) {
            const int c = o / PP, bin = o - c * PP;
            const float *fp = fb + (int64_t)c * plane;
            float sum = 0.0f;
            for (int s = 0; s < SS; s++) {
                const Tap t = taps[bin * SS + s];
                float v = add(mul(t.w1, __ldg(fp + t.o1)), mul(t.w2, __ldg(fp + t.o2)));
                v = add(v, mul(t.w3, __ldg(fp + t.o3)));
                v = add(v, mul(t.w4, __ldg(fp + t.o4)));
                sum = add(sum, v);
            }
            out[(int64_t)r * f.C * PP + o] = div(sum, cnt);
        }
        __syncthreads();                                   // taps are rebuilt for the next RoI
    }
}

__global__ void __launch_bounds__(kRoiThreads)
roialign_bwd_gather_kernel(const RoiFeat f, const float *__restrict__ rois5, int R, int P, int csplit,
                           const float *__restrict__ dout, const int32_t *__restrict__ only_flagged, const int32_t *__restrict__ nflagged)
{
    pdl_entry();
    if (nflagged && *nflagged == 0) return;                // the tile-stationary kernel counted the RoIs it left to this one
    __shared__ Tap taps[kRoiMaxTaps];
    for (int r = blockIdx.x; r < R; r += gridDim.x) {      // see roialign_fwd_gather_kernel
        if (only_flagged && !only_flagged[r]) continue;
        const int S = (int)__ldg(f.cfg + 1);
        const RoiGeom g = roi_geometry(f, rois5 + (int64_t)r * 5, P);
        const int PP = P * P, SS = S * S;
        if (!g.ok || S < 1 || PP * SS > kRoiMaxTaps) continue;   // bad batch index / oversized tap table: no gradient
        build_taps(g, P, S, taps);
        __syncthreads();
        const float cnt = (float)SS;
        const int cper = (f.C + csplit - 1) / csplit;
        const int c0 = blockIdx.y * cper, c1 = min(f.C, c0 + cper);
        const int64_t plane = (int64_t)g.H * g.W;
        float *fb = f.feat[g.l] + (int64_t)g.b * f.C * plane;
        for (int o = c0 * PP + threadIdx.x; o < c1 * PP; o += kRoiThreads) {
            const int c = o / PP, bin = o - c * PP;
            float *fp = fb + (int64_t)c * plane;
            const float gr = div(__ldg(dout + (int64_t)r * f.C * PP + o), cnt);
            for (int s = 0; s < SS; s++) {
                const Tap t = taps[bin * SS + s];
                if (t.w1 != 0.0f) atomicAdd(fp + t.o1, mul(gr, t.w1));
                if (t.w2 != 0.0f) atomicAdd(fp + t.o2, mul(gr, t.w2));
                if (t.w3 != 0.0f) atomicAdd(fp + t.o3, mul(gr, t.w3));
                if (t.w4 != 0.0f) atomicAdd(fp + t.o4, mul(gr, t.w4));
            }
        }
        __syncthreads();
    }
}

// ---- a13: mask targets = RoIAlign (scale 1) of the assigned gt's uint8 mask plane, binarised with >= 0.5 --------------
// One thread per output pixel; op order identical to oracle/region_oracle.c:o_mask_targets (bit-exact).
__global__ void __launch_bounds__(256)
mask_target_kernel(const uint8_t *__restrict__ masks, int B, int G, int H, int W, const float *__restrict__ rois5,
                   const int32_t *__restrict__ gt_idx, int R, int M, const float *__restrict__ cfg, uint8_t *__restrict__ out)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (int64_t)R * M * M) return;
    const int S = (int)__ldg(cfg);
    const int r = (int)(i / (M * M)), pq = (int)(i - (int64_t)r * M * M), ph = pq / M, pw = pq - ph * M;
    const float *roi = rois5 + (int64_t)r * 5;
    const int b = (int)__ldg(roi), g = __ldg(gt_idx + r);
    uint8_t res = 0;
    if (g >= 0 && g < G && b >= 0 && b < B) {
        const uint8_t *plane = masks + ((int64_t)b * G + g) * H * W;
        const float sw = mul(__ldg(roi + 1), 1.0f), sh = mul(__ldg(roi + 2), 1.0f);
        const float ew = mul(add(__ldg(roi + 3), 0.0f), 1.0f), eh = mul(add(__ldg(roi + 4), 0.0f), 1.0f);
        const float rw = fmaxf(sub(ew, sw), 1.0f), rh = fmaxf(sub(eh, sh), 1.0f);
        const float bw = div(rw, (float)M), bh = div(rh, (float)M);
        float sum = 0.0f;
        for (int iy = 0; iy < S; iy++)
            for (int ix = 0; ix < S; ix++) {
                float y = sample_coord(sh, bh, ph, iy, S), x = sample_coord(sw, bw, pw, ix, S);
                if (y < -1.0f || y > (float)H || x < -1.0f || x > (float)W) continue;
                const Tap t = make_tap(y, x, H, W);
                float v = add(mul(t.w1, (float)plane[t.o1]), mul(t.w2, (float)plane[t.o2]));
                v = add(v, mul(t.w3, (float)plane[t.o3]));
                v = add(v, mul(t.w4, (float)plane[t.o4]));
                sum = add(sum, v);
            }
        res = div(sum, (float)(S * S)) >= 0.5f ? 1 : 0;
    }
    out[i] = res;
}

cudaError_t launch_mask_targets(const uint8_t *masks, int B, int G, int H, int W, const float *rois5, const int32_t *gt_idx,
                                int R, int M, const float *cfg, uint8_t *out, cudaStream_t s)
{
    const int64_t n = (int64_t)R * M * M;
    if (n == 0) return cudaSuccess;
    mask_target_kernel<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(masks, B, G, H, W, rois5, gt_idx, R, M, cfg, out);
    return cudaGetLastError();
}

static RoiFeat to_roifeat(const FeatSet &fs, const float *cfg)
{
    RoiFeat f{};
    f.L = fs.L; f.B = fs.B; f.C = fs.C; f.cfg = cfg;
    for (int l = 0; l < fs.L; l++) { f.H[l] = fs.H[l]; f.W[l] = fs.W[l]; f.feat[l] = fs.feat[l]; }
    return f;
}

cudaError_t launch_roialign_fwd_tma(const FeatSet &fs, const RoiFeat &f, const float *rois5, int R, int P,
                                    float *out, int32_t *fallback_flag, cudaStream_t s, bool *launched);
cudaError_t launch_roialign_bwd_tma(const FeatSet &fs, const RoiFeat &f, const float *rois5, int R, int P,
                                    const float *dout, int32_t *fallback_flag, cudaStream_t s, bool *launched);
// roialign_ch.cu: channel-per-lane kernels (7x7, S = 2, C % 32 == 0); RoIs they decline are flagged for the gather kernels
cudaError_t launch_roialign_fwd_ch(const FeatSet &fs, const RoiFeat &f, const float *rois5, int R, int P, float *out,
                                   int32_t *fallback_flag, void *plan_ws, cudaStream_t s, bool *launched);
cudaError_t launch_roialign_bwd_ch(const FeatSet &fs, const RoiFeat &f, const float *rois5, int R, int P, const float *dout,
                                   int32_t *fallback_flag, void *plan_ws, cudaStream_t s, bool *launched);
size_t roialign_ch_workspace_bytes(int R);
// roialign_tile.cu: tile-stationary backward (7x7, S = 2, C % 32 == 0): every dX byte written once, no zero-fill
cudaError_t launch_roialign_bwd_tile(const FeatSet &fs, const RoiFeat &f, const float *rois5, int R, int P, const float *dout,
                                     void *tile_ws, bool accumulate, cudaStream_t s, bool *launched, const int32_t **flags,
                                     const int32_t **ndecl);
cudaError_t roialign_tile_prepare(const FeatSet &fs, const RoiFeat &f, const float *rois5, int R, void *buf, cudaStream_t s);
cudaError_t roialign_tile_run(const FeatSet &fs, int R, const float *dout, void *buf, bool accumulate, bool rearm, cudaStream_t s,
                              const int32_t **flags, const int32_t **ndecl);
bool roialign_tile_supports(const FeatSet &fs, int R, int P);
size_t roialign_tile_workspace_bytes(const FeatSet &fs, int R);

// flags (R ints, padded) + the channel-lane kernels' per-RoI plans
static size_t roi_flag_bytes(int R) { return ((size_t)(R > 0 ? R : 0) * sizeof(int32_t) + 255) & ~(size_t)255; }
size_t roialign_workspace_bytes(int R) { return roi_flag_bytes(R) + roialign_ch_workspace_bytes(R) + 256; }
size_t roialign_bwd_workspace_bytes(const FeatSet &fs, int R) { return roialign_workspace_bytes(R) + roialign_tile_workspace_bytes(fs, R); }

// mode: 0 = TMA separable kernel + gather for the RoIs it declines (default); 1 = gather only (bit-exact fwd)
cudaError_t launch_roialign_fwd(const FeatSet &fs, const float *rois5, int R, int P, const float *cfg,
                                float *out, void *ws, int *ctl, int mode, cudaStream_t s)
{
    if (R == 0) return cudaSuccess;
    if (P * P * 4 > kRoiMaxTaps || fs.L > kMaxLv) return cudaErrorInvalidValue;
    const RoiFeat f = to_roifeat(fs, cfg);
    int32_t *flags = reinterpret_cast<int32_t *>(ws);
    bool tma = false;
    if (mode == 0 && flags) {
        cudaError_t e = cudaSuccess;
        e = launch_roialign_fwd_ch(fs, f, rois5, R, P, out, flags, reinterpret_cast<unsigned char *>(ws) + roi_flag_bytes(R), s, &tma);
        if (e != cudaSuccess) return e;
        if (!tma) e = launch_roialign_fwd_tma(fs, f, rois5, R, P, out, flags, s, &tma);
        if (e != cudaSuccess) return e;
    }
    const int csplit = fs.C >= 64 ? 4 : 1;
    return launch_pdl(roialign_fwd_gather_kernel, dim3(tma ? (R < 148 ? R : 148) : R, csplit), dim3(kRoiThreads), 0, s, f, rois5, R, P, csplit, out,
                      tma ? (const int32_t *)flags : (const int32_t *)nullptr);
}

cudaError_t launch_roialign_bwd(const FeatSet &fs, const float *rois5, int R, int P, const float *cfg,
                                const float *dout, void *ws, int *ctl, int mode, cudaStream_t s, bool accumulate)
{
    if (P * P * 4 > kRoiMaxTaps || fs.L > kMaxLv) return cudaErrorInvalidValue;
    const int csplit = fs.C >= 64 ? 4 : 1;
    // the tile-stationary kernel writes every dX byte once: no zero-fill, no reduce-adds.  MdRoiAlignBwdAcc (accumulate into the
    // caller's tensors) keeps the scatter-add kernels unless MD_ROI_TILE_ACC=1: adding means reading all of dX back.
    const char *tile_acc_env = getenv("MD_ROI_TILE_ACC");
    const bool tile_acc = tile_acc_env && atoi(tile_acc_env) != 0;
    if (R > 0 && mode == 0 && ws && (!accumulate || tile_acc)) {
        const RoiFeat f = to_roifeat(fs, cfg);
        bool tile = false;
        const int32_t *flags = nullptr, *ndecl = nullptr;
        cudaError_t e = launch_roialign_bwd_tile(fs, f, rois5, R, P, dout, reinterpret_cast<unsigned char *>(ws) + roialign_workspace_bytes(R),
                                                 accumulate, s, &tile, &flags, &ndecl);
        if (e != cudaSuccess) return e;
        if (tile)
            return launch_pdl(roialign_bwd_gather_kernel, dim3(R < 148 ? R : 148, csplit), dim3(kRoiThreads), 0, s, f, rois5, R, P, csplit, dout,
                              flags, ndecl);
    }
    for (int l = 0; l < fs.L && !accumulate; l++) {
        cudaError_t e = cudaMemsetAsync(fs.feat[l], 0, (size_t)fs.B * fs.C * fs.H[l] * fs.W[l] * sizeof(float), s);
        if (e != cudaSuccess) return e;
    }
    if (R == 0) return cudaSuccess;
    const RoiFeat f = to_roifeat(fs, cfg);
    int32_t *flags = reinterpret_cast<int32_t *>(ws);
    bool tma = false;
    if (mode == 0 && flags) {
        cudaError_t e = cudaSuccess;
        e = launch_roialign_bwd_ch(fs, f, rois5, R, P, dout, flags, reinterpret_cast<unsigned char *>(ws) + roi_flag_bytes(R), s, &tma);
        if (e != cudaSuccess) return e;
        if (!tma) e = launch_roialign_bwd_tma(fs, f, rois5, R, P, dout, flags, s, &tma);
        if (e != cudaSuccess) return e;
    }
    return launch_pdl(roialign_bwd_gather_kernel, dim3(tma ? (R < 148 ? R : 148) : R, csplit), dim3(kRoiThreads), 0, s, f, rois5, R, P, csplit, dout,
                      tma ? (const int32_t *)flags : (const int32_t *)nullptr, (const int32_t *)nullptr);
}


// ---- two-op form of the tile-stationary backward: the plan depends on the RoIs only, so a graph can run it beside the forward --------
size_t roialign_plan_bytes(const FeatSet &fs, int R) { return roialign_tile_workspace_bytes(fs, R); }

cudaError_t launch_roialign_bwd_prepare(const FeatSet &fs, const float *rois5, int R, int P, const float *cfg, void *plan, cudaStream_t s)
{
    if (!roialign_tile_supports(fs, R, P)) return cudaErrorInvalidValue;
    const RoiFeat f = to_roifeat(fs, cfg);
    return roialign_tile_prepare(fs, f, rois5, R, plan, s);
}

cudaError_t launch_roialign_bwd_planned(const FeatSet &fs, const float *rois5, int R, int P, const float *cfg, const float *dout, void *plan,
                                        cudaStream_t s)
{
    if (!roialign_tile_supports(fs, R, P) || (reinterpret_cast<uintptr_t>(dout) & 15) != 0) return cudaErrorInvalidValue;
    const RoiFeat f = to_roifeat(fs, cfg);
    const int32_t *flags = nullptr, *ndecl = nullptr;
    cudaError_t e = roialign_tile_run(fs, R, dout, plan, false, true, s, &flags, &ndecl);
    if (e != cudaSuccess) return e;
    const int csplit = fs.C >= 64 ? 4 : 1;
    return launch_pdl(roialign_bwd_gather_kernel, dim3(R < 148 ? R : 148, csplit), dim3(kRoiThreads), 0, s, f, rois5, R, P, csplit, dout, flags,
                      ndecl);
}

}  // namespace md
