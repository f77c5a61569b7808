// roialign_ch.cu -- a10/a11 fast path for 7x7 / S = 2: CHANNEL-PER-LANE RoIAlign over TMA-staged slabs.
//
// RoIAlign is the same small linear operator for every channel of a RoI: Out_c = Ay . F_c . Ax^T, with the bilinear
// sample matrices Ay (7 x rows) and Ax (7 x cols) shared by all channels.  A warp therefore takes 32 CHANNELS of one
// RoI, one per lane: every row / column index, weight, loop bound and branch is warp-uniform (no idle lanes, no
// divergence, no per-lane index arithmetic), and the only per-lane quantity is the channel's base address.
//
//   plan     one thread per RoI (ch_plan_kernel) writes a ChPlan: the 7 bins' row taps (duplicates merged), the 28 column
//            taps, and the STAGES the RoI is cut into.  A stage = (bin range, column-quad range, feature rows y0..y0+rows)
//            that fits one shared-memory slab; a compact RoI is one stage, a tall one several (only the rows its bins
//            touch are ever staged), a wide one is cut into quad ranges bin by bin.
//   staging  tensor maps view a level as {W, B*C, H}; the box {4*kn columns, 32 channels, R rows} lands in shared memory
//            as [row][channel][4*kn]: lane c reads its channel with LDS.128.  With a channel pitch of 16*kn bytes a
//            quarter warp is bank-conflict-free when kn is odd; for even kn every lane walks the quads in a rotated
//            order (ch_rot), which restores it without padding.
//   forward  per bin: y-step U[x] = sum_taps wy . F[row][x] (<= 4 merged row taps, LDS.128), U parked in a per-warp
//            scratch [column][lane]; x-step Out[p][q] = 4 column taps read back from the scratch (dynamic column ->
//            plain shared-memory address, conflict-free).  The 32 x 49 outputs leave as ONE 6272-byte bulk store.
//   backward the transpose: dY arrives by one bulk load per item; T[x] = sum_q dY[p][q] Ax[q][x] (dense Ax in shared
//            memory), D[row][x] += wy . T[x] into the zeroed slab; the slab is folded into dX by
//            cp.reduce.async.bulk.tensor (.add.f32, performed at L2).
//   schedule persistent warps, items (RoI, 32-channel group) dealt round-robin in (image, channel group, RoI) order so the
//            warps resident at any time read the same 32 planes of one image out of L2; every warp keeps kChSlots slabs
//            in flight across item boundaries.
//
// RoIs this file declines (S != 2, a level without a tensor map, footprint wider than 64 columns, a bin taller than 16
// rows, ...) are flagged and taken by the row-streaming kernels of roialign_tma.cu / the gather kernels in the same call.
//
// No reference code exists for this op; semantics: oracle/CONVENTIONS.md #14-16, checked against
// oracle/region_oracle.c:o_roialign_fwd / o_roialign_bwd (rtol 1e-5: separable summation order, FMAs).
#include <cuda.h>

#include <climits>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <type_traits>

#include "kernels.h"
#include "roialign_common.cuh"
#include "tma_host.h"
#include "tma_ptx.cuh"

namespace md {

#ifndef MD_CH_WARPS
#define MD_CH_WARPS 4            // forward: warps per CTA, one CTA per SM
#endif
#ifndef MD_CH_BWD_WARPS
#define MD_CH_BWD_WARPS 4        // backward
#endif
#ifndef MD_CH_SLOT_QUADS
#define MD_CH_SLOT_QUADS 32      // slab capacity in (row, column quad) units of 512 bytes
#endif
#ifndef MD_CH_MAX_NQ
#define MD_CH_MAX_NQ 16          // aligned footprint width <= 64 columns (wider RoIs take the in-kernel gather path)
#endif
#ifndef MD_CH_SLOTS
#define MD_CH_SLOTS 2
#endif
constexpr int kChP = 7;
constexpr int kChWarps = MD_CH_WARPS;
constexpr int kChBwdWarps = MD_CH_BWD_WARPS;
constexpr int kChSlots = MD_CH_SLOTS;
constexpr int kChSlotQuads = MD_CH_SLOT_QUADS;
constexpr int kChSlotBytes = kChSlotQuads * 512;
constexpr int kChMaxNq = MD_CH_MAX_NQ;
constexpr int kChMaxRows = 16;               // rows of one stage
constexpr int kChMaxStages = 16;
constexpr int kChLevels = 4;
constexpr int kChItemBytes = 32 * kChP * kChP * 4;      // 32 channels x 49 outputs = 6272
constexpr int kChPlanRing = 4;

enum { CH_OK = 0, CH_ZERO = 1, CH_DECLINE = 2, CH_GATHER = 3 };   // GATHER: taken by the kernels' own per-lane gather path

struct __align__(16) ChBin { int row[4]; float w[4]; };          // merged row taps, rows relative to the stage's y0; w == 0: unused
struct ChStage { int y0; unsigned char rows, p0, p1, k0, kn, pad[3]; };   // bins [p0, p1), quads [k0, k0 + kn), rows y0 .. y0 + rows
struct __align__(16) ChPlan {
    ChBin bin[kChP];                 // 224
    int xoff[4 * kChP];              // column taps: column * 128 (byte offset of the column in the [column][lane] scratch)
    float xw[4 * kChP];
    ChStage stage[kChMaxStages];     // 192
    float axd[kChP][16];             // dense Ax (1/S folded in) of a RoI with nq <= 4: the register-resident x-step
    int status, b, l, x_lo, nq, nst, dense, pad1;
};
static_assert(sizeof(ChStage) == 12, "stage layout");
static_assert(sizeof(ChPlan) % 16 == 0, "plans are copied 16 bytes at a time");
constexpr int kChPlanBytes = (int)sizeof(ChPlan);

struct ChMaps { CUtensorMap m[kChLevels * kChMaxNq * 2]; };      // [level][kn - 1][R == 4]: box {4 kn, 32, 1 or 4}

MD_DEVINL float4 lds128f(uint32_t a)
{
    float4 v;
    asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(a));
    return v;
}
MD_DEVINL float lds32f(uint32_t a) { float v; asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(a)); return v; }
MD_DEVINL void sts32f(uint32_t a, float v) { asm volatile("st.shared.f32 [%0], %1;" :: "r"(a), "f"(v) : "memory"); }
MD_DEVINL void sts128f(uint32_t a, float4 v)
{
    asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" :: "r"(a), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}
MD_DEVINL void cp_async16(uint32_t sdst, const void *gsrc)
{
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" :: "r"(sdst), "l"(gsrc) : "memory");
}

// quad rotation that makes lane-per-channel LDS.128 conflict-free at a channel pitch of 16*kn bytes: lanes whose channel
// offsets share a 16-byte slot (gcd(kn, 8) of every 8) start at different quads
MD_DEVINL int ch_rot(int kn, int lane)
{
    int g = kn & -kn;
    g = g > 8 ? 8 : g;
    return ((lane & 7) * g) >> 3;
}

// =====================================================================================================
// plan
// =====================================================================================================
__global__ void __launch_bounds__(128)
ch_plan_kernel(const RoiFeat f, const int tma_mask, const float *__restrict__ rois5, const int R, const int slot_quads,
               ChPlan *__restrict__ plans, int32_t *__restrict__ flag)
{
    constexpr int P = kChP, S = 2, NS = P * S;
    const int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= R) return;
    ChPlan &pl = plans[r];
    const RoiGeom g = roi_geometry(f, rois5 + (int64_t)r * 5, P);
    int status = CH_OK;
    if ((int)__ldg(f.cfg + 1) != S) status = CH_DECLINE;                                  // the gather kernels take it
    else if (!g.ok) status = CH_ZERO;
    else if (g.l >= kChLevels || !((tma_mask >> g.l) & 1)) status = CH_GATHER;            // level without a tensor map

    int nq = 0, x_lo = 0, nst = 0;
    if (status == CH_OK) {
        // ---- columns ----
        int lo[NS], hi[NS];
        float wl[NS], wh[NS];
        bool ok[NS];
        int mn = INT_MAX, mx = -1;
#pragma unroll
        for (int i = 0; i < NS; i++) {
            const float v = sample_coord(g.sw, g.bw, i / S, i % S, S);
            lo[i] = hi[i] = 0; wl[i] = wh[i] = 0.0f;
            ok[i] = sample_1d(v, g.W, lo[i], hi[i], wl[i], wh[i]);
            if (ok[i]) { mn = min(mn, lo[i]); mx = max(mx, hi[i]); }
        }
        if (mx < 0) status = CH_ZERO;
        else {
            x_lo = mn & ~3;                       // the innermost TMA coordinate stays 16-byte aligned
            nq = (mx - x_lo + 4) >> 2;
            if (nq > kChMaxNq) status = CH_GATHER;
#pragma unroll
            for (int i = 0; i < NS; i++) {
                pl.xoff[2 * i] = ok[i] ? (lo[i] - x_lo) * 128 : 0;
                pl.xoff[2 * i + 1] = ok[i] ? (hi[i] - x_lo) * 128 : 0;
                pl.xw[2 * i] = ok[i] ? mul(wl[i], 0.5f) : 0.0f;
                pl.xw[2 * i + 1] = ok[i] ? mul(wh[i], 0.5f) : 0.0f;
            }
        }
    }
    if (status == CH_OK) {
        // ---- rows: per-bin (row, weight) taps, duplicates merged, zero weights dropped ----
        int trow[P][4], nt[P], rmin[P], rmax[P];
        float tw[P][4];
        bool any = false;
#pragma unroll
        for (int p = 0; p < P; p++) {
            nt[p] = 0; rmin[p] = INT_MAX; rmax[p] = -1;
#pragma unroll
            for (int i = 0; i < S; i++) {
                const float v = sample_coord(g.sh, g.bh, p, i, S);
                int lo = 0, hi = 0;
                float wl = 0.0f, wh = 0.0f;
                if (!sample_1d(v, g.H, lo, hi, wl, wh)) continue;
                any = true;
                rmin[p] = min(rmin[p], lo); rmax[p] = max(rmax[p], hi);
                const int rr[2] = { lo, hi };
                const float ww[2] = { mul(wl, 0.5f), mul(wh, 0.5f) };
#pragma unroll
                for (int k = 0; k < 2; k++) {
                    if (ww[k] == 0.0f) continue;
                    bool merged = false;
                    for (int t = 0; t < nt[p]; t++)
                        if (trow[p][t] == rr[k]) { tw[p][t] = add(tw[p][t], ww[k]); merged = true; }
                    if (!merged) { trow[p][nt[p]] = rr[k]; tw[p][nt[p]] = ww[k]; nt[p]++; }
                }
            }
        }
        if (!any) status = CH_ZERO;
        else {
            // bins without a valid sample borrow the row range of a neighbour (they contribute nothing)
            int last = -1;
#pragma unroll
            for (int p = 0; p < P; p++) {
                if (rmax[p] >= 0) last = rmax[p];
                else if (last >= 0) { rmin[p] = rmax[p] = last; }
            }
            int nxt = -1;
#pragma unroll
            for (int p = P - 1; p >= 0; p--) {
                if (rmax[p] >= 0) nxt = rmin[p];
                else { rmin[p] = rmax[p] = nxt; }
            }
            // ---- stages ----
            int y0_of[P];
            int p = 0;
            while (p < P && status == CH_OK) {
                const int span = rmax[p] - rmin[p] + 1;
                if (span > kChMaxRows) { status = CH_GATHER; break; }
                if (span * nq <= slot_quads) {
                    const int y0 = rmin[p];
                    int y1 = rmax[p], p1 = p + 1;
                    while (p1 < P) {
                        const int rows = max(y1, rmax[p1]) - y0 + 1;
                        if (rows > kChMaxRows || rows * nq > slot_quads) break;
                        y1 = max(y1, rmax[p1]); p1++;
                    }
                    if (nst >= kChMaxStages) { status = CH_GATHER; break; }
                    ChStage st;
                    st.y0 = y0; st.rows = (unsigned char)(y1 - y0 + 1); st.p0 = (unsigned char)p; st.p1 = (unsigned char)p1;
                    st.k0 = 0; st.kn = (unsigned char)nq; st.pad[0] = st.pad[1] = st.pad[2] = 0;
                    pl.stage[nst++] = st;
                    for (int q = p; q < p1; q++) y0_of[q] = y0;
                    p = p1;
                } else {
                    const int kmax = slot_quads / span;
                    if (kmax < 1) { status = CH_GATHER; break; }
                    for (int k0 = 0; k0 < nq; k0 += kmax) {
                        if (nst >= kChMaxStages) { status = CH_GATHER; break; }
                        ChStage st;
                        st.y0 = rmin[p]; st.rows = (unsigned char)span; st.p0 = (unsigned char)p; st.p1 = (unsigned char)(p + 1);
                        st.k0 = (unsigned char)k0; st.kn = (unsigned char)min(kmax, nq - k0); st.pad[0] = st.pad[1] = st.pad[2] = 0;
                        pl.stage[nst++] = st;
                    }
                    y0_of[p] = rmin[p];
                    p++;
                }
            }
            if (status == CH_OK) {
#pragma unroll
                for (int q = 0; q < P; q++) {
                    ChBin b;
#pragma unroll
                    for (int t = 0; t < 4; t++) {
                        const bool on = t < nt[q];
                        b.row[t] = on ? trow[q][t] - y0_of[q] : 0;
                        b.w[t] = on ? tw[q][t] : 0.0f;
                    }
                    pl.bin[q] = b;
                }
            }
        }
    }
    // register-resident x-step: footprint at most 16 columns wide and every stage full width
    int dense = status == CH_OK && nq <= 4;
    for (int k = 0; dense && k < nst; k++) dense = pl.stage[k].kn == nq;
    if (dense) {
        float ax[P][16];
#pragma unroll
        for (int q = 0; q < P; q++)
#pragma unroll
            for (int x = 0; x < 16; x++) ax[q][x] = 0.0f;
#pragma unroll
        for (int i = 0; i < 4 * P; i++) {
            const float w = pl.xw[i];
            if (w != 0.0f) ax[i >> 2][pl.xoff[i] >> 7] = add(ax[i >> 2][pl.xoff[i] >> 7], w);
        }
#pragma unroll
        for (int q = 0; q < P; q++)
#pragma unroll
            for (int x = 0; x < 16; x++) pl.axd[q][x] = ax[q][x];
    }
    pl.status = status; pl.b = g.b; pl.l = g.l; pl.x_lo = x_lo; pl.nq = nq; pl.nst = status == CH_OK ? nst : 0;
    pl.dense = dense; pl.pad1 = 0;
    flag[r] = status == CH_DECLINE ? 1 : 0;
}

// =====================================================================================================
// shared pieces of the two persistent kernels
// =====================================================================================================
struct ChItemId { int r, g; };
// item t -> (RoI, channel group): (segment of `seg` RoIs, group, RoI in segment) order, see work_item()
MD_DEVINL ChItemId ch_item(int t, int R, int seg, int ngroups)
{
    ChItemId id;
    if (seg == 512 && ngroups == 8 && (R & 511) == 0) {       // config-2 shape: no divisions
        id.r = ((t >> 12) << 9) | (t & 511); id.g = (t >> 9) & 7;
        return id;
    }
    const WorkItem w = work_item(t, R, seg, ngroups);
    id.r = w.r; id.g = w.chunk;
    return id;
}

// per-lane gather of one (RoI, channel) -- the path of RoIs that do not fit the staged kernels (wide / tall footprints,
// levels without a tensor map).  Arithmetic and order as roialign_fwd_gather_kernel.
__device__ __noinline__ void ch_fwd_gather_item(const RoiFeat &f, const float *__restrict__ rois5, int r, int c, uint32_t ob_lane)
{
    constexpr int P = kChP, S = 2, PW = 4;
    const RoiGeom g = roi_geometry(f, rois5 + (int64_t)r * 5, P);
    const float *fp = f.feat[g.l] + ((int64_t)g.b * f.C + c) * g.H * g.W;
#pragma unroll 1
    for (int it = 0; it < 2 * P; it++) {               // (output row, half a row): 64 loads in flight
        const int ph = it >> 1, pw0 = (it & 1) * PW;
        float v[PW][S * S][4], w[PW][S * S][4];
#pragma unroll
        for (int j = 0; j < PW; j++)
#pragma unroll
            for (int k = 0; k < S * S; k++) {
                const float y = sample_coord(g.sh, g.bh, ph, k >> 1, S), x = sample_coord(g.sw, g.bw, min(pw0 + j, P - 1), k & 1, S);
                const Tap t = make_tap(y, x, g.H, g.W);               // all-zero weights and offset 0 when out of range
                v[j][k][0] = __ldg(fp + t.o1); v[j][k][1] = __ldg(fp + t.o2); v[j][k][2] = __ldg(fp + t.o3); v[j][k][3] = __ldg(fp + t.o4);
                w[j][k][0] = t.w1; w[j][k][1] = t.w2; w[j][k][2] = t.w3; w[j][k][3] = t.w4;
            }
#pragma unroll
        for (int j = 0; j < PW; j++) {
            float sum = 0.0f;
#pragma unroll
            for (int k = 0; k < S * S; k++) {
                float a = add(mul(w[j][k][0], v[j][k][0]), mul(w[j][k][1], v[j][k][1]));
                a = add(a, mul(w[j][k][2], v[j][k][2]));
                a = add(a, mul(w[j][k][3], v[j][k][3]));
                sum = add(sum, a);
            }
            if (pw0 + j < P) sts32f(ob_lane + (ph * P + pw0 + j) * 4, mul(sum, 0.25f));
        }
    }
}
__device__ __noinline__ void ch_bwd_gather_item(const RoiFeat &f, const float *__restrict__ rois5, int r, int c, uint32_t dy_lane)
{
    constexpr int P = kChP, S = 2;
    const RoiGeom g = roi_geometry(f, rois5 + (int64_t)r * 5, P);
    float *fp = f.feat[g.l] + ((int64_t)g.b * f.C + c) * g.H * g.W;
#pragma unroll 1
    for (int bin = 0; bin < P * P; bin++) {
        const int ph = bin / P, pw = bin - ph * P;
        const float gr = mul(lds32f(dy_lane + bin * 4), 0.25f);
#pragma unroll
        for (int k = 0; k < S * S; k++) {
            const float y = sample_coord(g.sh, g.bh, ph, k >> 1, S), x = sample_coord(g.sw, g.bw, pw, k & 1, S);
            if (y < -1.0f || y > (float)g.H || x < -1.0f || x > (float)g.W) continue;
            const Tap t = make_tap(y, x, g.H, g.W);
            if (t.w1 != 0.0f) atomicAdd(fp + t.o1, mul(gr, t.w1));
            if (t.w2 != 0.0f) atomicAdd(fp + t.o2, mul(gr, t.w2));
            if (t.w3 != 0.0f) atomicAdd(fp + t.o3, mul(gr, t.w3));
            if (t.w4 != 0.0f) atomicAdd(fp + t.o4, mul(gr, t.w4));
        }
    }
}

// copy one plan into the ring (16 bytes per lane and step); the caller commits / waits the cp.async group
MD_DEVINL void ch_prefetch_plan(uint32_t dst, const ChPlan *src, int lane)
{
    const char *s = reinterpret_cast<const char *>(src);
    for (int o = lane * 16; o < kChPlanBytes; o += 512) cp_async16(dst + o, s + o);
}

// =====================================================================================================
// forward / backward building blocks
// =====================================================================================================
#ifndef MD_CH_DYBUFS
#define MD_CH_DYBUFS 2
#endif
constexpr int kChDyBufs = MD_CH_DYBUFS;
constexpr int kChAxPitch = 4 * kChMaxNq * 4;                                       // bytes per Ax row of the wide path
constexpr int kChAxBytes = kChP * kChAxPitch;                                      // dense Ax[7][4 * max nq], built per wide item
constexpr int kChFwdWarpBytes = ((kChSlots * kChSlotBytes + kChItemBytes + kChAxBytes + kChPlanRing * kChPlanBytes + 64) + 127) & ~127;
constexpr int kChBwdWarpBytes = ((kChSlots * kChSlotBytes + kChDyBufs * kChItemBytes + kChAxBytes + kChPlanRing * kChPlanBytes + 64) + 127) & ~127;
static_assert(kChWarps * kChFwdWarpBytes <= 232448 && kChBwdWarps * kChBwdWarpBytes <= 232448, "227 KB of shared memory per CTA");

// TMA wrappers on shared-memory addresses (tensor dims are {W, B*C, H}: coordinates {x, plane, row})
MD_DEVINL void ch_tma_load(uint32_t dst, const CUtensorMap *map, int x, int z, int y, uint32_t bar)
{
    asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
                 :: "r"(dst), "l"(map), "r"(x), "r"(z), "r"(y), "r"(bar) : "memory");
}
MD_DEVINL void ch_tma_reduce_add(const CUtensorMap *map, int x, int z, int y, uint32_t src)
{
    asm volatile("cp.reduce.async.bulk.tensor.3d.global.shared::cta.add.tile.bulk_group [%0, {%2, %3, %4}], [%1];"
                 :: "l"(map), "r"(src), "r"(x), "r"(z), "r"(y) : "memory");
}

// dense Ax[7][4 nq] of a wide item, from its 28 column taps (whole warp)
MD_DEVINL void ch_build_ax(float *AxS, const ChPlan &pl, int lane)
{
    for (int i = lane; i < kChP * 4 * kChMaxNq; i += 32) AxS[i] = 0.0f;
    __syncwarp();
    if (lane < 4 * kChP) {
        const float w = pl.xw[lane];
        if (w != 0.0f) atomicAdd(&AxS[(lane >> 2) * 4 * kChMaxNq + (pl.xoff[lane] >> 7)], w);
    }
    __syncwarp();
}

// y-step of one bin for GW quads of the lane's channel: u[j] = sum over the (<= 4, merged) row taps.  An unused tap has
// weight 0 and row 0 of the slab; all loads are issued before the first FMA.
template <int GW>
MD_DEVINL void ch_ystep(const uint32_t a0, const uint32_t a1, const uint32_t a2, const uint32_t a3, const float4 ww,
                        const int (&qo)[GW], float4 (&u)[GW])
{
    float4 v0[GW], v1[GW], v2[GW], v3[GW];
#pragma unroll
    for (int j = 0; j < GW; j++) {
        v0[j] = lds128f(a0 + qo[j]); v1[j] = lds128f(a1 + qo[j]); v2[j] = lds128f(a2 + qo[j]); v3[j] = lds128f(a3 + qo[j]);
    }
#pragma unroll
    for (int j = 0; j < GW; j++) {
        u[j].x = __fmaf_rn(ww.w, v3[j].x, __fmaf_rn(ww.z, v2[j].x, __fmaf_rn(ww.y, v1[j].x, mul(ww.x, v0[j].x))));
        u[j].y = __fmaf_rn(ww.w, v3[j].y, __fmaf_rn(ww.z, v2[j].y, __fmaf_rn(ww.y, v1[j].y, mul(ww.x, v0[j].y))));
        u[j].z = __fmaf_rn(ww.w, v3[j].z, __fmaf_rn(ww.z, v2[j].z, __fmaf_rn(ww.y, v1[j].z, mul(ww.x, v0[j].z))));
        u[j].w = __fmaf_rn(ww.w, v3[j].w, __fmaf_rn(ww.z, v2[j].w, __fmaf_rn(ww.y, v1[j].w, mul(ww.x, v0[j].w))));
    }
}

// wide path, one group of GW quads (j0 .. j0 + GW of the chunk's rotated order): y-step, then acc[q] += Ax[q][cols] . u with
// Ax read from shared memory
template <int GW>
MD_DEVINL void ch_fwd_wide_group(const uint32_t a0, const uint32_t a1, const uint32_t a2, const uint32_t a3, const float4 ww,
                                 const int j0, const int kn, const int rot, const uint32_t axk, float (&acc)[kChP])
{
    int qo[GW];
#pragma unroll
    for (int j = 0; j < GW; j++) {
        int q = j0 + j + rot;
        if (q >= kn) q -= kn;
        qo[j] = 16 * q;
    }
    float4 u[GW];
    ch_ystep<GW>(a0, a1, a2, a3, ww, qo, u);
#pragma unroll
    for (int px = 0; px < kChP; px++)
#pragma unroll
        for (int j = 0; j < GW; j++) {
            const float4 a = lds128f(axk + px * kChAxPitch + qo[j]);
            acc[px] = __fmaf_rn(a.w, u[j].w, __fmaf_rn(a.z, u[j].z, __fmaf_rn(a.y, u[j].y, __fmaf_rn(a.x, u[j].x, acc[px]))));
        }
}

// backward, one group of GW quads: T = dY[p][.] Ax (Ax from shared memory), then the slab rows of the active taps += wy . T.
// Active taps of a bin are distinct rows (the plan merges duplicates), so every load may precede every store.
template <int GW, bool AXREG>
MD_DEVINL void ch_bwd_group(const uint32_t a0, const uint32_t a1, const uint32_t a2, const uint32_t a3, const float4 ww,
                            const int (&qo)[GW], const float (&dy)[kChP], const uint32_t axk, const float (*axr)[4 * GW])
{
    float4 t[GW], v0[GW], v1[GW], v2[GW], v3[GW];
#pragma unroll
    for (int j = 0; j < GW; j++) {
        v0[j] = lds128f(a0 + qo[j]); v1[j] = lds128f(a1 + qo[j]); v2[j] = lds128f(a2 + qo[j]); v3[j] = lds128f(a3 + qo[j]);
        t[j] = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
    }
#pragma unroll
    for (int px = 0; px < kChP; px++)
#pragma unroll
        for (int j = 0; j < GW; j++) {
            float4 a;
            if (AXREG) a = make_float4(axr[px][4 * j], axr[px][4 * j + 1], axr[px][4 * j + 2], axr[px][4 * j + 3]);
            else a = lds128f(axk + px * kChAxPitch + qo[j]);
            t[j].x = __fmaf_rn(dy[px], a.x, t[j].x); t[j].y = __fmaf_rn(dy[px], a.y, t[j].y);
            t[j].z = __fmaf_rn(dy[px], a.z, t[j].z); t[j].w = __fmaf_rn(dy[px], a.w, t[j].w);
        }
    const bool t0 = ww.x != 0.0f, t1 = ww.y != 0.0f, t2 = ww.z != 0.0f, t3 = ww.w != 0.0f;       // warp-uniform
#pragma unroll
    for (int j = 0; j < GW; j++) {
        if (t0) {
            v0[j].x = __fmaf_rn(ww.x, t[j].x, v0[j].x); v0[j].y = __fmaf_rn(ww.x, t[j].y, v0[j].y);
            v0[j].z = __fmaf_rn(ww.x, t[j].z, v0[j].z); v0[j].w = __fmaf_rn(ww.x, t[j].w, v0[j].w);
            sts128f(a0 + qo[j], v0[j]);
        }
        if (t1) {
            v1[j].x = __fmaf_rn(ww.y, t[j].x, v1[j].x); v1[j].y = __fmaf_rn(ww.y, t[j].y, v1[j].y);
            v1[j].z = __fmaf_rn(ww.y, t[j].z, v1[j].z); v1[j].w = __fmaf_rn(ww.y, t[j].w, v1[j].w);
            sts128f(a1 + qo[j], v1[j]);
        }
        if (t2) {
            v2[j].x = __fmaf_rn(ww.z, t[j].x, v2[j].x); v2[j].y = __fmaf_rn(ww.z, t[j].y, v2[j].y);
            v2[j].z = __fmaf_rn(ww.z, t[j].z, v2[j].z); v2[j].w = __fmaf_rn(ww.z, t[j].w, v2[j].w);
            sts128f(a2 + qo[j], v2[j]);
        }
        if (t3) {
            v3[j].x = __fmaf_rn(ww.w, t[j].x, v3[j].x); v3[j].y = __fmaf_rn(ww.w, t[j].y, v3[j].y);
            v3[j].z = __fmaf_rn(ww.w, t[j].z, v3[j].z); v3[j].w = __fmaf_rn(ww.w, t[j].w, v3[j].w);
            sts128f(a3 + qo[j], v3[j]);
        }
    }
}

// =====================================================================================================
// forward
// =====================================================================================================
__global__ void __launch_bounds__(kChWarps * 32, 1)
roialign_fwd_ch_kernel(const __grid_constant__ ChMaps maps, const RoiFeat f, const float *__restrict__ rois5,
                       const ChPlan *__restrict__ plans, const int R, const int C, const int seg, float *__restrict__ out)
{
    constexpr int P = kChP, PP = P * P;
    extern __shared__ __align__(128) unsigned char dsm[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    unsigned char *base = dsm + warp * kChFwdWarpBytes;
    const uint32_t slots = smem_u32(base);
    unsigned char *ob = base + kChSlots * kChSlotBytes;
    float *AxS = reinterpret_cast<float *>(base + kChSlots * kChSlotBytes + kChItemBytes);
    ChPlan *ring = reinterpret_cast<ChPlan *>(base + kChSlots * kChSlotBytes + kChItemBytes + kChAxBytes);
    unsigned long long *full = reinterpret_cast<unsigned long long *>(base + kChSlots * kChSlotBytes + kChItemBytes + kChAxBytes +
                                                                      kChPlanRing * kChPlanBytes);
    const int ngroups = C >> 5;
    const int total = R * ngroups;
    const int NW = gridDim.x * kChWarps, w0 = blockIdx.x * kChWarps + warp;
    if (w0 >= total) return;
    if (lane == 0) {
        for (int i = 0; i < kChSlots; i++) mbar_init(&full[i], 1);
        fence_barrier_init();
    }
    __syncwarp();
    auto item_of = [&](int k) { return w0 + k * NW; };
    auto plan_of = [&](int k) -> const ChPlan & { return ring[k & (kChPlanRing - 1)]; };
    // (RoI, channel group) of items ci, ci + 1, ci + 2: computed once per item
    ChItemId id0 = { 0, 0 }, id1 = { 0, 0 }, id2 = { 0, 0 };
    auto prefetch = [&](int k) {
        const int t = item_of(k);
        if (t < total) {
            id2 = ch_item(t, R, seg, ngroups);
            ch_prefetch_plan(smem_u32(&ring[k & (kChPlanRing - 1)]), plans + id2.r, lane);
        }
        cp_async_commit_group();
    };
    prefetch(0);
    id0 = id2;
    prefetch(1);
    id1 = id2;
    cp_async_wait_group<0>();
    __syncwarp();

    // ---- producer: stages are issued in order across items, at most kChSlots in flight, never beyond item ci + 1 ----
    int ci = 0, pi = 0, ps = 0, n_iss = 0, n_con = 0;
    auto try_issue = [&]() -> bool {
        for (;;) {
            if (pi > ci + 1 || item_of(pi) >= total) return false;
            if (ps < plan_of(pi).nst) break;
            pi++; ps = 0;
        }
        const ChPlan &p = plan_of(pi);
        const ChStage st = p.stage[ps];
        const int slot = n_iss % kChSlots;
        const int rows = st.rows, kn = st.kn;
        const uint32_t bar = smem_u32(&full[slot]);
        if (lane == 0) {
            mbar_expect_tx(&full[slot], (uint32_t)(rows * kn) * 512u);
            const CUtensorMap *m1 = &maps.m[(p.l * kChMaxNq + kn - 1) * 2], *m4 = m1 + 1;
            const int z = p.b * C + (pi == ci ? id0.g : id1.g) * 32, x = p.x_lo + 4 * st.k0;
            const uint32_t dst = slots + slot * kChSlotBytes;
            int row = 0;
#pragma unroll 1
            for (; row + 4 <= rows; row += 4) ch_tma_load(dst + row * kn * 512, m4, x, z, st.y0 + row, bar);
#pragma unroll 1
            for (; row < rows; row++) ch_tma_load(dst + row * kn * 512, m1, x, z, st.y0 + row, bar);
        }
        n_iss++; ps++;
        return true;
    };
    auto refill = [&]() { while (n_iss - n_con < kChSlots && try_issue()) {} };
    refill();
    const uint32_t ob_lane = smem_u32(ob) + lane * PP * 4;
    auto out_ready = [&]() {                       // the previous item's bulk store has read the staging buffer
        if (lane == 0) bulk_wait_read<0>();
        __syncwarp();
    };

#pragma unroll 1
    for (ci = 0; item_of(ci) < total; ci++, id0 = id1, id1 = id2) {
        prefetch(ci + 2);
        cp_async_wait_group<1>();                 // plan ci + 1 has landed
        __syncwarp();
        refill();
        const ChPlan &pl = plan_of(ci);
        const int status = pl.status;
        if (status == CH_DECLINE) continue;

        // ---- compact RoIs (<= 16 columns, full-width stages): Ax (7 x 4 NQ, permuted like this lane's quads) in registers ----
        auto dense_item = [&](auto tag) {
            constexpr int NQ = decltype(tag)::value;
            const int rot = ch_rot(NQ, lane);
            int qo[NQ];
#pragma unroll
            for (int j = 0; j < NQ; j++) {
                int q = j + rot;
                if (q >= NQ) q -= NQ;
                qo[j] = 16 * q;
            }
            float ax[P][4 * NQ];
            const uint32_t axb = smem_u32(&pl.axd[0][0]);
#pragma unroll
            for (int px = 0; px < P; px++)
#pragma unroll
                for (int j = 0; j < NQ; j++) {
                    const float4 a = lds128f(axb + px * 64 + qo[j]);
                    ax[px][4 * j] = a.x; ax[px][4 * j + 1] = a.y; ax[px][4 * j + 2] = a.z; ax[px][4 * j + 3] = a.w;
                }
            int s = 0, p1 = 0;
            bool have = false;
            uint32_t slab_lane = 0;
            constexpr int rowpitch = 512 * NQ;
#pragma unroll 1
            for (int py = 0; py < P; py++) {
                if (!have) {
                    p1 = pl.stage[s].p1;
                    mbar_wait(&full[n_con % kChSlots], (uint32_t)(n_con / kChSlots) & 1u);
                    slab_lane = slots + (n_con % kChSlots) * kChSlotBytes + lane * 16 * NQ;
                    have = true;
                }
                const int4 rr = *reinterpret_cast<const int4 *>(pl.bin[py].row);
                const float4 ww = *reinterpret_cast<const float4 *>(pl.bin[py].w);
                float4 u[NQ];
                ch_ystep<NQ>(slab_lane + rr.x * rowpitch, slab_lane + rr.y * rowpitch, slab_lane + rr.z * rowpitch,
                             slab_lane + rr.w * rowpitch, ww, qo, u);
                if (py + 1 == p1) {                        // last bin of the stage: hand the slab back
                    __syncwarp();
                    n_con++; s++; have = false;
                    refill();
                }
                if (py == 0) out_ready();
                float acc[P];
#pragma unroll
                for (int px = 0; px < P; px++) {
                    acc[px] = mul(ax[px][0], u[0].x);
                    acc[px] = __fmaf_rn(ax[px][1], u[0].y, acc[px]);
                    acc[px] = __fmaf_rn(ax[px][2], u[0].z, acc[px]);
                    acc[px] = __fmaf_rn(ax[px][3], u[0].w, acc[px]);
#pragma unroll
                    for (int j = 1; j < NQ; j++) {
                        acc[px] = __fmaf_rn(ax[px][4 * j], u[j].x, acc[px]);
                        acc[px] = __fmaf_rn(ax[px][4 * j + 1], u[j].y, acc[px]);
                        acc[px] = __fmaf_rn(ax[px][4 * j + 2], u[j].z, acc[px]);
                        acc[px] = __fmaf_rn(ax[px][4 * j + 3], u[j].w, acc[px]);
                    }
                }
                const uint32_t orow = ob_lane + py * P * 4;
#pragma unroll
                for (int px = 0; px < P; px++) sts32f(orow + px * 4, acc[px]);
            }
        };
        // ---- wide RoIs / chunked stages: quads in groups of four, Ax from shared memory, acc carried across chunks ----
        auto wide_item = [&]() {
            ch_build_ax(AxS, pl, lane);
            const uint32_t axb = smem_u32(AxS);
            const int nq = pl.nq;
            int s = 0, kn = 0, k0 = 0, p1 = 0, rot = 0;
            bool have = false;
            uint32_t slab_lane = 0;
#pragma unroll 1
            for (int py = 0; py < P; py++) {
                float acc[P];
#pragma unroll
                for (int px = 0; px < P; px++) acc[px] = 0.0f;
                const int4 rr = *reinterpret_cast<const int4 *>(pl.bin[py].row);
                const float4 ww = *reinterpret_cast<const float4 *>(pl.bin[py].w);
                bool more;
#pragma unroll 1
                do {
                    if (!have) {
                        const ChStage st = pl.stage[s];
                        kn = st.kn; k0 = st.k0; p1 = st.p1;
                        mbar_wait(&full[n_con % kChSlots], (uint32_t)(n_con / kChSlots) & 1u);
                        slab_lane = slots + (n_con % kChSlots) * kChSlotBytes + lane * 16 * kn;
                        rot = ch_rot(kn, lane);
                        have = true;
                    }
                    const int rowpitch = 512 * kn;
                    const uint32_t a0 = slab_lane + rr.x * rowpitch, a1 = slab_lane + rr.y * rowpitch;
                    const uint32_t a2 = slab_lane + rr.z * rowpitch, a3 = slab_lane + rr.w * rowpitch;
                    const uint32_t axk = axb + 16 * k0;
#pragma unroll 1
                    for (int j0 = 0; j0 < kn; j0 += 4) {
                        switch (min(4, kn - j0)) {
                            case 1: ch_fwd_wide_group<1>(a0, a1, a2, a3, ww, j0, kn, rot, axk, acc); break;
                            case 2: ch_fwd_wide_group<2>(a0, a1, a2, a3, ww, j0, kn, rot, axk, acc); break;
                            case 3: ch_fwd_wide_group<3>(a0, a1, a2, a3, ww, j0, kn, rot, axk, acc); break;
                            default: ch_fwd_wide_group<4>(a0, a1, a2, a3, ww, j0, kn, rot, axk, acc); break;
                        }
                    }
                    more = k0 + kn < nq;
                    if (py + 1 == p1) {                    // last bin of the stage: hand the slab back
                        __syncwarp();
                        n_con++; s++; have = false;
                        refill();
                    }
                } while (more);
                if (py == 0) out_ready();
                const uint32_t orow = ob_lane + py * P * 4;
#pragma unroll
                for (int px = 0; px < P; px++) sts32f(orow + px * 4, acc[px]);
            }
        };
        if (status == CH_OK) {
            if (pl.dense) {
                switch (pl.nq) {
                    case 1: dense_item(std::integral_constant<int, 1>()); break;
                    case 2: dense_item(std::integral_constant<int, 2>()); break;
                    case 3: dense_item(std::integral_constant<int, 3>()); break;
                    default: dense_item(std::integral_constant<int, 4>()); break;
                }
            } else wide_item();
        } else {
            out_ready();
            if (status == CH_GATHER) ch_fwd_gather_item(f, rois5, id0.r, id0.g * 32 + lane, ob_lane);
            else {
#pragma unroll 1
                for (int j = 0; j < PP; j++) sts32f(ob_lane + j * 4, 0.0f);
            }
        }
        // ---- 32 x 49 outputs: one bulk store ----
        fence_proxy_async();
        __syncwarp();
        if (lane == 0) {
            bulk_store_1d(out + ((int64_t)id0.r * C + id0.g * 32) * PP, ob, kChItemBytes);
            bulk_commit();
        }
    }
    if (lane == 0) bulk_wait_all<0>();
}

// =====================================================================================================
// backward
// =====================================================================================================
__global__ void __launch_bounds__(kChBwdWarps * 32, 1)
roialign_bwd_ch_kernel(const __grid_constant__ ChMaps maps, const RoiFeat f, const float *__restrict__ rois5,
                       const ChPlan *__restrict__ plans, const int R, const int C, const int seg, const float *__restrict__ dout)
{
    constexpr int P = kChP, PP = P * P;
    extern __shared__ __align__(128) unsigned char dsm[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    unsigned char *base = dsm + warp * kChBwdWarpBytes;
    const uint32_t slots = smem_u32(base);
    unsigned char *dybuf = base + kChSlots * kChSlotBytes;
    float *AxS = reinterpret_cast<float *>(base + kChSlots * kChSlotBytes + kChDyBufs * kChItemBytes);
    ChPlan *ring = reinterpret_cast<ChPlan *>(base + kChSlots * kChSlotBytes + kChDyBufs * kChItemBytes + kChAxBytes);
    unsigned long long *dyfull = reinterpret_cast<unsigned long long *>(base + kChSlots * kChSlotBytes + kChDyBufs * kChItemBytes + kChAxBytes +
                                                                        kChPlanRing * kChPlanBytes);
    const int ngroups = C >> 5;
    const int total = R * ngroups;
    const int NW = gridDim.x * kChBwdWarps, w0 = blockIdx.x * kChBwdWarps + warp;
    if (w0 >= total) return;
    if (lane == 0) {
        for (int i = 0; i < kChDyBufs; i++) mbar_init(&dyfull[i], 1);
        fence_barrier_init();
    }
    __syncwarp();
    auto item_of = [&](int k) { return w0 + k * NW; };
    auto plan_of = [&](int k) -> const ChPlan & { return ring[k & (kChPlanRing - 1)]; };
    ChItemId id0 = { 0, 0 }, id1 = { 0, 0 }, id2 = { 0, 0 };       // items ci, ci + 1, ci + 2
    auto prefetch = [&](int k) {
        const int t = item_of(k);
        if (t < total) {
            id2 = ch_item(t, R, seg, ngroups);
            ch_prefetch_plan(smem_u32(&ring[k & (kChPlanRing - 1)]), plans + id2.r, lane);
        }
        cp_async_commit_group();
    };
    auto issue_dy = [&](int k, const ChItemId id) {           // dY of item k -> buffer k % kChDyBufs
        if (lane == 0) {
            fence_proxy_async();
            mbar_expect_tx(&dyfull[k % kChDyBufs], kChItemBytes);
            bulk_load_1d(dybuf + (k % kChDyBufs) * kChItemBytes, dout + ((int64_t)id.r * C + id.g * 32) * PP, kChItemBytes,
                         &dyfull[k % kChDyBufs]);
        }
    };
    prefetch(0);
    id0 = id2;
    prefetch(1);
    id1 = id2;
    issue_dy(0, id0);
    cp_async_wait_group<0>();
    __syncwarp();

    int n_red = 0;                                   // slabs handed to the reduce engine so far
#pragma unroll 1
    for (int ci = 0; item_of(ci) < total; ci++, id0 = id1, id1 = id2) {
        prefetch(ci + 2);
        cp_async_wait_group<1>();
        __syncwarp();                                // also: every lane is done with item ci - 1 (its dY buffer is free)
        if (kChDyBufs > 1 && item_of(ci + 1) < total) issue_dy(ci + 1, id1);
        mbar_wait(&dyfull[ci % kChDyBufs], (uint32_t)(ci / kChDyBufs) & 1u);
        const ChPlan &pl = plan_of(ci);
        const uint32_t dyb = smem_u32(dybuf + (ci % kChDyBufs) * kChItemBytes) + lane * PP * 4;
        if (pl.status == CH_GATHER) ch_bwd_gather_item(f, rois5, id0.r, id0.g * 32 + lane, dyb);
        if (pl.status == CH_OK) {
            const int nq = pl.nq, nst = pl.nst;
            const int z = pl.b * C + id0.g * 32;
            const bool dense = pl.dense != 0;
            if (!dense) ch_build_ax(AxS, pl, lane);
            const uint32_t axb = smem_u32(AxS);

            // one stage: zero the slab, accumulate its bins, hand it to the reduce engine.  NQ > 0: compact RoI, Ax in registers.
            auto run_stages = [&](auto tag) {
                constexpr int NQ = decltype(tag)::value;
                constexpr int AW = NQ > 0 ? NQ : 1;
                float ax[P][4 * AW];
                int qd[AW];
                if (NQ > 0) {
                    const int rot = ch_rot(NQ, lane);
#pragma unroll
                    for (int j = 0; j < AW; j++) {
                        int q = j + rot;
                        if (q >= NQ) q -= NQ;
                        qd[j] = 16 * q;
                    }
                    const uint32_t axd = smem_u32(&pl.axd[0][0]);
#pragma unroll
                    for (int px = 0; px < P; px++)
#pragma unroll
                        for (int j = 0; j < AW; j++) {
                            const float4 a = lds128f(axd + px * 64 + qd[j]);
                            ax[px][4 * j] = a.x; ax[px][4 * j + 1] = a.y; ax[px][4 * j + 2] = a.z; ax[px][4 * j + 3] = a.w;
                        }
                }
#pragma unroll 1
                for (int s = 0; s < nst; s++) {
                    const ChStage st = pl.stage[s];
                    const int kn = st.kn, rows = st.rows;
                    const uint32_t slab = slots + (n_red % kChSlots) * kChSlotBytes;
                    if (lane == 0) bulk_wait_read<kChSlots - 1>();       // the reduce that last used this slab has read it
                    __syncwarp();
                    const int nz = rows * kn;
                    const float4 zero = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
#pragma unroll 4
                    for (int i = 0; i < nz; i++) sts128f(slab + (i * 32 + lane) * 16, zero);
                    __syncwarp();
                    const uint32_t slab_lane = slab + lane * 16 * kn;
                    const int rowpitch = 512 * kn;
#pragma unroll 1
                    for (int py = st.p0; py < st.p1; py++) {
                        float dy[P];
#pragma unroll
                        for (int px = 0; px < P; px++) dy[px] = lds32f(dyb + (py * P + px) * 4);
                        const int4 rr = *reinterpret_cast<const int4 *>(pl.bin[py].row);
                        const float4 ww = *reinterpret_cast<const float4 *>(pl.bin[py].w);
                        const uint32_t a0 = slab_lane + rr.x * rowpitch, a1 = slab_lane + rr.y * rowpitch;
                        const uint32_t a2 = slab_lane + rr.z * rowpitch, a3 = slab_lane + rr.w * rowpitch;
                        if (NQ > 0) {
                            ch_bwd_group<AW, true>(a0, a1, a2, a3, ww, qd, dy, 0u, ax);
                        } else {
                            const int rot = ch_rot(kn, lane);
                            const uint32_t axk = axb + 16 * st.k0;
#pragma unroll 1
                            for (int j0 = 0; j0 < kn; j0 += 2) {
                                int q0 = j0 + rot, q1 = j0 + 1 + rot;
                                if (q0 >= kn) q0 -= kn;
                                if (q1 >= kn) q1 -= kn;
                                if (j0 + 1 < kn) {
                                    const int qo[2] = { 16 * q0, 16 * q1 };
                                    ch_bwd_group<2, false>(a0, a1, a2, a3, ww, qo, dy, axk, nullptr);
                                } else {
                                    const int qo[1] = { 16 * q0 };
                                    ch_bwd_group<1, false>(a0, a1, a2, a3, ww, qo, dy, axk, nullptr);
                                }
                            }
                        }
                    }
                    fence_proxy_async();
                    __syncwarp();
                    if (lane == 0) {
                        const CUtensorMap *m1 = &maps.m[(pl.l * kChMaxNq + kn - 1) * 2], *m4 = m1 + 1;
                        const int x = pl.x_lo + 4 * st.k0;
                        int row = 0;
#pragma unroll 1
                        for (; row + 4 <= rows; row += 4) ch_tma_reduce_add(m4, x, z, st.y0 + row, slab + row * kn * 512);
#pragma unroll 1
                        for (; row < rows; row++) ch_tma_reduce_add(m1, x, z, st.y0 + row, slab + row * kn * 512);
                        bulk_commit();
                    }
                    n_red++;
                }
            };
            if (dense) {
                switch (nq) {
                    case 1: run_stages(std::integral_constant<int, 1>()); break;
                    case 2: run_stages(std::integral_constant<int, 2>()); break;
                    case 3: run_stages(std::integral_constant<int, 3>()); break;
                    default: run_stages(std::integral_constant<int, 4>()); break;
                }
            } else run_stages(std::integral_constant<int, 0>());
        }
        if (kChDyBufs == 1) {                        // single dY buffer: the next item's dY is fetched when this one is done
            __syncwarp();
            if (item_of(ci + 1) < total) issue_dy(ci + 1, id1);
        }
    }
    if (lane == 0) bulk_wait_all<0>();
}

// =====================================================================================================
// host
// =====================================================================================================
struct ChMapCache {
    void *ptr[kChLevels]; int H[kChLevels], W[kChLevels], BC, L, mask; ChMaps maps; bool valid; unsigned long long stamp;
};
constexpr int kChCacheEntries = 32;
static ChMapCache g_ch_cache[kChCacheEntries];
static unsigned long long g_ch_stamp = 0;
static std::mutex g_ch_mutex;

// Returns the mask of levels that have tensor maps (row pitch and base 16-byte aligned, driver entry point present).
static int ch_build_maps(const FeatSet &fs, ChMaps *out)
{
    EncodeTiledFn enc = get_encode();
    std::lock_guard<std::mutex> lock(g_ch_mutex);
    const int L = fs.L < kChLevels ? fs.L : kChLevels;
    ChMapCache *hit = nullptr, *victim = &g_ch_cache[0];
    for (int e = 0; e < kChCacheEntries; e++) {
        ChMapCache &c = g_ch_cache[e];
        bool same = c.valid && c.L == L && c.BC == fs.B * fs.C;
        for (int l = 0; same && l < L; l++) same = c.ptr[l] == fs.feat[l] && c.H[l] == fs.H[l] && c.W[l] == fs.W[l];
        if (same) { hit = &c; break; }
        if (!c.valid) { if (victim->valid) victim = &c; }
        else if (victim->valid && c.stamp < victim->stamp) victim = &c;
    }
    if (!hit) {
        ChMapCache &c = *victim;
        std::memset(&c.maps, 0, sizeof(c.maps));
        c.mask = 0;
        for (int l = 0; l < L; l++) {
            c.ptr[l] = fs.feat[l]; c.H[l] = fs.H[l]; c.W[l] = fs.W[l];
            if (!enc || (fs.W[l] & 3) || (reinterpret_cast<uintptr_t>(fs.feat[l]) & 15)) continue;
            bool ok = true;
            for (int kn = 1; kn <= kChMaxNq && ok; kn++)
                for (int r4 = 0; r4 < 2 && ok; r4++) {
                    const cuuint64_t dims[3] = { (cuuint64_t)fs.W[l], (cuuint64_t)fs.B * fs.C, (cuuint64_t)fs.H[l] };
                    const cuuint64_t strides[2] = { (cuuint64_t)fs.W[l] * fs.H[l] * 4, (cuuint64_t)fs.W[l] * 4 };
                    const cuuint32_t box[3] = { (cuuint32_t)(4 * kn), 32u, r4 ? 4u : 1u };
                    const cuuint32_t estr[3] = { 1, 1, 1 };
                    ok = enc(&c.maps.m[(l * kChMaxNq + kn - 1) * 2 + r4], CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, fs.feat[l], dims, strides,
                             box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                             CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
                }
            if (ok) c.mask |= 1 << l;
        }
        c.L = L; c.BC = fs.B * fs.C; c.valid = true;
        hit = &c;
    }
    hit->stamp = ++g_ch_stamp;
    *out = hit->maps;
    return hit->mask;
}

// MD_ROI_CH=1 selects the channel-lane kernels.  They are parity-green but not faster than the row-streaming kernels of
// roialign_tma.cu at config 2 (both designs sit on the same wall, ~28 M 48-byte TMA row requests per launch at
// one CTA-resident pipeline per SM), so the row-streaming kernels stay the default.
static bool ch_enabled()
{
    const char *e = getenv("MD_ROI_CH");
    return e && atoi(e) != 0;
}

static int ch_sm_count()
{
    int dev = 0, sms = 148;
    if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    return sms;
}

size_t roialign_ch_workspace_bytes(int R) { return (size_t)(R > 0 ? R : 0) * sizeof(ChPlan) + 256; }

constexpr int kChSegRois = 512;       // RoIs per L2 sweep: one image of config 2

template <bool FWD>
static cudaError_t ch_launch(const FeatSet &fs, const RoiFeat &f, const float *rois5, int R, int P, float *out, const float *dout,
                             int32_t *flag, void *plan_ws, cudaStream_t s, bool *launched)
{
    *launched = false;
    if ((fs.C & 31) || P != kChP || R <= 0 || !plan_ws || !ch_enabled()) return cudaSuccess;
    if ((reinterpret_cast<uintptr_t>(FWD ? (const void *)out : (const void *)dout) & 15)) return cudaSuccess;
    ChMaps maps;
    const int mask = ch_build_maps(fs, &maps);
    ChPlan *plans = reinterpret_cast<ChPlan *>((reinterpret_cast<uintptr_t>(plan_ws) + 255) & ~(uintptr_t)255);
    ch_plan_kernel<<<(R + 127) / 128, 128, 0, s>>>(f, mask, rois5, R, kChSlotQuads, plans, flag);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return e;
    const int total = R * (fs.C >> 5);
    const int sms = ch_sm_count();
    constexpr int W = FWD ? kChWarps : kChBwdWarps;
    const int grid = (total + W - 1) / W < sms ? (total + W - 1) / W : sms;
    if (FWD) {
        auto kern = roialign_fwd_ch_kernel;
        const int smem = kChWarps * kChFwdWarpBytes;
        e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);      // per device: set every time
        if (e != cudaSuccess) return e;
        kern<<<grid, kChWarps * 32, smem, s>>>(maps, f, rois5, plans, R, fs.C, kChSegRois, out);
    } else {
        auto kern = roialign_bwd_ch_kernel;
        const int smem = kChBwdWarps * kChBwdWarpBytes;
        e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
        if (e != cudaSuccess) return e;
        kern<<<grid, kChBwdWarps * 32, smem, s>>>(maps, f, rois5, plans, R, fs.C, kChSegRois, dout);
    }
    e = cudaGetLastError();
    *launched = e == cudaSuccess;
    return e;
}

cudaError_t launch_roialign_fwd_ch(const FeatSet &fs, const RoiFeat &f, const float *rois5, int R, int P, float *out,
                                   int32_t *fallback_flag, void *plan_ws, cudaStream_t s, bool *launched)
{
    return ch_launch<true>(fs, f, rois5, R, P, out, nullptr, fallback_flag, plan_ws, s, launched);
}

cudaError_t launch_roialign_bwd_ch(const FeatSet &fs, const RoiFeat &f, const float *rois5, int R, int P, const float *dout,
                                   int32_t *fallback_flag, void *plan_ws, cudaStream_t s, bool *launched)
{
    return ch_launch<false>(fs, f, rois5, R, P, nullptr, dout, fallback_flag, plan_ws, s, launched);
}

}  // namespace md
