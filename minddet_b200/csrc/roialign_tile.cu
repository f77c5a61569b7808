// roialign_tile.cu -- a11: TILE-STATIONARY (gather-form) RoIAlign backward for 7x7 / S = 2.  sm_100a.
//
// The scatter-add backward (roialign_tma.cu) needs dX zero-filled first (731 MB of memset at config 2) and then
// read-modify-writes every footprint at L2.  Here every dX byte is written exactly ONCE:
//
//   plan   tile_plan_kernel, one thread per RoI: the separable operators of the RoI (roialign_tile_plan.h: <= 28 touched
//          feature rows x 7 bin weights; columns dense for footprints <= 16 wide, per bin otherwise), a 16-byte header, and
//          one count per 8 x 32-pixel dX tile the footprint overlaps.
//   fill   tile_fill_kernel, one warp per non-empty tile: the RoIs overlapping the tile, in RoI order (ballot + prefix over the
//          headers, so the order -- and with it the floating-point sum -- is deterministic run to run).
//   main   tile_bwd_kernel, persistent single-warp CTAs, work item = (tile, 32 channels), LANE = CHANNEL: the 8 x 32 x 32-channel
//          accumulator tile lives in shared memory as [row][channel][33] (odd pitch: conflict-free for lane = channel and for the
//          lane = column read-out; slot 32 of every row is a dump slot for columns outside the tile).  Per RoI of the tile's list
//          ("visit") the warp gets dY (32 channels x 49 = 6272 contiguous bytes) and the plan (1664 bytes) by two bulk copies
//          on one mbarrier, issued one visit ahead; every index, weight, loop bound and branch is warp-uniform.  Per touched
//          row: V[q] = sum_p Wy[y][p] dY[p][q] (49 FMAs on registers), then per in-tile column one shared-memory
//          read-modify-write with 7 FMAs (dense plans: column weights in registers), or per bin 4 read-modify-writes (wide
//          plans; bins {0,2,4,6} then {1,3,5} as two batches of independent updates).  After the last visit the tile leaves
//          as 128-byte rows of streaming stores; tiles no RoI touches are written as zeros without going through shared memory.
//
// RoIs the plan declines (S != 2, ...) are flagged for the gather kernel (roialign.cu), which runs after this kernel and
// adds onto the finished dX.  Algorithmic bytes: dY (R*C*49*4) + dX written once; nothing is zero-filled, nothing is re-read.
//
// No reference code exists for this op; semantics: oracle/CONVENTIONS.md #14-16, #23, checked against
// oracle/region_oracle.c:o_roialign_bwd (1e-5 of the gradient scale: separable summation order, FMAs).
#include <cstdio>
#include <cstdlib>

#include "kernels.h"
#include "launch.cuh"
#include "roialign_common.cuh"
#include "roialign_tile_plan.h"
#include "tma_ptx.cuh"

namespace md {

using namespace tile;

constexpr int kTilePitch = kTW + 1;                              // 32 columns + dump slot
constexpr int kTileRowFloats = kTC * kTilePitch;                 // 1056
constexpr int kTileFloats = kTH * kTileRowFloats;                // 8448
constexpr int kTileBytes = kTileFloats * 4;                      // 33792
constexpr int kDyBytes = kTC * kP * kP * 4;                      // 6272
constexpr int kPlanBytes = (int)sizeof(Plan);                    // 1664
constexpr int kStageBytes = kDyBytes + kPlanBytes;
constexpr int kWytBytes = kTH * 8 * 4;                           // 256
constexpr int kCotBytes = 32 * 4;
constexpr int kTileSmem = kTileBytes + kStageBytes + kWytBytes + kCotBytes + 16;
static_assert(kTileBytes % 16 == 0 && kDyBytes % 16 == 0 && kPlanBytes % 16 == 0, "bulk copies are 16-byte granular");

struct TileArgs {
    Grid g;
    int C, R, ncg, nitems;
    float *feat[kMaxLv];
    const float *dout;
    const Plan *plans;
    const int2 *tiles;          // per tile: {offset into lists, count}
    const int *lists;
    int *ticket;
    long long feat_elems[kMaxLv], lists_cap;
    int dbg;     // MD_TILE_CHECK builds: extents for the device-side range checks
};

// ---- plan + count ---------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128)
tile_plan_kernel(const __grid_constant__ RoiFeat f, const __grid_constant__ Grid g, const float *__restrict__ rois5, const int R, Plan *__restrict__ plans,
                 Hdr *__restrict__ hdr, int *__restrict__ cnt, int32_t *__restrict__ flag)
{
    pdl_entry();
    const int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= R) return;
    float roi[5];
#pragma unroll
    for (int k = 0; k < 5; k++) roi[k] = __ldg(rois5 + (int64_t)r * 5 + k);
    int b, l;
    Plan &pl = plans[r];
    plan_roi(roi, f.B, f.L, f.H, f.W, f.cfg, pl, b, l);
    hdr[r] = make_hdr(pl, b, l);
    flag[r] = pl.status == ST_DECLINE;
    if (pl.status != ST_OK) return;
    for (int ty = pl.y0 / kTH; ty <= pl.y1 / kTH; ty++)
        for (int tx = pl.x0 / kTW; tx <= pl.x1 / kTW; tx++) atomicAdd(cnt + tile_id(g, l, b, ty, tx), 1);
}

MD_DEVINL void decode_tile(const Grid &g, int t, int &l, int &b, int &ty, int &tx)
{
    l = 0;
    while (l + 1 < g.L && t >= g.base[l + 1]) l++;
    int rem = t - g.base[l];
    const int per = g.nty[l] * g.ntx[l];
    b = rem / per; rem -= b * per;
    ty = rem / g.ntx[l]; tx = rem - ty * g.ntx[l];
}

// ---- lists ----------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128)
tile_fill_kernel(const __grid_constant__ Grid g, const Hdr *__restrict__ hdr, const int R, const int *__restrict__ cnt, int2 *__restrict__ tiles,
                 int *__restrict__ cursor, int *__restrict__ lists)
{
    pdl_entry();
    const int lane = threadIdx.x & 31;
    const int T = g.base[g.L];
    for (int t = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); t < T; t += gridDim.x * (blockDim.x >> 5)) {
        const int c = cnt[t];
        int off = 0;
        if (lane == 0 && c) off = atomicAdd(cursor, c);
        off = __shfl_sync(0xffffffffu, off, 0);
        if (lane == 0) tiles[t] = make_int2(off, c);
        if (!c) continue;
        int l, b, ty, tx;
        decode_tile(g, t, l, b, ty, tx);
        const int key = ST_OK | (l << 8) | (b << 16);
        int pos = off;
        for (int base = 0; base < R; base += 32) {
            const int r = base + lane;
            bool hit = false;
            if (r < R) {
                const int4 h = __ldg(reinterpret_cast<const int4 *>(hdr) + r);
                hit = h.x == key && (h.y & 0xffff) / kTW <= tx && (h.y >> 16) / kTW >= tx && (h.z & 0xffff) / kTH <= ty &&
                      (h.z >> 16) / kTH >= ty;
            }
            const unsigned m = __ballot_sync(0xffffffffu, hit);
            if (hit) lists[pos + __popc(m & ((1u << lane) - 1))] = r;
            pos += __popc(m);
        }
    }
}

// ---- main -----------------------------------------------------------------------------------------------------------
MD_DEVINL float fma_(float a, float b, float c) { return __fmaf_rn(a, b, c); }


// Dense plans: the touched rows of one visit.  G = ceil(n / 4) column groups; the tile loads of a row are issued before its
// 49-FMA y-step (which hides their latency), the x-step is G x 4 independent 7-FMA chains, stores are predicated on k < n.
template <int G>
MD_DEVINL void dense_rows(unsigned m, const int n, const float *__restrict__ wyt, float *__restrict__ tp0, const float (&d)[kP * kP],
                          const float (&wk)[kDenseCols][kP])
{
    while (m) {
        const int i = __ffs(m) - 1;
        m &= m - 1;
        float *const tp = tp0 + i * kTileRowFloats;
        float acc[4 * G];
#pragma unroll
        for (int k = 0; k < 4 * G; k++) acc[k] = (k < 4 * (G - 1) || k < n) ? tp[k] : 0.0f;
        const float4 wa = *reinterpret_cast<const float4 *>(wyt + i * 8), wb = *reinterpret_cast<const float4 *>(wyt + i * 8 + 4);
        const float wy[kP] = {wa.x, wa.y, wa.z, wa.w, wb.x, wb.y, wb.z};
        float V[kP];
#pragma unroll
        for (int q = 0; q < kP; q++) {
            V[q] = wy[0] * d[q];
#pragma unroll
            for (int p = 1; p < kP; p++) V[q] = fma_(wy[p], d[p * kP + q], V[q]);
        }
#pragma unroll
        for (int q = 0; q < kP; q++)
#pragma unroll
            for (int k = 0; k < 4 * G; k++) acc[k] = fma_(wk[k][q], V[q], acc[k]);
#pragma unroll
        for (int k = 0; k < 4 * G; k++)
            if (k < 4 * (G - 1) || k < n) tp[k] = acc[k];
    }
}

template <bool ACC>
__global__ void __launch_bounds__(32, 5)
tile_bwd_kernel(const __grid_constant__ TileArgs a)
{
    pdl_entry();
    extern __shared__ __align__(128) unsigned char smem[];
    float *const tile_s = reinterpret_cast<float *>(smem);
    unsigned char *const stage = smem + kTileBytes;
    const float *const sdy = reinterpret_cast<const float *>(stage);
    const Plan *const spl = reinterpret_cast<const Plan *>(stage + kDyBytes);
    float *const wyt = reinterpret_cast<float *>(stage + kStageBytes);
    int *const cot = reinterpret_cast<int *>(stage + kStageBytes + kWytBytes);
    unsigned long long *const bar = reinterpret_cast<unsigned long long *>(stage + kStageBytes + kWytBytes + kCotBytes);
    const int lane = threadIdx.x;
    if (lane == 0) { mbar_init(bar, 1); fence_barrier_init(); }
    __syncwarp();
    uint32_t parity = 0;
    float *const tlane = tile_s + lane * kTilePitch;

    for (;;) {
        int item = 0;
        if (lane == 0) item = atomicAdd(a.ticket, 1);
        item = __shfl_sync(0xffffffffu, item, 0);
        if (item >= a.nitems) break;
        const int t = item / a.ncg, c0 = (item - t * a.ncg) * kTC;
        int l, b, ty, tx;
        decode_tile(a.g, t, l, b, ty, tx);
        const int H = a.g.H[l], W = a.g.W[l];
        const int ty0 = ty * kTH, tx0 = tx * kTW;
        int2 rec = __ldg(a.tiles + t);
        if (a.dbg == 11) rec.y = 0;
#ifdef MD_TILE_CHECK
        if (l < 0 || l >= a.g.L || b < 0 || b >= a.g.B || ty < 0 || ty >= a.g.nty[l] || tx < 0 || tx >= a.g.ntx[l] || rec.y < 0 || rec.x < 0 ||
            rec.y > a.R || c0 + kTC > a.C) {
            if (lane == 0) printf("tile check: item %d t %d l %d b %d ty %d tx %d rec %d %d c0 %d\n", item, t, l, b, ty, tx, rec.x, rec.y, c0);
            continue;
        }
#endif
        float *const gbase = a.feat[l] + (((int64_t)b * a.C + c0) * H + ty0) * W + tx0;
        const int nrow = min(kTH, H - ty0), ncol = min(kTW, W - tx0);

#ifdef MD_TILE_CHECK
        {
            const long long first = (((long long)b * a.C + c0) * H + ty0) * W + tx0;
            const long long last = first + ((long long)(kTC - 1) * H + (nrow - 1)) * W + (ncol - 1);
            if (first < 0 || last >= a.feat_elems[l] || nrow < 1 || ncol < 1 || (rec.y > 0 && rec.x + (long long)rec.y > a.lists_cap)) {
                if (lane == 0) printf("tile check: extent t %d l %d b %d ty %d tx %d first %lld last %lld of %lld nrow %d ncol %d rec %d %d\n", t, l, b, ty, tx, first, last, a.feat_elems[l], nrow, ncol, rec.x, rec.y);
                continue;
            }
        }
#endif
        if (rec.y == 0) {                                   // no RoI touches this tile: zeros (or nothing, when accumulating)
            if (!ACC) {
                if (ncol == kTW && (W & 3) == 0) {          // 128-byte rows, 16-byte aligned: lane = (row of 4, 16-byte piece)
                    float4 *p = reinterpret_cast<float4 *>(gbase + (int64_t)(lane >> 3) * W) + (lane & 7);
                    const bool lo = (lane >> 3) < nrow, hi = (lane >> 3) + 4 < nrow;
                    for (int c = 0; c < kTC; c++, p += (int64_t)H * W / 4) {
                        if (lo) __stcs(p, make_float4(0.f, 0.f, 0.f, 0.f));
                        if (hi) __stcs(p + W, make_float4(0.f, 0.f, 0.f, 0.f));
                    }
                } else {
                    for (int c = 0; c < kTC; c++)
                        for (int row = 0; row < nrow; row++)
                            if (lane < ncol) __stcs(gbase + ((int64_t)c * H + row) * W + lane, 0.0f);
                }
            }
            continue;
        }

        if (a.dbg == 12) rec.y = 0;
        const int *const list = a.lists + rec.x;
        auto issue = [&](int r) {
#ifdef MD_TILE_CHECK
            if (r < 0 || r >= a.R) { if (lane == 0) printf("tile check: list entry %d (t %d rec %d %d)\n", r, t, rec.x, rec.y); r = 0; }
#endif
            if (lane == 0) {
                mbar_expect_tx(bar, kStageBytes);
                bulk_load_1d(stage, a.dout + ((int64_t)r * a.C + c0) * (kP * kP), kDyBytes, bar);
                bulk_load_1d(stage + kDyBytes, a.plans + r, kPlanBytes, bar);
            }
        };
        if (rec.y) issue(__ldg(list));
        {
            float4 *z = reinterpret_cast<float4 *>(tile_s);
#pragma unroll 6
            for (int i = lane; i < kTileBytes / 16; i += 32) z[i] = make_float4(0.f, 0.f, 0.f, 0.f);
        }
        __syncwarp();

        for (int v = 0; v < rec.y; v++) {
            mbar_wait(bar, parity);
            parity ^= 1;
            // ---- everything the visit needs leaves the stage now, so the next visit's copies can start
            float d[kP * kP];
#pragma unroll
            for (int k = 0; k < kP * kP; k++) d[k] = sdy[lane * (kP * kP) + k];
            const int wide = spl->wide, nrows = spl->nrows, ncols = spl->ncols, x0 = spl->x0;
            if (a.dbg == 13) {
                __syncwarp();
                if (v + 1 < rec.y) issue(__ldg(list + v + 1));
                continue;
            }
#ifdef MD_TILE_CHECK
            if (nrows < 0 || nrows > kMaxRows || ncols < 1 || x0 < 0 || x0 >= W || (!wide && ncols > kDenseCols) || spl->status != ST_OK ||
                x0 >= tx0 + kTW || x0 + ncols <= tx0) {
                if (lane == 0) printf("tile check: plan t %d v %d r %d wide %d nrows %d ncols %d x0 %d status %d tx0 %d\n", t, v, list[v], wide, nrows, ncols, x0, spl->status, tx0);
                __syncwarp();
                if (v + 1 < rec.y) issue(__ldg(list + v + 1));
                continue;
            }
#endif
            unsigned rowmask;
            {
                unsigned bit = 0;
                if (lane < nrows) {
                    const int yy = spl->row[lane].y - ty0;
                    if (yy >= 0 && yy < kTH) {
                        bit = 1u << yy;
                        const float4 *src = reinterpret_cast<const float4 *>(&spl->row[lane]);
                        float4 w0 = src[0], w1 = src[1];
                        float4 *dst = reinterpret_cast<float4 *>(wyt + yy * 8);
                        dst[0] = make_float4(w0.y, w0.z, w0.w, w1.x);          // w[0..3]  (src[0].x is the row index)
                        dst[1] = make_float4(w1.y, w1.z, w1.w, 0.0f);          // w[4..6]
                    }
                }
                rowmask = __reduce_or_sync(0xffffffffu, bit);        // bit = tile row
            }
            if (!wide) {
                const int ja = max(0, tx0 - x0), n = min(ncols, tx0 + kTW - x0) - ja;
                float wk[kDenseCols][kP];
#pragma unroll
                for (int k = 0; k < kDenseCols; k++) {
                    if (k < n) {
                        const float4 *src = reinterpret_cast<const float4 *>(spl->xw[ja + k]);
                        const float4 w0 = src[0], w1 = src[1];
                        wk[k][0] = w0.x; wk[k][1] = w0.y; wk[k][2] = w0.z; wk[k][3] = w0.w;
                        wk[k][4] = w1.x; wk[k][5] = w1.y; wk[k][6] = w1.z;
                    } else {
#pragma unroll
                        for (int q = 0; q < kP; q++) wk[k][q] = 0.0f;
                    }
                }
                __syncwarp();
                if (v + 1 < rec.y) issue(__ldg(list + v + 1));
                float *const tp0 = tlane + (x0 + ja - tx0);
                switch ((n + 3) >> 2) {                       // row loop specialised on the number of 4-column groups
                case 1: dense_rows<1>(rowmask, n, wyt, tp0, d, wk); break;
                case 2: dense_rows<2>(rowmask, n, wyt, tp0, d, wk); break;
                case 3: dense_rows<3>(rowmask, n, wyt, tp0, d, wk); break;
                default: dense_rows<4>(rowmask, n, wyt, tp0, d, wk); break;
                }
            } else {
                if (lane < 4 * kP) {
                    const int col = spl->bin[lane >> 2].col[lane & 3];
                    cot[lane] = (col >= tx0 && col < tx0 + kTW) ? col - tx0 : kTW;
                }
                float cw[4 * kP];
#pragma unroll
                for (int q = 0; q < kP; q++) {
                    const float4 w = *reinterpret_cast<const float4 *>(spl->bin[q].w);
                    cw[q * 4 + 0] = w.x; cw[q * 4 + 1] = w.y; cw[q * 4 + 2] = w.z; cw[q * 4 + 3] = w.w;
                }
                __syncwarp();
                int co[4 * kP];
#pragma unroll
                for (int q = 0; q < kP; q++) {
                    const int4 o = *reinterpret_cast<const int4 *>(cot + q * 4);
                    co[q * 4 + 0] = o.x; co[q * 4 + 1] = o.y; co[q * 4 + 2] = o.z; co[q * 4 + 3] = o.w;
                }
                __syncwarp();
#ifdef MD_TILE_CHECK
                if (a.dbg == 21 && lane == 0 && c0 == 0) {
                    printf("wide visit t %d r %d rowmask %x tx0 %d ty0 %d\n", t, list[v], rowmask, tx0, ty0);
                    for (int q = 0; q < kP; q++)
                        printf("  q %d co %d %d %d %d cw %.3f %.3f %.3f %.3f plan col %d %d %d %d\n", q, co[q * 4], co[q * 4 + 1], co[q * 4 + 2], co[q * 4 + 3],
                               cw[q * 4], cw[q * 4 + 1], cw[q * 4 + 2], cw[q * 4 + 3], spl->bin[q].col[0], spl->bin[q].col[1], spl->bin[q].col[2], spl->bin[q].col[3]);
                }
                __syncwarp();
#endif
                if (v + 1 < rec.y) issue(__ldg(list + v + 1));
#pragma unroll
                for (int i = 0; i < kTH; i++) {
                    if (!((rowmask >> i) & 1u)) continue;
                    const float4 wa = *reinterpret_cast<const float4 *>(wyt + i * 8), wb = *reinterpret_cast<const float4 *>(wyt + i * 8 + 4);
                    const float wy[kP] = {wa.x, wa.y, wa.z, wa.w, wb.x, wb.y, wb.z};
                    float V[kP];
#pragma unroll
                    for (int q = 0; q < kP; q++) {
                        V[q] = wy[0] * d[q];
#pragma unroll
                        for (int p = 1; p < kP; p++) V[q] = fma_(wy[p], d[p * kP + q], V[q]);
                    }
                    float *const tp = tlane + i * kTileRowFloats;
#pragma unroll
                    for (int ph = 0; ph < 2; ph++) {              // bins {0,2,4,6}, then {1,3,5}: distinct columns inside a batch
                        float tv[16];
#pragma unroll
                        for (int q = ph; q < kP; q += 2)
#pragma unroll
                            for (int s = 0; s < 4; s++) tv[(q >> 1) * 4 + s] = tp[co[q * 4 + s]];
#pragma unroll
                        for (int q = ph; q < kP; q += 2)
#pragma unroll
                            for (int s = 0; s < 4; s++) tv[(q >> 1) * 4 + s] = fma_(cw[q * 4 + s], V[q], tv[(q >> 1) * 4 + s]);
#pragma unroll
                        for (int q = ph; q < kP; q += 2)
#pragma unroll
                            for (int s = 0; s < 4; s++) tp[co[q * 4 + s]] = tv[(q >> 1) * 4 + s];
                    }
                }
            }
            __syncwarp();                                    // wyt / cot are rewritten by the next visit
        }

        // ---- read-out: lane = column, 128-byte rows of streaming stores
        for (int c = 0; c < kTC; c++) {
            const float *src = tile_s + c * kTilePitch + lane;
            float *dst = gbase + (int64_t)c * H * W + lane;
#pragma unroll
            for (int i = 0; i < kTH; i++) {
                if (i < nrow && lane < ncol) {
                    float val = src[i * kTileRowFloats];
                    if (ACC) val += __ldcs(dst + i * W);
                    __stcs(dst + i * W, val);
                }
            }
        }
        __syncwarp();
    }
}

// ---- host -----------------------------------------------------------------------------------------------------------
static size_t al256(size_t n) { return (n + 255) & ~(size_t)255; }

struct TileLayout { size_t ctl, cnt, tiles, hdr, plans, lists, total; int T, cap; };
static TileLayout tile_layout(const FeatSet &fs, int R)
{
    Grid g;
    make_grid(g, fs.L, fs.B, fs.H, fs.W);
    TileLayout o{};
    o.T = g.base[fs.L];
    int per_img_max = 1;
    for (int l = 0; l < fs.L; l++) per_img_max = per_img_max > g.nty[l] * g.ntx[l] ? per_img_max : g.nty[l] * g.ntx[l];
    o.cap = R * per_img_max;                                 // a RoI lies on one (image, level): it overlaps at most every tile of it
    size_t off = 0;
    o.ctl = off; off += 256;                                 // cursor, ticket
    o.cnt = off; off += al256((size_t)o.T * sizeof(int));
    o.tiles = off; off += al256((size_t)o.T * sizeof(int2));
    o.hdr = off; off += al256((size_t)R * sizeof(Hdr));
    o.plans = off; off += al256((size_t)R * sizeof(Plan));
    o.lists = off; off += al256((size_t)o.cap * sizeof(int));
    o.total = off;
    return o;
}
size_t roialign_tile_workspace_bytes(const FeatSet &fs, int R) { return tile_layout(fs, R > 0 ? R : 0).total + 256; }

bool roialign_tile_enabled()
{
    const char *e = getenv("MD_ROI_TILE");       // read on every call (tests flip it)
    return !e || atoi(e) != 0;
}

// dX written once (accumulate = false) or dX += (true).  `flags` (R ints) receives 1 for the RoIs left to the gather kernel.
cudaError_t launch_roialign_bwd_tile(const FeatSet &fs, const RoiFeat &f, const float *rois5, int R, int P, const float *dout,
                                     int32_t *flags, void *tile_ws, bool accumulate, cudaStream_t s, bool *launched)
{
    *launched = false;
    if (!roialign_tile_enabled() || P != kP || fs.C % kTC != 0 || fs.L > kMaxLv || R <= 0 || !tile_ws) return cudaSuccess;
    if ((reinterpret_cast<uintptr_t>(dout) & 15) != 0) return cudaSuccess;
    for (int l = 0; l < fs.L; l++)
        if (fs.W[l] >= 65536 || fs.H[l] >= 65536) return cudaSuccess;
    if (fs.B >= 32768) return cudaSuccess;
    static int sms[64] = {0};
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return e;
    if (dev >= 64) return cudaSuccess;
    if (!sms[dev]) {
        int n = 0;
        if ((e = cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev)) != cudaSuccess) return e;
        if ((e = cudaFuncSetAttribute(tile_bwd_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, kTileSmem)) != cudaSuccess) return e;
        if ((e = cudaFuncSetAttribute(tile_bwd_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kTileSmem)) != cudaSuccess) return e;
        sms[dev] = n;
    }
    const char *dbg_env = getenv("MD_TILE_DEBUG");
    const int dbg = dbg_env ? atoi(dbg_env) : 0;
    const TileLayout lo = tile_layout(fs, R);
    unsigned char *w = reinterpret_cast<unsigned char *>((reinterpret_cast<uintptr_t>(tile_ws) + 255) & ~(uintptr_t)255);
    TileArgs a{};
    make_grid(a.g, fs.L, fs.B, fs.H, fs.W);
    a.C = fs.C; a.R = R; a.ncg = fs.C / kTC; a.nitems = lo.T * a.ncg;
    for (int l = 0; l < fs.L; l++) a.feat[l] = fs.feat[l];
    a.dout = dout;
    a.plans = reinterpret_cast<const Plan *>(w + lo.plans);
    a.tiles = reinterpret_cast<const int2 *>(w + lo.tiles);
    a.lists = reinterpret_cast<const int *>(w + lo.lists);
    int *ctl = reinterpret_cast<int *>(w + lo.ctl);
    a.ticket = ctl + 1;
    for (int l = 0; l < fs.L; l++) a.feat_elems[l] = (long long)fs.B * fs.C * fs.H[l] * fs.W[l];
    a.lists_cap = lo.cap;
    a.dbg = dbg;
    if ((e = cudaMemsetAsync(w + lo.ctl, 0, lo.cnt + al256((size_t)lo.T * sizeof(int)) - lo.ctl, s)) != cudaSuccess) return e;
    tile_plan_kernel<<<(R + 127) / 128, 128, 0, s>>>(f, a.g, rois5, R, reinterpret_cast<Plan *>(w + lo.plans),
                                                      reinterpret_cast<Hdr *>(w + lo.hdr), reinterpret_cast<int *>(w + lo.cnt), flags);
    if ((e = cudaGetLastError()) != cudaSuccess) return e;
    if (dbg == 1) { *launched = true; return cudaSuccess; }
    const int fill_blocks = (lo.T + 3) / 4;
    if ((e = launch_pdl(tile_fill_kernel, dim3(fill_blocks), dim3(128), 0, s, a.g, reinterpret_cast<const Hdr *>(w + lo.hdr), R,
                        reinterpret_cast<const int *>(w + lo.cnt), reinterpret_cast<int2 *>(w + lo.tiles), ctl,
                        reinterpret_cast<int *>(w + lo.lists))) != cudaSuccess) return e;
    if (dbg == 2) { *launched = true; return cudaSuccess; }
    const int grid = a.nitems < sms[dev] * 5 ? a.nitems : sms[dev] * 5;
    if (getenv("MD_VERBOSE")) fprintf(stderr, "[mdregion] tile bwd: T %d items %d grid %d smem %d R %d\n", lo.T, a.nitems, grid, kTileSmem, R);
    if (accumulate) e = launch_pdl(tile_bwd_kernel<true>, dim3(grid), dim3(32), (size_t)kTileSmem, s, a);
    else e = launch_pdl(tile_bwd_kernel<false>, dim3(grid), dim3(32), (size_t)kTileSmem, s, a);
    if (e != cudaSuccess) return e;
    *launched = true;
    return cudaSuccess;
}

}  // namespace md
