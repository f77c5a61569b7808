// roialign_tile.cu -- a11: TILE-STATIONARY (gather-form) RoIAlign backward for 7x7 / S = 2.  sm_100a.
//
// The scatter-add backward (roialign_tma.cu) needs dX zero-filled first (731 MB of memset at config 2) and then
// read-modify-writes every footprint at L2.  Here every dX byte is written exactly ONCE:
//
//   plan   tile_plan_kernel, one thread per RoI: the separable operators of the RoI (roialign_tile_plan.h: <= 28 touched
//          feature rows x 7 bin weights; columns dense for footprints <= 16 wide, per bin otherwise), a 16-byte header, a mask
//          of the tile rows that hold a touched feature row, and one count per kTH x 32-pixel dX tile of the footprint (kTH = 4).
//   lists  tile_offsets_kernel (one thread per tile: its slice of the list buffer and its work-item records),
//          tile_scatter_kernel (one warp per RoI drops its index into every tile it overlaps, in any order), tile_sort_kernel
//          (one warp per tile: ascending RoI order, so the floating-point sum of a tile is taken in a fixed order).
//   items  a work item = (tile, <= kChunk consecutive visits of its list) x (32 channels).  A crowd puts > 100 RoIs on one
//          tile; such a tile is cut into several items, zero-filled by tile_zero_kernel and every item ADDS its partial tile
//          with vector reductions at L2 (the only place where the order of a sum is not fixed).  Crowded tiles' items get the
//          early tickets.
//   main   tile_bwd_kernel, persistent single-warp CTAs (8 per SM), LANE = CHANNEL: the kTH x 32 x 32-channel accumulator tile
//          lives in shared memory as [row][channel][33] (odd pitch: conflict-free for lane = channel and for the 16-byte
//          read-out; slot 32 of every row is a dump slot for columns outside the tile).  Per RoI of the item ("visit") the warp
//          gets dY (32 channels x 49 = 6272 contiguous bytes) and the plan (1664 bytes) by two bulk copies on one mbarrier,
//          issued one visit ahead -- at the last visit of an item for the first visit of the NEXT item (tickets are taken two
//          items ahead, item records one item ahead); every index, weight, loop bound and branch is warp-uniform.  Per
//          touched row: V[q] = sum_p Wy[y][p] dY[p][q] (49 FMAs on registers), then per in-tile column one shared-memory
//          read-modify-write with 7 FMAs (dense plans: column weights in registers), or per bin 4 read-modify-writes (wide
//          plans; bins {0,2,4,6} then {1,3,5} as two batches of independent updates).  After the last visit the tile leaves
//          as 128-byte rows of streaming stores; tiles no RoI touches are written as zeros without going through shared memory.
//   forms  one call (MdRoiAlignBwd: prepare + run on the (device, stream) workspace) or two ops (MdRoiAlignBwdPrepare ->
//          plan tensor owned by the framework -> MdRoiAlignBwdPlanned), so that a graph can prepare beside the forward.
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
constexpr int kTileCtasPerSm = (227 * 1024) / (kTileSmem + 1024) < 16 ? (227 * 1024) / (kTileSmem + 1024) : 16;   // 1 KB per CTA is reserved by the system
static_assert(kTileBytes % 16 == 0 && kDyBytes % 16 == 0 && kPlanBytes % 16 == 0, "bulk copies are 16-byte granular");

// a work item of the main kernel: kChunk (default) consecutive visits of one tile's sorted list, for every channel group
struct __align__(16) Item {
    int t, chunk, nchunk, nv;          // tile, chunk index, chunks of the tile, visits in this chunk (0: nobody touches the tile)
    int first, list, lb, tyx;          // first RoI of the chunk, offset of its list slice, level | image << 8, tile row | column << 16
};
static_assert(sizeof(Item) == 32, "item layout");

struct TileArgs {
    Grid g;
    int C, R, ncg, chunk, item_cap, list_cap, sig;   // sig: what the buffer was prepared for (counters[5] must match)
    float *feat[kMaxLv];
    const float *dout;
    const Plan *plans;
    const int4 *tiles;          // per tile: {offset into lists, count, first item, items}
    const int *lists;           // per tile: RoI indices, ascending
    const Item *items;          // crowded tiles (more than one chunk) from the front, the others from the back: the long items
                                // get the early tickets
    const int *counters;        // [0] list cursor, [1] ticket, [2] items at the front, [3] items at the back, [4] declined RoIs,
                                // [5] signature of the preparation
    int *ticket;
};
constexpr int kMaxTilesPerRoi = 128;   // RoIs overlapping more tiles are left to the gather kernel (bounds the list buffer)
constexpr int kChunk = 16;      // default visits per work item: bounds the longest item (a crowd puts > 100 RoIs on one tile)

// ---- plan + count ---------------------------------------------------------------------------------------------------
// One thread per RoI.  The plan is assembled in LOCAL memory (L1-resident: 32-thread blocks keep it at 52 KB per SM) and
// copied out once; building it in place cost ~200 dependent global read-modify-writes per thread (50 us for 4096 RoIs).
constexpr int kPlanThreads = 32;
__global__ void __launch_bounds__(kPlanThreads)
tile_plan_kernel(const __grid_constant__ RoiFeat f, const __grid_constant__ Grid g, const float *__restrict__ rois5, const int R,
                 Plan *__restrict__ plans, Hdr *__restrict__ hdr, unsigned long long *__restrict__ rowmask, int *__restrict__ cnt,
                 int32_t *__restrict__ flag, int *__restrict__ ndecl)
{
    pdl_entry();
    const int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= R) return;
    float roi[5];
#pragma unroll
    for (int k = 0; k < 5; k++) roi[k] = __ldg(rois5 + (int64_t)r * 5 + k);
    int b, l;
    Plan pl;
    plan_roi(roi, f.B, f.L, f.H, f.W, f.cfg, pl, b, l);
    if (pl.status == ST_OK && (pl.y1 / kTH - pl.y0 / kTH + 1) * (pl.x1 / kTW - pl.x0 / kTW + 1) > kMaxTilesPerRoi) pl.status = ST_DECLINE;
    flag[r] = pl.status == ST_DECLINE;
    if (pl.status == ST_DECLINE) atomicAdd(ndecl, 1);
    hdr[r] = make_hdr(pl, b, l);
    if (pl.status != ST_OK) return;
    {
        const int4 *src = reinterpret_cast<const int4 *>(&pl);
        int4 *dst = reinterpret_cast<int4 *>(plans + r);
        constexpr int kHead = (int)((sizeof(Plan) - sizeof(pl.xw) - sizeof(pl.bin)) / 16);
        const int nhead = 2 + 2 * pl.nrows;                                  // header + the rows in use
        for (int i = 0; i < nhead; i++) dst[i] = src[i];
        if (pl.wide)
            for (int i = kHead + (int)sizeof(pl.xw) / 16; i < (int)sizeof(Plan) / 16; i++) dst[i] = src[i];
        else
            for (int i = kHead; i < kHead + 2 * pl.ncols; i++) dst[i] = src[i];
    }
    // tile rows that hold at least one touched feature row (the bins of a tall RoI leave gaps of several rows)
    const int ty_lo = pl.y0 / kTH;
    unsigned long long rmask = 0;
    for (int i = 0; i < pl.nrows; i++) {
        const int k = pl.row[i].y / kTH - ty_lo;
        rmask |= k < 64 ? 1ull << k : 0ull;
    }
    rowmask[r] = rmask;
    for (int ty = ty_lo; ty <= pl.y1 / kTH; ty++) {
        if (ty - ty_lo < 64 && !((rmask >> (ty - ty_lo)) & 1ull)) continue;
        for (int tx = pl.x0 / kTW; tx <= pl.x1 / kTW; tx++) atomicAdd(cnt + tile_id(g, l, b, ty, tx), 1);
    }
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

// ---- lists: offsets (one thread per tile), then one thread per RoI drops its index into every tile it overlaps -------
__global__ void __launch_bounds__(256)
tile_offsets_kernel(const __grid_constant__ TileArgs a, int *__restrict__ cnt, int4 *__restrict__ tiles, int *__restrict__ counters,
                    Item *__restrict__ items)
{
    pdl_entry();
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= a.g.base[a.g.L]) return;
    if (t == 0) counters[5] = a.sig;                                         // the buffer now belongs to this (shapes, R, chunk)
    const int c = cnt[t];
    cnt[t] = 0;                                                              // re-used as the tile's fill cursor
    const int off = c ? atomicAdd(counters + 0, c) : 0;
    const int nch = c ? (c + a.chunk - 1) / a.chunk : 1;
    // items of a crowded tile go to the front of the array (early tickets), everything else fills it from the back
    const int it = nch > 1 ? atomicAdd(counters + 2, nch) : a.item_cap - 1 - atomicAdd(counters + 3, 1);
    tiles[t] = make_int4(off, c, it, nch);
    int l, b, ty, tx;
    decode_tile(a.g, t, l, b, ty, tx);
    for (int k = 0; k < nch; k++) {
        Item im;
        im.t = t; im.chunk = k; im.nchunk = nch;
        im.nv = min(a.chunk, c - k * a.chunk);
        im.first = 0;                                                        // filled by tile_sort_kernel
        im.list = off + k * a.chunk;
        im.lb = l | (b << 8);
        im.tyx = ty | (tx << 16);
        items[it + k] = im;
    }
}

__global__ void __launch_bounds__(128)
tile_scatter_kernel(const __grid_constant__ Grid g, const Hdr *__restrict__ hdr, const unsigned long long *__restrict__ rowmask, const int R,
                    int *__restrict__ cnt, const int4 *__restrict__ tiles, int *__restrict__ lists)
{
    pdl_entry();
    // one warp per RoI, one lane per overlapped tile (a chain of dependent L2 round trips per tile: lanes run them side by side)
    const int r = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (r >= R) return;
    const int4 h = __ldg(reinterpret_cast<const int4 *>(hdr) + r);
    if ((h.x & 0xff) != ST_OK) return;
    const int l = (h.x >> 8) & 0xff, b = h.x >> 16;
    const int ty_lo = (h.z & 0xffff) / kTH, nty = (h.z >> 16) / kTH - ty_lo + 1;
    const int tx_lo = (h.y & 0xffff) / kTW, ntx = (h.y >> 16) / kTW - tx_lo + 1;
    const unsigned long long rmask = __ldg(rowmask + r);
    for (int j = lane; j < nty * ntx; j += 32) {
        const int dy = j / ntx, dx = j - dy * ntx;
        if (dy < 64 && !((rmask >> dy) & 1ull)) continue;                    // same rule as the count in tile_plan_kernel
        const int t = tile_id(g, l, b, ty_lo + dy, tx_lo + dx);
        lists[tiles[t].x + atomicAdd(cnt + t, 1)] = r;                       // any order: tile_sort_kernel puts it into RoI order
    }
}

// Tiles that were cut into several items (their partial sums are added at L2) get zeros first: the blocks walk the front of the
// item array (the crowded tiles' items) and take the tiles whose first chunk they meet.  Part of the backward proper (it writes
// dX), not of the preparation.
__global__ void __launch_bounds__(256)
tile_zero_kernel(const __grid_constant__ TileArgs a)
{
    pdl_entry();
    if (__ldg(a.counters + 5) != a.sig) return;                              // not a plan made for these shapes: touch nothing
    const int nfront = __ldg(a.counters + 2);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int i = blockIdx.x; i < nfront; i += gridDim.x) {
        const int4 u = __ldg(reinterpret_cast<const int4 *>(a.items + i));      // t, chunk, nchunk, nv
        if (u.y != 0) continue;
        int l, b, ty, tx;
        decode_tile(a.g, u.x, l, b, ty, tx);
        const int H = a.g.H[l], W = a.g.W[l], ty0 = ty * kTH, tx0 = tx * kTW;
        const int nrow = min(kTH, H - ty0), ncol = min(kTW, W - tx0);
        float *const base = a.feat[l] + ((int64_t)b * a.C * H + ty0) * W + tx0;
        if (ncol == kTW && (W & 3) == 0) {                                   // lane = (row of 4, 16-byte piece)
            for (int c = warp; c < a.C; c += 8) {
                float4 *p = reinterpret_cast<float4 *>(base + ((int64_t)c * H + (lane >> 3)) * W) + (lane & 7);
                for (int r4 = 0; r4 < kTH; r4 += 4)
                    if ((lane >> 3) + r4 < nrow) p[(int64_t)r4 * W / 4] = make_float4(0.f, 0.f, 0.f, 0.f);
            }
        } else {
            for (int c = warp; c < a.C; c += 8)
                for (int row = 0; row < nrow; row++)
                    if (lane < ncol) base[((int64_t)c * H + row) * W + lane] = 0.0f;
        }
    }
}

// One warp per tile: its list into ascending RoI order (rank = number of smaller entries; entries are distinct) and the first
// RoI of every chunk into the item records.
constexpr int kSortThreads = 256;
__global__ void __launch_bounds__(kSortThreads)
tile_sort_kernel(const __grid_constant__ TileArgs a, int *__restrict__ lists, Item *__restrict__ items)
{
    pdl_entry();
    const int t = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (t >= a.g.base[a.g.L]) return;
    const int4 rec = a.tiles[t];
    if (rec.y == 0) return;
    int *list = lists + rec.x;
    if (rec.y <= 32) {
        const int e = lane < rec.y ? list[lane] : 0x7fffffff;
        int rank = 0;
        for (int j = 0; j < rec.y; j++) rank += __shfl_sync(0xffffffffu, e, j) < e;
        __syncwarp();
        if (lane < rec.y) list[rank] = e;
    } else {
        // longer lists: every lane ranks its entries against the whole list (read through L1); ranks first, writes after
        int *const scratch = lists + a.list_cap + rec.x;                     // second half of the list buffer: the sorted copy
        for (int i = lane; i < rec.y; i += 32) {
            const int e = list[i];
            int rank = 0;
            for (int j = 0; j < rec.y; j++) rank += list[j] < e;
            scratch[rank] = e;
        }
        __syncwarp();
        for (int i = lane; i < rec.y; i += 32) list[i] = scratch[i];
    }
    __syncwarp();
    for (int k = lane; k < rec.w; k += 32) items[rec.z + k].first = list[k * a.chunk];
}

// ---- main -----------------------------------------------------------------------------------------------------------
MD_DEVINL float fma_(float a, float b, float c) { return __fmaf_rn(a, b, c); }

// Dense plans: one visit.  G = ceil(n / 4) column groups.  The constructor pulls the column weights of the 4 G in-tile columns
// into registers (after it the stage can be refilled); rows() walks the touched rows: the tile loads of a row are issued
// before its 49-FMA y-step (which hides their latency), the x-step is 4 G independent 7-FMA chains, loads / stores of the
// last group are predicated on k < n.
template <int G>
struct DenseVisit {
    float wk[4 * G][kP];
    MD_DEVINL DenseVisit(const Plan *__restrict__ spl, const int ja)
    {
#pragma unroll
        for (int k = 0; k < 4 * G; k++) {
            const float4 *src = reinterpret_cast<const float4 *>(spl->xw[min(ja + k, kDenseCols - 1)]);
            const float4 w0 = src[0], w1 = src[1];
            wk[k][0] = w0.x; wk[k][1] = w0.y; wk[k][2] = w0.z; wk[k][3] = w0.w;
            wk[k][4] = w1.x; wk[k][5] = w1.y; wk[k][6] = w1.z;
        }
    }
    // y-step of one row: V[q] = sum_p wy[p] d[p][q].  (Skipping the bins whose weight is zero for the row -- usually 4-5 of the
    // 7 -- behind warp-uniform branches measured slower, 473 vs 452 us per call; two rows per iteration too, 512 us.)
    static MD_DEVINL void ystep(const float4 wa, const float4 wb, const float (&d)[kP * kP], float (&V)[kP])
    {
        const float wy[kP] = {wa.x, wa.y, wa.z, wa.w, wb.x, wb.y, wb.z};
#pragma unroll
        for (int q = 0; q < kP; q++) {
            V[q] = wy[0] * d[q];
#pragma unroll
            for (int p = 1; p < kP; p++) V[q] = fma_(wy[p], d[p * kP + q], V[q]);
        }
    }
    MD_DEVINL void rows(unsigned m, const int n, const float *__restrict__ wyt, float *__restrict__ tp0, const float (&d)[kP * kP]) const
    {
        while (m) {
            const int i = __ffs(m) - 1;
            m &= m - 1;
            float *const tp = tp0 + i * kTileRowFloats;
            float acc[4 * G];
#pragma unroll
            for (int k = 0; k < 4 * G; k++) acc[k] = (k < 4 * (G - 1) || k < n) ? tp[k] : 0.0f;
            const float4 wa = *reinterpret_cast<const float4 *>(wyt + i * 8), wb = *reinterpret_cast<const float4 *>(wyt + i * 8 + 4);
            float V[kP];
            ystep(wa, wb, d, V);
#pragma unroll
            for (int q = 0; q < kP; q++)
#pragma unroll
                for (int k = 0; k < 4 * G; k++) acc[k] = fma_(wk[k][q], V[q], acc[k]);
#pragma unroll
            for (int k = 0; k < 4 * G; k++)
                if (k < 4 * (G - 1) || k < n) tp[k] = acc[k];
        }
    }
};

MD_DEVINL void red_add_v4(float4 *p, const float4 v)
{
    asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" :: "l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}

// 128-byte rows out of the tile.  MODE 0: streaming stores; 1: dX += tile (this warp owns the pixels); 2: reductions at L2
// (several items share the pixels)
template <int MODE>
MD_DEVINL void tile_readout(const float *__restrict__ tile_s, float *__restrict__ gbase, const int H, const int W, const int nrow, const int ncol,
                            const bool vec, const int lane)
{
    if (vec) {
        // lane = (channel of 4, 16-byte piece): four pitch-33 loads (conflict-free: banks c + 4 piece + j) -> one 16-byte access
        const int cq = lane >> 3, piece = lane & 7;
        const float *src = tile_s + cq * kTilePitch + piece * 4;
        float *dst = gbase + (int64_t)cq * H * W + piece * 4;
        const int64_t cstep = (int64_t)4 * H * W;
        for (int cb = 0; cb < kTC / 4; cb++, src += 4 * kTilePitch, dst += cstep) {
#pragma unroll
            for (int i = 0; i < kTH; i++) {
                if (i < nrow) {
                    const float *sp = src + i * kTileRowFloats;
                    float4 val = make_float4(sp[0], sp[1], sp[2], sp[3]);
                    float4 *gp = reinterpret_cast<float4 *>(dst + (int64_t)i * W);
                    if (MODE == 2) {
                        red_add_v4(gp, val);
                    } else {
                        if (MODE == 1) {
                            const float4 old = __ldcs(gp);
                            val.x += old.x; val.y += old.y; val.z += old.z; val.w += old.w;
                        }
                        __stcs(gp, val);
                    }
                }
            }
        }
    } else {
        for (int c = 0; c < kTC; c++) {
            const float *src = tile_s + c * kTilePitch + lane;
            float *dst = gbase + (int64_t)c * H * W + lane;
#pragma unroll
            for (int i = 0; i < kTH; i++) {
                if (i < nrow && lane < ncol) {
                    float val = src[i * kTileRowFloats];
                    if (MODE == 2) {
                        atomicAdd(dst + i * W, val);
                    } else {
                        if (MODE == 1) val += __ldcs(dst + i * W);
                        __stcs(dst + i * W, val);
                    }
                }
            }
        }
    }
}

template <bool ACC>
__global__ void __launch_bounds__(32, kTileCtasPerSm)
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
    // a plan tensor that was not prepared for these shapes / RoI count / chunk length (two-op form: it comes from the caller) is
    // refused as a whole: nothing is read through it, dX stays unwritten
    if (__ldg(a.counters + 5) != a.sig) return;
    const int nfront = __ldg(a.counters + 2), nback = __ldg(a.counters + 3);
    const int ntickets = (nfront + nback) * a.ncg;

    // Software pipeline over work items: the ticket of item i + 2 is requested and the record of item i + 1 is loaded while
    // item i is being worked on, and the first copies of item i + 1 start as soon as the last visit of item i has left the
    // stage -- the chain ticket -> record -> copies (three dependent round trips to L2) is off the critical path.
    auto take = [&]() {
        int tk = 0;
        if (lane == 0) tk = atomicAdd(a.ticket, 1);
        return __shfl_sync(0xffffffffu, tk, 0);
    };
    auto record = [&](int tk, Item &im, int &c0) {
        const int di = tk / a.ncg;
        c0 = (tk - di * a.ncg) * kTC;
        const int idx = di < nfront ? di : a.item_cap - 1 - (di - nfront);
        const int4 *p = reinterpret_cast<const int4 *>(a.items + idx);
        const int4 u = __ldg(p), w = __ldg(p + 1);
        im.t = u.x; im.chunk = u.y; im.nchunk = u.z; im.nv = u.w;
        im.first = w.x; im.list = w.y; im.lb = w.z; im.tyx = w.w;
    };
    auto issue = [&](int r, int c0) {
        if (lane == 0) {
            mbar_expect_tx(bar, kStageBytes);
            bulk_load_1d(stage, a.dout + ((int64_t)r * a.C + c0) * (kP * kP), kDyBytes, bar);
            bulk_load_1d(stage + kDyBytes, a.plans + r, kPlanBytes, bar);
        }
    };

    int tk_cur = take(), tk_n1 = take();
    Item cur{}, nxt{};
    int c0 = 0, c0n = 0;
    if (tk_cur < ntickets) {
        record(tk_cur, cur, c0);
        if (cur.nv > 0) issue(cur.first, c0);
    }
    while (tk_cur < ntickets) {
        const int tk_n2 = take();
        const bool have_next = tk_n1 < ntickets;
        if (have_next) record(tk_n1, nxt, c0n);
        const int l = cur.lb & 0xff, b = cur.lb >> 8, ty = cur.tyx & 0xffff, tx = cur.tyx >> 16;
        const int H = a.g.H[l], W = a.g.W[l];
        const int ty0 = ty * kTH, tx0 = tx * kTW;
        float *const gbase = a.feat[l] + (((int64_t)b * a.C + c0) * H + ty0) * W + tx0;
        const int nrow = min(kTH, H - ty0), ncol = min(kTW, W - tx0);
        const bool vec = ncol == kTW && (W & 3) == 0;       // 128-byte rows, 16-byte aligned
        const int nv = cur.nv;

        if (nv == 0) {                                      // no RoI touches this tile: zeros (or nothing, when accumulating)
            if (have_next && nxt.nv > 0) issue(nxt.first, c0n);
            if (!ACC) {
                if (vec) {                                  // lane = (row of 4, 16-byte piece)
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
        } else {
            const int *const list = a.lists + cur.list;
            {
                float4 *z = reinterpret_cast<float4 *>(tile_s);
#pragma unroll 6
                for (int i = lane; i < kTileBytes / 16; i += 32) z[i] = make_float4(0.f, 0.f, 0.f, 0.f);
            }
            __syncwarp();

            for (int v = 0; v < nv; v++) {
                const bool last = v + 1 == nv;
                const int rnext = last ? 0 : __ldg(list + v + 1);
                // once the visit has read everything it needs from the stage: the next visit's copies, or the next item's first
                auto refill = [&]() {
                    if (!last) issue(rnext, c0);
                    else if (have_next && nxt.nv > 0) issue(nxt.first, c0n);
                };
                mbar_wait(bar, parity);
                parity ^= 1;
                float d[kP * kP];
#pragma unroll
                for (int k = 0; k < kP * kP; k++) d[k] = sdy[lane * (kP * kP) + k];
                const int wide = spl->wide, nrows = spl->nrows, ncols = spl->ncols, x0 = spl->x0;
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
                    rowmask = __reduce_or_sync(0xffffffffu, bit);                  // bit = tile row
                }
                if (!wide) {
                    const int ja = max(0, tx0 - x0), n = min(ncols, tx0 + kTW - x0) - ja;
                    float *const tp0 = tlane + (x0 + ja - tx0);
                    switch ((n + 3) >> 2) {                       // specialised on the number of 4-column groups
                    case 1: { DenseVisit<1> dv(spl, ja); __syncwarp(); refill(); dv.rows(rowmask, n, wyt, tp0, d); } break;
                    case 2: { DenseVisit<2> dv(spl, ja); __syncwarp(); refill(); dv.rows(rowmask, n, wyt, tp0, d); } break;
                    case 3: { DenseVisit<3> dv(spl, ja); __syncwarp(); refill(); dv.rows(rowmask, n, wyt, tp0, d); } break;
                    default: { DenseVisit<4> dv(spl, ja); __syncwarp(); refill(); dv.rows(rowmask, n, wyt, tp0, d); } break;
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
                    refill();
                    unsigned m = rowmask;
                    while (m) {
                        const int i = __ffs(m) - 1;
                        m &= m - 1;
                        const float4 wa = *reinterpret_cast<const float4 *>(wyt + i * 8), wb = *reinterpret_cast<const float4 *>(wyt + i * 8 + 4);
                        float V[kP];
                        DenseVisit<1>::ystep(wa, wb, d, V);
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

            // ---- read-out.  A tile whose list was cut into several items was zero-filled by tile_sort_kernel (unless the
            // call accumulates anyway) and every item adds its partial sums with vector reductions at L2; other tiles are
            // stored once.
            if (cur.nchunk > 1) tile_readout<2>(tile_s, gbase, H, W, nrow, ncol, vec, lane);
            else if (ACC) tile_readout<1>(tile_s, gbase, H, W, nrow, ncol, vec, lane);
            else tile_readout<0>(tile_s, gbase, H, W, nrow, ncol, vec, lane);
            __syncwarp();
        }
        cur = nxt; c0 = c0n;
        tk_cur = tk_n1; tk_n1 = tk_n2;
    }
}

// ---- host -----------------------------------------------------------------------------------------------------------
static size_t al256(size_t n) { return (n + 255) & ~(size_t)255; }

static int tile_chunk()
{
    const char *ce = getenv("MD_TILE_CHUNK");
    int c = ce && atoi(ce) > 0 ? atoi(ce) : kChunk;
    return c > 32767 ? 32767 : c;
}

// One buffer holds everything the backward needs besides dY: in the one-call form it lives in the (device, stream) workspace,
// in the two-op form (MdRoiAlignBwdPrepare -> MdRoiAlignBwdPlanned) it is a tensor the framework owns.
struct TileLayout { size_t ctl, cnt, zero_end, tiles, hdr, rmask, flags, plans, lists, items, total; int T, cap, item_cap; };
static TileLayout tile_layout(const FeatSet &fs, int R)
{
    Grid g;
    make_grid(g, fs.L, fs.B, fs.H, fs.W);
    TileLayout o{};
    o.T = g.base[fs.L];
    int per_img_max = 1;
    for (int l = 0; l < fs.L; l++) per_img_max = per_img_max > g.nty[l] * g.ntx[l] ? per_img_max : g.nty[l] * g.ntx[l];
    // a RoI lies on one (image, level) and the plan leaves RoIs that overlap more than kMaxTilesPerRoi tiles to the gather kernel
    o.cap = R * (per_img_max < kMaxTilesPerRoi ? per_img_max : kMaxTilesPerRoi);
    size_t off = 0;
    o.ctl = off; off += 256;                                 // counters (TileArgs::counters), ticket
    o.cnt = off; off += al256((size_t)o.T * sizeof(int));
    o.zero_end = off;                                        // [ctl, zero_end) is cleared at the top of every prepare
    o.tiles = off; off += al256((size_t)o.T * sizeof(int4));
    o.hdr = off; off += al256((size_t)R * sizeof(Hdr));
    o.rmask = off; off += al256((size_t)R * sizeof(unsigned long long));
    o.flags = off; off += al256((size_t)R * sizeof(int32_t));
    o.plans = off; off += al256((size_t)R * sizeof(Plan));
    o.lists = off; off += al256((size_t)o.cap * 2 * sizeof(int));   // lists + the sort's scratch copy
    o.item_cap = o.T + o.cap / tile_chunk() + 1;
    o.items = off; off += al256((size_t)o.item_cap * sizeof(Item));
    o.total = off;
    return o;
}
size_t roialign_tile_workspace_bytes(const FeatSet &fs, int R) { return tile_layout(fs, R > 0 ? R : 0).total + 256; }

bool roialign_tile_enabled()
{
    const char *e = getenv("MD_ROI_TILE");       // read on every call (tests flip it)
    return !e || atoi(e) != 0;
}

bool roialign_tile_supports(const FeatSet &fs, int R, int P)
{
    if (P != kP || fs.C % kTC != 0 || fs.L > kMaxLv || fs.L < 1 || R <= 0 || fs.B >= 32768) return false;
    for (int l = 0; l < fs.L; l++)
        if (fs.W[l] >= 65536 || fs.H[l] >= 65536) return false;
    return true;
}

static unsigned char *tile_base(void *buf) { return reinterpret_cast<unsigned char *>((reinterpret_cast<uintptr_t>(buf) + 255) & ~(uintptr_t)255); }

static void tile_args(const FeatSet &fs, int R, unsigned char *w, const TileLayout &lo, TileArgs &a)
{
    make_grid(a.g, fs.L, fs.B, fs.H, fs.W);
    a.C = fs.C; a.R = R; a.ncg = fs.C / kTC; a.item_cap = lo.item_cap; a.list_cap = lo.cap;
    for (int l = 0; l < fs.L; l++) a.feat[l] = fs.feat[l];
    a.plans = reinterpret_cast<const Plan *>(w + lo.plans);
    a.tiles = reinterpret_cast<const int4 *>(w + lo.tiles);
    a.lists = reinterpret_cast<const int *>(w + lo.lists);
    a.items = reinterpret_cast<const Item *>(w + lo.items);
    a.counters = reinterpret_cast<int *>(w + lo.ctl);
    a.ticket = reinterpret_cast<int *>(w + lo.ctl) + 1;
    a.chunk = tile_chunk();
    unsigned h = 2166136261u;                                                // FNV-1a over everything the layout depends on
    auto mix = [&h](int v) { h = (h ^ (unsigned)v) * 16777619u; };
    mix(R); mix(fs.B); mix(fs.C); mix(fs.L); mix(a.chunk); mix(kTH);
    for (int l = 0; l < fs.L; l++) { mix(fs.H[l]); mix(fs.W[l]); }
    a.sig = (int)(h | 1u);                                                   // never 0 (a cleared buffer is not a plan)
}

// plans, lists, work items: everything that depends on the RoIs only (fs supplies the level shapes; its pointers are not used)
cudaError_t roialign_tile_prepare(const FeatSet &fs, const RoiFeat &f, const float *rois5, int R, void *buf, cudaStream_t s)
{
    const TileLayout lo = tile_layout(fs, R);
    unsigned char *w = tile_base(buf);
    TileArgs a{};
    tile_args(fs, R, w, lo, a);
    int *ctl = reinterpret_cast<int *>(w + lo.ctl);
    cudaError_t e;
    if ((e = cudaMemsetAsync(w + lo.ctl, 0, lo.zero_end - lo.ctl, s)) != cudaSuccess) return e;
    tile_plan_kernel<<<(R + kPlanThreads - 1) / kPlanThreads, kPlanThreads, 0, s>>>(f, a.g, rois5, R, reinterpret_cast<Plan *>(w + lo.plans),
                                                                                   reinterpret_cast<Hdr *>(w + lo.hdr),
                                                                                   reinterpret_cast<unsigned long long *>(w + lo.rmask),
                                                                                   reinterpret_cast<int *>(w + lo.cnt),
                                                                                   reinterpret_cast<int32_t *>(w + lo.flags), ctl + 4);
    if ((e = cudaGetLastError()) != cudaSuccess) return e;
    if ((e = launch_pdl(tile_offsets_kernel, dim3((lo.T + 255) / 256), dim3(256), 0, s, a, reinterpret_cast<int *>(w + lo.cnt),
                        reinterpret_cast<int4 *>(w + lo.tiles), ctl, reinterpret_cast<Item *>(w + lo.items))) != cudaSuccess) return e;
    if ((e = launch_pdl(tile_scatter_kernel, dim3((R + 3) / 4), dim3(128), 0, s, a.g, reinterpret_cast<const Hdr *>(w + lo.hdr),
                        reinterpret_cast<const unsigned long long *>(w + lo.rmask), R, reinterpret_cast<int *>(w + lo.cnt),
                        reinterpret_cast<const int4 *>(w + lo.tiles), reinterpret_cast<int *>(w + lo.lists))) != cudaSuccess) return e;
    const int nsort = (lo.T + kSortThreads / 32 - 1) / (kSortThreads / 32);
    return launch_pdl(tile_sort_kernel, dim3(nsort), dim3(kSortThreads), 0, s, a, reinterpret_cast<int *>(w + lo.lists),
                      reinterpret_cast<Item *>(w + lo.items));
}

// the backward proper on a prepared buffer: zeros for the crowded tiles, the tile kernel; *flags / *ndecl: the RoIs left to the
// gather kernel (which the caller launches next).  rearm: the buffer may have been consumed before (ticket back to zero).
cudaError_t roialign_tile_run(const FeatSet &fs, int R, const float *dout, void *buf, bool accumulate, bool rearm, cudaStream_t s,
                              const int32_t **flags, const int32_t **ndecl)
{
    static_assert(kTileSmem <= 48 * 1024, "no opt-in for large dynamic shared memory needed");
    int dev = 0, nsm = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return e;
    if ((e = cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, dev)) != cudaSuccess) return e;
    const TileLayout lo = tile_layout(fs, R);
    unsigned char *w = tile_base(buf);
    TileArgs a{};
    tile_args(fs, R, w, lo, a);
    a.dout = dout;
    if (rearm && (e = cudaMemsetAsync(a.ticket, 0, sizeof(int), s)) != cudaSuccess) return e;
    if (!accumulate && (e = launch_pdl(tile_zero_kernel, dim3(lo.T < 592 ? lo.T : 592), dim3(256), 0, s, a)) != cudaSuccess) return e;
    const int grid = lo.T * a.ncg < nsm * kTileCtasPerSm ? lo.T * a.ncg : nsm * kTileCtasPerSm;
    if (accumulate) e = launch_pdl(tile_bwd_kernel<true>, dim3(grid), dim3(32), (size_t)kTileSmem, s, a);
    else e = launch_pdl(tile_bwd_kernel<false>, dim3(grid), dim3(32), (size_t)kTileSmem, s, a);
    if (e != cudaSuccess) return e;
    *flags = reinterpret_cast<const int32_t *>(w + lo.flags);
    *ndecl = reinterpret_cast<const int32_t *>(w + lo.ctl) + 4;
    return cudaSuccess;
}

// one-call form: dX written once (accumulate = false) or dX += (true)
cudaError_t launch_roialign_bwd_tile(const FeatSet &fs, const RoiFeat &f, const float *rois5, int R, int P, const float *dout,
                                     void *tile_ws, bool accumulate, cudaStream_t s, bool *launched, const int32_t **flags,
                                     const int32_t **ndecl)
{
    *launched = false;
    *ndecl = nullptr;
    *flags = nullptr;
    if (!roialign_tile_enabled() || !roialign_tile_supports(fs, R, P) || !tile_ws) return cudaSuccess;
    if ((reinterpret_cast<uintptr_t>(dout) & 15) != 0) return cudaSuccess;
    cudaError_t e = roialign_tile_prepare(fs, f, rois5, R, tile_ws, s);
    if (e != cudaSuccess) return e;
    if ((e = roialign_tile_run(fs, R, dout, tile_ws, accumulate, false, s, flags, ndecl)) != cudaSuccess) return e;
    *launched = true;
    return cudaSuccess;
}

}  // namespace md
