// roialign_tma.cu -- a10/a11 fast path: per-warp ROW-STREAMING TMA rings + separable bilinear operators.
//
// RoIAlign (aligned=False, S x S samples averaged) is separable per (RoI, channel):
//        Out[p][q] = sum_{sy in bin p} sum_{sx in bin q}  wy(sy) . F[rows(sy)][cols(sx)] . wx(sx)
// where every 1-D sample touches two adjacent rows / columns (validity and edge clamping are per axis).
//
// Work decomposition (one CTA of 4 warps per RoI, warps fully decoupled -- no CTA barrier after the
// prologue):
//   * lane = (channel-in-group, column quad): a warp owns CPW = 32/LPC channels at a time, LPC lanes per
//     channel, 4 feature columns per lane (LPC in {4,8,16,32} chosen from the footprint width BW <= 128).
//   * the footprint of a channel group is streamed through a per-warp shared-memory ring of ROW BLOCKS
//     (box = BW columns x 4 rows x CPW channels, NCHW -> smem by ONE cp.async.bulk.tensor.3d each; the
//     box x-origin is rounded down to 4 floats because TMA needs a 16-byte aligned innermost coordinate).
//     Blocks of consecutive channel groups follow each other in the ring, so the pipeline never drains;
//     any footprint height works (the window slides with the samples, which are sorted in y).
//   forward : step 1 (registers)  U[p][4 cols] += wy . row  for the 2*S*P (sample,row) pairs, 128-bit LDS
//             step 2              U -> smem, 4 column taps per output, outputs stored fully coalesced.
//   backward: step 1 (registers)  T[p][4 cols] = sum_q dY[p][q] . Ax[q][cols]      (dY staged by cp.async)
//             step 2              D[row][4 cols] = sum_p Ay[p][row] . T[p]  (<= 4 non-zero bins per row),
//             written conflict-free into the ring and folded into dX by ONE TMA reduce-add per row block
//             (cp.reduce.async.bulk.tensor .add.f32, performed at L2; no smem or per-element atomics).
//   Levels whose row pitch is not a multiple of 16 bytes (e.g. 25x42) cannot be described by a tensor map:
//   the same ring is filled with 4-byte cp.async (completion through the same mbarriers), and the
//   backward uses red.global.add.f32 per element.  RoIs the kernel declines (footprint wider than 128
//   columns, S != 2, ...) are flagged and handled by the gather kernels of roialign.cu.
//
// Forward here is NOT bit-identical to the oracle (separable summation order, FMA): tolerance
// rtol 1e-5 (north_star) + atol 1e-6 for cancellation; tests/test_gpu_parity.py states it.
#include <cuda.h>

#include <cstdlib>
#include <cstring>
#include <mutex>

#include "kernels.h"
#include "roialign_common.cuh"
#include "tma_host.h"
#include "tma_ptx.cuh"
#include "launch.cuh"

namespace md {

constexpr int kStWarps = 4;
constexpr int kStThreads = kStWarps * 32;
#ifndef MD_RING
#define MD_RING 4608
#endif
#ifndef MD_FWD_CTAS
#define MD_FWD_CTAS 3
#endif
constexpr int kRingFloats = MD_RING;       // forward ring per warp (18 KB; a consumed stage doubles as the U buffer)
constexpr int kMaxBW = 128;                // footprint width limit (floats)
constexpr int kMaxSlots = 16;
constexpr int kBwdSlots = 4;               // fixed: cp.async.bulk.wait_group needs an immediate
constexpr int kBwdRingFloats = kBwdSlots * 512;
constexpr int kNumBW = kMaxBW / 4;         // tensor maps per level
constexpr int kTmaLevels = 4;
constexpr int kMaxRowsBwd = 128;           // backward row-table capacity

constexpr int kMapsPerLevel = kNumBW + 1;  // + one padded map, see fwd_box_width()
struct TmaMaps { CUtensorMap m[kTmaLevels * kMapsPerLevel]; };   // [level][bw/4 - 1], box = {bw, 4, CPW(bw)}; [level][kNumBW] = {20, 4, 8}

struct __align__(16) SampleTap { int lo, hi; float wl, wh; };   // rows/cols relative to the footprint origin

// Forward only: row pitch of the staged box.  With 4 lanes per channel a quarter-warp's LDS.128 covers two channels,
// and when the channel stride 4 rows * bw * 4 B is a multiple of 128 B (bw = 8, 16) both land in the same banks: every
// row read was a 2-way conflict (1.7x the ideal wavefronts over the kernel, which is bound by that pipe).  Four unused
// columns of padding move the second channel to the other half of the banks.
__host__ __device__ inline int fwd_box_width(int bw) { return (bw == 8 || bw == 16) ? bw + 4 : bw; }
__host__ __device__ inline int lanes_per_channel(int bw) { return bw <= 16 ? 4 : (bw <= 32 ? 8 : (bw <= 64 ? 16 : 32)); }

// ---- per-RoI prologue: sample tables and footprint --------------------------------------------------
template <int P>
struct StreamShared {
    unsigned long long full[kStWarps][kMaxSlots];
    SampleTap ytab[2 * P], xtab[2 * P], yoff[2 * P];
    int x_lo, y_lo, bw, h_fp, any_x, any_y;
};

// warp 0 builds the y table, warp 1 the x table (2*P <= 32 samples per axis); weights carry the 1/S factor.
template <int P>
MD_DEVINL void build_tables(StreamShared<P> &sh, const RoiGeom &g)
{
    constexpr int S = 2, NS = P * S;
    static_assert(NS <= 32, "one warp builds one axis");
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (warp < 2) {
        const bool is_y = warp == 0;
        const int extent = is_y ? g.H : g.W;
        int lo = 0, hi = 0;
        float wl = 0.0f, wh = 0.0f, v = 0.0f;
        bool ok = false;
        if (lane < NS) {
            v = is_y ? sample_coord(g.sh, g.bh, lane / S, lane % S, S) : sample_coord(g.sw, g.bw, lane / S, lane % S, S);
            ok = sample_1d(v, extent, lo, hi, wl, wh);
        }
        int mn = ok ? lo : (1 << 30), mx = ok ? hi : -1;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            mn = min(mn, __shfl_xor_sync(0xffffffffu, mn, o));
            mx = max(mx, __shfl_xor_sync(0xffffffffu, mx, o));
        }
        const bool any = mx >= 0;
        if (!any) { mn = 0; mx = 0; }
        if (!is_y) mn &= ~3;   // the innermost TMA coordinate must be 16-byte aligned
        if (lane < NS) {
            SampleTap t;
            if (ok) { t.lo = lo - mn; t.hi = hi - mn; t.wl = mul(wl, 0.5f); t.wh = mul(wh, 0.5f); }
            else { t.lo = t.hi = (v < -1.0f) ? 0 : (mx - mn); t.wl = t.wh = 0.0f; }   // keeps the table monotone
            (is_y ? sh.ytab : sh.xtab)[lane] = t;
        }
        if (lane == 0) {
            if (is_y) { sh.y_lo = mn; sh.h_fp = mx - mn + 1; sh.any_y = any; }
            else { sh.x_lo = mn; sh.bw = (mx - mn + 1 + 3) & ~3; sh.any_x = any; }
        }
    }
}

// =====================================================================================================
// forward
// =====================================================================================================
// step 2 of the forward: U[cs][p][BWU] (smem) -> Out[cs][p][q].  lane = (channel-in-round, q); the lane's four
// column taps stay in registers, rows advance by an immediate (BWU is a template parameter).
template <int P, int BWU>
MD_DEVINL void fwd_step2(const float *U, int CPW, const SampleTap t0, const SampleTap t1, float *o, int lane)
{
    constexpr int CSG = 32 / P, PP = P * P;                 // channels per round
    const int csl = lane / P, q = lane - csl * P;
    if (csl >= CSG) return;
    for (int cb = 0; cb < CPW; cb += CSG) {
        const int cs = cb + csl;
        if (cs < CPW) {
            const float *uc = U + cs * P * BWU;
            float *oc = o + cs * PP + q;
#pragma unroll
            for (int p = 0; p < P; p++) {
                float acc = mul(t0.wl, uc[p * BWU + t0.lo]);
                acc = __fmaf_rn(t0.wh, uc[p * BWU + t0.hi], acc);
                acc = __fmaf_rn(t1.wl, uc[p * BWU + t1.lo], acc);
                acc = __fmaf_rn(t1.wh, uc[p * BWU + t1.hi], acc);
                oc[p * P] = acc;
            }
        }
    }
}

template <int P>
__global__ void __launch_bounds__(kStThreads, MD_FWD_CTAS)
roialign_fwd_stream_kernel(const __grid_constant__ TmaMaps maps, const RoiFeat f, const int tma_mask,
                           const float *__restrict__ rois5, const int R, const int seg, const int nchunk,
                           float *__restrict__ out, int32_t *__restrict__ fallback_flag)
{
    pdl_entry();
    constexpr int S = 2, NS = P * S, PP = P * P;
    constexpr int kUFloats = P * 160;                     // max over LPC of CPW * P * (4*LPC + 4)
    extern __shared__ __align__(128) unsigned char dsm[];
    float *ring_all = reinterpret_cast<float *>(dsm);
    StreamShared<P> &sh = *reinterpret_cast<StreamShared<P> *>(ring_all + kStWarps * kRingFloats);

    const WorkItem wi = work_item(blockIdx.x, R, seg, nchunk);
    const int r = wi.r, tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const RoiGeom g = roi_geometry(f, rois5 + (int64_t)r * 5, P);
    const int C = f.C, CH = C / nchunk, cbase = wi.chunk * CH;
    const bool use_tma = (tma_mask >> g.l) & 1;
    const bool flag_writer = tid == 0 && wi.chunk == 0;
    if ((int)__ldg(f.cfg + 1) != S || g.l >= kTmaLevels) {        // uniform: gather path handles this RoI
        if (flag_writer) fallback_flag[r] = 1;
        return;
    }
    if (lane == 0) {
        for (int i = 0; i < kMaxSlots; i++) mbar_init(&sh.full[warp][i], use_tma ? 1 : 32);
        fence_barrier_init();
    }
    build_tables<P>(sh, g);
    __syncthreads();
    const int BW = sh.bw;
    if (BW > kMaxBW) {
        if (flag_writer) fallback_flag[r] = 1;
        return;
    }
    if (flag_writer) fallback_flag[r] = 0;
    float *orow = out + ((int64_t)r * C + cbase) * PP;
    if (!g.ok || !sh.any_x || !sh.any_y) {                          // bad batch index, or every sample out of range -> zeros
        for (int i = tid; i < CH * PP; i += kStThreads) orow[i] = 0.0f;
        return;
    }
    const int LPC = lanes_per_channel(BW), CPW = 32 / LPC, BWU = 4 * LPC + 4;
    const int BWS = fwd_box_width(BW);                               // row pitch of the staged box (>= BW)
    const int lshift = LPC == 4 ? 2 : (LPC == 8 ? 3 : (LPC == 16 ? 4 : 5));
    const int csub = lane >> lshift, xq = lane & (LPC - 1);
    const bool col_ok = 4 * xq < BW;
    const int x_lo = sh.x_lo, y_lo = sh.y_lo, h_fp = sh.h_fp;
    const int nblk = (h_fp + 3) >> 2;
    const int BLK = CPW * 4 * BWS, SLOT = (BLK + 31) & ~31;
    const int ngroups = CH / CPW;                                     // host guarantees CH % 8 == 0
    const int ng_w = (ngroups - warp + kStWarps - 1) / kStWarps;      // groups warp, warp+4, ...
    float *ring = ring_all + warp * kRingFloats;
    unsigned long long *full = sh.full[warp];
    const CUtensorMap *map = &maps.m[g.l * kMapsPerLevel + (BWS == 20 && BW == 16 ? kNumBW : (BWS >> 2) - 1)];
    const int H = g.H, W = g.W;
    const float *fplane = f.feat[g.l] + ((int64_t)g.b * C + cbase) * H * W;
    const int zbase = g.b * C + cbase;
    const int lane_off = csub * 4 * BWS + 4 * xq;          // this lane's float offset inside a row block

    // one row block (BW x 4 rows x CPW channels) of channel group `grp` -> dst, completion on `bar`
    auto load_block = [&](float *dst, int grp, int j, unsigned long long *bar) {
        const int c0 = (warp + kStWarps * grp) * CPW;
        if (use_tma) {
            if (lane == 0) tma_load_3d(dst, map, x_lo, y_lo + 4 * j, zbase + c0, bar);
        } else {
            // elements outside the map are never used with a non-zero weight (rows/cols clamp): left unwritten
            for (int e = lane; e < BLK; e += 32) {
                const int c = e / (4 * BWS), rem = e - c * (4 * BWS);
                const int rr = rem / BWS, xx = rem - rr * BWS;
                const int y = y_lo + 4 * j + rr, x = x_lo + xx;
                if (y < H && x < W) cp_async4(dst + e, fplane + ((int64_t)(c0 + c) * H + y) * W + x);
            }
        }
    };

    // ---- step 2: the lane's output column q and its four taps -------------------------------------------
    const int q2 = lane % P;
    const SampleTap xt0 = sh.xtab[q2 * S], xt1 = sh.xtab[q2 * S + 1];
    auto step2 = [&](float *U, const float (&u)[P][4], int grp) {
        if (col_ok) {
#pragma unroll
            for (int p = 0; p < P; p++)
                *reinterpret_cast<float4 *>(U + (csub * P + p) * BWU + 4 * xq) = make_float4(u[p][0], u[p][1], u[p][2], u[p][3]);
        }
        __syncwarp();
        float *o = orow + (int64_t)(warp + kStWarps * grp) * CPW * PP;
        switch (LPC) {
            case 4: fwd_step2<P, 20>(U, CPW, xt0, xt1, o, lane); break;
            case 8: fwd_step2<P, 36>(U, CPW, xt0, xt1, o, lane); break;
            case 16: fwd_step2<P, 68>(U, CPW, xt0, xt1, o, lane); break;
            default: fwd_step2<P, 132>(U, CPW, xt0, xt1, o, lane); break;
        }
        __syncwarp();
    };

    const int TILE = nblk * SLOT;
    const int SP = (max(TILE, CPW * P * BWU) + 31) & ~31;   // stage pitch: a consumed stage doubles as the U buffer
    // tile mode even when only ONE stage fits: the other warps of the SM cover the exposed load latency, and the
    // row-block bookkeeping of stream mode costs ~2.3x the instructions per pass (measured 322 -> 287 us)
    if (SP <= kRingFloats) {
        // =============== tile mode: the whole footprint of a channel group is one pipeline stage ===============
        const int NST = min(4, kRingFloats / SP);
        // per-sample float offsets inside a tile (row block, row in block); same for every group
        SampleTap *yoff = sh.yoff;
        if (warp == 0 && lane < NS) {
            const SampleTap t = sh.ytab[lane];
            SampleTap o;
            o.lo = (t.lo >> 2) * SLOT + (t.lo & 3) * BWS; o.hi = (t.hi >> 2) * SLOT + (t.hi & 3) * BWS; o.wl = t.wl; o.wh = t.wh;
            yoff[lane] = o;
        }
        __syncthreads();
        auto issue_tile = [&](int grp, int st) {
            if (use_tma && lane == 0) mbar_expect_tx(&full[st], (uint32_t)(nblk * BLK) * 4u);
            for (int j = 0; j < nblk; j++) load_block(ring + st * SP + j * SLOT, grp, j, &full[st]);
            if (!use_tma) cp_async_mbar_arrive(&full[st]);
        };
        for (int gi = 0; gi < NST && gi < ng_w; gi++) issue_tile(gi, gi);
        int st = 0;
        uint32_t par = 0;
        for (int gi = 0; gi < ng_w; gi++) {
            mbar_wait(&full[st], par);
            float *stage = ring + st * SP;
            const float *tile = stage + lane_off;
            float u[P][4];
#pragma unroll
            for (int p = 0; p < P; p++) u[p][0] = u[p][1] = u[p][2] = u[p][3] = 0.0f;
            if (col_ok) {
#pragma unroll
                for (int s = 0; s < NS; s++) {
                    const SampleTap t = yoff[s];
                    const float4 a = *reinterpret_cast<const float4 *>(tile + t.lo);
                    const float4 b = *reinterpret_cast<const float4 *>(tile + t.hi);
                    u[s / S][0] = __fmaf_rn(t.wl, a.x, __fmaf_rn(t.wh, b.x, u[s / S][0]));
                    u[s / S][1] = __fmaf_rn(t.wl, a.y, __fmaf_rn(t.wh, b.y, u[s / S][1]));
                    u[s / S][2] = __fmaf_rn(t.wl, a.z, __fmaf_rn(t.wh, b.z, u[s / S][2]));
                    u[s / S][3] = __fmaf_rn(t.wl, a.w, __fmaf_rn(t.wh, b.w, u[s / S][3]));
                }
            }
            __syncwarp();                                   // every lane is done reading this stage
            step2(stage, u, gi);                            // the consumed stage is the U buffer
            if (gi + NST < ng_w) issue_tile(gi + NST, st);
            if (++st == NST) { st = 0; par ^= 1u; }
        }
        return;
    }

    // =============== stream mode: row blocks slide through the ring (any footprint height) ===============
    float *Ubuf = ring + kRingFloats - kUFloats;
    const int NB = min(kMaxSlots, (kRingFloats - kUFloats) / SLOT);
    const int total = ng_w * nblk;
    int iss = 0, iss_g = 0, iss_j = 0, iss_slot = 0;
    auto issue_one = [&]() {
        __syncwarp();                                   // every lane is done reading the slot's previous block
        if (use_tma && lane == 0) mbar_expect_tx(&full[iss_slot], (uint32_t)BLK * 4u);
        load_block(ring + iss_slot * SLOT, iss_g, iss_j, &full[iss_slot]);
        if (!use_tma) cp_async_mbar_arrive(&full[iss_slot]);
        iss++;
        if (++iss_j == nblk) { iss_j = 0; iss_g++; }
        if (++iss_slot == NB) iss_slot = 0;
    };
    int wt = 0, wt_slot = 0;
    uint32_t wt_par = 0;
    int grp_slot = 0;                                   // slot of the current group's block 0
    // Make every block <= need visible.  A slot may be refilled only when its previous block has been both
    // waited for (its mbarrier phase observed) and consumed (every block below `freed` is dead).
    auto advance = [&](int need, int freed) {
        for (;;) {
            const int lim = min(freed, wt) + NB;
            while (iss < total && iss < lim) issue_one();
            if (wt > need) break;
            mbar_wait(&full[wt_slot], wt_par);
            wt++;
            if (++wt_slot == NB) { wt_slot = 0; wt_par ^= 1u; }
        }
    };
    for (int gi = 0; gi < ng_w; gi++) {
        const int gbase = gi * nblk;
        float u[P][4];
#pragma unroll
        for (int p = 0; p < P; p++) u[p][0] = u[p][1] = u[p][2] = u[p][3] = 0.0f;
        int cur_blk = 0, cur_slot = grp_slot;
#pragma unroll 1
        for (int p = 0; p < P; p++) {
            float a0 = 0.0f, a1 = 0.0f, a2 = 0.0f, a3 = 0.0f;
#pragma unroll
            for (int i = 0; i < S; i++) {
                const SampleTap t = sh.ytab[p * S + i];
                const int blo = t.lo >> 2, bhi = t.hi >> 2;
                while (cur_blk < blo) { cur_blk++; if (++cur_slot == NB) cur_slot = 0; }
                advance(gbase + bhi, gbase + blo);
                int hi_slot = cur_slot;
                if (bhi != blo) { hi_slot = cur_slot + 1; if (hi_slot == NB) hi_slot = 0; }
                if (col_ok) {
                    const float4 a = *reinterpret_cast<const float4 *>(ring + cur_slot * SLOT + (t.lo & 3) * BWS + lane_off);
                    const float4 b = *reinterpret_cast<const float4 *>(ring + hi_slot * SLOT + (t.hi & 3) * BWS + lane_off);
                    a0 = __fmaf_rn(t.wl, a.x, __fmaf_rn(t.wh, b.x, a0));
                    a1 = __fmaf_rn(t.wl, a.y, __fmaf_rn(t.wh, b.y, a1));
                    a2 = __fmaf_rn(t.wl, a.z, __fmaf_rn(t.wh, b.z, a2));
                    a3 = __fmaf_rn(t.wl, a.w, __fmaf_rn(t.wh, b.w, a3));
                }
            }
            // static register index without unrolling the bookkeeping P times
#pragma unroll
            for (int pp = 0; pp < P; pp++)
                if (pp == p) { u[pp][0] = a0; u[pp][1] = a1; u[pp][2] = a2; u[pp][3] = a3; }
        }
        grp_slot += nblk;
        while (grp_slot >= NB) grp_slot -= NB;
        advance(gbase + nblk - 1, gbase + nblk);           // the whole group is consumed: keep the producer ahead
        step2(Ubuf, u, gi);
    }
}

// =====================================================================================================
// backward
// =====================================================================================================
template <int P>
struct BwdShared {
    StreamShared<P> st;
    float4 row_w[kMaxRowsBwd];         // weights of bins row_p .. row_p+3
    int row_p[kMaxRowsBwd];            // first contributing bin of each footprint row (-1: none)
    int row_p2[kMaxRowsBwd];           // same with empty rows filled in
    float ay_dense[16][16];            // rows x bins (padded), only when some row has more than 4 contributing bins
    int rs[P + 1];                     // rows [rs[p], rs[p+1]) have first contributing bin p (row_p is non-decreasing)
    int dense;
};

// d += w * T[PA + K] when that bin exists (compile-time guard keeps the register index static)
#define MD_ACC(K, WV)                                                          \
    if (PA + K < P) {                                                          \
        d0 = __fmaf_rn(WV, T[PA + K < P ? PA + K : 0][0], d0);                 \
        d1 = __fmaf_rn(WV, T[PA + K < P ? PA + K : 0][1], d1);                 \
        d2 = __fmaf_rn(WV, T[PA + K < P ? PA + K : 0][2], d2);                 \
        d3 = __fmaf_rn(WV, T[PA + K < P ? PA + K : 0][3], d3);                 \
    }
template <int P, int PA>
MD_DEVINL void row_from_bins(const float (&T)[P][4], const float4 w, float &d0, float &d1, float &d2, float &d3)
{
    d0 = mul(w.x, T[PA][0]); d1 = mul(w.x, T[PA][1]); d2 = mul(w.x, T[PA][2]); d3 = mul(w.x, T[PA][3]);
    MD_ACC(1, w.y)
    MD_ACC(2, w.z)
    MD_ACC(3, w.w)
}
#undef MD_ACC

// walks the footprint rows bin by bin: rows [rs[PA], rs[PA+1]) have first contributing bin PA (static register index)
template <int P, int PA>
struct BinWalk {
    template <class Emit>
    static MD_DEVINL void run(const float (&T)[P][4], const int *rs, const float4 *row_w, int &y, Emit &emit)
    {
        for (const int ye = rs[PA + 1]; y < ye; y++) {
            const float4 w = row_w[y];
            float d0, d1, d2, d3;
            row_from_bins<P, PA>(T, w, d0, d1, d2, d3);
            emit(y, d0, d1, d2, d3);
        }
        BinWalk<P, PA + 1>::run(T, rs, row_w, y, emit);
    }
};
template <int P>
struct BinWalk<P, P> {
    template <class Emit>
    static MD_DEVINL void run(const float (&)[P][4], const int *, const float4 *, int &, Emit &) {}
};

template <int P>
__global__ void __launch_bounds__(kStThreads, P <= 8 ? 4 : 2)
roialign_bwd_stream_kernel(const __grid_constant__ TmaMaps maps, const RoiFeat f, const int tma_mask,
                           const float *__restrict__ rois5, const int R, const int seg, const int nchunk,
                           const float *__restrict__ dout, int32_t *__restrict__ fallback_flag)
{
    pdl_entry();
    static_assert(P == 7 || P == 14, "backward stream kernel: 7x7 (box head) and 14x14 (mask head)");
    constexpr int S = 2, NS = P * S, PP = P * P;
    constexpr int QP = P <= 8 ? 8 : 16;                   // dY rows padded to a multiple of 4 floats
    constexpr int NBUF = P <= 8 ? 2 : 1;                  // dY staging buffers per warp (one for 14x14: shared memory)
    constexpr int kGFloats = 8 * P * QP;                  // CPW(max 8) x P x QP
    constexpr int kAxFloats = P * 132;                    // Ax dense [q][<=128 + 4]
    extern __shared__ __align__(128) unsigned char dsm[];
    float *ring_all = reinterpret_cast<float *>(dsm);
    float *G_all = ring_all + kStWarps * kBwdRingFloats; // [warp][NBUF][kGFloats]
    float *AxD = G_all + kStWarps * NBUF * kGFloats;     // [q][BWA]
    BwdShared<P> &bs = *reinterpret_cast<BwdShared<P> *>(AxD + kAxFloats);
    StreamShared<P> &sh = bs.st;

    const WorkItem wi = work_item(blockIdx.x, R, seg, nchunk);
    const int r = wi.r, tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const RoiGeom g = roi_geometry(f, rois5 + (int64_t)r * 5, P);
    const int C = f.C, CH = C / nchunk, cbase = wi.chunk * CH;
    const bool use_tma = (tma_mask >> g.l) & 1;
    const bool flag_writer = tid == 0 && wi.chunk == 0;
    if ((int)__ldg(f.cfg + 1) != S || g.l >= kTmaLevels) {
        if (flag_writer) fallback_flag[r] = 1;
        return;
    }
    build_tables<P>(sh, g);
    if (tid == 0) bs.dense = 0;
    __syncthreads();
    const int BW = sh.bw, h_fp = sh.h_fp;
    if (BW > kMaxBW || h_fp > kMaxRowsBwd) {
        if (flag_writer) fallback_flag[r] = 1;
        return;
    }
    if (!g.ok || !sh.any_x || !sh.any_y) {                          // bad batch index / no sample in range -> no gradient
        if (flag_writer) fallback_flag[r] = 0;
        return;
    }
    const int BWA = BW + 4;
    // ---- dense Ax[q][x] (x relative to x_lo) and the per-row bin lists --------------------------------
    for (int i = tid; i < P * BWA; i += kStThreads) AxD[i] = 0.0f;
    __syncthreads();
    if (tid < P) {
        const SampleTap t0 = sh.xtab[tid * S], t1 = sh.xtab[tid * S + 1];
        float *a = AxD + tid * BWA;
        a[t0.lo] += t0.wl; a[t0.hi] += t0.wh; a[t1.lo] += t1.wl; a[t1.hi] += t1.wh;
    }
    for (int y = tid; y < h_fp; y += kStThreads) {
        float w[P];
#pragma unroll
        for (int p = 0; p < P; p++) w[p] = 0.0f;
#pragma unroll
        for (int s = 0; s < NS; s++) {
            const SampleTap t = sh.ytab[s];
            if (t.lo == y) w[s / S] += t.wl;
            if (t.hi == y) w[s / S] += t.wh;
        }
        int pa = P, pz = -1;
#pragma unroll
        for (int p = P - 1; p >= 0; p--) if (w[p] != 0.0f) pa = p;
#pragma unroll
        for (int p = 0; p < P; p++) if (w[p] != 0.0f) pz = p;
        if (pz < 0) { pa = -1; pz = -1; }                            // row without contribution (fixed up below)
        float4 ww = make_float4(0, 0, 0, 0);
#pragma unroll
        for (int p = 0; p < P; p++) {
            if (p == pa) ww.x = w[p];
            if (p == pa + 1) ww.y = w[p];
            if (p == pa + 2) ww.z = w[p];
            if (p == pa + 3) ww.w = w[p];
        }
        bs.row_p[y] = pa;
        bs.row_w[y] = ww;
        if (pz - pa > 3) bs.dense = 1;                               // benign race: every writer stores 1
        if (y < 16) {
#pragma unroll
            for (int p = 0; p < 16; p++) bs.ay_dense[y][p] = p < P ? w[p] : 0.0f;
        }
    }
    if (tid <= P) bs.rs[tid] = h_fp;
    __syncthreads();
    // empty rows inherit the previous row's first bin (weights are zero), which keeps row_p non-decreasing;
    // rs[p] = first row whose first bin is >= p
    for (int y = tid; y < h_fp; y += kStThreads) {
        int pa = bs.row_p[y], yy = y;
        while (pa < 0 && yy > 0) pa = bs.row_p[--yy];
        if (pa < 0) pa = 0;
        int prev = -1, yp = y - 1;
        if (yp >= 0) { prev = bs.row_p[yp]; while (prev < 0 && yp > 0) prev = bs.row_p[--yp]; if (prev < 0) prev = 0; }
        for (int q = prev + 1; q <= pa; q++) bs.rs[q] = y;
        bs.row_p2[y] = pa;
    }
    __syncthreads();
    const bool dense = bs.dense != 0;
    if (flag_writer) fallback_flag[r] = (dense && h_fp > 16) ? 1 : 0;
    if (dense && h_fp > 16) return;                                   // cannot happen for bins < 1 row; be safe

    const int LPC = lanes_per_channel(BW), CPW = 32 / LPC;
    const int lshift = LPC == 4 ? 2 : (LPC == 8 ? 3 : (LPC == 16 ? 4 : 5));
    const int csub = lane >> lshift, xq = lane & (LPC - 1);
    const bool col_ok = 4 * xq < BW;
    const int x_lo = sh.x_lo, y_lo = sh.y_lo;
    const int nblk = (h_fp + 3) >> 2;
    const int BLK = CPW * 4 * BW, SLOT = (BLK + 31) & ~31;
    const int ngroups = CH / CPW;
    const int ng_w = (ngroups - warp + kStWarps - 1) / kStWarps;
    float *ring = ring_all + warp * kBwdRingFloats;
    float *Gs = G_all + warp * NBUF * kGFloats;
    const CUtensorMap *map = &maps.m[g.l * kMapsPerLevel + (BW >> 2) - 1];
    const int H = g.H, W = g.W;
    float *dplane = f.feat[g.l] + ((int64_t)g.b * C + cbase) * H * W;
    const float *grow = dout + ((int64_t)r * C + cbase) * PP;
    const int zbase = g.b * C + cbase;

    const int lane_off = csub * 4 * BW + 4 * xq;

    float ax[P][4];
#pragma unroll
    for (int q = 0; q < P; q++) {
        const float4 v = col_ok ? *reinterpret_cast<const float4 *>(AxD + q * BWA + 4 * xq) : make_float4(0, 0, 0, 0);
        ax[q][0] = v.x; ax[q][1] = v.y; ax[q][2] = v.z; ax[q][3] = v.w;
    }

    auto stage_g = [&](int gi) {                          // dY of group gi -> Gs[gi % NBUF][cs][p][QP]
        const int c0 = (warp + kStWarps * gi) * CPW;
        float *dst = Gs + (gi % NBUF) * kGFloats;
        const float *src = grow + (int64_t)c0 * PP;
        for (int e = lane; e < CPW * PP; e += 32) cp_async4(dst + e + (e / P) * (QP - P), src + e);   // rows of P padded to QP
        cp_async_commit_group();
    };

    int slot = 0;
    if (ng_w > 0) stage_g(0);
    for (int gi = 0; gi < ng_w; gi++) {
        const int c0 = (warp + kStWarps * gi) * CPW;
        if (NBUF == 2 && gi + 1 < ng_w) { stage_g(gi + 1); cp_async_wait_group<1>(); } else { cp_async_wait_group<0>(); }
        __syncwarp();
        // ---- step 1: T[p][4 cols] = sum_q dY[p][q] * Ax[q][cols] -----------------------------------------
        float T[P][4];
        {
            const float *gs = Gs + (gi % NBUF) * kGFloats + csub * P * QP;
#pragma unroll
            for (int p = 0; p < P; p++) {
                float gq[QP];
#pragma unroll
                for (int v4 = 0; v4 < QP / 4; v4++) {
                    const float4 t = *reinterpret_cast<const float4 *>(gs + p * QP + 4 * v4);
                    gq[4 * v4] = t.x; gq[4 * v4 + 1] = t.y; gq[4 * v4 + 2] = t.z; gq[4 * v4 + 3] = t.w;
                }
#pragma unroll
                for (int k = 0; k < 4; k++) {
                    float acc = mul(gq[0], ax[0][k]);
#pragma unroll
                    for (int q = 1; q < P; q++) acc = __fmaf_rn(gq[q], ax[q][k], acc);
                    T[p][k] = acc;
                }
            }
        }
        if (NBUF == 1 && gi + 1 < ng_w) { __syncwarp(); stage_g(gi + 1); }      // single buffer: refill once step 1 has read it
        // ---- step 2: D[row][4 cols] = sum_p Ay[p][row] * T[p].  Rows are walked bin by bin (row_p is
        //      non-decreasing, so the register index of T stays static and no per-row dispatch is needed).
        //      TMA levels: rows go conflict-free into a ring slot; every completed 4-row block is folded into
        //      dX by ONE TMA reduce-add (at most kBwdSlots in flight).  Other levels: red.global per element. ----
        auto emit_row = [&](int y, float d0, float d1, float d2, float d3) {
            if (use_tma) {
                float *dst = ring + slot * SLOT + lane_off;
                if (col_ok) *reinterpret_cast<float4 *>(dst + (y & 3) * BW) = make_float4(d0, d1, d2, d3);
                if ((y & 3) == 3 || y == h_fp - 1) {
                    for (int z = (y & 3) + 1; z < 4; z++)
                        if (col_ok) *reinterpret_cast<float4 *>(dst + z * BW) = make_float4(0, 0, 0, 0);
                    fence_proxy_async();
                    __syncwarp();
                    if (++slot == kBwdSlots) slot = 0;
                    if (lane == 0) {
                        tma_reduce_add_3d(map, x_lo, y_lo + (y & ~3), zbase + c0, dst - lane_off);
                        bulk_commit();
                        bulk_wait_read<kBwdSlots - 1>();             // the reduce that last read the NEXT slot is done
                    }
                    __syncwarp();
                }
            } else if (col_ok && y_lo + y < H) {
                const int xa = x_lo + 4 * xq;
                float *dp = dplane + ((int64_t)(c0 + csub) * H + (y_lo + y)) * W + xa;
                if (xa < W && d0 != 0.0f) atomicAdd(dp, d0);
                if (xa + 1 < W && d1 != 0.0f) atomicAdd(dp + 1, d1);
                if (xa + 2 < W && d2 != 0.0f) atomicAdd(dp + 2, d2);
                if (xa + 3 < W && d3 != 0.0f) atomicAdd(dp + 3, d3);
            }
        };
        if (!dense) {
            int y = 0;
            BinWalk<P, 0>::run(T, bs.rs, bs.row_w, y, emit_row);
        } else {
            for (int y = 0; y < h_fp; y++) {
                float wp[16];
#pragma unroll
                for (int v4 = 0; v4 < 4; v4++) {
                    const float4 t = *reinterpret_cast<const float4 *>(&bs.ay_dense[y][4 * v4]);
                    wp[4 * v4] = t.x; wp[4 * v4 + 1] = t.y; wp[4 * v4 + 2] = t.z; wp[4 * v4 + 3] = t.w;
                }
                float d0 = 0.0f, d1 = 0.0f, d2 = 0.0f, d3 = 0.0f;
#pragma unroll
                for (int p = 0; p < P; p++) {
                    d0 = __fmaf_rn(wp[p], T[p][0], d0); d1 = __fmaf_rn(wp[p], T[p][1], d1);
                    d2 = __fmaf_rn(wp[p], T[p][2], d2); d3 = __fmaf_rn(wp[p], T[p][3], d3);
                }
                emit_row(y, d0, d1, d2, d3);
            }
        }
        __syncwarp();                                     // the dY buffer may be overwritten by the next stage_g
    }
    if (use_tma && lane == 0) bulk_wait_all<0>();
}

// =====================================================================================================
// host: tensor maps (one per (level, box width)); small LRU keyed by (pointers, dims)
// =====================================================================================================
EncodeTiledFn get_encode()
{
    static EncodeTiledFn fn = nullptr;
    static std::once_flag once;
    std::call_once(once, [] {
        void *p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
            q == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiledFn>(p);
    });
    return fn;
}

struct MapCache {
    void *ptr[kTmaLevels]; int H[kTmaLevels], W[kTmaLevels], BC, L, mask; TmaMaps maps; bool valid; unsigned long long stamp;
};
constexpr int kCacheEntries = 32;
static MapCache g_cache[kCacheEntries];
static unsigned long long g_stamp = 0;
static std::mutex g_cache_mutex;

// Returns the TMA-eligible level mask (0 when the driver entry point is missing: cp.async path for all).
static int build_maps(const FeatSet &fs, TmaMaps *out)
{
    EncodeTiledFn enc = get_encode();
    const char *pe = getenv("MD_ROI_L2PROMO");
    const int pv = pe ? atoi(pe) : 128;
    const CUtensorMapL2promotion l2promo = pv == 0 ? CU_TENSOR_MAP_L2_PROMOTION_NONE : (pv == 64 ? CU_TENSOR_MAP_L2_PROMOTION_L2_64B : (pv == 256 ? CU_TENSOR_MAP_L2_PROMOTION_L2_256B : CU_TENSOR_MAP_L2_PROMOTION_L2_128B));
    std::lock_guard<std::mutex> lock(g_cache_mutex);
    const int L = fs.L < kTmaLevels ? fs.L : kTmaLevels;
    MapCache *hit = nullptr, *victim = &g_cache[0];
    for (int e = 0; e < kCacheEntries; e++) {
        MapCache &c = g_cache[e];
        bool same = c.valid && c.L == L && c.BC == fs.B * fs.C;
        for (int l = 0; same && l < L; l++) same = c.ptr[l] == fs.feat[l] && c.H[l] == fs.H[l] && c.W[l] == fs.W[l];
        if (same) { hit = &c; break; }
        if (!c.valid) { if (victim->valid) victim = &c; }
        else if (victim->valid && c.stamp < victim->stamp) victim = &c;
    }
    if (!hit) {
        MapCache &c = *victim;
        std::memset(&c.maps, 0, sizeof(c.maps));
        c.mask = 0;
        for (int l = 0; l < L; l++) {
            c.ptr[l] = fs.feat[l]; c.H[l] = fs.H[l]; c.W[l] = fs.W[l];
            if (!enc || (fs.W[l] & 3) || (reinterpret_cast<uintptr_t>(fs.feat[l]) & 15)) continue;   // cp.async path
            bool ok = true;
            for (int k = 0; k < kNumBW && ok; k++) {
                const int bw = 4 * (k + 1);
                const cuuint64_t dims[3] = { (cuuint64_t)fs.W[l], (cuuint64_t)fs.H[l], (cuuint64_t)fs.B * fs.C };
                const cuuint64_t strides[2] = { (cuuint64_t)fs.W[l] * 4, (cuuint64_t)fs.W[l] * fs.H[l] * 4 };
                const cuuint32_t box[3] = { (cuuint32_t)bw, 4u, (cuuint32_t)(32 / lanes_per_channel(bw)) };
                const cuuint32_t estr[3] = { 1, 1, 1 };
                ok = enc(&c.maps.m[l * kMapsPerLevel + k], CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, fs.feat[l], dims, strides, box, estr,
                         CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, l2promo,
                         CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
            }
            if (ok) {
                // the padded forward box for 16-column footprints: 20 columns but still 8 channels (see fwd_box_width)
                const cuuint64_t dims[3] = { (cuuint64_t)fs.W[l], (cuuint64_t)fs.H[l], (cuuint64_t)fs.B * fs.C };
                const cuuint64_t strides[2] = { (cuuint64_t)fs.W[l] * 4, (cuuint64_t)fs.W[l] * fs.H[l] * 4 };
                const cuuint32_t box[3] = { 20u, 4u, 8u };
                const cuuint32_t estr[3] = { 1, 1, 1 };
                ok = enc(&c.maps.m[l * kMapsPerLevel + kNumBW], CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, fs.feat[l], dims, strides, box,
                         estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, l2promo,
                         CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
            }
            if (ok) c.mask |= 1 << l;
        }
        c.L = L; c.BC = fs.B * fs.C; c.valid = true;
        hit = &c;
    }
    hit->stamp = ++g_stamp;
    *out = hit->maps;
    return hit->mask;
}

// channel chunks per RoI: 128 channels per CTA when C allows it (working set of one (image, chunk) sweep
// = 128 planes of every level ~ 46 MB at config 2, inside the 126 MB L2; measured best of 32/64/128/256)
constexpr int kSegRois = 512;
static int chunks_for(int C)
{
    const char *e = getenv("MD_ROI_CHUNK");
    const int want = e ? atoi(e) : 128;
    int n = C / want;
    while (n > 1 && (C % n != 0 || (C / n) % 8 != 0)) n--;
    return n < 1 ? 1 : n;
}

template <int P> static size_t fwd_smem()
{
    return (size_t)(kStWarps * kRingFloats) * sizeof(float) + sizeof(StreamShared<P>) + 128;
}
template <int P> static size_t bwd_smem()
{
    return (size_t)(kStWarps * kBwdRingFloats + kStWarps * (P <= 8 ? 2 : 1) * 8 * P * (P <= 8 ? 8 : 16) + P * 132) * sizeof(float) +
           sizeof(BwdShared<P>) + 128;
}

template <int P>
static cudaError_t launch_fwd(const TmaMaps &maps, const RoiFeat &f, int mask, const float *rois5, int R, float *out,
                              int32_t *flag, cudaStream_t s)
{
    auto kern = roialign_fwd_stream_kernel<P>;
    {   // per device, constant value: set on every launch
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)fwd_smem<P>());
        if (e != cudaSuccess) return e;
    }
    const int nchunk = chunks_for(f.C);
    return launch_pdl(kern, dim3(R * nchunk), dim3(kStThreads), fwd_smem<P>(), s, maps, f, mask, rois5, R, kSegRois, nchunk, out, flag);
}

cudaError_t launch_roialign_fwd_tma(const FeatSet &fs, const RoiFeat &f, const float *rois5, int R, int P,
                                    float *out, int32_t *fallback_flag, cudaStream_t s, bool *launched)
{
    *launched = false;
    if ((fs.C & 7) || (P != 7 && P != 14)) return cudaSuccess;        // gather path
    TmaMaps maps;
    const int mask = build_maps(fs, &maps);
    cudaError_t e = P == 7 ? launch_fwd<7>(maps, f, mask, rois5, R, out, fallback_flag, s)
                           : launch_fwd<14>(maps, f, mask, rois5, R, out, fallback_flag, s);
    *launched = e == cudaSuccess;
    return e;
}

template <int P>
static cudaError_t launch_bwd(const TmaMaps &maps, const RoiFeat &f, int mask, const float *rois5, int R, const float *dout,
                              int32_t *flag, cudaStream_t s)
{
    auto kern = roialign_bwd_stream_kernel<P>;
    {   // per device, constant value: set on every launch
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bwd_smem<P>());
        if (e != cudaSuccess) return e;
    }
    const int nchunk = chunks_for(f.C);
    return launch_pdl(kern, dim3(R * nchunk), dim3(kStThreads), bwd_smem<P>(), s, maps, f, mask, rois5, R, kSegRois, nchunk, dout, flag);
}

cudaError_t launch_roialign_bwd_tma(const FeatSet &fs, const RoiFeat &f, const float *rois5, int R, int P,
                                    const float *dout, int32_t *fallback_flag, cudaStream_t s, bool *launched)
{
    *launched = false;
    if ((fs.C & 7) || (P != 7 && P != 14)) return cudaSuccess;
    TmaMaps maps;
    const int mask = build_maps(fs, &maps);
    cudaError_t e = P == 7 ? launch_bwd<7>(maps, f, mask, rois5, R, dout, fallback_flag, s)
                           : launch_bwd<14>(maps, f, mask, rois5, R, dout, fallback_flag, s);
    *launched = e == cudaSuccess;
    return e;
}

}  // namespace md
