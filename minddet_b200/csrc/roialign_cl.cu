// roialign_cl.cu -- a10/a11 fast path, second generation: CHANNEL-PER-LANE RoIAlign (7x7, S = 2).
//
// RoIAlign is the same small linear operator for every channel of a RoI: Out_c = Ay . F_c . Ax^T with the bilinear
// sample matrices Ay (7 x rows), Ax (7 x cols) shared by all C channels.  So a warp takes 32 CHANNELS of one RoI, one per
// lane: every tap position, weight, loop bound and branch is warp-uniform (no idle lanes, no divergence, no per-lane
// address arithmetic), and each lane runs two tiny dense products out of registers.
//
//   * footprint staging: one cp.async.bulk.tensor.3d per footprint row, box = {4*NQ columns, 1 row, 32 channels}, NCHW ->
//     shared memory [row][channel][4*NQ'].  Lane c reads its own channel with LDS.128; the channel pitch is kept an odd
//     number of quads (NQ' = NQ or NQ + 1), which makes those reads bank-conflict-free.  Only rows that carry a
//     bilinear tap are loaded (<= 2 per sample = 28 per RoI, whatever its height).
//   * forward: row-stationary y-step U[p][x] += Ay[p][row] * F[row][x] (each staged row is read ONCE; rows are walked in
//     the order of their first bin so the register index p stays static), then x-step Out[p][q] = sum_x Ax[q][x] U[p][x]
//     straight out of registers; the 32 x 49 outputs of a warp leave through shared memory as ONE 6272-byte bulk store.
//   * backward: the transpose -- dY arrives by one bulk load, T[p][x] = sum_q dY[p][q] Ax[q][x], rows D = Ay^T T are written
//     into the staging slot and folded into dX by one cp.reduce.async.bulk.tensor (add, at L2) per row.
//   * big RoIs are cut into pieces (bins p0..p1 x column chunks of <= 16) that fit a 14 KB stage; a persistent 1-warp CTA
//     keeps 3 stages in flight and takes (RoI, channel chunk) items from a global ticket in image order, so the RoIs
//     resident at any time read the same planes out of L2.
//
// No reference code exists for this op; semantics: oracle/CONVENTIONS.md #14-16, checked against
// oracle/region_oracle.c:o_roialign_fwd / o_roialign_bwd (rtol 1e-5: summation order differs, FMAs are used).
#include <cuda.h>

#include <climits>
#include <cstdlib>
#include <cstring>
#include <mutex>

#include "kernels.h"
#include "roialign_common.cuh"
#include "tma_host.h"
#include "tma_ptx.cuh"

namespace md {

#ifndef MD_CL_SLOTS
#define MD_CL_SLOTS 3
#endif
#ifndef MD_CL_SLOT_KB
#define MD_CL_SLOT_KB 13
#endif
#ifndef MD_CL_CTAS
#define MD_CL_CTAS 4
#endif
constexpr int kClSlots = MD_CL_SLOTS;
constexpr int kClSlotBytes = MD_CL_SLOT_KB * 1024;
constexpr int kClRows = 28;                 // compact footprint rows per RoI: 2 per sample, 14 samples
constexpr int kClMaxQuads = 16;             // aligned footprint width <= 64 columns
constexpr int kClLevels = 4;
constexpr int kClP = 7;
constexpr int kClStageFloats = 32 * kClP * kClP;       // 32 channels x 49 outputs
constexpr int kClStageBytes = kClStageFloats * 4;      // 6272


// Staged rows are [row][channel][4*NQ floats], unpadded (the TMA box).  Lane c reads channel c with LDS.128; a quarter
// warp (8 lanes) must hit 8 distinct 16-byte slots of a 128-byte bank row.  With a channel pitch of 32 / 64 bytes lanes
// c and c+4 / c+2 would collide, so every lane walks the quads in a ROTATED order, quad (k + rot(c)) % NQ at step k:
// conflict-free without padding or swizzling (the 128-byte TMA swizzle returned wrong data for 32-byte rows, measured).
// The accumulators therefore hold the columns in a lane-dependent order; they are un-permuted for free when they are
// parked in the lane's scratch column (computed addresses).
__host__ __device__ constexpr int cl_pitch(int nq) { return 16 * nq; }
__host__ __device__ constexpr int cl_row_bytes(int nq) { return 32 * cl_pitch(nq); }
MD_DEVINL int cl_rot(int nq, int lane) { return nq == 4 ? (lane >> 1) & 3 : (nq == 2 ? (lane >> 2) & 1 : 0); }

// Tensor maps view a level as {W, B*C, H} (x, channel plane, row): the box {4*nq, 32 channels, R rows} lands in shared
// memory as [row][channel][4*nq] -- R consecutive footprint rows of 32 channels in ONE bulk-tensor operation.
constexpr int kClRsel = 4;                  // R = 1, 2, 4, 8
struct ClMaps { CUtensorMap m[kClLevels * 4 * kClRsel]; };       // [level][nq - 1][log2 R]

constexpr int kClMaxChunks = 4;             // column chunks of <= 4 quads
constexpr int kClMaxPieces = 28;            // (chunk, bin range) pieces per (RoI, channel group)
constexpr int kClMaxOps = 36;               // bulk-tensor row operations per RoI (rows + one per extra piece)

// one bin: compact rows [ja, ja + nr), nr <= 4, and their weights (1/S folded in, both samples summed)
struct __align__(16) ClBin { int ja, nr, pad0, pad1; float w[4]; };
// one bilinear output column q of one column chunk: 4 taps = byte offsets into the lane's scratch column + weights
struct __align__(16) ClTap { int o[4]; float w[4]; };
// one staged piece: columns [x0, x0 + 4 nqc) of compact rows [jA, jA + rows), producing bins [p0, p1);
// its rows arrive as the bulk-tensor operations op[op0 .. op0 + nop)
struct __align__(16) ClPiece { int x0, nqc, jA, rows, p0, p1, chunk, ops; };   // ops = op0 | nop << 16
struct ClOp { short j, lr; };               // rows j .. j + (1 << lr) - 1 of the compact list (consecutive feature rows)

enum { CL_OK = 0, CL_ZERO = 1, CL_DECLINE = 2 };

// Everything the kernels need to know about one RoI, built ONCE by roialign_plan_kernel (one warp per RoI) and then
// pulled into shared memory by one bulk copy per (RoI, channel chunk) item, one item ahead of its use.
struct __align__(16) ClItem {
    ClBin bin[8];
    ClTap tap[kClMaxChunks][8];      // Ax, sparse, per column chunk
    ClPiece piece[kClMaxPieces];
    ClOp op[kClMaxOps];
    int yof[32];                     // feature row of compact row j
    int xlo[16], xhi[16];            // x samples, columns relative to x_lo (the backward's dense Ax is built from these)
    float xwl[16], xwh[16];
    int status, b, l, x_lo, nq, nrows, npiece, nxc;
};
static_assert(sizeof(ClItem) % 16 == 0, "bulk copies move multiples of 16 bytes");
struct ClBuild { float ew[kClRows][8]; int ja[8], jb[8]; };     // scratch of cl_build_item

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

// ---- per-RoI geometry (one whole warp): lanes 0..13 own the y samples, lanes 16..29 the x samples ------------------
template <int P>
MD_DEVINL int cl_build_item(ClItem &it, ClBuild &bs, const RoiFeat &f, int tma_mask, const float *__restrict__ rois5, int r,
                            int slot_bytes, int lane)
{
    constexpr int S = 2, NS = P * S;
    static_assert(NS <= 16, "one half-warp per axis");
    const float *roi = rois5 + (int64_t)r * 5;
    const RoiGeom g = roi_geometry(f, roi, P);
    if ((int)__ldg(f.cfg + 1) != S || g.l >= kClLevels || !((tma_mask >> g.l) & 1)) return CL_DECLINE;
    if (!g.ok) return CL_ZERO;                                      // batch index out of range (or NaN): no data
    const bool is_y = lane < 16;
    const int s = lane & 15;
    bool ok = false;
    int lo = 0, hi = 0;
    float wl = 0.0f, wh = 0.0f;
    if (s < NS) {
        const float v = sample_coord(is_y ? g.sh : g.sw, is_y ? g.bh : g.bw, s / S, s % S, S);
        ok = sample_1d(v, is_y ? g.H : g.W, lo, hi, wl, wh);
        wl = mul(wl, 0.5f); wh = mul(wh, 0.5f);
    }
    int mn = ok ? lo : INT_MAX, mx = ok ? hi : -1;
#pragma unroll
    for (int o = 8; o > 0; o >>= 1) {
        mn = min(mn, __shfl_xor_sync(0xffffffffu, mn, o));
        mx = max(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    }
    const int any_y = __shfl_sync(0xffffffffu, mx, 0) >= 0, any_x = __shfl_sync(0xffffffffu, mx, 16) >= 0;
    if (!any_x || !any_y) return CL_ZERO;
    const int x_lo = __shfl_sync(0xffffffffu, mn, 16) & ~3;
    const int nq = (__shfl_sync(0xffffffffu, mx, 16) - x_lo + 4) >> 2;
    if (nq > kClMaxQuads) return CL_DECLINE;

    // compact row index: samples are sorted, so the rows met up to sample s are a prefix of the compact list
    const int prev_hi = __shfl_up_sync(0xffffffffu, hi, 1, 16);
    const int prev_ok = __shfl_up_sync(0xffffffffu, (int)ok, 1, 16);
    int cnt = 0;
    if (ok) {
        if (s == 0 || !prev_ok) cnt = 1 + (hi > lo);
        else cnt = (lo > prev_hi) + ((hi > lo) && (hi > prev_hi));
    }
    int tot = cnt;
#pragma unroll
    for (int o = 1; o < 16; o <<= 1) {
        const int t = __shfl_up_sync(0xffffffffu, tot, o, 16);
        if (s >= o) tot += t;
    }
    const int idx_hi = tot - 1, idx_lo = (hi > lo) ? tot - 2 : tot - 1;
    const int nrows = __shfl_sync(0xffffffffu, tot, 15);

    float *ew = &bs.ew[0][0];
    for (int i = lane; i < kClRows * 8; i += 32) ew[i] = 0.0f;
    it.yof[lane] = 0;
    __syncwarp();
    if (is_y && ok) { it.yof[idx_lo] = lo; it.yof[idx_hi] = hi; }
    if (!is_y) {
        it.xlo[s] = ok ? lo - x_lo : 0; it.xhi[s] = ok ? hi - x_lo : 0;
        it.xwl[s] = ok ? wl : 0.0f; it.xwh[s] = ok ? wh : 0.0f;
    }
    // the two samples of a bin may share a row: add them in a fixed order (even sample, then odd) -> deterministic sums
#pragma unroll
    for (int ph = 0; ph < S; ph++) {
        if (is_y && ok && (s % S) == ph) {
            bs.ew[idx_lo][s / S] += wl;
            bs.ew[idx_hi][s / S] += wh;
        }
        __syncwarp();
    }
    {   // per-bin compact row range and weights
        const int p = min(lane, P - 1);
        const int ok0 = __shfl_sync(0xffffffffu, (int)ok, 2 * p), ok1 = __shfl_sync(0xffffffffu, (int)ok, 2 * p + 1);
        const int lo0 = __shfl_sync(0xffffffffu, idx_lo, 2 * p), lo1 = __shfl_sync(0xffffffffu, idx_lo, 2 * p + 1);
        const int hi0 = __shfl_sync(0xffffffffu, idx_hi, 2 * p), hi1 = __shfl_sync(0xffffffffu, idx_hi, 2 * p + 1);
        const int tot1 = __shfl_sync(0xffffffffu, tot, 2 * p + 1);
        if (lane < 8) {
            const int ja = lane < P ? (ok0 ? lo0 : (ok1 ? lo1 : tot1)) : 0, jb = lane < P ? (ok1 ? hi1 + 1 : (ok0 ? hi0 + 1 : tot1)) : 0;
            bs.ja[lane] = ja; bs.jb[lane] = jb;
            ClBin bn;
            bn.ja = ja; bn.nr = jb - ja; bn.pad0 = bn.pad1 = 0;
#pragma unroll
            for (int i = 0; i < 4; i++) bn.w[i] = (lane < P && ja + i < jb) ? bs.ew[ja + i][lane] : 0.0f;
            it.bin[lane] = bn;
        }
    }
    // sparse Ax per column chunk: lane = (chunk, q)
    const int nxc = (nq + 3) >> 2, nqc_max = (nq + nxc - 1) / nxc;
    __syncwarp();
    const unsigned runmask = __ballot_sync(0xffffffffu, lane + 1 < nrows && it.yof[min(lane + 1, 31)] == it.yof[lane] + 1);
    {
        const int ch = lane >> 3, q = lane & 7;
        ClTap t;
#pragma unroll
        for (int i = 0; i < 4; i++) { t.o[i] = 0; t.w[i] = 0.0f; }
        if (ch < nxc && q < P) {
            const int qa = ch * nq / nxc, c0 = 4 * qa, c1 = 4 * ((ch + 1) * nq / nxc);
#pragma unroll
            for (int i = 0; i < 4; i++) {
                const int sx = q * S + (i >> 1);
                const int col = (i & 1) ? it.xhi[sx] : it.xlo[sx];
                const float w = (i & 1) ? it.xwh[sx] : it.xwl[sx];
                const bool in = col >= c0 && col < c1;
                t.o[i] = in ? (col - c0) * 128 : 0;
                t.w[i] = in ? w : 0.0f;
            }
        }
        it.tap[ch][q] = t;
    }
    // pieces: as many whole bins as fit one stage slot, the same bin ranges for every column chunk; the rows of a piece
    // travel as bulk-tensor operations of 8 / 4 / 2 / 1 consecutive feature rows (uniform code; lane 0 writes)
    const int maxrows = min(kClRows, slot_bytes / cl_row_bytes(nqc_max));
    int k = 0, nop = 0;
    {
        int p = 0;
        while (p < P) {
            const int jA = bs.ja[p];
            int jB = bs.jb[p];
            const int p0 = p;
            p++;
            while (p < P && bs.jb[p] - jA <= maxrows) { jB = max(jB, bs.jb[p]); p++; }
            const int op0 = nop;
            int j = jA;
            while (j < jB) {
                const unsigned cont = ~(runmask >> j);              // first zero bit = end of the run that starts at row j
                const int run = min(cont ? __ffs(cont) : 32, jB - j);
                const int lr = run >= 8 ? 3 : (run >= 4 ? 2 : (run >= 2 ? 1 : 0));
                if (lane == 0 && nop < kClMaxOps) { ClOp o; o.j = (short)j; o.lr = (short)lr; it.op[nop] = o; }
                nop++;
                j += 1 << lr;
            }
            for (int ch = 0; ch < nxc; ch++) {
                const int kk = ch * 8 + k;                             // provisional slot: compacted below
                (void)kk;
            }
            if (lane == 0 && k < 8) {
                ClPiece pc;
                pc.x0 = 0; pc.nqc = 0; pc.jA = jA; pc.rows = jB - jA; pc.p0 = p0; pc.p1 = p; pc.chunk = 0; pc.ops = op0 | ((nop - op0) << 16);
                it.piece[k] = pc;                                      // chunk 0 first; the other chunks are copies (below)
            }
            k++;
        }
    }
    if (k * nxc > kClMaxPieces || nop > kClMaxOps || k > 7) return CL_DECLINE;   // (wide AND tall: left to the gather kernel)
    __syncwarp();
    // chunk-major piece list: piece[ch * k + i]
    for (int idx = lane; idx < k * nxc; idx += 32) {
        const int ch = idx / k, i = idx - ch * k;
        ClPiece pc = it.piece[i];
        const int qa = ch * nq / nxc;
        pc.x0 = x_lo + 4 * qa; pc.nqc = (ch + 1) * nq / nxc - qa; pc.chunk = ch;
        __syncwarp(__activemask());
        it.piece[idx] = pc;
    }
    if (lane == 0) {
        it.status = CL_OK; it.b = g.b; it.l = g.l; it.x_lo = x_lo; it.nq = nq; it.nrows = nrows; it.npiece = k * nxc; it.nxc = nxc;
    }
    __syncwarp();
    return CL_OK;
}

// one warp per RoI -> items[r] (global); flags[r] = 1 for the RoIs left to the gather kernels
constexpr int kPlanWarps = 4;
template <int P>
__global__ void __launch_bounds__(32 * kPlanWarps)
roialign_plan_kernel(const RoiFeat f, const int tma_mask, const float *__restrict__ rois5, const int R, const int slot_bytes,
                     ClItem *__restrict__ items, int32_t *__restrict__ flag)
{
    __shared__ ClItem its[kPlanWarps];
    __shared__ ClBuild bld[kPlanWarps];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int r = blockIdx.x * kPlanWarps + warp;
    if (r >= R) return;
    ClItem &it = its[warp];
    const int st = cl_build_item<P>(it, bld[warp], f, tma_mask, rois5, r, slot_bytes, lane);
    if (st != CL_OK && lane == 0) it.status = st;
    if (lane == 0) flag[r] = st == CL_DECLINE ? 1 : 0;
    __syncwarp();
    const uint4 *src = reinterpret_cast<const uint4 *>(&it);
    uint4 *dst = reinterpret_cast<uint4 *>(items + r);
    if (st == CL_OK) {
        for (int i = lane; i < (int)(sizeof(ClItem) / 16); i += 32) dst[i] = src[i];
    } else if (lane == 0) {
        items[r].status = st;
    }
}

// byte offsets of the lane's quads inside a staged row (rotated walk) and of the matching scratch columns
template <int NQ>
MD_DEVINL void cl_lane_offsets(uint32_t (&off)[NQ], uint32_t (&col)[NQ], int lane)
{
    const int rot = cl_rot(NQ, lane);
#pragma unroll
    for (int k = 0; k < NQ; k++) {
        const int qd = (k + rot) % NQ;
        off[k] = (uint32_t)(lane * cl_pitch(NQ) + 16 * qd);
        col[k] = (uint32_t)(4 * qd) * 128u;
    }
}

// ---- forward piece: bin by bin, software-pipelined ----------------------------------------------------------------
// Bin p needs at most 4 staged rows: A[x] = sum_i w_i row_i[x] (y-step, registers); the lane parks its 4*NQ column sums
// in its private scratch column, and the x-step reads back just the 4 taps of each output q:
// Out[p][q] (+)= sum_i w_i A[col_i] -> staging tile (stride 49 floats per lane: conflict-free).
// The shared-memory accesses are volatile asm, so they stay in program order; the loop is arranged so that the row loads
// of bin p+1 are in flight during the x-step of bin p and every batch of loads is issued before its first use.
template <int NQ>
MD_DEVINL void cl_fwd_piece(const ClItem &it, const ClTap (&tp)[kClP], uint32_t slot, uint32_t stg, uint32_t scr, int p0, int p1,
                            int jA, bool first, int lane)
{
    constexpr int P = kClP, X = 4 * NQ;
    constexpr uint32_t RB = (uint32_t)cl_row_bytes(NQ);
    uint32_t off[NQ], col[NQ];
    cl_lane_offsets<NQ>(off, col, lane);
    const uint32_t bins = smem_u32(&it.bin[0]);
    const uint32_t so = stg + (uint32_t)lane * (P * P * 4), sl = scr + (uint32_t)lane * 4u;
#pragma unroll 1
    for (int p = p0 - 1; p < p1; p++) {
        const bool nxt = p + 1 < p1;
        // (a) rows of bin p + 1
        int nr = 0;
        float4 bw = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
        float4 v0[NQ], v1[NQ], v2[NQ], v3[NQ];
        if (nxt) {
            const uint4 bi = lds128(bins + 32u * (p + 1));
            bw = lds128f(bins + 32u * (p + 1) + 16u);
            nr = (int)bi.y;
            const uint32_t ra = slot + (uint32_t)((int)bi.x - jA) * RB;
            if (nr > 0) {
#pragma unroll
                for (int k = 0; k < NQ; k++) { v0[k] = lds128f(ra + off[k]); v1[k] = lds128f(ra + (nr > 1 ? RB : 0u) + off[k]); }
            }
            if (nr > 2) {
#pragma unroll
                for (int k = 0; k < NQ; k++) v2[k] = lds128f(ra + 2u * RB + off[k]);
            }
            if (nr > 3) {
#pragma unroll
                for (int k = 0; k < NQ; k++) v3[k] = lds128f(ra + 3u * RB + off[k]);
            }
        }
        // (b, c) x-step of bin p from the scratch column parked one iteration ago
        if (p >= p0) {
            const uint32_t oa = so + (uint32_t)p * (P * 4);
            float t[P][4], old[P];
#pragma unroll
            for (int q = 0; q < P; q++) {
                t[q][0] = lds32f(sl + tp[q].o[0]); t[q][1] = lds32f(sl + tp[q].o[1]);
                t[q][2] = lds32f(sl + tp[q].o[2]); t[q][3] = lds32f(sl + tp[q].o[3]);
            }
            if (!first) {
#pragma unroll
                for (int q = 0; q < P; q++) old[q] = lds32f(oa + 4 * q);
            }
#pragma unroll
            for (int q = 0; q < P; q++) {
                float acc = mul(tp[q].w[0], t[q][0]);
                acc = __fmaf_rn(tp[q].w[1], t[q][1], acc);
                acc = __fmaf_rn(tp[q].w[2], t[q][2], acc);
                acc = __fmaf_rn(tp[q].w[3], t[q][3], acc);
                if (!first) acc = add(acc, old[q]);
                sts32f(oa + 4 * q, acc);
            }
        }
        // (d, e) y-step of bin p + 1, parked for the next iteration
        if (nxt) {
            float A[X];
            if (nr > 0) {
#pragma unroll
                for (int k = 0; k < NQ; k++) {
                    A[4 * k] = __fmaf_rn(bw.y, v1[k].x, mul(bw.x, v0[k].x)); A[4 * k + 1] = __fmaf_rn(bw.y, v1[k].y, mul(bw.x, v0[k].y));
                    A[4 * k + 2] = __fmaf_rn(bw.y, v1[k].z, mul(bw.x, v0[k].z)); A[4 * k + 3] = __fmaf_rn(bw.y, v1[k].w, mul(bw.x, v0[k].w));
                }
            } else {
#pragma unroll
                for (int x = 0; x < X; x++) A[x] = 0.0f;
            }
            if (nr > 2) {
#pragma unroll
                for (int k = 0; k < NQ; k++) {
                    A[4 * k] = __fmaf_rn(bw.z, v2[k].x, A[4 * k]); A[4 * k + 1] = __fmaf_rn(bw.z, v2[k].y, A[4 * k + 1]);
                    A[4 * k + 2] = __fmaf_rn(bw.z, v2[k].z, A[4 * k + 2]); A[4 * k + 3] = __fmaf_rn(bw.z, v2[k].w, A[4 * k + 3]);
                }
            }
            if (nr > 3) {
#pragma unroll
                for (int k = 0; k < NQ; k++) {
                    A[4 * k] = __fmaf_rn(bw.w, v3[k].x, A[4 * k]); A[4 * k + 1] = __fmaf_rn(bw.w, v3[k].y, A[4 * k + 1]);
                    A[4 * k + 2] = __fmaf_rn(bw.w, v3[k].z, A[4 * k + 2]); A[4 * k + 3] = __fmaf_rn(bw.w, v3[k].w, A[4 * k + 3]);
                }
            }
#pragma unroll
            for (int k = 0; k < NQ; k++) {
                sts32f(sl + col[k], A[4 * k]); sts32f(sl + col[k] + 128u, A[4 * k + 1]);
                sts32f(sl + col[k] + 256u, A[4 * k + 2]); sts32f(sl + col[k] + 384u, A[4 * k + 3]);
            }
        }
    }
}

// ---- the item stream shared by the forward and the backward kernel ------------------------------------------------
// A persistent 1-warp CTA takes (RoI, channel chunk) items from a global ticket (the NEXT ticket is requested while the
// current item is being worked on), pulls each item's table into one of two shared-memory buffers with a bulk copy, and
// walks three cursors over the same item sequence: l (tables requested) >= p (producer) >= c (consumer), l <= c + 2.
struct ClStream {
    int l_seq, p_seq, c_seq;
    int tick_next;
    bool exhausted;
    int tr[2], tc[2];                       // RoI / channel chunk of the items in the two table buffers
};

constexpr int kClStgBytes = 6400;           // 32 x 49 floats, padded
constexpr int kClScrBytes = 2048;           // 16 columns x 32 lanes
constexpr size_t kClSmemBytes = 1024 + (size_t)kClSlots * kClSlotBytes + kClStgBytes + kClScrBytes + 2 * sizeof(ClItem) + 64;

template <int P>
__global__ void __launch_bounds__(32, MD_CL_CTAS)
roialign_fwd_cl_kernel(const __grid_constant__ ClMaps maps, const ClItem *__restrict__ items, const int R, const int C,
                       const int seg, const int nchunk, float *__restrict__ out, int *__restrict__ ctr)
{
    static_assert(P == kClP, "7x7 only");
    extern __shared__ unsigned char dsm_raw[];
    unsigned char *sp = dsm_raw + ((1024u - (smem_u32(dsm_raw) & 1023u)) & 1023u);
    unsigned char *slots = sp; sp += kClSlots * kClSlotBytes;
    unsigned char *stgp = sp; sp += kClStgBytes;
    unsigned char *scrp = sp; sp += kClScrBytes;
    ClItem *tabs = reinterpret_cast<ClItem *>(sp); sp += 2 * sizeof(ClItem);
    unsigned long long *full = reinterpret_cast<unsigned long long *>(sp);
    unsigned long long *tfull = full + kClSlots;

    const int lane = threadIdx.x;
    const int CH = C / nchunk, ngroups = CH / 32;
    const int total = R * nchunk;
    if (lane == 0) {
        for (int i = 0; i < kClSlots; i++) mbar_init(&full[i], 1);
        mbar_init(&tfull[0], 1); mbar_init(&tfull[1], 1);
        fence_barrier_init();
    }
    __syncwarp();
    const uint32_t stg = smem_u32(stgp), scr = smem_u32(scrp);

    ClStream st;
    st.l_seq = st.p_seq = st.c_seq = 0;
    st.exhausted = false;
    {
        int t = 0;
        if (lane == 0) t = atomicAdd(ctr, 1);
        st.tick_next = __shfl_sync(0xffffffffu, t, 0);
    }
    bool p_ready = false, c_ready = false;    // the table of item p_seq / c_seq has landed and holds work
    int p_g = 0, p_k = 0, c_g = 0, c_k = 0;
    int n_iss = 0, n_con = 0;
    int tap_seq = -1, tap_ch = -1;
    bool store_pending = false;
    ClTap tp[P];

#pragma unroll 1
    for (;;) {
        // ---- request tables: at most two items beyond the consumer's ---------------------------------------------
        while (!st.exhausted && st.l_seq < st.c_seq + 2) {
            const int t = st.tick_next;
            if (t >= total) { st.exhausted = true; break; }
            int tn = 0;
            if (lane == 0) tn = atomicAdd(ctr, 1);                   // the NEXT ticket: its latency hides behind this item
            st.tick_next = __shfl_sync(0xffffffffu, tn, 0);
            const WorkItem wi = work_item(t, R, seg, nchunk);
            const int b = st.l_seq & 1;
            st.tr[b] = wi.r; st.tc[b] = wi.chunk;
            if (lane == 0) {
                fence_proxy_async();
                mbar_expect_tx(&tfull[b], (uint32_t)sizeof(ClItem));
                bulk_load_1d(&tabs[b], items + wi.r, (uint32_t)sizeof(ClItem), &tfull[b]);
            }
            st.l_seq++;
        }
        // ---- producer: keep the stage slots full ------------------------------------------------------------------
#pragma unroll 1
        for (;;) {
            if (st.p_seq >= st.l_seq) break;
            const int b = st.p_seq & 1;
            if (!p_ready) {
                mbar_wait(&tfull[b], (uint32_t)(st.p_seq >> 1) & 1u);
                if (tabs[b].status != CL_OK) { st.p_seq++; continue; }   // no pieces (the consumer writes the zeros)
                p_ready = true; p_g = p_k = 0;
            }
            if (n_iss - n_con >= kClSlots) break;
            const ClItem &it = tabs[b];
            const ClPiece pc = it.piece[p_k];
            const int slot = n_iss % kClSlots;
            const int rb = 512 * pc.nqc;
            if (lane == 0) {
                fence_proxy_async();
                mbar_expect_tx(&full[slot], (uint32_t)(pc.rows * rb));
                const CUtensorMap *mp = &maps.m[(it.l * 4 + pc.nqc - 1) * kClRsel];
                const int z = it.b * C + st.tc[b] * CH + 32 * p_g;
                unsigned char *dst = slots + slot * kClSlotBytes - pc.jA * rb;
                const int op0 = pc.ops & 0xffff, nop = pc.ops >> 16;
                for (int o = op0; o < op0 + nop; o++) {
                    const ClOp op = it.op[o];
                    tma_load_3d(dst + op.j * rb, mp + op.lr, pc.x0, z, it.yof[op.j], &full[slot]);
                }
            }
            n_iss++;
            if (++p_k == it.npiece) {
                p_k = 0;
                if (++p_g == ngroups) { p_ready = false; st.p_seq++; }
            }
        }
        // ---- consumer -----------------------------------------------------------------------------------------------
        if (st.c_seq >= st.l_seq) break;                         // every ticket taken, every item done
        const int cb = st.c_seq & 1;
        if (!c_ready) {
            mbar_wait(&tfull[cb], (uint32_t)(st.c_seq >> 1) & 1u);
            const int status = tabs[cb].status;
            if (status != CL_OK) {
                if (status == CL_ZERO) {
                    float *o = out + ((int64_t)st.tr[cb] * C + (int64_t)st.tc[cb] * CH) * (P * P);
                    for (int i = lane; i < CH * P * P; i += 32) o[i] = 0.0f;
                }
                __syncwarp();
                st.c_seq++;
                continue;
            }
            c_ready = true; c_g = c_k = 0;
        }
        const ClItem &it = tabs[cb];
        const ClPiece pc = it.piece[c_k];
        if (tap_seq != st.c_seq || tap_ch != pc.chunk) {
#pragma unroll
            for (int q = 0; q < P; q++) tp[q] = it.tap[pc.chunk][q];
            tap_seq = st.c_seq; tap_ch = pc.chunk;
        }
        if (store_pending && c_k == 0) {                         // the staging tile is about to be overwritten
            if (lane == 0) bulk_wait_read<0>();
            __syncwarp();
            store_pending = false;
        }
        const int slot = n_con % kClSlots;
        mbar_wait(&full[slot], (uint32_t)(n_con / kClSlots) & 1u);
        const uint32_t sa = smem_u32(slots + slot * kClSlotBytes);
        switch (pc.nqc) {
            case 1: cl_fwd_piece<1>(it, tp, sa, stg, scr, pc.p0, pc.p1, pc.jA, pc.chunk == 0, lane); break;
            case 2: cl_fwd_piece<2>(it, tp, sa, stg, scr, pc.p0, pc.p1, pc.jA, pc.chunk == 0, lane); break;
            case 3: cl_fwd_piece<3>(it, tp, sa, stg, scr, pc.p0, pc.p1, pc.jA, pc.chunk == 0, lane); break;
            default: cl_fwd_piece<4>(it, tp, sa, stg, scr, pc.p0, pc.p1, pc.jA, pc.chunk == 0, lane); break;
        }
        __syncwarp();
        n_con++;
        const int npiece = it.npiece;
        if (c_k == npiece - 1) {                                 // the 32 x 49 outputs of this (RoI, channel group) are complete
            fence_proxy_async();
            __syncwarp();
            if (lane == 0) {
                bulk_store_1d(out + ((int64_t)st.tr[cb] * C + st.tc[cb] * CH + 32 * c_g) * (P * P), stgp, kClStageBytes);
                bulk_commit();
            }
            store_pending = true;
        }
        if (++c_k == npiece) {
            c_k = 0;
            if (++c_g == ngroups) { c_ready = false; st.c_seq++; }
        }
    }
    if (lane == 0) {
        bulk_wait_all<0>();
        // the last CTA out re-arms the ticket for the next launch on this stream
        __threadfence();
        if (atomicAdd(ctr + 1, 1) == (int)gridDim.x - 1) { ctr[0] = 0; ctr[1] = 0; __threadfence(); }
    }
}

// dense Ax of one column chunk in the lane's (rotated) quad order, in registers: ax[q][4k + i] = Ax[q][4 quad(k) + i]
template <int NQ>
MD_DEVINL void cl_load_ax(float (&ax)[kClP][4 * NQ], const ClItem &it, int chunk, int lane)
{
    constexpr int P = kClP, S = 2;
    const int nq = it.nq, nxc = it.nxc;
    const int c0 = 4 * (chunk * nq / nxc);
    const int rot = cl_rot(NQ, lane);
#pragma unroll
    for (int q = 0; q < P; q++) {
#pragma unroll
        for (int x = 0; x < 4 * NQ; x++) ax[q][x] = 0.0f;
#pragma unroll
        for (int ph = 0; ph < S; ph++) {
            const int s = q * S + ph;
            const int lo = it.xlo[s] - c0, hi = it.xhi[s] - c0;
            const float wl = it.xwl[s], wh = it.xwh[s];
#pragma unroll
            for (int k = 0; k < NQ; k++) {
                const int base = 4 * ((k + rot) % NQ);
#pragma unroll
                for (int i = 0; i < 4; i++) {
                    if (lo == base + i) ax[q][4 * k + i] = add(ax[q][4 * k + i], wl);
                    if (hi == base + i) ax[q][4 * k + i] = add(ax[q][4 * k + i], wh);
                }
            }
        }
    }
}

// =====================================================================================================
// backward: dX += Ay^T (dY Ax) per channel; rows leave through cp.reduce.async.bulk.tensor (add, at L2)
// =====================================================================================================
#ifndef MD_CLB_SLOT_KB
#define MD_CLB_SLOT_KB 12
#endif
#ifndef MD_CLB_CTAS
#define MD_CLB_CTAS 4
#endif
constexpr int kClbSlotBytes = MD_CLB_SLOT_KB * 1024;
constexpr int kClbGBufBytes = 6400;
constexpr size_t kClbSmemBytes = 1024 + (size_t)kClSlots * kClbSlotBytes + 2 * kClbGBufBytes + 2 * sizeof(ClItem) + 64;

// one piece, bin by bin: T[x] = sum_q dY[p][q] Ax[q][x] (registers, Ax register-resident), then the bin's <= 4 rows of
// the slot += w_i T (the rows of a piece are zeroed first; the lane owns its channel's bytes of every row, so plain
// read-modify-write; loads of row i+1 are issued before the stores of row i)
template <int NQ>
MD_DEVINL void cl_bwd_piece(const ClItem &it, const float (&ax)[kClP][4 * NQ], uint32_t slot, uint32_t gbuf, int p0, int p1,
                            int jA, int rows, int lane)
{
    constexpr int P = kClP, X = 4 * NQ;
    constexpr uint32_t RB = (uint32_t)cl_row_bytes(NQ);
    uint32_t off[NQ];
#pragma unroll
    for (int k = 0; k < NQ; k++) off[k] = (uint32_t)(lane * cl_pitch(NQ) + 16 * ((k + cl_rot(NQ, lane)) % NQ));
#pragma unroll 1
    for (int j = 0; j < rows; j++)
#pragma unroll
        for (int k = 0; k < NQ; k++) sts128f(slot + (uint32_t)j * RB + off[k], make_float4(0.0f, 0.0f, 0.0f, 0.0f));
    const uint32_t bins = smem_u32(&it.bin[0]);
    const uint32_t go = gbuf + (uint32_t)lane * (P * P * 4);
#pragma unroll 1
    for (int p = p0; p < p1; p++) {
        const uint4 bi = lds128(bins + 32u * p);
        const float4 bw = lds128f(bins + 32u * p + 16u);
        const int nr = (int)bi.y;
        if (nr <= 0) continue;
        const uint32_t ga = go + (uint32_t)p * (P * 4);
        float g[P];
#pragma unroll
        for (int q = 0; q < P; q++) g[q] = lds32f(ga + 4 * q);
        const uint32_t ra = slot + (uint32_t)((int)bi.x - jA) * RB;
        float4 d0[NQ], d1[NQ];
#pragma unroll
        for (int k = 0; k < NQ; k++) { d0[k] = lds128f(ra + off[k]); d1[k] = lds128f(ra + (nr > 1 ? RB : 0u) + off[k]); }
        float T[X];
#pragma unroll
        for (int x = 0; x < X; x++) {
            float acc = mul(g[0], ax[0][x]);
#pragma unroll
            for (int q = 1; q < P; q++) acc = __fmaf_rn(g[q], ax[q][x], acc);
            T[x] = acc;
        }
        float4 d2[NQ], d3[NQ];
        if (nr > 2) {
#pragma unroll
            for (int k = 0; k < NQ; k++) d2[k] = lds128f(ra + 2u * RB + off[k]);
        }
        if (nr > 3) {
#pragma unroll
            for (int k = 0; k < NQ; k++) d3[k] = lds128f(ra + 3u * RB + off[k]);
        }
#pragma unroll
        for (int k = 0; k < NQ; k++) {
            d0[k].x = __fmaf_rn(bw.x, T[4 * k], d0[k].x); d0[k].y = __fmaf_rn(bw.x, T[4 * k + 1], d0[k].y);
            d0[k].z = __fmaf_rn(bw.x, T[4 * k + 2], d0[k].z); d0[k].w = __fmaf_rn(bw.x, T[4 * k + 3], d0[k].w);
            sts128f(ra + off[k], d0[k]);
        }
        if (nr > 1) {
#pragma unroll
            for (int k = 0; k < NQ; k++) {
                d1[k].x = __fmaf_rn(bw.y, T[4 * k], d1[k].x); d1[k].y = __fmaf_rn(bw.y, T[4 * k + 1], d1[k].y);
                d1[k].z = __fmaf_rn(bw.y, T[4 * k + 2], d1[k].z); d1[k].w = __fmaf_rn(bw.y, T[4 * k + 3], d1[k].w);
                sts128f(ra + RB + off[k], d1[k]);
            }
        }
        if (nr > 2) {
#pragma unroll
            for (int k = 0; k < NQ; k++) {
                d2[k].x = __fmaf_rn(bw.z, T[4 * k], d2[k].x); d2[k].y = __fmaf_rn(bw.z, T[4 * k + 1], d2[k].y);
                d2[k].z = __fmaf_rn(bw.z, T[4 * k + 2], d2[k].z); d2[k].w = __fmaf_rn(bw.z, T[4 * k + 3], d2[k].w);
                sts128f(ra + 2u * RB + off[k], d2[k]);
            }
        }
        if (nr > 3) {
#pragma unroll
            for (int k = 0; k < NQ; k++) {
                d3[k].x = __fmaf_rn(bw.w, T[4 * k], d3[k].x); d3[k].y = __fmaf_rn(bw.w, T[4 * k + 1], d3[k].y);
                d3[k].z = __fmaf_rn(bw.w, T[4 * k + 2], d3[k].z); d3[k].w = __fmaf_rn(bw.w, T[4 * k + 3], d3[k].w);
                sts128f(ra + 3u * RB + off[k], d3[k]);
            }
        }
    }
}

template <int NQ>
MD_DEVINL void cl_bwd_chunk(const ClItem &it, int chunk, int k0, int k1, uint32_t gbuf, unsigned char *slots, int &n_piece,
                            const CUtensorMap *maps_l, int z, int lane)
{
    float ax[kClP][4 * NQ];
    cl_load_ax<NQ>(ax, it, chunk, lane);
#pragma unroll 1
    for (int k = k0; k < k1; k++) {
        const ClPiece pc = it.piece[k];
        if (pc.rows <= 0) continue;
        const int rb = 512 * NQ;
        const int slot = n_piece % kClSlots;
        bulk_wait_read<kClSlots - 1>();          // this lane's reduce that last read the slot has finished reading
        __syncwarp();
        cl_bwd_piece<NQ>(it, ax, smem_u32(slots + slot * kClbSlotBytes), gbuf, pc.p0, pc.p1, pc.jA, pc.rows, lane);
        fence_proxy_async();
        __syncwarp();
        if (lane == 0) {
            const CUtensorMap *mp = maps_l + (NQ - 1) * kClRsel;
            const unsigned char *src = slots + slot * kClbSlotBytes - pc.jA * rb;
            const int op0 = pc.ops & 0xffff, nop = pc.ops >> 16;
            for (int o = op0; o < op0 + nop; o++) {
                const ClOp op = it.op[o];
                tma_reduce_add_3d(mp + op.lr, pc.x0, z, it.yof[op.j], src + op.j * rb);
            }
        }
        bulk_commit();
        n_piece++;
    }
}

template <int P>
__global__ void __launch_bounds__(32, MD_CLB_CTAS)
roialign_bwd_cl_kernel(const __grid_constant__ ClMaps maps, const ClItem *__restrict__ items, const int R, const int C,
                       const int seg, const int nchunk, const float *__restrict__ dout, int *__restrict__ ctr)
{
    static_assert(P == kClP, "7x7 only");
    extern __shared__ unsigned char dsm_raw[];
    unsigned char *sp = dsm_raw + ((1024u - (smem_u32(dsm_raw) & 1023u)) & 1023u);
    unsigned char *slots = sp; sp += kClSlots * kClbSlotBytes;
    unsigned char *gbufs = sp; sp += 2 * kClbGBufBytes;
    ClItem *tabs = reinterpret_cast<ClItem *>(sp); sp += 2 * sizeof(ClItem);
    unsigned long long *gfull = reinterpret_cast<unsigned long long *>(sp);
    unsigned long long *tfull = gfull + 2;

    const int lane = threadIdx.x;
    const int CH = C / nchunk, ngroups = CH / 32;
    const int total = R * nchunk;
    if (lane == 0) {
        mbar_init(&gfull[0], 1); mbar_init(&gfull[1], 1);
        mbar_init(&tfull[0], 1); mbar_init(&tfull[1], 1);
        fence_barrier_init();
    }
    __syncwarp();

    ClStream st;
    st.l_seq = st.p_seq = st.c_seq = 0;
    st.exhausted = false;
    {
        int t = 0;
        if (lane == 0) t = atomicAdd(ctr, 1);
        st.tick_next = __shfl_sync(0xffffffffu, t, 0);
    }
    bool p_ready = false, c_ready = false;
    int p_g = 0, c_g = 0;
    int g_iss = 0, g_con = 0;             // dY buffers issued / released
    int n_piece = 0;                      // pieces written so far (slot ring position)

#pragma unroll 1
    for (;;) {
        while (!st.exhausted && st.l_seq < st.c_seq + 2) {
            const int t = st.tick_next;
            if (t >= total) { st.exhausted = true; break; }
            int tn = 0;
            if (lane == 0) tn = atomicAdd(ctr, 1);
            st.tick_next = __shfl_sync(0xffffffffu, tn, 0);
            const WorkItem wi = work_item(t, R, seg, nchunk);
            const int b = st.l_seq & 1;
            st.tr[b] = wi.r; st.tc[b] = wi.chunk;
            if (lane == 0) {
                fence_proxy_async();
                mbar_expect_tx(&tfull[b], (uint32_t)sizeof(ClItem));
                bulk_load_1d(&tabs[b], items + wi.r, (uint32_t)sizeof(ClItem), &tfull[b]);
            }
            st.l_seq++;
        }
        // producer = the dY loads, at most one (item, channel group) ahead of the consumer
#pragma unroll 1
        for (;;) {
            if (st.p_seq >= st.l_seq) break;
            const int b = st.p_seq & 1;
            if (!p_ready) {
                mbar_wait(&tfull[b], (uint32_t)(st.p_seq >> 1) & 1u);
                if (tabs[b].status != CL_OK) { st.p_seq++; continue; }
                p_ready = true; p_g = 0;
            }
            if (g_iss - g_con >= 2) break;
            const int gb = g_iss & 1;
            if (lane == 0) {
                fence_proxy_async();
                mbar_expect_tx(&gfull[gb], kClStageBytes);
                bulk_load_1d(gbufs + gb * kClbGBufBytes, dout + ((int64_t)st.tr[b] * C + st.tc[b] * CH + 32 * p_g) * (P * P), kClStageBytes, &gfull[gb]);
            }
            g_iss++;
            if (++p_g == ngroups) { p_ready = false; st.p_seq++; }
        }
        if (st.c_seq >= st.l_seq) break;
        const int cb = st.c_seq & 1;
        if (!c_ready) {
            mbar_wait(&tfull[cb], (uint32_t)(st.c_seq >> 1) & 1u);
            if (tabs[cb].status != CL_OK) { __syncwarp(); st.c_seq++; continue; }      // no sample in range / declined: no gradient here
            c_ready = true; c_g = 0;
        }
        // ---- one (item, channel group): all its pieces, chunk by chunk ----------------------------------------------
        const ClItem &it = tabs[cb];
        const int gb = g_con & 1;
        mbar_wait(&gfull[gb], (uint32_t)(g_con >> 1) & 1u);
        const uint32_t gbuf = smem_u32(gbufs + gb * kClbGBufBytes);
        const int nxc = it.nxc, per = it.npiece / nxc;
        const int z = it.b * C + st.tc[cb] * CH + 32 * c_g;
        const CUtensorMap *maps_l = &maps.m[it.l * 4 * kClRsel];
#pragma unroll 1
        for (int ch = 0; ch < nxc; ch++) {
            const int nqc = it.piece[ch * per].nqc;
            switch (nqc) {
                case 1: cl_bwd_chunk<1>(it, ch, ch * per, (ch + 1) * per, gbuf, slots, n_piece, maps_l, z, lane); break;
                case 2: cl_bwd_chunk<2>(it, ch, ch * per, (ch + 1) * per, gbuf, slots, n_piece, maps_l, z, lane); break;
                case 3: cl_bwd_chunk<3>(it, ch, ch * per, (ch + 1) * per, gbuf, slots, n_piece, maps_l, z, lane); break;
                default: cl_bwd_chunk<4>(it, ch, ch * per, (ch + 1) * per, gbuf, slots, n_piece, maps_l, z, lane); break;
            }
        }
        __syncwarp();                                    // every lane is done with this dY buffer
        g_con++;
        if (++c_g == ngroups) { c_ready = false; st.c_seq++; }
    }
    bulk_wait_all<0>();
    __syncwarp();
    if (lane == 0) {
        __threadfence();
        if (atomicAdd(ctr + 1, 1) == (int)gridDim.x - 1) { ctr[0] = 0; ctr[1] = 0; __threadfence(); }
    }
}

// =====================================================================================================
// host
// =====================================================================================================
struct ClMapCache {
    void *ptr[kClLevels]; int H[kClLevels], W[kClLevels], BC, L, mask; ClMaps maps; bool valid; unsigned long long stamp;
};
constexpr int kClCacheEntries = 32;
static ClMapCache g_cl_cache[kClCacheEntries];
static unsigned long long g_cl_stamp = 0;
static std::mutex g_cl_mutex;

static int cl_build_maps(const FeatSet &fs, ClMaps *out)
{
    EncodeTiledFn enc = get_encode();
    std::lock_guard<std::mutex> lock(g_cl_mutex);
    const int L = fs.L < kClLevels ? fs.L : kClLevels;
    ClMapCache *hit = nullptr, *victim = &g_cl_cache[0];
    for (int e = 0; e < kClCacheEntries; e++) {
        ClMapCache &c = g_cl_cache[e];
        bool same = c.valid && c.L == L && c.BC == fs.B * fs.C;
        for (int l = 0; same && l < L; l++) same = c.ptr[l] == fs.feat[l] && c.H[l] == fs.H[l] && c.W[l] == fs.W[l];
        if (same) { hit = &c; break; }
        if (!c.valid) { if (victim->valid) victim = &c; }
        else if (victim->valid && c.stamp < victim->stamp) victim = &c;
    }
    if (!hit) {
        ClMapCache &c = *victim;
        std::memset(&c.maps, 0, sizeof(c.maps));
        c.mask = 0;
        for (int l = 0; l < L; l++) {
            c.ptr[l] = fs.feat[l]; c.H[l] = fs.H[l]; c.W[l] = fs.W[l];
            if (!enc || (fs.W[l] & 3) || (reinterpret_cast<uintptr_t>(fs.feat[l]) & 15)) continue;
            bool ok = true;
            for (int nq = 1; nq <= 4 && ok; nq++)
                for (int lr = 0; lr < kClRsel && ok; lr++) {
                    const cuuint64_t dims[3] = { (cuuint64_t)fs.W[l], (cuuint64_t)fs.B * fs.C, (cuuint64_t)fs.H[l] };
                    const cuuint64_t strides[2] = { (cuuint64_t)fs.W[l] * fs.H[l] * 4, (cuuint64_t)fs.W[l] * 4 };
                    const cuuint32_t box[3] = { (cuuint32_t)(4 * nq), 32u, (cuuint32_t)(1 << lr) };
                    const cuuint32_t estr[3] = { 1, 1, 1 };
                    ok = enc(&c.maps.m[(l * 4 + nq - 1) * kClRsel + lr], CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, fs.feat[l], dims, strides, box,
                             estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                             CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
                }
            if (ok) c.mask |= 1 << l;
        }
        c.L = L; c.BC = fs.B * fs.C; c.valid = true;
        hit = &c;
    }
    hit->stamp = ++g_cl_stamp;
    *out = hit->maps;
    return hit->mask;
}

static int cl_chunks_for(int C)
{
    const char *e = getenv("MD_ROI_CHUNK");
    const int want = e ? atoi(e) : 128;
    int n = C / (want < 32 ? 32 : want);
    while (n > 1 && (C % n != 0 || (C / n) % 32 != 0)) n--;
    return n < 1 ? 1 : n;
}

// experiment switch while the channel-lane kernels are being tuned: MD_ROI_CL=1 selects them
static bool cl_enabled()
{
    const char *e = getenv("MD_ROI_CL");
    return e && atoi(e) != 0;
}

static int cl_grid(int ctas_per_sm)
{
    int dev = 0, sms = 148;
    if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    return sms * ctas_per_sm;
}

// ctr: two ints, zero before the first launch (the kernel re-arms them)
size_t roialign_cl_workspace_bytes(int R) { return (size_t)(R > 0 ? R : 0) * sizeof(ClItem) + 256; }

// plan (one warp per RoI) -> items; then the persistent kernel.  ctr: two ints, zero before the first launch (the kernel
// re-arms them); items: roialign_cl_workspace_bytes(R) bytes of scratch
template <bool FWD>
static cudaError_t cl_launch(const FeatSet &fs, const RoiFeat &f, const float *rois5, int R, int P, float *out, const float *dout,
                             int32_t *fallback_flag, void *items_ws, int *ctr, cudaStream_t s, bool *launched)
{
    *launched = false;
    if ((fs.C & 31) || P != kClP || R <= 0 || !items_ws) return cudaSuccess;
    if (!cl_enabled()) return cudaSuccess;
    ClMaps maps;
    const int mask = cl_build_maps(fs, &maps);
    if (!mask) return cudaSuccess;
    ClItem *items = reinterpret_cast<ClItem *>((reinterpret_cast<uintptr_t>(items_ws) + 255) & ~(uintptr_t)255);
    roialign_plan_kernel<kClP><<<(R + kPlanWarps - 1) / kPlanWarps, 32 * kPlanWarps, 0, s>>>(
        f, mask, rois5, R, FWD ? kClSlotBytes : kClbSlotBytes, items, fallback_flag);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return e;
    const int nchunk = cl_chunks_for(fs.C);
    const int total = R * nchunk;
    if (FWD) {
        auto kern = roialign_fwd_cl_kernel<kClP>;
        e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kClSmemBytes);   // per device: set every time
        if (e != cudaSuccess) return e;
        cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
        const int grid = cl_grid(MD_CL_CTAS);
        kern<<<total < grid ? total : grid, 32, kClSmemBytes, s>>>(maps, items, R, fs.C, 512, nchunk, out, ctr);
    } else {
        auto kern = roialign_bwd_cl_kernel<kClP>;
        e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kClbSmemBytes);
        if (e != cudaSuccess) return e;
        cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
        const int grid = cl_grid(MD_CLB_CTAS);
        kern<<<total < grid ? total : grid, 32, kClbSmemBytes, s>>>(maps, items, R, fs.C, 512, nchunk, dout, ctr);
    }
    *launched = true;
    return cudaGetLastError();
}

cudaError_t launch_roialign_fwd_cl(const FeatSet &fs, const RoiFeat &f, const float *rois5, int R, int P, float *out,
                                   int32_t *fallback_flag, void *items_ws, int *ctr, cudaStream_t s, bool *launched)
{
    return cl_launch<true>(fs, f, rois5, R, P, out, nullptr, fallback_flag, items_ws, ctr, s, launched);
}

cudaError_t launch_roialign_bwd_cl(const FeatSet &fs, const RoiFeat &f, const float *rois5, int R, int P, const float *dout,
                                   int32_t *fallback_flag, void *items_ws, int *ctr, cudaStream_t s, bool *launched)
{
    return cl_launch<false>(fs, f, rois5, R, P, nullptr, dout, fallback_flag, items_ws, ctr, s, launched);
}

}  // namespace md
