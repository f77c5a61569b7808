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

enum { CL_OK = 0, CL_ZERO = 1, CL_DECLINE = 2 };

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

// one bin: compact rows [ja, ja + nr), nr <= 4, and their weights (1/S folded in, both samples summed)
struct __align__(16) ClBin { int ja, nr, pad0, pad1; float w[4]; };
// one bilinear output column q of one column chunk: 4 taps = byte offsets into the lane's scratch column + weights
struct __align__(16) ClTap { int o[4]; float w[4]; };
// one staged piece: columns [x0, x0 + 4 nqc) of compact rows [jA, jA + rows), producing bins [p0, p1)
struct __align__(16) ClPiece { int x0, nqc, jA, rows, p0, p1, chunk, first; };

struct __align__(16) ClItem {
    ClBin bin[8];
    ClTap tap[kClMaxChunks][8];      // Ax, sparse, per column chunk
    ClPiece piece[kClMaxPieces];
    int yof[32];                     // feature row of compact row j
    int xlo[16], xhi[16];            // x samples, columns relative to x_lo
    float xwl[16], xwh[16];
    int r, chunk, b, l, x_lo, nq, nrows, npiece, nxc;
    unsigned runmask;                // bit j: compact row j+1 is the feature row right below compact row j
    int pad_[2];
};
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

// ---- per-item geometry (one whole warp): lanes 0..13 own the y samples, lanes 16..29 the x samples ----------------
template <int P, int SLOTB>
MD_DEVINL int cl_build_item(ClItem &it, ClBuild &bs, const RoiFeat &f, int tma_mask, const float *__restrict__ rois5, int r,
                            int chunk, int lane)
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
        if (lane < P) {
            const int ja = ok0 ? lo0 : (ok1 ? lo1 : tot1), jb = ok1 ? hi1 + 1 : (ok0 ? hi0 + 1 : tot1);
            bs.ja[lane] = ja; bs.jb[lane] = jb;
            ClBin bn;
            bn.ja = ja; bn.nr = jb - ja; bn.pad0 = bn.pad1 = 0;
#pragma unroll
            for (int i = 0; i < 4; i++) bn.w[i] = ja + i < jb ? bs.ew[ja + i][lane] : 0.0f;
            it.bin[lane] = bn;
        }
    }
    // sparse Ax per column chunk: lane = (chunk, q)
    const int nxc = (nq + 3) >> 2, nqc_max = (nq + nxc - 1) / nxc;
    __syncwarp();
    const unsigned runmask = __ballot_sync(0xffffffffu, lane + 1 < nrows && it.yof[min(lane + 1, 31)] == it.yof[lane] + 1);
    {
        const int ch = lane >> 3, q = lane & 7;
        if (ch < nxc && q < P) {
            const int qa = ch * nq / nxc, c0 = 4 * qa, c1 = 4 * ((ch + 1) * nq / nxc);
            ClTap t;
#pragma unroll
            for (int i = 0; i < 4; i++) {
                const int sx = q * S + (i >> 1);
                const int col = (i & 1) ? it.xhi[sx] : it.xlo[sx];
                const float w = (i & 1) ? it.xwh[sx] : it.xwl[sx];
                const bool in = col >= c0 && col < c1;
                t.o[i] = in ? (col - c0) * 128 : 0;
                t.w[i] = in ? w : 0.0f;
            }
            it.tap[ch][q] = t;
        }
    }
    // pieces: per column chunk, as many whole bins as fit one stage slot (uniform; lane 0 writes)
    const int maxrows = min(kClRows, SLOTB / cl_row_bytes(nqc_max));
    int k = 0;
    for (int ch = 0; ch < nxc; ch++) {
        const int qa = ch * nq / nxc, nqc = (ch + 1) * nq / nxc - qa;
        int p = 0;
        while (p < P) {
            const int jA = bs.ja[p];
            int jB = bs.jb[p];
            const int p0 = p;
            p++;
            while (p < P && bs.jb[p] - jA <= maxrows) { jB = max(jB, bs.jb[p]); p++; }
            if (lane == 0 && k < kClMaxPieces) {
                ClPiece pc;
                pc.x0 = x_lo + 4 * qa; pc.nqc = nqc; pc.jA = jA; pc.rows = jB - jA; pc.p0 = p0; pc.p1 = p; pc.chunk = ch; pc.first = ch == 0;
                it.piece[k] = pc;
            }
            k++;
        }
    }
    if (k > kClMaxPieces) return CL_DECLINE;                          // (wide AND tall: left to the gather kernel)
    if (lane == 0) {
        it.r = r; it.chunk = chunk; it.b = g.b; it.l = g.l; it.x_lo = x_lo; it.nq = nq; it.nrows = nrows;
        it.npiece = k; it.nxc = nxc; it.runmask = runmask;
    }
    __syncwarp();
    return CL_OK;
}


// Rows [jA, jA + rows) of a piece as bulk-tensor operations of 8 / 4 / 2 / 1 consecutive feature rows.  fn(j, lr): rows
// j .. j + (1 << lr) - 1.  Uniform over the warp; the caller elects the issuing lane.
template <class Fn>
MD_DEVINL void cl_for_row_ops(unsigned runmask, int jA, int rows, Fn fn)
{
    int j = jA;
    const int jB = jA + rows;
    while (j < jB) {
        const unsigned cont = ~(runmask >> j);                  // first zero bit = end of the run that starts at row j
        const int run = min(cont ? __ffs(cont) : 32, jB - j);
        const int lr = run >= 8 ? 3 : (run >= 4 ? 2 : (run >= 2 ? 1 : 0));
        fn(j, lr);
        j += 1 << lr;
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

// ---- forward piece: bin by bin, straight-line ---------------------------------------------------------------------
// Bin p needs at most 4 staged rows: A[x] = sum_i w_i row_i[x] (y-step, registers), the lane parks its 4*NQ column sums
// in its private scratch column, and the x-step reads back just the 4 taps of each output q:
// Out[p][q] (+)= sum_i w_i A[col_i] -> staging tile (stride 49 floats per lane: conflict-free).
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
    for (int p = p0; p < p1; p++) {
        const uint4 bi = lds128(bins + 32u * p);
        const float4 bw = lds128f(bins + 32u * p + 16u);
        const int nr = (int)bi.y;
        const uint32_t ra = slot + (uint32_t)((int)bi.x - jA) * RB;
        float A[X];
        if (nr > 0) {
            float4 v0[NQ], v1[NQ];
#pragma unroll
            for (int k = 0; k < NQ; k++) { v0[k] = lds128f(ra + off[k]); v1[k] = lds128f(ra + (nr > 1 ? RB : 0u) + off[k]); }
#pragma unroll
            for (int k = 0; k < NQ; k++) {
                A[4 * k] = __fmaf_rn(bw.y, v1[k].x, mul(bw.x, v0[k].x)); A[4 * k + 1] = __fmaf_rn(bw.y, v1[k].y, mul(bw.x, v0[k].y));
                A[4 * k + 2] = __fmaf_rn(bw.y, v1[k].z, mul(bw.x, v0[k].z)); A[4 * k + 3] = __fmaf_rn(bw.y, v1[k].w, mul(bw.x, v0[k].w));
            }
            if (nr > 2) {
#pragma unroll
                for (int k = 0; k < NQ; k++) {
                    const float4 v = lds128f(ra + 2u * RB + off[k]);
                    A[4 * k] = __fmaf_rn(bw.z, v.x, A[4 * k]); A[4 * k + 1] = __fmaf_rn(bw.z, v.y, A[4 * k + 1]);
                    A[4 * k + 2] = __fmaf_rn(bw.z, v.z, A[4 * k + 2]); A[4 * k + 3] = __fmaf_rn(bw.z, v.w, A[4 * k + 3]);
                }
            }
            if (nr > 3) {
#pragma unroll
                for (int k = 0; k < NQ; k++) {
                    const float4 v = lds128f(ra + 3u * RB + off[k]);
                    A[4 * k] = __fmaf_rn(bw.w, v.x, A[4 * k]); A[4 * k + 1] = __fmaf_rn(bw.w, v.y, A[4 * k + 1]);
                    A[4 * k + 2] = __fmaf_rn(bw.w, v.z, A[4 * k + 2]); A[4 * k + 3] = __fmaf_rn(bw.w, v.w, A[4 * k + 3]);
                }
            }
        } else {
#pragma unroll
            for (int x = 0; x < X; x++) A[x] = 0.0f;
        }
#pragma unroll
        for (int k = 0; k < NQ; k++) {
            sts32f(sl + col[k], A[4 * k]); sts32f(sl + col[k] + 128u, A[4 * k + 1]);
            sts32f(sl + col[k] + 256u, A[4 * k + 2]); sts32f(sl + col[k] + 384u, A[4 * k + 3]);
        }
        // x-step: every tap load first (the shared-memory accesses are volatile asm and keep their order: interleaving
        // loads with the stores below would expose one shared-memory round trip per output)
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
}

constexpr int kClStgBytes = 6400;           // 32 x 49 floats, padded
constexpr int kClScrBytes = 2048;           // 16 columns x 32 lanes
constexpr size_t kClSmemBytes = 1024 + (size_t)kClSlots * kClSlotBytes + kClStgBytes + kClScrBytes + 2 * sizeof(ClItem) + sizeof(ClBuild) + 64;

template <int P>
__global__ void __launch_bounds__(32, MD_CL_CTAS)
roialign_fwd_cl_kernel(const __grid_constant__ ClMaps maps, const RoiFeat f, const int tma_mask,
                       const float *__restrict__ rois5, const int R, const int seg, const int nchunk,
                       float *__restrict__ out, int32_t *__restrict__ flag, int *__restrict__ ctr)
{
    static_assert(P == kClP, "7x7 only");
    extern __shared__ unsigned char dsm_raw[];
    unsigned char *sp = dsm_raw + ((1024u - (smem_u32(dsm_raw) & 1023u)) & 1023u);
    unsigned char *slots = sp; sp += kClSlots * kClSlotBytes;
    unsigned char *stgp = sp; sp += kClStgBytes;
    unsigned char *scrp = sp; sp += kClScrBytes;
    ClItem *tabs = reinterpret_cast<ClItem *>(sp); sp += 2 * sizeof(ClItem);
    ClBuild *bld = reinterpret_cast<ClBuild *>(sp); sp += sizeof(ClBuild);
    unsigned long long *full = reinterpret_cast<unsigned long long *>(sp);

    const int lane = threadIdx.x;
    const int C = f.C, CH = C / nchunk, ngroups = CH / 32;
    const int total = R * nchunk;
    if (lane == 0) {
        for (int i = 0; i < kClSlots; i++) mbar_init(&full[i], 1);
        fence_barrier_init();
    }
    __syncwarp();
    const uint32_t stg = smem_u32(stgp), scr = smem_u32(scrp);

    // the warp is producer (takes items from the ticket, builds their tables, issues the TMA loads up to kClSlots pieces
    // ahead) and consumer (one staged piece at a time)
    int p_seq = 0, c_seq = 0;                 // sequence numbers (of items with work) the producer / consumer are in
    bool p_have = false, exhausted = false;
    int p_g = 0, p_k = 0, c_g = 0, c_k = 0;
    int n_iss = 0, n_con = 0;
    int tap_seq = -1, tap_ch = -1;
    bool store_pending = false;
    ClTap tp[P];

    auto fetch = [&]() {
        for (;;) {
            int t = 0;
            if (lane == 0) t = atomicAdd(ctr, 1);
            t = __shfl_sync(0xffffffffu, t, 0);
            if (t >= total) { exhausted = true; p_have = false; return; }
            const WorkItem wi = work_item(t, R, seg, nchunk);
            const int st = cl_build_item<P, kClSlotBytes>(tabs[p_seq & 1], *bld, f, tma_mask, rois5, wi.r, wi.chunk, lane);
            if (wi.chunk == 0 && lane == 0) flag[wi.r] = st == CL_DECLINE ? 1 : 0;
            if (st == CL_OK) { p_have = true; p_g = p_k = 0; return; }
            if (st == CL_ZERO) {
                float *o = out + ((int64_t)wi.r * C + (int64_t)wi.chunk * CH) * (P * P);
                for (int i = lane; i < CH * P * P; i += 32) o[i] = 0.0f;
            }
        }
    };

#pragma unroll 1
    for (;;) {
#pragma unroll 1
        for (;;) {
            if (!p_have) {
                if (exhausted || c_seq < p_seq - 1) break;       // (the consumer still reads the table this item would take)
                fetch();
                if (!p_have) break;
            }
            if (n_iss - n_con >= kClSlots) break;
            const ClItem &it = tabs[p_seq & 1];
            const ClPiece pc = it.piece[p_k];
            const int slot = n_iss % kClSlots;
            const int rb = 512 * pc.nqc;
            fence_proxy_async();
            if (lane == 0) mbar_expect_tx(&full[slot], (uint32_t)(pc.rows * rb));
            __syncwarp();
            if (lane == 0) {
                const CUtensorMap *mp = &maps.m[(it.l * 4 + pc.nqc - 1) * kClRsel];
                const int z = it.b * C + it.chunk * CH + 32 * p_g;
                unsigned char *dst = slots + slot * kClSlotBytes - pc.jA * rb;
                cl_for_row_ops(it.runmask, pc.jA, pc.rows, [&](int j, int lr) {
                    tma_load_3d(dst + j * rb, mp + lr, pc.x0, z, it.yof[j], &full[slot]);
                });
            }
            n_iss++;
            if (++p_k == it.npiece) {
                p_k = 0;
                if (++p_g == ngroups) { p_have = false; p_seq++; }
            }
        }
        if (n_con == n_iss) break;
        // ---- consume one staged piece -----------------------------------------------------------------------------
        const ClItem &it = tabs[c_seq & 1];
        const ClPiece pc = it.piece[c_k];
        if (tap_seq != c_seq || tap_ch != pc.chunk) {
#pragma unroll
            for (int q = 0; q < P; q++) tp[q] = it.tap[pc.chunk][q];
            tap_seq = c_seq; tap_ch = pc.chunk;
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
            case 1: cl_fwd_piece<1>(it, tp, sa, stg, scr, pc.p0, pc.p1, pc.jA, pc.first != 0, lane); break;
            case 2: cl_fwd_piece<2>(it, tp, sa, stg, scr, pc.p0, pc.p1, pc.jA, pc.first != 0, lane); break;
            case 3: cl_fwd_piece<3>(it, tp, sa, stg, scr, pc.p0, pc.p1, pc.jA, pc.first != 0, lane); break;
            default: cl_fwd_piece<4>(it, tp, sa, stg, scr, pc.p0, pc.p1, pc.jA, pc.first != 0, lane); break;
        }
        __syncwarp();
        n_con++;
        const int npiece = it.npiece;
        if (c_k == npiece - 1) {                                 // the 32 x 49 outputs of this (RoI, channel group) are complete
            fence_proxy_async();
            __syncwarp();
            if (lane == 0) {
                bulk_store_1d(out + ((int64_t)it.r * C + it.chunk * CH + 32 * c_g) * (P * P), stgp, kClStageBytes);
                bulk_commit();
            }
            store_pending = true;
        }
        if (++c_k == npiece) {
            c_k = 0;
            if (++c_g == ngroups) { c_g = 0; c_seq++; }
        }
    }
    if (lane == 0) {
        bulk_wait_all<0>();
        // the last CTA out re-arms the ticket for the next launch on this stream
        __threadfence();
        if (atomicAdd(ctr + 1, 1) == (int)gridDim.x - 1) { ctr[0] = 0; ctr[1] = 0; __threadfence(); }
    }
}

// dense Ax[q][16] of one column chunk (columns [4*qa, 4*qa + 4*nqc) of the footprint) -- the backward's x operator
MD_DEVINL void cl_build_ax(float *axw, const ClItem &it, int qa, int nqc, int lane)
{
    constexpr int P = kClP, S = 2;
#pragma unroll
    for (int i = 0; i < 4; i++) axw[lane + 32 * i] = 0.0f;          // 7 x 16 = 112 floats (128 reserved)
    __syncwarp();
    const int c0 = 4 * qa, c1 = c0 + 4 * nqc;
#pragma unroll
    for (int ph = 0; ph < S; ph++) {
        if (lane < P) {
            const int s = lane * S + ph;
            const int lo = it.xlo[s], hi = it.xhi[s];
            const float wl = it.xwl[s], wh = it.xwh[s];
            if (lo >= c0 && lo < c1) axw[lane * 16 + lo - c0] += wl;
            if (hi >= c0 && hi < c1) axw[lane * 16 + hi - c0] += wh;
        }
    }
    __syncwarp();
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
constexpr size_t kClbSmemBytes = 1024 + (size_t)kClSlots * kClbSlotBytes + 2 * kClbGBufBytes + 512 + 2 * sizeof(ClItem) + sizeof(ClBuild) + 64;

// one piece, bin by bin: T[x] = sum_q dY[p][q] Ax[q][x] (registers), then the bin's <= 4 rows of the slot += w_i T
// (the rows of a piece are zeroed first; the lane owns its channel's bytes of every row, so plain read-modify-write)
template <int NQ>
MD_DEVINL void cl_bwd_piece(const ClItem &it, uint32_t slot, uint32_t gbuf, uint32_t axw, int p0, int p1, int jA, int rows, int lane)
{
    constexpr int P = kClP, X = 4 * NQ;
    constexpr uint32_t RB = (uint32_t)cl_row_bytes(NQ);
    uint32_t off[NQ];
#pragma unroll
    for (int k = 0; k < NQ; k++) off[k] = (uint32_t)(lane * cl_pitch(NQ) + 16 * ((k + cl_rot(NQ, lane)) % NQ));
    // Ax in the lane's (rotated) quad order: ax[q][4k + i] = Ax[q][4 * quad(k) + i]
    const int rot = cl_rot(NQ, lane);
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
        float T[X];
#pragma unroll
        for (int x = 0; x < X; x++) T[x] = 0.0f;
        const uint32_t ga = go + (uint32_t)p * (P * 4);
#pragma unroll
        for (int q = 0; q < P; q++) {
            const float g = lds32f(ga + 4 * q);
#pragma unroll
            for (int k = 0; k < NQ; k++) {
                const float4 a = lds128f(axw + (uint32_t)(q * 16 + 4 * ((k + rot) % NQ)) * 4u);
                T[4 * k] = __fmaf_rn(g, a.x, T[4 * k]); T[4 * k + 1] = __fmaf_rn(g, a.y, T[4 * k + 1]);
                T[4 * k + 2] = __fmaf_rn(g, a.z, T[4 * k + 2]); T[4 * k + 3] = __fmaf_rn(g, a.w, T[4 * k + 3]);
            }
        }
        const uint32_t ra = slot + (uint32_t)((int)bi.x - jA) * RB;
        const float w[4] = { bw.x, bw.y, bw.z, bw.w };
        // all loads, then all FMAs, then all stores (volatile shared-memory accesses keep their order)
        float4 d[4][NQ];
#pragma unroll
        for (int i = 0; i < 4; i++) {
            if (i < nr) {
#pragma unroll
                for (int k = 0; k < NQ; k++) d[i][k] = lds128f(ra + (uint32_t)i * RB + off[k]);
            }
        }
#pragma unroll
        for (int i = 0; i < 4; i++) {
            if (i < nr) {
#pragma unroll
                for (int k = 0; k < NQ; k++) {
                    d[i][k].x = __fmaf_rn(w[i], T[4 * k], d[i][k].x); d[i][k].y = __fmaf_rn(w[i], T[4 * k + 1], d[i][k].y);
                    d[i][k].z = __fmaf_rn(w[i], T[4 * k + 2], d[i][k].z); d[i][k].w = __fmaf_rn(w[i], T[4 * k + 3], d[i][k].w);
                    sts128f(ra + (uint32_t)i * RB + off[k], d[i][k]);
                }
            }
        }
    }
}

template <int P>
__global__ void __launch_bounds__(32, MD_CLB_CTAS)
roialign_bwd_cl_kernel(const __grid_constant__ ClMaps maps, const RoiFeat f, const int tma_mask,
                       const float *__restrict__ rois5, const int R, const int seg, const int nchunk,
                       const float *__restrict__ dout, int32_t *__restrict__ flag, int *__restrict__ ctr)
{
    static_assert(P == kClP, "7x7 only");
    extern __shared__ unsigned char dsm_raw[];
    unsigned char *sp = dsm_raw + ((1024u - (smem_u32(dsm_raw) & 1023u)) & 1023u);
    unsigned char *slots = sp; sp += kClSlots * kClbSlotBytes;
    unsigned char *gbufs = sp; sp += 2 * kClbGBufBytes;
    float *axw_p = reinterpret_cast<float *>(sp); sp += 512;
    ClItem *tabs = reinterpret_cast<ClItem *>(sp); sp += 2 * sizeof(ClItem);
    ClBuild *bld = reinterpret_cast<ClBuild *>(sp); sp += sizeof(ClBuild);
    unsigned long long *gfull = reinterpret_cast<unsigned long long *>(sp);

    const int lane = threadIdx.x;
    const int C = f.C, CH = C / nchunk, ngroups = CH / 32;
    const int total = R * nchunk;
    if (lane == 0) {
        mbar_init(&gfull[0], 1); mbar_init(&gfull[1], 1);
        fence_barrier_init();
    }
    __syncwarp();
    const uint32_t axw = smem_u32(axw_p);

    // producer = the dY loads, one (item, channel group) ahead of the consumer; consumer = everything else
    int p_seq = 0, c_seq = 0;
    bool p_have = false, exhausted = false;
    int p_g = 0, c_g = 0;
    int g_iss = 0, g_con = 0;             // dY buffers issued / released
    int n_piece = 0;                      // pieces written so far (slot ring position)
    int ax_seq = -1, ax_xc = -1;

    auto fetch = [&]() {
        for (;;) {
            int t = 0;
            if (lane == 0) t = atomicAdd(ctr, 1);
            t = __shfl_sync(0xffffffffu, t, 0);
            if (t >= total) { exhausted = true; p_have = false; return; }
            const WorkItem wi = work_item(t, R, seg, nchunk);
            const int st = cl_build_item<P, kClbSlotBytes>(tabs[p_seq & 1], *bld, f, tma_mask, rois5, wi.r, wi.chunk, lane);
            if (wi.chunk == 0 && lane == 0) flag[wi.r] = st == CL_DECLINE ? 1 : 0;
            if (st == CL_OK) { p_have = true; p_g = 0; return; }      // CL_ZERO: no sample in range -> no gradient
        }
    };

#pragma unroll 1
    for (;;) {
#pragma unroll 1
        for (;;) {
            if (!p_have) {
                if (exhausted || c_seq < p_seq - 1) break;
                fetch();
                if (!p_have) break;
            }
            if (g_iss - g_con >= 2) break;
            const ClItem &it = tabs[p_seq & 1];
            const int b = g_iss & 1;
            if (lane == 0) {
                fence_proxy_async();
                mbar_expect_tx(&gfull[b], kClStageBytes);
                bulk_load_1d(gbufs + b * kClbGBufBytes, dout + ((int64_t)it.r * C + it.chunk * CH + 32 * p_g) * (P * P), kClStageBytes, &gfull[b]);
            }
            g_iss++;
            if (++p_g == ngroups) { p_have = false; p_seq++; }
        }
        if (g_con == g_iss) break;
        // ---- one (item, channel group): all its pieces ----------------------------------------------------------
        const ClItem &it = tabs[c_seq & 1];
        const int gb = g_con & 1;
        mbar_wait(&gfull[gb], (uint32_t)(g_con >> 1) & 1u);
        const uint32_t gbuf = smem_u32(gbufs + gb * kClbGBufBytes);
        const int nq = it.nq, nxc = it.nxc, npiece = it.npiece;
        const int z = it.b * C + it.chunk * CH + 32 * c_g;
#pragma unroll 1
        for (int k = 0; k < npiece; k++) {
            const ClPiece pc = it.piece[k];
            if (pc.rows <= 0) continue;
            if (ax_seq != c_seq || ax_xc != pc.chunk) {
                cl_build_ax(axw_p, it, pc.chunk * nq / nxc, pc.nqc, lane);
                ax_seq = c_seq; ax_xc = pc.chunk;
            }
            const int rb = 512 * pc.nqc;
            const int slot = n_piece % kClSlots;
            bulk_wait_read<kClSlots - 1>();          // this lane's reduce that last read the slot has finished reading
            __syncwarp();
            const uint32_t sa = smem_u32(slots + slot * kClbSlotBytes);
            switch (pc.nqc) {
                case 1: cl_bwd_piece<1>(it, sa, gbuf, axw, pc.p0, pc.p1, pc.jA, pc.rows, lane); break;
                case 2: cl_bwd_piece<2>(it, sa, gbuf, axw, pc.p0, pc.p1, pc.jA, pc.rows, lane); break;
                case 3: cl_bwd_piece<3>(it, sa, gbuf, axw, pc.p0, pc.p1, pc.jA, pc.rows, lane); break;
                default: cl_bwd_piece<4>(it, sa, gbuf, axw, pc.p0, pc.p1, pc.jA, pc.rows, lane); break;
            }
            fence_proxy_async();
            __syncwarp();
            if (lane == 0) {
                const CUtensorMap *mp = &maps.m[(it.l * 4 + pc.nqc - 1) * kClRsel];
                const unsigned char *src = slots + slot * kClbSlotBytes - pc.jA * rb;
                cl_for_row_ops(it.runmask, pc.jA, pc.rows, [&](int j, int lr) {
                    tma_reduce_add_3d(mp + lr, pc.x0, z, it.yof[j], src + j * rb);
                });
            }
            bulk_commit();
            n_piece++;
        }
        __syncwarp();                                    // every lane is done with this dY buffer
        g_con++;
        if (++c_g == ngroups) { c_g = 0; c_seq++; }
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
cudaError_t launch_roialign_fwd_cl(const FeatSet &fs, const RoiFeat &f, const float *rois5, int R, int P, float *out,
                                   int32_t *fallback_flag, int *ctr, cudaStream_t s, bool *launched)
{
    *launched = false;
    if ((fs.C & 31) || P != kClP || R <= 0) return cudaSuccess;
    if (!cl_enabled()) return cudaSuccess;
    ClMaps maps;
    const int mask = cl_build_maps(fs, &maps);
    if (!mask) return cudaSuccess;
    auto kern = roialign_fwd_cl_kernel<kClP>;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kClSmemBytes);   // per device: set every time
    if (e != cudaSuccess) return e;
    cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
    const int nchunk = cl_chunks_for(fs.C);
    const int total = R * nchunk, grid = cl_grid(MD_CL_CTAS);
    kern<<<total < grid ? total : grid, 32, kClSmemBytes, s>>>(maps, f, mask, rois5, R, 512, nchunk, out, fallback_flag, ctr);
    *launched = true;
    return cudaGetLastError();
}


cudaError_t launch_roialign_bwd_cl(const FeatSet &fs, const RoiFeat &f, const float *rois5, int R, int P, const float *dout,
                                   int32_t *fallback_flag, int *ctr, cudaStream_t s, bool *launched)
{
    *launched = false;
    if ((fs.C & 31) || P != kClP || R <= 0) return cudaSuccess;
    if (!cl_enabled()) return cudaSuccess;
    ClMaps maps;
    const int mask = cl_build_maps(fs, &maps);
    if (!mask) return cudaSuccess;
    auto kern = roialign_bwd_cl_kernel<kClP>;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kClbSmemBytes);
    if (e != cudaSuccess) return e;
    cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
    const int nchunk = cl_chunks_for(fs.C);
    const int total = R * nchunk, grid = cl_grid(MD_CLB_CTAS);
    kern<<<total < grid ? total : grid, 32, kClbSmemBytes, s>>>(maps, f, mask, rois5, R, 512, nchunk, dout, fallback_flag, ctr);
    *launched = true;
    return cudaGetLastError();
}

}  // namespace md
