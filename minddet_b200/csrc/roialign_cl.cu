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
#define MD_CL_SLOT_KB 14
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

// Row pitch of one channel inside a staged row.  Lane c reads channel c with LDS.128, which is conflict-free when the
// pitch is an ODD number of 16-byte quads; even quad counts are padded by one quad (the box is one quad wider: 16 bytes
// more per row and channel are fetched and never used).  The 128-byte TMA swizzle would avoid the padding, but it
// returned wrong data for 32-byte rows and on the level whose row pitch is not a multiple of 128 bytes (measured, r2).
__host__ __device__ constexpr int cl_box_quads(int nq) { return (nq & 1) ? nq : nq + 1; }
__host__ __device__ constexpr bool cl_swizzled(int) { return false; }
__host__ __device__ constexpr int cl_pitch(int nq) { return 16 * cl_box_quads(nq); }
__host__ __device__ constexpr int cl_row_bytes(int nq) { return 32 * cl_pitch(nq); }

struct ClMaps { CUtensorMap m[kClLevels * 4]; };       // [level][nq - 1]: box = {4 * cl_box_quads(nq), 1, 32}

constexpr int kClEnt = 28;                  // y entries per RoI (rows, or (sample, row) pairs in dense mode)

struct __align__(16) ClItem {
    float ew[kClEnt][12];            // Ay^T: weight of bins 0..6 of entry e (1/S folded in; columns 7..11 stay zero)
    int erow[32], epz[32];           // entry -> compact row, last bin the entry touches
    int xlo[16], xhi[16];            // x samples, columns relative to x_lo
    float xwl[16], xwh[16];
    int yof[32];                     // feature row of compact row j
    int pa[32], pz[32];              // first / last bin touching compact row j
    int ja[8], jb[8];                // compact rows [ja[p], jb[p]) carry bin p
    int pp0[8], pj0[8], pj1[8], pe0[8], pe1[8];   // piece k: bins [pp0[k], pp0[k+1]), rows [pj0, pj1), entries [pe0, pe1)
    int r, chunk, b, l, x_lo, nq, nrows, dense, npp, nxc;
    int pad_[2];
};

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

// ---- per-item geometry (whole warp): lanes 0..13 own the y samples, lanes 16..29 the x samples -------------------
// Entries: normally one per compact row, carrying that row's weight for every bin (at most 3 consecutive bins are
// non-zero).  RoIs lower than ~5 feature rows can put more than 3 bins on one row ("dense"): their entries are the 28
// (sample, row) pairs, one bin each, so the 3-bin window of the kernels below still covers every entry.
template <int P, int SLOTB>
MD_DEVINL int cl_build_item(ClItem &it, const RoiFeat &f, int tma_mask, const float *__restrict__ rois5, int r, int chunk, int lane)
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

    float *ew = &it.ew[0][0];
    for (int i = lane; i < kClEnt * 12; i += 32) ew[i] = 0.0f;
    it.pa[lane] = P; it.pz[lane] = -1;
    __syncwarp();
    if (is_y && ok) {
        it.yof[idx_lo] = lo; it.yof[idx_hi] = hi;
        atomicMin(&it.pa[idx_lo], s / S); atomicMin(&it.pa[idx_hi], s / S);
        atomicMax(&it.pz[idx_lo], s / S); atomicMax(&it.pz[idx_hi], s / S);
    }
    if (!is_y) {
        it.xlo[s] = ok ? lo - x_lo : 0; it.xhi[s] = ok ? hi - x_lo : 0;
        it.xwl[s] = ok ? wl : 0.0f; it.xwh[s] = ok ? wh : 0.0f;
    }
    __syncwarp();
    const int paj = it.pa[lane], pzj = it.pz[lane];
    const int dense = __any_sync(0xffffffffu, lane < nrows && pzj - paj > 2);
    {   // per-bin compact row range
        const int p = min(lane, P - 1);
        const int ok0 = __shfl_sync(0xffffffffu, (int)ok, 2 * p), ok1 = __shfl_sync(0xffffffffu, (int)ok, 2 * p + 1);
        const int lo0 = __shfl_sync(0xffffffffu, idx_lo, 2 * p), lo1 = __shfl_sync(0xffffffffu, idx_lo, 2 * p + 1);
        const int hi0 = __shfl_sync(0xffffffffu, idx_hi, 2 * p), hi1 = __shfl_sync(0xffffffffu, idx_hi, 2 * p + 1);
        const int tot1 = __shfl_sync(0xffffffffu, tot, 2 * p + 1);
        if (lane < P) {
            it.ja[lane] = ok0 ? lo0 : (ok1 ? lo1 : tot1);
            it.jb[lane] = ok1 ? hi1 + 1 : (ok0 ? hi0 + 1 : tot1);
        }
    }
    if (!dense) {
        // entry = compact row.  The two samples of a bin may share a row: add them in a fixed order (even sample, then odd)
        if (lane < kClEnt) { it.erow[lane] = lane; it.epz[lane] = pzj; }
#pragma unroll
        for (int ph = 0; ph < S; ph++) {
            if (is_y && ok && (s % S) == ph) {
                it.ew[idx_lo][s / S] += wl;
                it.ew[idx_hi][s / S] += wh;
            }
            __syncwarp();
        }
    } else if (is_y && s < NS) {
        // entry 2s / 2s+1 = (sample s, its low / high row); rows of invalid samples are never read with a non-zero weight
        const int jl = ok ? idx_lo : 0, jh = ok ? idx_hi : 0;
        it.erow[2 * s] = jl; it.erow[2 * s + 1] = jh;
        it.epz[2 * s] = s / S; it.epz[2 * s + 1] = s / S;
        it.ew[2 * s][s / S] = ok ? wl : 0.0f;
        it.ew[2 * s + 1][s / S] = ok ? wh : 0.0f;
    }
    __syncwarp();
    // pieces along p: as many whole bins as fit one stage slot (uniform; lane 0 writes)
    const int nxc = (nq + 3) >> 2, nqc_max = (nq + nxc - 1) / nxc;
    const int maxrows = min(kClRows, SLOTB / cl_row_bytes(nqc_max));
    int k = 0, p = 0;
    while (p < P) {
        const int jA = it.ja[p];
        int jB = it.jb[p];
        const int p0 = p;
        p++;
        while (p < P && it.jb[p] - jA <= maxrows) { jB = max(jB, it.jb[p]); p++; }
        if (lane == 0) {
            it.pp0[k] = p0; it.pj0[k] = jA; it.pj1[k] = jB;
            it.pe0[k] = dense ? 4 * p0 : jA; it.pe1[k] = dense ? 4 * p : jB;
        }
        k++;
    }
    if (lane == 0) {
        it.pp0[k] = P;
        it.r = r; it.chunk = chunk; it.b = g.b; it.l = g.l; it.x_lo = x_lo; it.nq = nq; it.nrows = nrows; it.dense = dense;
        it.npp = k; it.nxc = nxc;
    }
    __syncwarp();
    return CL_OK;
}

// dense Ax[q][16] of one column chunk (columns [4*qa, 4*qa + 4*nqc) of the footprint)
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

template <int NQ>
MD_DEVINL void cl_lane_offsets(uint32_t (&off)[NQ], int lane)
{
#pragma unroll
    for (int k = 0; k < NQ; k++) {
        uint32_t o = (uint32_t)(lane * cl_pitch(NQ) + 16 * k);
        if (cl_swizzled(NQ)) o ^= ((o >> 7) & 7u) << 4;
        off[k] = o;
    }
}

// ---- forward piece ------------------------------------------------------------------------------------------------
// Entries are walked once, in order; a window of three bin accumulators A[0..2] = bins cur..cur+2 follows them (an entry
// touches at most 3 consecutive bins and the first bin never decreases).  A bin that leaves the window is complete: its
// x-step  Out[p][q] (+)= sum_x Ax[q][x] A[x]  runs at once and goes to the staging tile (stride 49 floats per lane:
// conflict-free).  One copy of the row body and one of the x-step per NQ keeps the kernel inside the instruction cache
// (a first version with every bin's code unrolled was 190 KB of SASS and ran 4x slower, fetch-bound).
template <int NQ>
MD_DEVINL void cl_fwd_piece(const ClItem &it, uint32_t slot, uint32_t stg, uint32_t axw, int p0, int p1, int jA, int eA, int eB,
                            bool first, int lane)
{
    constexpr int P = kClP, X = 4 * NQ;
    float A[3][X];
#pragma unroll
    for (int k = 0; k < 3; k++)
#pragma unroll
        for (int x = 0; x < X; x++) A[k][x] = 0.0f;
    uint32_t off[NQ];
    cl_lane_offsets<NQ>(off, lane);
    const uint32_t ew = smem_u32(&it.ew[0][0]), erow = smem_u32(&it.erow[0]), epz = smem_u32(&it.epz[0]);
    const uint32_t rowbase = slot - (uint32_t)jA * (uint32_t)cl_row_bytes(NQ);
    const uint32_t so = stg + (uint32_t)lane * (P * P * 4);
    int cur = p0, e = eA;
    float4 v[NQ];
    {
        const uint32_t ra = rowbase + (e < eB ? lds32(erow + 4 * e) : (uint32_t)jA) * (uint32_t)cl_row_bytes(NQ);
#pragma unroll
        for (int k = 0; k < NQ; k++) v[k] = lds128f(ra + off[k]);
    }
#pragma unroll 1
    for (;;) {
        const int pze = e < eB ? (int)lds32(epz + 4 * e) : 64;
#pragma unroll 1
        while (cur + 2 < pze && cur < p1) {
            // bin `cur` is complete: x-step, then slide the window
            const uint32_t oa = so + (uint32_t)cur * (P * 4);
#pragma unroll
            for (int q = 0; q < P; q++) {
                float acc = 0.0f;
#pragma unroll
                for (int k = 0; k < NQ; k++) {
                    const float4 a = lds128f(axw + (uint32_t)(q * 16 + 4 * k) * 4u);
                    acc = __fmaf_rn(a.x, A[0][4 * k], acc); acc = __fmaf_rn(a.y, A[0][4 * k + 1], acc);
                    acc = __fmaf_rn(a.z, A[0][4 * k + 2], acc); acc = __fmaf_rn(a.w, A[0][4 * k + 3], acc);
                }
                if (!first) acc = add(acc, lds32f(oa + 4 * q));
                sts32f(oa + 4 * q, acc);
            }
#pragma unroll
            for (int x = 0; x < X; x++) { A[0][x] = A[1][x]; A[1][x] = A[2][x]; A[2][x] = 0.0f; }
            cur++;
        }
        if (e >= eB || cur >= p1) break;
        float4 vn[NQ];
        {
            const uint32_t ra = rowbase + (e + 1 < eB ? lds32(erow + 4 * (e + 1)) : (uint32_t)jA) * (uint32_t)cl_row_bytes(NQ);
#pragma unroll
            for (int k = 0; k < NQ; k++) vn[k] = lds128f(ra + off[k]);
        }
        const uint32_t wa = ew + (uint32_t)(e * 12 + cur) * 4u;
        const float w0 = lds32f(wa), w1 = lds32f(wa + 4), w2 = lds32f(wa + 8);
        if (w0 != 0.0f || w1 != 0.0f || w2 != 0.0f) {
#pragma unroll
            for (int k = 0; k < NQ; k++) {
                A[0][4 * k] = __fmaf_rn(w0, v[k].x, A[0][4 * k]); A[0][4 * k + 1] = __fmaf_rn(w0, v[k].y, A[0][4 * k + 1]);
                A[0][4 * k + 2] = __fmaf_rn(w0, v[k].z, A[0][4 * k + 2]); A[0][4 * k + 3] = __fmaf_rn(w0, v[k].w, A[0][4 * k + 3]);
                A[1][4 * k] = __fmaf_rn(w1, v[k].x, A[1][4 * k]); A[1][4 * k + 1] = __fmaf_rn(w1, v[k].y, A[1][4 * k + 1]);
                A[1][4 * k + 2] = __fmaf_rn(w1, v[k].z, A[1][4 * k + 2]); A[1][4 * k + 3] = __fmaf_rn(w1, v[k].w, A[1][4 * k + 3]);
                A[2][4 * k] = __fmaf_rn(w2, v[k].x, A[2][4 * k]); A[2][4 * k + 1] = __fmaf_rn(w2, v[k].y, A[2][4 * k + 1]);
                A[2][4 * k + 2] = __fmaf_rn(w2, v[k].z, A[2][4 * k + 2]); A[2][4 * k + 3] = __fmaf_rn(w2, v[k].w, A[2][4 * k + 3]);
            }
        }
#pragma unroll
        for (int k = 0; k < NQ; k++) v[k] = vn[k];
        e++;
    }
}

struct ClShared {
    unsigned char *slots;
    float *stg, *axw;
    ClItem *tabs;
    unsigned long long *full;
};
constexpr int kClTailBytes = 6400 + 512 + 2 * (int)sizeof(ClItem) + 64;
constexpr size_t kClSmemBytes = 1024 + (size_t)kClSlots * kClSlotBytes + kClTailBytes;

MD_DEVINL ClShared cl_carve(unsigned char *raw)
{
    const uint32_t a = smem_u32(raw);
    unsigned char *p = raw + ((1024u - (a & 1023u)) & 1023u);
    ClShared s;
    s.slots = p; p += kClSlots * kClSlotBytes;
    s.stg = reinterpret_cast<float *>(p); p += 6400;
    s.axw = reinterpret_cast<float *>(p); p += 512;
    s.tabs = reinterpret_cast<ClItem *>(p); p += 2 * sizeof(ClItem);
    s.full = reinterpret_cast<unsigned long long *>(p);
    return s;
}

template <int P>
__global__ void __launch_bounds__(32, MD_CL_CTAS)
roialign_fwd_cl_kernel(const __grid_constant__ ClMaps maps, const RoiFeat f, const int tma_mask,
                       const float *__restrict__ rois5, const int R, const int seg, const int nchunk,
                       float *__restrict__ out, int32_t *__restrict__ flag, int *__restrict__ ctr)
{
    static_assert(P == kClP, "7x7 only");
    extern __shared__ unsigned char dsm_raw[];
    const ClShared sh = cl_carve(dsm_raw);
    const int lane = threadIdx.x;
    const int C = f.C, CH = C / nchunk, ngroups = CH / 32;
    const int total = R * nchunk;
    if (lane == 0) {
        for (int i = 0; i < kClSlots; i++) mbar_init(&sh.full[i], 1);
        fence_barrier_init();
    }
    __syncwarp();
    const uint32_t stg = smem_u32(sh.stg), axw = smem_u32(sh.axw);

    int p_seq = 0, c_seq = 0;                 // sequence numbers (of items with work) the producer / consumer are in
    bool p_have = false, exhausted = false;
    int p_g = 0, p_xc = 0, p_pp = 0, c_g = 0, c_xc = 0, c_pp = 0;
    int n_iss = 0, n_con = 0;
    int ax_seq = -1, ax_xc = -1;
    bool store_pending = false;

    auto fetch = [&]() {
        for (;;) {
            int t = 0;
            if (lane == 0) t = atomicAdd(ctr, 1);
            t = __shfl_sync(0xffffffffu, t, 0);
            if (t >= total) { exhausted = true; p_have = false; return; }
            const WorkItem wi = work_item(t, R, seg, nchunk);
            const int st = cl_build_item<P, kClSlotBytes>(sh.tabs[p_seq & 1], f, tma_mask, rois5, wi.r, wi.chunk, lane);
            if (wi.chunk == 0 && lane == 0) flag[wi.r] = st == CL_DECLINE ? 1 : 0;
            if (st == CL_OK) { p_have = true; p_g = p_xc = p_pp = 0; return; }
            if (st == CL_ZERO) {
                float *o = out + ((int64_t)wi.r * C + (int64_t)wi.chunk * CH) * (P * P);
                for (int i = lane; i < CH * P * P; i += 32) o[i] = 0.0f;
            }
        }
    };

#pragma unroll 1
    for (;;) {
        // ---- producer: keep the stage slots full ------------------------------------------------------------------
#pragma unroll 1
        for (;;) {
            if (!p_have) {
                if (exhausted || c_seq < p_seq - 1) break;       // (the consumer still reads the table this item would take)
                fetch();
                if (!p_have) break;
            }
            if (n_iss - n_con >= kClSlots) break;
            const ClItem &it = sh.tabs[p_seq & 1];
            const int nq = it.nq, nxc = it.nxc;
            const int qa = p_xc * nq / nxc, nqc = (p_xc + 1) * nq / nxc - qa;
            const int jA = it.pj0[p_pp], rows = it.pj1[p_pp] - jA;
            const int slot = n_iss % kClSlots;
            const int rb = nqc == 1 ? cl_row_bytes(1) : (nqc == 2 ? cl_row_bytes(2) : (nqc == 3 ? cl_row_bytes(3) : cl_row_bytes(4)));
            fence_proxy_async();
            if (lane == 0) mbar_expect_tx(&sh.full[slot], (uint32_t)(rows * rb));
            __syncwarp();
            if (lane < rows)
                tma_load_3d(sh.slots + slot * kClSlotBytes + lane * rb, &maps.m[it.l * 4 + nqc - 1], it.x_lo + 4 * qa,
                            it.yof[jA + lane], it.b * C + it.chunk * CH + 32 * p_g, &sh.full[slot]);
            n_iss++;
            if (++p_pp == it.npp) {
                p_pp = 0;
                if (++p_xc == nxc) {
                    p_xc = 0;
                    if (++p_g == ngroups) { p_have = false; p_seq++; }
                }
            }
        }
        if (n_con == n_iss) break;
        // ---- consumer: one staged piece ---------------------------------------------------------------------------
        const ClItem &it = sh.tabs[c_seq & 1];
        const int nq = it.nq, nxc = it.nxc, npp = it.npp;
        const int qa = c_xc * nq / nxc, nqc = (c_xc + 1) * nq / nxc - qa;
        if (ax_seq != c_seq || ax_xc != c_xc) {
            cl_build_ax(sh.axw, it, qa, nqc, lane);
            ax_seq = c_seq; ax_xc = c_xc;
        }
        if (store_pending && c_xc == 0 && c_pp == 0) {           // the staging tile is about to be overwritten
            if (lane == 0) bulk_wait_read<0>();
            __syncwarp();
            store_pending = false;
        }
        const int slot = n_con % kClSlots;
        mbar_wait(&sh.full[slot], (uint32_t)(n_con / kClSlots) & 1u);
        const uint32_t sa = smem_u32(sh.slots + slot * kClSlotBytes);
        const int p0 = it.pp0[c_pp], p1 = it.pp0[c_pp + 1], jA = it.pj0[c_pp], eA = it.pe0[c_pp], eB = it.pe1[c_pp];
        const bool first = c_xc == 0;
        switch (nqc) {
            case 1: cl_fwd_piece<1>(it, sa, stg, axw, p0, p1, jA, eA, eB, first, lane); break;
            case 2: cl_fwd_piece<2>(it, sa, stg, axw, p0, p1, jA, eA, eB, first, lane); break;
            case 3: cl_fwd_piece<3>(it, sa, stg, axw, p0, p1, jA, eA, eB, first, lane); break;
            default: cl_fwd_piece<4>(it, sa, stg, axw, p0, p1, jA, eA, eB, first, lane); break;
        }
        __syncwarp();
        n_con++;
        if (c_pp == npp - 1 && c_xc == nxc - 1) {                // the 32 x 49 outputs of this (RoI, channel group) are complete
            fence_proxy_async();
            __syncwarp();
            if (lane == 0) {
                bulk_store_1d(out + ((int64_t)it.r * C + it.chunk * CH + 32 * c_g) * (P * P), sh.stg, kClStageBytes);
                bulk_commit();
            }
            store_pending = true;
        }
        if (++c_pp == npp) {
            c_pp = 0;
            if (++c_xc == nxc) {
                c_xc = 0;
                if (++c_g == ngroups) { c_g = 0; c_seq++; }
            }
        }
    }
    if (lane == 0) {
        bulk_wait_all<0>();
        // the last CTA out re-arms the ticket for the next launch on this stream
        __threadfence();
        if (atomicAdd(ctr + 1, 1) == (int)gridDim.x - 1) { ctr[0] = 0; ctr[1] = 0; __threadfence(); }
    }
}

// =====================================================================================================
// backward: dX += Ay^T (dY Ax) per channel; rows leave through cp.reduce.async.bulk.tensor (add, at L2)
// =====================================================================================================
#ifndef MD_CLB_SLOT_KB
#define MD_CLB_SLOT_KB 12
#endif
constexpr int kClbSlotBytes = MD_CLB_SLOT_KB * 1024;
constexpr int kClbGBufBytes = 6400;
constexpr size_t kClbSmemBytes = 1024 + (size_t)kClSlots * kClbSlotBytes + 2 * kClbGBufBytes + 512 + 2 * sizeof(ClItem) + 64;

// one piece: the same window walk, transposed.  T[k] = (dY Ax) of bin cur+k (zero outside [p0, p1)) is built when the bin
// enters the window; every entry writes D = sum_k w_k T[k] into its row of the slot (dense mode: adds, rows zeroed first).
template <int NQ>
MD_DEVINL void cl_bwd_piece(const ClItem &it, uint32_t slot, uint32_t gbuf, uint32_t axw, int p0, int p1, int jA, int jB,
                            int eA, int eB, int lane)
{
    constexpr int P = kClP, X = 4 * NQ;
    float T[3][X];
#pragma unroll
    for (int k = 0; k < 3; k++)
#pragma unroll
        for (int x = 0; x < X; x++) T[k][x] = 0.0f;
    uint32_t off[NQ];
    cl_lane_offsets<NQ>(off, lane);
    const uint32_t ew = smem_u32(&it.ew[0][0]), erow = smem_u32(&it.erow[0]), epz = smem_u32(&it.epz[0]);
    const uint32_t rowbase = slot - (uint32_t)jA * (uint32_t)cl_row_bytes(NQ);
    const uint32_t go = gbuf + (uint32_t)lane * (P * P * 4);
    const bool dense = it.dense != 0;
    if (dense || cl_box_quads(NQ) != NQ) {
        // the padding quad of every row is part of the reduce box: it must add zero.  Dense mode accumulates: zero all.
        for (int j = jA; j < jB; j++) {
            const uint32_t ra = rowbase + (uint32_t)j * (uint32_t)cl_row_bytes(NQ);
            if (dense) {
#pragma unroll
                for (int k = 0; k < NQ; k++) sts128f(ra + off[k], make_float4(0.0f, 0.0f, 0.0f, 0.0f));
            }
            if (cl_box_quads(NQ) != NQ) sts128f(ra + (uint32_t)(lane * cl_pitch(NQ) + 16 * NQ), make_float4(0.0f, 0.0f, 0.0f, 0.0f));
        }
    }
    int cur = p0 - 3;
#pragma unroll 1
    for (int e = eA; e < eB; e++) {
        const int pze = (int)lds32(epz + 4 * e);
#pragma unroll 1
        while ((cur < p0 || cur + 2 < pze) && cur < p1) {
            // slide the window: bin cur+3 enters
#pragma unroll
            for (int x = 0; x < X; x++) { T[0][x] = T[1][x]; T[1][x] = T[2][x]; T[2][x] = 0.0f; }
            cur++;
            const int pn = cur + 2;
            if (pn >= p0 && pn < p1) {
                const uint32_t ga = go + (uint32_t)pn * (P * 4);
#pragma unroll
                for (int q = 0; q < P; q++) {
                    const float g = lds32f(ga + 4 * q);
#pragma unroll
                    for (int k = 0; k < NQ; k++) {
                        const float4 a = lds128f(axw + (uint32_t)(q * 16 + 4 * k) * 4u);
                        T[2][4 * k] = __fmaf_rn(g, a.x, T[2][4 * k]); T[2][4 * k + 1] = __fmaf_rn(g, a.y, T[2][4 * k + 1]);
                        T[2][4 * k + 2] = __fmaf_rn(g, a.z, T[2][4 * k + 2]); T[2][4 * k + 3] = __fmaf_rn(g, a.w, T[2][4 * k + 3]);
                    }
                }
            }
        }
        if (cur >= p1) break;
        const uint32_t wa = ew + (uint32_t)(e * 12 + cur) * 4u;
        const float w0 = lds32f(wa), w1 = lds32f(wa + 4), w2 = lds32f(wa + 8);
        const uint32_t ra = rowbase + lds32(erow + 4 * e) * (uint32_t)cl_row_bytes(NQ);
        if (dense && w0 == 0.0f && w1 == 0.0f && w2 == 0.0f) continue;
#pragma unroll
        for (int k = 0; k < NQ; k++) {
            float4 d;
            d.x = __fmaf_rn(w2, T[2][4 * k], __fmaf_rn(w1, T[1][4 * k], mul(w0, T[0][4 * k])));
            d.y = __fmaf_rn(w2, T[2][4 * k + 1], __fmaf_rn(w1, T[1][4 * k + 1], mul(w0, T[0][4 * k + 1])));
            d.z = __fmaf_rn(w2, T[2][4 * k + 2], __fmaf_rn(w1, T[1][4 * k + 2], mul(w0, T[0][4 * k + 2])));
            d.w = __fmaf_rn(w2, T[2][4 * k + 3], __fmaf_rn(w1, T[1][4 * k + 3], mul(w0, T[0][4 * k + 3])));
            if (dense) {
                const float4 o = lds128f(ra + off[k]);
                d.x = add(d.x, o.x); d.y = add(d.y, o.y); d.z = add(d.z, o.z); d.w = add(d.w, o.w);
            }
            sts128f(ra + off[k], d);
        }
    }
}

template <int P>
__global__ void __launch_bounds__(32, MD_CL_CTAS)
roialign_bwd_cl_kernel(const __grid_constant__ ClMaps maps, const RoiFeat f, const int tma_mask,
                       const float *__restrict__ rois5, const int R, const int seg, const int nchunk,
                       const float *__restrict__ dout, int32_t *__restrict__ flag, int *__restrict__ ctr)
{
    static_assert(P == kClP, "7x7 only");
    extern __shared__ unsigned char dsm_raw[];
    const uint32_t a0 = smem_u32(dsm_raw);
    unsigned char *base = dsm_raw + ((1024u - (a0 & 1023u)) & 1023u);
    unsigned char *slots = base;
    unsigned char *gbufs = slots + kClSlots * kClbSlotBytes;
    float *axw_p = reinterpret_cast<float *>(gbufs + 2 * kClbGBufBytes);
    ClItem *tabs = reinterpret_cast<ClItem *>(reinterpret_cast<unsigned char *>(axw_p) + 512);
    unsigned long long *gfull = reinterpret_cast<unsigned long long *>(tabs + 2);

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
            const int st = cl_build_item<P, kClbSlotBytes>(tabs[p_seq & 1], f, tma_mask, rois5, wi.r, wi.chunk, lane);
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
        const int nq = it.nq, nxc = it.nxc, npp = it.npp;
        const int z = it.b * C + it.chunk * CH + 32 * c_g;
#pragma unroll 1
        for (int xc = 0; xc < nxc; xc++) {
            const int qa = xc * nq / nxc, nqc = (xc + 1) * nq / nxc - qa;
            if (ax_seq != c_seq || ax_xc != xc) {
                cl_build_ax(axw_p, it, qa, nqc, lane);
                ax_seq = c_seq; ax_xc = xc;
            }
            const int rb = nqc == 1 ? cl_row_bytes(1) : (nqc == 2 ? cl_row_bytes(2) : (nqc == 3 ? cl_row_bytes(3) : cl_row_bytes(4)));
#pragma unroll 1
            for (int pp = 0; pp < npp; pp++) {
                const int p0 = it.pp0[pp], p1 = it.pp0[pp + 1], jA = it.pj0[pp], jB = it.pj1[pp];
                if (jB <= jA) continue;
                const int slot = n_piece % kClSlots;
                bulk_wait_read<kClSlots - 1>();          // this lane's reduce that last read the slot has finished reading
                __syncwarp();
                const uint32_t sa = smem_u32(slots + slot * kClbSlotBytes);
                switch (nqc) {
                    case 1: cl_bwd_piece<1>(it, sa, gbuf, axw, p0, p1, jA, jB, it.pe0[pp], it.pe1[pp], lane); break;
                    case 2: cl_bwd_piece<2>(it, sa, gbuf, axw, p0, p1, jA, jB, it.pe0[pp], it.pe1[pp], lane); break;
                    case 3: cl_bwd_piece<3>(it, sa, gbuf, axw, p0, p1, jA, jB, it.pe0[pp], it.pe1[pp], lane); break;
                    default: cl_bwd_piece<4>(it, sa, gbuf, axw, p0, p1, jA, jB, it.pe0[pp], it.pe1[pp], lane); break;
                }
                fence_proxy_async();
                __syncwarp();
                if (lane < jB - jA)
                    tma_reduce_add_3d(&maps.m[it.l * 4 + nqc - 1], it.x_lo + 4 * qa, it.yof[jA + lane], z,
                                      slots + slot * kClbSlotBytes + lane * rb);
                bulk_commit();
                n_piece++;
            }
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
            for (int nq = 1; nq <= 4 && ok; nq++) {
                const cuuint64_t dims[3] = { (cuuint64_t)fs.W[l], (cuuint64_t)fs.H[l], (cuuint64_t)fs.B * fs.C };
                const cuuint64_t strides[2] = { (cuuint64_t)fs.W[l] * 4, (cuuint64_t)fs.W[l] * fs.H[l] * 4 };
                const cuuint32_t box[3] = { (cuuint32_t)(4 * cl_box_quads(nq)), 1u, 32u };
                const cuuint32_t estr[3] = { 1, 1, 1 };
                ok = enc(&c.maps.m[l * 4 + nq - 1], CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, fs.feat[l], dims, strides, box, estr,
                         CU_TENSOR_MAP_INTERLEAVE_NONE, cl_swizzled(nq) ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_NONE,
                         CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
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

static int cl_grid()
{
    int dev = 0, sms = 148;
    if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    return sms * MD_CL_CTAS;
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
    const int total = R * nchunk, grid = cl_grid();
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
    const int total = R * nchunk, grid = cl_grid();
    kern<<<total < grid ? total : grid, 32, kClbSmemBytes, s>>>(maps, f, mask, rois5, R, 512, nchunk, dout, fallback_flag, ctr);
    *launched = true;
    return cudaGetLastError();
}

}  // namespace md
