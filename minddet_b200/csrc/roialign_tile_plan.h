// roialign_tile_plan.h -- per-RoI plan of the tile-stationary RoIAlign backward (roialign_tile.cu).
//
// RoIAlign of a 7x7 / S = 2 RoI is separable: dX[y][x] += sum_p sum_q Wy[y][p] . dY[p][q] . Wx[x][q], where
// Wy (<= 28 touched feature rows x 7 bins) and Wx (touched columns x 7 bins) collect the bilinear weights of the 14
// sample rows / 14 sample columns (1/S folded into each).  The plan stores exactly that: the touched rows with their
// dense bin weights, and the column side either dense (footprints <= 16 columns: the usual case) or per bin (<= 4
// distinct columns each).  Host + device code: the CPU check of the plan logic (scripts/tile_plan_check.cpp) compiles the
// same functions with g++ and compares the accumulated gradient with the oracle.
//
// Arithmetic follows roialign_common.cuh / oracle/region_oracle.c (roi_level, roi_geometry, sample_coord, make_tap)
// operation by operation; the library is built with -fmad=false, so the plain operators below are individually rounded
// on the device as on the host.  No reference code exists for this op; semantics:
// oracle/CONVENTIONS.md #14-16, #23.
#pragma once
#include <math.h>
#include <stdint.h>

#if defined(__CUDACC__)
#define MD_TILE_HD __host__ __device__ __forceinline__
#else
#define MD_TILE_HD inline
#endif

namespace md {
namespace tile {

constexpr int kP = 7, kS = 2, kNS = kP * kS;
#ifndef MD_TILE_TH
#define MD_TILE_TH 4
#endif
constexpr int kTH = MD_TILE_TH, kTW = 32, kTC = 32;   // dX tile: rows x columns, channels per warp
constexpr int kMaxRows = 2 * kNS;                 // every sample row touches <= 2 feature rows
constexpr int kDenseCols = 16;

enum { ST_OK = 0, ST_NONE = 1, ST_DECLINE = 2 };  // NONE: no gradient (bad batch index / no valid sample); DECLINE: gather kernel

struct alignas(16) Row { int y; float w[kP]; };               // a touched feature row and its weight per output bin row
struct alignas(16) Bin { int col[4]; float w[4]; };           // the distinct columns of one output bin column (-1: unused)
struct alignas(16) Plan {
    int status, wide, nrows, ncols;
    int x0, x1, y0, y1;                                       // inclusive footprint bounds (feature pixels)
    Row row[kMaxRows];
    float xw[kDenseCols][8];                                  // !wide: xw[x - x0][q]
    Bin bin[kP];                                              // wide: per-bin columns
};
static_assert(sizeof(Plan) == 1664, "plan layout (copied by one bulk load)");

struct alignas(16) Hdr { int key, xr, yr, pad; };             // key = status | level << 8 | image << 16; xr = x0 | x1 << 16
MD_TILE_HD Hdr make_hdr(const Plan &p, int b, int l)
{
    Hdr h;
    h.key = p.status | (l << 8) | (b << 16);
    h.xr = p.x0 | (p.x1 << 16);
    h.yr = p.y0 | (p.y1 << 16);
    h.pad = 0;
    return h;
}

MD_TILE_HD int level_of(const float *r, float finest, int num_levels)
{
    const float w = (r[2] - r[0]) + 1.0f;
    const float h = (r[3] - r[1]) + 1.0f;
    const float s = sqrtf(w * h);
    const float t = s / finest + 1e-6f;
    int l = (t >= 2.0f) + (t >= 4.0f) + (t >= 8.0f);
    for (int k = 4; k < num_levels; k++) l += (t >= (float)(1 << k));
    return l < num_levels - 1 ? l : num_levels - 1;
}

MD_TILE_HD float sample_at(float start, float bin, int p, int i)
{
    const float base = start + (float)p * bin;
    const float o = (((float)i + 0.5f) * bin) / (float)kS;
    return base + o;
}

// 1-D bilinear sample -> (low index, high index, low weight, high weight); mirrors make_tap
MD_TILE_HD bool sample_1d(float v, int extent, int &lo, int &hi, float &wl, float &wh)
{
    if (v < -1.0f || v > (float)extent) return false;
    if (v <= 0.0f) v = 0.0f;
    lo = (int)v;
    if (lo >= extent - 1) { hi = lo = extent - 1; v = (float)lo; } else hi = lo + 1;
    wh = v - (float)lo;
    wl = 1.0f - wh;
    return true;
}

MD_TILE_HD void add_row(Plan &pl, int y, int p, float w)
{
    for (int k = pl.nrows - 1; k >= 0 && k >= pl.nrows - 2; k--)
        if (pl.row[k].y == y) { pl.row[k].w[p] += w; return; }
    Row &r = pl.row[pl.nrows++];
    r.y = y;
    for (int k = 0; k < kP; k++) r.w[k] = 0.0f;
    r.w[p] = w;
}

// roi = {batch, x1, y1, x2, y2}; cfg = MD_CFG_ROI (finest, sample_num, end_mode, -, stride per level).
// Returns image and level through b / l; pl.status says what to do with the RoI.
MD_TILE_HD void plan_roi(const float *roi, int B, int L, const int *H_, const int *W_, const float *cfg, Plan &pl, int &b, int &l)
{
    pl.status = ST_OK; pl.wide = 0; pl.nrows = 0; pl.ncols = 0;
    pl.x0 = pl.x1 = pl.y0 = pl.y1 = 0;
    const float bf = roi[0];
    const bool ok = bf >= 0.0f && bf < (float)B;                 // false for NaN too (CONVENTIONS #23)
    b = ok ? (int)bf : 0;
    l = level_of(roi + 1, cfg[0], L);
    if (!ok) { pl.status = ST_NONE; return; }
    if ((int)cfg[1] != kS) { pl.status = ST_DECLINE; return; }
    const int H = H_[l], W = W_[l];
    const float scale = 1.0f / cfg[4 + l];
    const float em = cfg[2];
    const float sw = roi[1] * scale, sh = roi[2] * scale;
    const float ew = (roi[3] + em) * scale, eh = (roi[4] + em) * scale;
    const float rw = fmaxf(ew - sw, 1.0f), rh = fmaxf(eh - sh, 1.0f);
    const float bw = rw / (float)kP, bh = rh / (float)kP;

    // ---- rows: the sample rows are visited in increasing order, so a feature row is one of the last two entries or new
    for (int p = 0; p < kP; p++)
        for (int i = 0; i < kS; i++) {
            int lo, hi;
            float wl, wh;
            if (!sample_1d(sample_at(sh, bh, p, i), H, lo, hi, wl, wh)) continue;
            add_row(pl, lo, p, 0.5f * wl);
            add_row(pl, hi, p, 0.5f * wh);
        }
    if (pl.nrows == 0) { pl.status = ST_NONE; return; }
    pl.y0 = pl.row[0].y;
    pl.y1 = pl.row[pl.nrows - 1].y;

    // ---- columns
    int lo[kNS], hi[kNS];
    float wl[kNS], wh[kNS];
    bool v[kNS];
    int x0 = 1 << 30, x1 = -1;
    for (int k = 0; k < kNS; k++) {
        v[k] = sample_1d(sample_at(sw, bw, k / kS, k % kS), W, lo[k], hi[k], wl[k], wh[k]);
        if (v[k]) { x0 = lo[k] < x0 ? lo[k] : x0; x1 = hi[k] > x1 ? hi[k] : x1; }
    }
    if (x1 < 0) { pl.status = ST_NONE; return; }
    pl.x0 = x0; pl.x1 = x1; pl.ncols = x1 - x0 + 1;
    if (pl.ncols <= kDenseCols) {
        for (int j = 0; j < kDenseCols; j++)
            for (int q = 0; q < 8; q++) pl.xw[j][q] = 0.0f;
        for (int k = 0; k < kNS; k++)
            if (v[k]) {
                pl.xw[lo[k] - x0][k / kS] += 0.5f * wl[k];
                pl.xw[hi[k] - x0][k / kS] += 0.5f * wh[k];
            }
    } else {
        pl.wide = 1;
        // every bin is assembled in registers and written once (<= 4 distinct columns: two samples x two taps)
        for (int q = 0; q < kP; q++) {
            int c[4] = {-1, -1, -1, -1};
            float w[4] = {0.0f, 0.0f, 0.0f, 0.0f};
            int n = 0;
            for (int i = 0; i < kS; i++) {
                const int k = q * kS + i;
                if (!v[k]) continue;
                for (int t = 0; t < 2; t++) {
                    const int col = t ? hi[k] : lo[k];
                    const float wt = 0.5f * (t ? wh[k] : wl[k]);
                    bool found = false;
                    for (int s = 0; s < 4; s++)
                        if (s < n && c[s] == col) { w[s] += wt; found = true; }
                    if (!found) {
                        for (int s = 0; s < 4; s++)
                            if (s == n) { c[s] = col; w[s] = wt; }
                        n++;
                    }
                }
            }
            for (int s = 0; s < 4; s++) { pl.bin[q].col[s] = c[s]; pl.bin[q].w[s] = w[s]; }
        }
        // the kernel updates bins {0,2,4,6} and then {1,3,5} as two batches of independent read-modify-writes: columns
        // must be distinct inside a batch (always true for bins wider than 4/3 pixel; checked, not assumed)
        for (int q = 0; q + 2 < kP; q++) {
            int mx = -1, mn = 1 << 30;
            for (int s = 0; s < 4; s++) {
                if (pl.bin[q].col[s] > mx) mx = pl.bin[q].col[s];
                if (pl.bin[q + 2].col[s] >= 0 && pl.bin[q + 2].col[s] < mn) mn = pl.bin[q + 2].col[s];
            }
            if (mx >= 0 && mn != (1 << 30) && mx >= mn) { pl.status = ST_DECLINE; return; }
        }
    }
}

// tile index space: level-major, then image, tile row, tile column
struct Grid {
    int L, B;
    int H[8], W[8], nty[8], ntx[8], base[9];
};
MD_TILE_HD void make_grid(Grid &g, int L, int B, const int *H, const int *W)
{
    g.L = L; g.B = B; g.base[0] = 0;
    for (int l = 0; l < L; l++) {
        g.H[l] = H[l]; g.W[l] = W[l];
        g.nty[l] = (H[l] + kTH - 1) / kTH;
        g.ntx[l] = (W[l] + kTW - 1) / kTW;
        g.base[l + 1] = g.base[l] + B * g.nty[l] * g.ntx[l];
    }
}
MD_TILE_HD int tile_id(const Grid &g, int l, int b, int ty, int tx) { return g.base[l] + (b * g.nty[l] + ty) * g.ntx[l] + tx; }

}  // namespace tile
}  // namespace md
