// CPU check of the tile-stationary RoIAlign backward's plan + visit logic (minddet_b200/csrc/roialign_tile_plan.h) against
// the oracle (oracle/region_oracle.c:o_roialign_bwd).  Development tool, not part of the product:
//   g++ -O2 -ffp-contract=off -o /tmp/tile_plan_check scripts/tile_plan_check.cpp oracle/_build/liboracle.so -Wl,-rpath,$PWD/oracle/_build
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>
#include <random>
#include <algorithm>
#include "../minddet_b200/csrc/roialign_tile_plan.h"

extern "C" void o_roialign_bwd(int L, float *const *dfeats, const int *H_, const int *W_, const float *scale_, int B, int C,
                               const float *rois5, int64_t R, const int32_t *lvl_in, float finest, int P, int S, float end_mode,
                               const float *dout);
using namespace md::tile;

int main(int argc, char **argv)
{
    const int L = 4, B = 2, C = 2, R = argc > 1 ? atoi(argv[1]) : 3000;
    const int H[4] = {200, 336 / 2 + 32, 50, 25}, W[4] = {336, 168, 84, 42};
    const float stride[4] = {4, 8, 16, 32};
    float cfg[8] = {56.0f, 2.0f, 0.0f, 0.0f, 4, 8, 16, 32};
    std::mt19937 rng(argc > 2 ? atoi(argv[2]) : 1);
    std::uniform_real_distribution<float> U(0, 1);
    std::vector<float> rois(R * 5);
    for (int r = 0; r < R; r++) {
        const float imw = W[0] * 4.0f, imh = H[0] * 4.0f;
        float cx = U(rng) * imw * 1.2f - 0.1f * imw, cy = U(rng) * imh * 1.2f - 0.1f * imh;
        float w = expf(U(rng) * logf(900.0f / 2.0f)) * 2.0f, h = expf(U(rng) * logf(900.0f / 2.0f)) * 2.0f;
        if (r % 7 == 0) { w = 1.0f + U(rng) * 40; h = 1.0f + U(rng) * 40; }
        rois[r * 5 + 0] = (float)(rng() % B);
        if (r % 97 == 0) rois[r * 5 + 0] = -1.0f;
        if (r % 101 == 0) rois[r * 5 + 0] = (float)B;
        rois[r * 5 + 1] = cx - w / 2; rois[r * 5 + 2] = cy - h / 2; rois[r * 5 + 3] = cx + w / 2; rois[r * 5 + 4] = cy + h / 2;
    }
    std::vector<float> dout((size_t)R * C * 49);
    for (auto &v : dout) v = U(rng) * 2 - 1;
    std::vector<std::vector<float>> ref(L), got(L);
    float *pref[4];
    float scale[4];
    for (int l = 0; l < L; l++) { ref[l].assign((size_t)B * C * H[l] * W[l], 0.0f); got[l] = ref[l]; pref[l] = ref[l].data(); scale[l] = 1.0f / stride[l]; }
    o_roialign_bwd(L, pref, H, W, scale, B, C, rois.data(), R, nullptr, 56.0f, 7, 2, 0.0f, dout.data());

    Grid g; make_grid(g, L, B, H, W);
    std::vector<Plan> plans(R);
    std::vector<Hdr> hdr(R);
    int nwide = 0, ndecl = 0, nnone = 0;
    for (int r = 0; r < R; r++) {
        int b, l;
        plan_roi(&rois[r * 5], B, L, H, W, cfg, plans[r], b, l);
        hdr[r] = make_hdr(plans[r], b, l);
        nwide += plans[r].status == ST_OK && plans[r].wide; ndecl += plans[r].status == ST_DECLINE; nnone += plans[r].status == ST_NONE;
    }
    printf("R %d wide %d declined %d none %d tiles %d\n", R, nwide, ndecl, nnone, g.base[L]);
    long visits = 0, emptyvis = 0;
    for (int l = 0; l < L; l++) for (int b = 0; b < B; b++) for (int ty = 0; ty < g.nty[l]; ty++) for (int tx = 0; tx < g.ntx[l]; tx++) {
        const int ty0 = ty * kTH, tx0 = tx * kTW;
        for (int c = 0; c < C; c++) {
            float tile[kTH][kTW + 1];
            memset(tile, 0, sizeof(tile));
            for (int r = 0; r < R; r++) {
                const Hdr &h = hdr[r];
                if ((h.key & 0xff) != ST_OK || ((h.key >> 8) & 0xff) != l || (h.key >> 16) != b) continue;
                const int x0 = h.xr & 0xffff, x1 = h.xr >> 16, y0 = h.yr & 0xffff, y1 = h.yr >> 16;
                if (x0 / kTW > tx || x1 / kTW < tx || y0 / kTH > ty || y1 / kTH < ty) continue;
                const Plan &pl = plans[r];
                if (c == 0) visits++;
                const float *d = &dout[((size_t)r * C + c) * 49];
                unsigned rowmask = 0;
                float wyt[kTH][8];
                for (int i = 0; i < pl.nrows; i++) {
                    const int yy = pl.row[i].y - ty0;
                    if (yy >= 0 && yy < kTH) { rowmask |= 1u << yy; for (int p = 0; p < 7; p++) wyt[yy][p] = pl.row[i].w[p]; }
                }
                if (c == 0 && !rowmask) emptyvis++;
                if (!pl.wide) {
                    const int ja = std::max(0, tx0 - pl.x0), jb = std::min(pl.ncols, tx0 + kTW - pl.x0), n = jb - ja;
                    if (n <= 0) { printf("n<=0?\n"); return 1; }
                    for (int i = 0; i < kTH; i++) if (rowmask >> i & 1) {
                        float V[7];
                        for (int q = 0; q < 7; q++) { V[q] = 0; for (int p = 0; p < 7; p++) V[q] += wyt[i][p] * d[p * 7 + q]; }
                        for (int k = 0; k < n; k++) { float a = tile[i][pl.x0 + ja - tx0 + k]; for (int q = 0; q < 7; q++) a += pl.xw[ja + k][q] * V[q]; tile[i][pl.x0 + ja - tx0 + k] = a; }
                    }
                } else {
                    int co[28]; float cw[28];
                    for (int k = 0; k < 28; k++) { const int col = pl.bin[k >> 2].col[k & 3]; co[k] = (col >= tx0 && col < tx0 + kTW) ? col - tx0 : kTW; cw[k] = pl.bin[k >> 2].w[k & 3]; }
                    for (int i = 0; i < kTH; i++) if (rowmask >> i & 1) {
                        float V[7];
                        for (int q = 0; q < 7; q++) { V[q] = 0; for (int p = 0; p < 7; p++) V[q] += wyt[i][p] * d[p * 7 + q]; }
                        for (int ph = 0; ph < 2; ph++) {
                            float t[16]; int n = 0;
                            for (int q = ph; q < 7; q += 2) for (int s = 0; s < 4; s++) t[n++] = tile[i][co[q * 4 + s]];
                            n = 0;
                            for (int q = ph; q < 7; q += 2) for (int s = 0; s < 4; s++) { t[n] += cw[q * 4 + s] * V[q]; n++; }
                            n = 0;
                            for (int q = ph; q < 7; q += 2) for (int s = 0; s < 4; s++) tile[i][co[q * 4 + s]] = t[n++];
                        }
                    }
                }
            }
            for (int i = 0; i < kTH && ty0 + i < H[l]; i++) for (int x = 0; x < kTW && tx0 + x < W[l]; x++)
                got[l][(((size_t)b * C + c) * H[l] + ty0 + i) * W[l] + tx0 + x] = tile[i][x];
        }
    }
    // declined RoIs: the gather kernel's job -- add them through the oracle for the comparison
    {
        std::vector<float> sub;
        for (int r = 0; r < R; r++) if (plans[r].status == ST_DECLINE) {
            float *pg[4]; for (int l = 0; l < L; l++) pg[l] = got[l].data();
            // o_roialign_bwd wants dout indexed by r: pass a 1-RoI view
            o_roialign_bwd(L, pg, H, W, scale, B, C, &rois[r * 5], 1, nullptr, 56.0f, 7, 2, 0.0f, &dout[(size_t)r * C * 49]);
        }
    }
    double maxerr = 0, maxref = 0;
    for (int l = 0; l < L; l++) for (size_t i = 0; i < ref[l].size(); i++) { maxerr = std::max(maxerr, (double)fabsf(ref[l][i] - got[l][i])); maxref = std::max(maxref, (double)fabsf(ref[l][i])); }
    printf("visits %ld (%.2f per RoI), empty-row visits %ld, max |err| %.3g of max |ref| %.3g -> %s\n", visits, (double)visits / R, emptyvis, maxerr, maxref, maxerr <= 1e-5 * maxref ? "OK" : "FAIL");
    return maxerr <= 1e-5 * maxref ? 0 : 1;
}
