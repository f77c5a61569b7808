// tma_rate.cu -- how fast does one B200 move small RoI-like footprints with cp.async.bulk.tensor, as a function of the
// box shape?  Tensor = (B*C = 2048 planes, 200, 336) fp32 (550 MB, level 0 of config 2); every warp (1-warp CTAs, 4 per SM)
// loads "footprints" of 12 rows x 16 columns x 32 channels at random 4-aligned positions into a 3-slot ring, waits, moves on.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/tma_rate scripts/tma_rate.cu && /tmp/tma_rate
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>

struct Maps { CUtensorMap m[8]; };
__device__ __forceinline__ uint32_t s32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_wait(unsigned long long *bar, uint32_t par)
{
    uint32_t done = 0;
    while (!done)
        asm volatile("{.reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0,1,0,p;}" : "=r"(done) : "r"(s32(bar)), "r"(par) : "memory");
}
__device__ __forceinline__ void tma3(void *dst, const CUtensorMap *map, int x, int y, int z, unsigned long long *bar)
{
    asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
                 :: "r"(s32(dst)), "l"(map), "r"(x), "r"(y), "r"(z), "r"(s32(bar)) : "memory");
}

// box = {16, BR rows, BC channels}; a footprint = 12 rows x 32 channels = (12/BR) x (32/BC) ops.  lanes issue ops in parallel when PAR.
__global__ void __launch_bounds__(32, 4) rate_kernel(const __grid_constant__ Maps maps, int mi, int BR, int BC, int par, int iters, int nplanes, float *sink)
{
    extern __shared__ __align__(1024) unsigned char sm[];
    __shared__ unsigned long long bar[3];
    const int lane = threadIdx.x;
    if (lane == 0) {
        for (int i = 0; i < 3; i++) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" :: "r"(s32(&bar[i])) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncwarp();
    const int nr = 12 / BR, nc = 32 / BC, nops = nr * nc;
    const uint32_t bytes = 12 * 32 * 64;
    uint32_t rng = blockIdx.x * 2654435761u + 12345u;
    auto issue = [&](int n) {
        rng = rng * 1664525u + 1013904223u;
        const int x = ((rng >> 8) % 80) * 4, y = (rng >> 16) % 188, z = ((rng >> 3) % (nplanes / 32)) * 32;
        const int slot = n % 3;
        if (lane == 0) asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(s32(&bar[slot])), "r"(bytes) : "memory");
        __syncwarp();
        unsigned char *dst = sm + slot * bytes;
        if (par) {
            if (lane < nops) { const int ir = lane % nr, ic = lane / nr; tma3(dst + (ic * nr + ir) * (BR * BC * 64), &maps.m[mi], x, y + ir * BR, z + ic * BC, &bar[slot]); }
        } else if (lane == 0) {
            for (int o = 0; o < nops; o++) { const int ir = o % nr, ic = o / nr; tma3(dst + o * (BR * BC * 64), &maps.m[mi], x, y + ir * BR, z + ic * BC, &bar[slot]); }
        }
    };
    float acc = 0.0f;
    issue(0); issue(1);
    for (int n = 0; n < iters; n++) {
        if (n + 2 < iters) issue(n + 2);
        mbar_wait(&bar[n % 3], (n / 3) & 1);
        acc += reinterpret_cast<float *>(sm + (n % 3) * bytes)[lane * 16];
        __syncwarp();
    }
    if (acc == 123.456f) sink[0] = acc;
}

typedef CUresult (*Enc)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int main()
{
    const int W = 336, H = 200, NP = 2048;
    float *d; cudaMalloc(&d, sizeof(float) * (size_t)W * H * NP);
    cudaMemset(d, 0, sizeof(float) * (size_t)W * H * NP);
    void *p = nullptr; cudaDriverEntryPointQueryResult q;
    cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q);
    Enc enc = (Enc)p;
    Maps maps; memset(&maps, 0, sizeof(maps));
    const int shapes[][2] = { {1, 32}, {4, 8}, {12, 32}, {3, 32}, {2, 16}, {4, 32}, {1, 8}, {12, 8} };
    for (int i = 0; i < 8; i++) {
        cuuint64_t dims[3] = { W, H, NP }; cuuint64_t str[2] = { W * 4, (cuuint64_t)W * H * 4 };
        cuuint32_t box[3] = { 16, (cuuint32_t)shapes[i][0], (cuuint32_t)shapes[i][1] }, es[3] = { 1, 1, 1 };
        CUresult rc = enc(&maps.m[i], CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, d, dims, str, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (rc) { printf("encode %d failed %d\n", i, rc); return 2; }
    }
    float *sink; cudaMalloc(&sink, 4);
    const size_t smem = 3 * 12 * 32 * 64;
    cudaFuncSetAttribute(rate_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    const int iters = 200, grid = 148 * 4;
    for (int np_i = 0; np_i < 2; np_i++) {
        const int nplanes = np_i ? 64 : NP;       // 64 planes = 17 MB: L2-resident
        for (int i = 0; i < 8; i++)
            for (int par = 0; par < 2; par++) {
                const int nops = (12 / shapes[i][0]) * (32 / shapes[i][1]);
                if (par && nops > 32) continue;
                cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
                rate_kernel<<<grid, 32, smem>>>(maps, i, shapes[i][0], shapes[i][1], par, 20, nplanes, sink);
                cudaEventRecord(e0);
                rate_kernel<<<grid, 32, smem>>>(maps, i, shapes[i][0], shapes[i][1], par, iters, nplanes, sink);
                cudaEventRecord(e1);
                cudaError_t e = cudaDeviceSynchronize();
                float ms = 0; cudaEventElapsedTime(&ms, e0, e1);
                const double fp = (double)grid * iters;
                printf("planes %4d box {16,%2d,%2d} ops/footprint %3d %s: %8.1f us  %6.2f Mfootprints/s  %7.1f GB/s  %6.2f Grows/s (%s)\n", nplanes, shapes[i][0], shapes[i][1], nops,
                       par ? "lanes " : "lane 0", ms * 1e3, fp / ms / 1e3, fp * 12 * 32 * 64 / ms / 1e6, fp * 12 * 32 / ms / 1e6, cudaGetErrorString(e));
            }
    }
    return 0;
}
