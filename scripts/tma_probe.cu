// Probe: which cp.async.bulk(.tensor) forms run on this B200 box?  One mode per process (a fault kills the context).
//   nvcc -gencode arch=compute_100a,code=sm_100a -o tma_probe tma_probe.cu
//   for m in 0 1 2 3 4 5 6 7; do ./tma_probe $m; done
// mode 0: 1-D cp.async.bulk global->shared (no tensor map)
// mode 1: 2-D map  f32 box {32,8}            __grid_constant__
// mode 2: 3-D map  f32 box {8,8,4}           __grid_constant__
// mode 3: 3-D map  f32 box {32,8,4}          __grid_constant__
// mode 4: 3-D map  f32 box {8,8,4}, L2 promotion NONE
// mode 5: 3-D map  f32 box {8,8,4}, map in global memory
// mode 6: 3-D map  f32 box {8,8,4}, cuda::ptx-free variant without ".tile" qualifier / shared::cta barrier
// mode 7: 3-D reduce-add store of a box {8,8,4}
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <cstring>
struct Maps { CUtensorMap m[4]; };
__device__ __forceinline__ uint32_t s32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void bar_init(unsigned long long* bar) {
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" :: "r"(s32(bar)) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
}
__device__ __forceinline__ void bar_wait(unsigned long long* bar) {
  uint32_t done = 0;
  while (!done) asm volatile("{.reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0,1,0,p;}" : "=r"(done) : "r"(s32(bar)), "r"(0) : "memory");
}

__global__ void k_bulk1d(const float* src, float* out) {
  __shared__ __align__(128) float tile[256];
  __shared__ unsigned long long bar;
  bar_init(&bar);
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(s32(&bar)), "r"(1024) : "memory");
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 :: "r"(s32(tile)), "l"(src), "r"(1024), "r"(s32(&bar)) : "memory");
  }
  bar_wait(&bar);
  out[threadIdx.x] = tile[threadIdx.x];
}

template <int RANK, bool TILEQ>
__device__ void run(const CUtensorMap* map, float* out, int x, int y, int z, int bytes) {
  __shared__ __align__(128) float tile[32 * 8 * 4];
  __shared__ unsigned long long bar;
  bar_init(&bar);
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(s32(&bar)), "r"(bytes) : "memory");
    if (RANK == 2)
      asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
                   :: "r"(s32(tile)), "l"(map), "r"(x), "r"(y), "r"(s32(&bar)) : "memory");
    else if (TILEQ)
      asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
                   :: "r"(s32(tile)), "l"(map), "r"(x), "r"(y), "r"(z), "r"(s32(&bar)) : "memory");
    else
      asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
                   :: "r"(s32(tile)), "l"(map), "r"(x), "r"(y), "r"(z), "r"(s32(&bar)) : "memory");
  }
  bar_wait(&bar);
  out[threadIdx.x] = tile[threadIdx.x];
}
__global__ void k_param2(const __grid_constant__ Maps maps, int idx, float* out, int bytes) { run<2, false>(&maps.m[idx], out, 4, 2, 0, bytes); }
__global__ void k_param3(const __grid_constant__ Maps maps, int idx, float* out, int bytes, int x) { run<3, true>(&maps.m[idx], out, x, 2, 1, bytes); }
__global__ void k_param3n(const __grid_constant__ Maps maps, int idx, float* out, int bytes) { run<3, false>(&maps.m[idx], out, 4, 2, 1, bytes); }
__global__ void k_global3(const CUtensorMap* maps, int idx, float* out, int bytes) { run<3, true>(&maps[idx], out, 4, 2, 1, bytes); }
__global__ void k_reduce(const __grid_constant__ Maps maps, int idx) {
  __shared__ __align__(128) float tile[8 * 8 * 4];
  tile[threadIdx.x] = 1.0f;
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  __syncthreads();
  if (threadIdx.x == 0) {
    asm volatile("cp.reduce.async.bulk.tensor.3d.global.shared::cta.add.tile.bulk_group [%0, {%2, %3, %4}], [%1];"
                 :: "l"(&maps.m[idx]), "r"(s32(tile)), "r"(4), "r"(2), "r"(1) : "memory");
    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
    asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
  }
  __syncthreads();
}

typedef CUresult (*Enc)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int main(int argc, char** argv) {
  int mode = atoi(argv[1]); int X = argc > 2 ? atoi(argv[2]) : 4;
  const int W = 336, H = 200, C = 16;
  float* d; cudaMalloc(&d, sizeof(float) * W * H * C);
  float* h = (float*)malloc(sizeof(float) * W * H * C);
  for (int i = 0; i < W * H * C; i++) h[i] = (float)(i % 1000);
  cudaMemcpy(d, h, sizeof(float) * W * H * C, cudaMemcpyHostToDevice);
  void* p = nullptr; cudaDriverEntryPointQueryResult q;
  cudaError_t ge = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q);
  if (ge != cudaSuccess || q != cudaDriverEntryPointSuccess || !p) { printf("mode %d: no encode entry point (%d, %d)\n", mode, (int)ge, (int)q); return 3; }
  Enc enc = (Enc)p;
  Maps maps; memset(&maps, 0, sizeof(maps));
  auto mk = [&](int i, int rank, int bx, int by, int bz, CUtensorMapL2promotion l2) {
    cuuint64_t dims[3] = {W, H, C}; cuuint64_t str[2] = {W * 4, (cuuint64_t)W * H * 4};
    cuuint32_t box[3] = {(cuuint32_t)bx, (cuuint32_t)by, (cuuint32_t)bz}, es[3] = {1, 1, 1};
    if (rank == 2) dims[1] = (cuuint64_t)H * C;
    CUresult rc = enc(&maps.m[i], CU_TENSOR_MAP_DATA_TYPE_FLOAT32, rank, d, dims, str, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, l2, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (rc) { printf("encode %d failed %d\n", i, rc); exit(2); }
  };
  mk(0, 2, 32, 8, 1, CU_TENSOR_MAP_L2_PROMOTION_L2_128B);
  mk(1, 3, 8, 8, 4, CU_TENSOR_MAP_L2_PROMOTION_L2_128B);
  mk(2, 3, 32, 8, 4, CU_TENSOR_MAP_L2_PROMOTION_L2_128B);
  mk(3, 3, 8, 8, 4, CU_TENSOR_MAP_L2_PROMOTION_NONE);
  float* out; cudaMalloc(&out, 256 * 4); cudaMemset(out, 0, 256 * 4);
  float expect = 0;
  if (mode == 0) { k_bulk1d<<<1, 256>>>(d + 64, out); expect = h[64]; }
  if (mode == 1) { k_param2<<<1, 256>>>(maps, 0, out, 32 * 8 * 4); expect = h[2 * W + 4]; }
  if (mode == 2) { k_param3<<<1, 256>>>(maps, 1, out, 8 * 8 * 4 * 4, X); expect = h[(1 * H + 2) * W + X]; }
  if (mode == 3) { k_param3<<<1, 256>>>(maps, 2, out, 32 * 8 * 4 * 4, X); expect = h[(1 * H + 2) * W + X]; }
  if (mode == 4) { k_param3<<<1, 256>>>(maps, 3, out, 8 * 8 * 4 * 4, X); expect = h[(1 * H + 2) * W + X]; }
  if (mode == 5) { CUtensorMap* g; cudaMalloc(&g, sizeof(maps)); cudaMemcpy(g, &maps, sizeof(maps), cudaMemcpyHostToDevice); k_global3<<<1, 256>>>(g, 1, out, 8 * 8 * 4 * 4); expect = h[(1 * H + 2) * W + 4]; }
  if (mode == 6) { k_param3n<<<1, 256>>>(maps, 1, out, 8 * 8 * 4 * 4); expect = h[(1 * H + 2) * W + 4]; }
  if (mode == 7) { k_reduce<<<1, 256>>>(maps, 1); }
  cudaError_t le = cudaGetLastError();
  cudaError_t e = cudaDeviceSynchronize();
  float ho[256]; memset(ho, 0, sizeof(ho));
  if (mode == 7) { cudaMemcpy(ho, d + (1 * H + 2) * W + 4, 4, cudaMemcpyDeviceToHost); expect = h[(1 * H + 2) * W + 4] + 1.0f; }
  else cudaMemcpy(ho, out, sizeof(ho), cudaMemcpyDeviceToHost);
  printf("mode %d: launch=%s sync=%s  got %.0f expect %.0f\n", mode, cudaGetErrorString(le), cudaGetErrorString(e), ho[0], expect);
  return e != cudaSuccess;
}
