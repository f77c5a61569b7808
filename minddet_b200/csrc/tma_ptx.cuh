// tma_ptx.cuh -- PTX wrappers shared by the RoIAlign kernels: mbarriers, TMA tile loads / reduce-adds, bulk copies,
// cp.async.  sm_100a only.
#pragma once
#include <cuda.h>

#include "common.cuh"

namespace md {

// ---- PTX wrappers -----------------------------------------------------------------------------------
MD_DEVINL uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
MD_DEVINL void mbar_init(unsigned long long *bar, int count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(smem_u32(bar)), "r"(count) : "memory");
}
MD_DEVINL void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
MD_DEVINL void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
MD_DEVINL void mbar_expect_tx(unsigned long long *bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(smem_u32(bar)), "r"(bytes) : "memory");
}
MD_DEVINL void mbar_wait(unsigned long long *bar, uint32_t parity)
{
    uint32_t done = 0;
    const uint32_t a = smem_u32(bar);
    while (!done) {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(done) : "r"(a), "r"(parity) : "memory");
    }
}
MD_DEVINL void tma_load_3d(void *dst, const CUtensorMap *map, int x, int y, int z, unsigned long long *bar)
{
    asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
                 :: "r"(smem_u32(dst)), "l"(map), "r"(x), "r"(y), "r"(z), "r"(smem_u32(bar)) : "memory");
}
MD_DEVINL void tma_reduce_add_3d(const CUtensorMap *map, int x, int y, int z, const void *src)
{
    asm volatile("cp.reduce.async.bulk.tensor.3d.global.shared::cta.add.tile.bulk_group [%0, {%2, %3, %4}], [%1];"
                 :: "l"(map), "r"(smem_u32(src)), "r"(x), "r"(y), "r"(z) : "memory");
}
MD_DEVINL void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N> MD_DEVINL void bulk_wait_read() { asm volatile("cp.async.bulk.wait_group.read %0;" :: "n"(N) : "memory"); }
template <int N> MD_DEVINL void bulk_wait_all() { asm volatile("cp.async.bulk.wait_group %0;" :: "n"(N) : "memory"); }
MD_DEVINL void cp_async4(void *smem, const void *gmem)
{
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" :: "r"(smem_u32(smem)), "l"(gmem) : "memory");
}
MD_DEVINL void cp_async_mbar_arrive(unsigned long long *bar)
{
    asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" :: "r"(smem_u32(bar)) : "memory");
}
MD_DEVINL void cp_async_commit_group() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N> MD_DEVINL void cp_async_wait_group() { asm volatile("cp.async.wait_group %0;" :: "n"(N) : "memory"); }


// 1-D bulk copy shared -> global (contiguous, 16-byte aligned, size a multiple of 16); joins the thread's bulk group
MD_DEVINL void bulk_store_1d(void *gdst, const void *ssrc, uint32_t bytes)
{
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" :: "l"(gdst), "r"(smem_u32(ssrc)), "r"(bytes) : "memory");
}
// 1-D bulk copy global -> shared, completion on an mbarrier
MD_DEVINL void bulk_load_1d(void *sdst, const void *gsrc, uint32_t bytes, unsigned long long *bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 :: "r"(smem_u32(sdst)), "l"(gsrc), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}

}  // namespace md
