// launch.cuh -- programmatic dependent launch (PDL) for the chains of short kernels of the region path.
//
// A step is ~25 kernels, most of them 5-40 us, each depending on its predecessor in the stream: between two of them the
// GPU pays a grid drain + launch latency of 3-5 us (chain alone: 837 us of kernels in a 922 us replay).  With the
// programmatic-stream-serialization attribute a kernel may be scheduled while its predecessor's last CTAs are still
// draining; it executes pdl_entry() first: `griddepcontrol.wait` blocks until the predecessor grid has completed and its
// writes are visible, so nothing a predecessor wrote is read early and nothing it reads is overwritten early.  Launched
// without the attribute, or behind a memset / event wait, pdl_entry() is a no-op.  Stream capture records the edge as a
// programmatic dependency.  MD_PDL=0 switches the attribute off.
// Measured (config 2): MdProposal alone 148.6 -> 141.6 us, the step 0.986 -> 0.981 ms.  An EARLY trigger
// (`griddepcontrol.launch_dependents` at the top of every kernel) was worse, 175.7 us / 1.05 ms: the waiting CTAs of the
// successors take the shared memory and thread slots the other Proposal lane's 8-CTA clusters need; early triggers in the
// target-assignment kernels alone: 0.990 vs 0.984 ms.
#pragma once
#include <cuda_runtime.h>

#include <cstdlib>
#include <utility>

#include "common.cuh"

namespace md {

inline bool pdl_enabled()
{
    static const bool on = [] { const char *e = getenv("MD_PDL"); return !e || atoi(e) != 0; }();
    return on;
}

template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t s, Args &&...args)
{
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = s;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = at;
    cfg.numAttrs = pdl_enabled() ? 1 : 0;
    return cudaLaunchKernelEx(&cfg, kern, std::forward<Args>(args)...);
}

// first statement of every kernel launched through launch_pdl
MD_DEVINL void pdl_entry()
{
    asm volatile("griddepcontrol.wait;" ::: "memory");
}

}  // namespace md
