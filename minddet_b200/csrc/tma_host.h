// tma_host.h -- host-side access to cuTensorMapEncodeTiled without linking libcuda (driver entry point lookup).
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>

namespace md {

typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                  const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiledFn get_encode();   // nullptr when the driver does not export it (roialign_tma.cu)

}  // namespace md
