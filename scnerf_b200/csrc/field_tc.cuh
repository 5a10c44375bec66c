// Fused tcgen05 field forward (placeholder until the kernel lands in this file).
#pragma once
#include "common.cuh"
#include "field_simt.cuh"
namespace scnerf {
inline int field_tc_fwd(const scnerf_mlp&, int, const float*, int, const float*, const float*,
                        const float*, int64_t, int, const FieldBufs&, float*, void*) {
  return fail(SCNERF_ERR_UNSUPPORTED, "tensor-core field path not built into this library yet");
}
}  // namespace scnerf
