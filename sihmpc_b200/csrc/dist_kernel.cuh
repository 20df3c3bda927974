// dist_kernel.cuh -- register-tiled fast path for small controllers (placeholder until the
// specialised kernel lands: every shape currently takes the general warp-per-scenario kernel).
#pragma once
#include <cuda_runtime.h>
#include "devplan.h"

namespace sih {
inline int dist_configure(const DevPlan &, int *fast_path, int *blocks_per_sm) {
    *fast_path = 0; *blocks_per_sm = 0; return 0;
}
inline bool dist_supported(const DevPlan &, const sihmpc_io &) { return false; }
inline cudaError_t dist_launch(const DevPlan &, const sihmpc_io &, long long, int, int, int, cudaStream_t) {
    return cudaErrorNotSupported;
}
}  // namespace sih
