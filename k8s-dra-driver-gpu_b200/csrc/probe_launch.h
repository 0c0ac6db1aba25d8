// probe_launch.h — host-callable launchers of the kernels in probe_kernels.cu.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "probe_types.h"

namespace cdp {

// Sets the dynamic shared memory attribute on the current device and reports
// how many CTAs of the persistent kernel fit one SM. Returns a cudaError_t.
int probe_kernel_prepare(int* max_ctas_per_sm);
// Launches the persistent probe kernel on `stream` of the current device.
int probe_kernel_launch(const ProbeParams* p, unsigned grid, bool cooperative, cudaStream_t stream);
// Fills `bytes` of source pattern for `rank` (bytes % 16 == 0).
int probe_fill_launch(void* dst, uint64_t bytes, uint64_t seed, uint32_t rank, unsigned grid, cudaStream_t stream);

}  // namespace cdp
