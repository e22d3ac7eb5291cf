// fixed_24x24.cu -- the generic step kernel (k_step: scripted policies, pathfinding, external actions, ...) compiled once more
// for ONE layout: 24x24 maps, 128 unit slots, scripted-policy unit words and the A*/BFS scratch in shared memory -- the
// configuration of BASELINE configs[2] (maps/24x24/basesWorkers24x24*.xml with WorkerRush / LightRush).  With MRTS_TU_FIXED the
// layout fields of `Game` are static constants (engine.cuh), so the out-of-line device functions of this copy address shared
// memory with immediates instead of reloading offsets from the Game object.  microrts_cuda.cu launches this kernel instead of
// k_step whenever a batch has exactly this layout (and is not a MRTS_FLAG_PO_POLICIES batch).
#define MRTS_TU_FIXED 1
#define MRTS_TU_W 24
#define MRTS_TU_H 24
#define MRTS_TU_CAP 128
#include <cuda_runtime.h>

#include "engine.cuh"

__global__ void __launch_bounds__(MRTS_WARPS_PER_CTA * 32, 3) k_step_fixed_24x24(StepParams p) {
    step_kernel_body<KERNEL_GENERIC>(p, mrts_smem, threadIdx.x, blockDim.x, blockIdx.x, gridDim.x);
}

// the variant's description for microrts_cuda.cu: returns the kernel; W, H, cap as compiled
extern "C" __attribute__((visibility("hidden"))) const void *mrts_fixed_generic_24x24(int *W, int *H, int *cap) {
    *W = MRTS_TU_W; *H = MRTS_TU_H; *cap = MRTS_TU_CAP;
    return (const void *)k_step_fixed_24x24;
}
