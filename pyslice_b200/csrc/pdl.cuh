// Programmatic dependent launch (PDL) helpers: the fused kernels of one stream chain are launched with
// cudaLaunchAttributeProgrammaticStreamSerialization so that kernel k+1's CTAs are scheduled as SMs drain and
// run their prologue (mbarrier init, twiddles, tables) while kernel k finishes; every kernel body starts with
// pdl_wait(), which returns once the preceding grid has completed and its writes are visible (so completion
// order stays transitive along the chain), and calls pdl_trigger() first thing.
#pragma once
#include <cuda_runtime.h>

namespace psb {

__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

template <class... KArgs, class... Args>
inline cudaError_t pdl_launch(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t s, Args&&... args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid;
    cfg.blockDim = block;
    cfg.dynamicSmemBytes = smem;
    cfg.stream = s;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    return cudaLaunchKernelEx(&cfg, kernel, std::forward<Args>(args)...);
}

}  // namespace psb
