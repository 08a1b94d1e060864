// cc_platform.h -- CUDA build glue.
//
// The product is compiled by nvcc for sm_100a and runs only on the GPU. The same kernel sources can ALSO be
// compiled by g++ with -DCC_EMU against tests/emu/cuda_emu.h, which executes every "kernel" as a plain loop
// with one-lane warps and one-thread blocks: that build exists purely so that the host logic and the kernel
// logic can be unit-tested in the GPU-less CI container (pytest -m "not gpu"). It is test infrastructure
// (lives under tests/emu, is never loaded by the package, has a different file name) and not a fallback:
// continuous_clustering_b200/_lib.py refuses to run without the nvcc-built library and a CUDA device.
#ifndef CC_PLATFORM_H
#define CC_PLATFORM_H

#ifdef CC_EMU
#include "cuda_emu.h"
#else
#include <cuda_runtime.h>
#define CC_WARP 32
// Every kernel is launched with programmatic stream serialization (programmatic dependent launch): a kernel's blocks may
// be scheduled while the kernel before it in the stream is still draining; each kernel starts with CC_PDL_ENTER(),
// which lets its own successor be scheduled early and then waits until the predecessor has completed and its memory
// operations are visible. Launch latency and block scheduling of the 14 dependent kernels of a push thereby overlap
// the tail of the kernel before.
template<typename... KArgs, typename... Args>
static inline cudaError_t cc_launch(void (*kernel)(KArgs...), int grid, int block, size_t smem, cudaStream_t stream, Args&&... args)
{
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(static_cast<unsigned>(grid), 1, 1);
    cfg.blockDim = dim3(static_cast<unsigned>(block), 1, 1);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}
#define CC_LAUNCH(kernel, grid, block, smem, stream, ...) cc_launch(kernel, (grid), (block), (smem), (stream), __VA_ARGS__)
#ifdef CC_NO_PDL /* profiling aid: kernels without the two griddepcontrol instructions */
#define CC_PDL_ENTER()                                                                                                 \
    do                                                                                                                 \
    {                                                                                                                  \
    } while (0)
#else
#define CC_PDL_ENTER()                                                                                                 \
    do                                                                                                                 \
    {                                                                                                                  \
        asm volatile("griddepcontrol.launch_dependents;" ::: "memory");                                                \
        asm volatile("griddepcontrol.wait;" ::: "memory");                                                             \
    } while (0)
#endif
#define CC_SMEM(name) extern __shared__ __align__(16) unsigned char name[]
#define CC_FULL_MASK 0xffffffffu
#endif

#endif
