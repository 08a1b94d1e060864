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
#define CC_LAUNCH(kernel, grid, block, smem, stream, ...) kernel<<<(grid), (block), (smem), (stream)>>>(__VA_ARGS__)
#define CC_SMEM(name) extern __shared__ __align__(16) unsigned char name[]
#define CC_FULL_MASK 0xffffffffu
#endif

#endif
