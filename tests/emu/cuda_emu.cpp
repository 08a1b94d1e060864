// TEST INFRASTRUCTURE: globals of the single-threaded CUDA stand-in (see cuda_emu.h).
#include "cuda_emu.h"

emu_dim3 threadIdx{0, 0, 0}, blockIdx{0, 0, 0}, blockDim{1, 1, 1}, gridDim{1, 1, 1};
unsigned char* cc_emu_smem = nullptr;
static size_t cc_emu_smem_size = 0;

void cc_emu_ensure_smem(size_t bytes)
{
    if (bytes > cc_emu_smem_size)
    {
        free(cc_emu_smem);
        cc_emu_smem = static_cast<unsigned char*>(aligned_alloc(64, (bytes + 63) / 64 * 64));
        cc_emu_smem_size = bytes;
    }
}
