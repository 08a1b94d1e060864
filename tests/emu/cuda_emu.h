// cuda_emu.h -- TEST INFRASTRUCTURE. Minimal single-threaded stand-in for the CUDA runtime + device
// intrinsics, so that continuous_clustering_b200/csrc/*.cu can be compiled by g++ (-DCC_EMU) into
// tests/emu/libcc_b200_emu_test.so and exercised by the CPU test-suite. Warps have ONE lane and blocks have ONE
// thread; a launch runs its blocks one after another. This checks kernel LOGIC (indices, state machines,
// host orchestration); it cannot find races or memory-ordering bugs -- those are covered by the -m gpu tests.
#ifndef CUDA_EMU_H
#define CUDA_EMU_H

#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdint>
#include <cstdlib>
#include <cstring>

#define CC_WARP 1
#define CC_PDL_ENTER() do { } while (0)
#define CC_FULL_MASK 0x1u
#define __global__
#define __device__
#define __host__
#define __forceinline__ inline
#define __restrict__
#define __launch_bounds__(...)
#define __shared__ static

struct emu_dim3
{
    unsigned x, y, z;
};
struct float4
{
    float x, y, z, w;
};
struct uchar4
{
    unsigned char x, y, z, w;
};
static inline float4 make_float4(float x, float y, float z, float w)
{
    return float4{x, y, z, w};
}
static inline uchar4 make_uchar4(unsigned char x, unsigned char y, unsigned char z, unsigned char w)
{
    return uchar4{x, y, z, w};
}

extern emu_dim3 threadIdx, blockIdx, blockDim, gridDim;
extern unsigned char* cc_emu_smem;
void cc_emu_ensure_smem(size_t bytes);

#define CC_SMEM(name) unsigned char* name = cc_emu_smem
#define CC_LAUNCH(kernel, grid, block, smem, stream, ...)                                                          \
    do                                                                                                               \
    {                                                                                                                \
        cc_emu_ensure_smem((smem) + 64);                                                                             \
        gridDim = emu_dim3{static_cast<unsigned>(grid), 1, 1};                                                       \
        blockDim = emu_dim3{1, 1, 1};                                                                                \
        threadIdx = emu_dim3{0, 0, 0};                                                                               \
        for (unsigned cc_b_ = 0; cc_b_ < static_cast<unsigned>(grid); cc_b_++)                                      \
        {                                                                                                            \
            blockIdx = emu_dim3{cc_b_, 0, 0};                                                                        \
            kernel(__VA_ARGS__);                                                                                     \
        }                                                                                                            \
    } while (0)

static inline void __syncthreads() {}
static inline void __syncwarp(unsigned = 1) {}
static inline void __threadfence() {}
template<typename T>
static inline T __shfl_sync(unsigned, T v, int)
{
    return v;
}
template<typename T>
static inline T __shfl_down_sync(unsigned, T v, int)
{
    return v;
}
template<typename T>
static inline T __shfl_up_sync(unsigned, T v, int)
{
    return v;
}
template<typename T>
static inline T __shfl_xor_sync(unsigned, T v, int)
{
    return v;
}
static inline unsigned __ballot_sync(unsigned, int p)
{
    return p ? 1u : 0u;
}
template<typename T>
static inline T __reduce_min_sync(unsigned, T v)
{
    return v;
}
template<typename T>
static inline T __reduce_add_sync(unsigned, T v)
{
    return v;
}
static inline int __reduce_max_sync(unsigned, int v)
{
    return v;
}
static inline unsigned __reduce_max_sync(unsigned, unsigned v)
{
    return v;
}
static inline int __clz(int v)
{
    return v ? __builtin_clz(static_cast<unsigned>(v)) : 32;
}
static inline int __reduce_or_sync(unsigned, int v)
{
    return v;
}
static inline int __ffs(unsigned v)
{
    return __builtin_ffs(static_cast<int>(v));
}
static inline int __popc(unsigned v)
{
    return __builtin_popcount(v);
}
static inline void __pipeline_memcpy_async(void* dst, const void* src, size_t n)
{
    memcpy(dst, src, n);
}
static inline void __pipeline_commit() {}
static inline void __pipeline_wait_prior(int) {}
template<typename T>
static inline T __ldg(const T* p)
{
    return *p;
}

template<typename T>
static inline T atomicAdd(T* a, T v)
{
    T o = *a;
    *a = o + v;
    return o;
}
template<typename T>
static inline T atomicMax(T* a, T v)
{
    T o = *a;
    if (v > o)
        *a = v;
    return o;
}
template<typename T>
static inline T atomicMin(T* a, T v)
{
    T o = *a;
    if (v < o)
        *a = v;
    return o;
}
template<typename T>
static inline T atomicCAS(T* a, T cmp, T v)
{
    T o = *a;
    if (o == cmp)
        *a = v;
    return o;
}
template<typename T>
static inline T atomicExch(T* a, T v)
{
    T o = *a;
    *a = v;
    return o;
}
template<typename T>
static inline T atomicOr(T* a, T v)
{
    T o = *a;
    *a = o | v;
    return o;
}

static inline double __longlong_as_double(long long v)
{
    double d;
    memcpy(&d, &v, 8);
    return d;
}
static inline long long __double_as_longlong(double d)
{
    long long v;
    memcpy(&v, &d, 8);
    return v;
}

// ---- runtime ----
typedef int cudaError_t;
typedef void* cudaStream_t;
struct emu_event
{
    std::chrono::steady_clock::time_point t;
};
typedef emu_event* cudaEvent_t;
enum
{
    cudaSuccess = 0,
    cudaMemcpyHostToDevice = 1,
    cudaMemcpyDeviceToHost = 2,
    cudaMemcpyDeviceToDevice = 3,
    cudaStreamNonBlocking = 1,
    cudaFuncAttributeMaxDynamicSharedMemorySize = 8
};
static inline const char* cudaGetErrorString(cudaError_t)
{
    return "emu";
}
static inline cudaError_t cudaGetLastError()
{
    return cudaSuccess;
}
static inline cudaError_t cudaSetDevice(int)
{
    return cudaSuccess;
}
static inline cudaError_t cudaGetDeviceCount(int* n)
{
    *n = 1;
    return cudaSuccess;
}
static inline cudaError_t cudaMalloc(void** p, size_t n)
{
    *p = malloc(n ? n : 1);
    return *p ? cudaSuccess : 2;
}
static inline cudaError_t cudaMallocHost(void** p, size_t n)
{
    return cudaMalloc(p, n);
}
static inline cudaError_t cudaFree(void* p)
{
    free(p);
    return cudaSuccess;
}
static inline cudaError_t cudaFreeHost(void* p)
{
    free(p);
    return cudaSuccess;
}
static inline cudaError_t cudaMemcpyAsync(void* d, const void* s, size_t n, int, cudaStream_t = nullptr)
{
    memcpy(d, s, n);
    return cudaSuccess;
}
static inline cudaError_t cudaMemcpy(void* d, const void* s, size_t n, int)
{
    memcpy(d, s, n);
    return cudaSuccess;
}
static inline cudaError_t cudaMemsetAsync(void* d, int v, size_t n, cudaStream_t = nullptr)
{
    memset(d, v, n);
    return cudaSuccess;
}
static inline cudaError_t cudaMemset(void* d, int v, size_t n)
{
    memset(d, v, n);
    return cudaSuccess;
}
static inline cudaError_t cudaStreamCreateWithFlags(cudaStream_t* s, int)
{
    *s = nullptr;
    return cudaSuccess;
}
static inline cudaError_t cudaStreamDestroy(cudaStream_t)
{
    return cudaSuccess;
}
static inline cudaError_t cudaStreamSynchronize(cudaStream_t)
{
    return cudaSuccess;
}
static inline cudaError_t cudaDeviceSynchronize()
{
    return cudaSuccess;
}
static inline cudaError_t cudaEventCreate(cudaEvent_t* e)
{
    *e = new emu_event();
    return cudaSuccess;
}
static inline cudaError_t cudaEventDestroy(cudaEvent_t e)
{
    delete e;
    return cudaSuccess;
}
static inline cudaError_t cudaEventRecord(cudaEvent_t e, cudaStream_t = nullptr)
{
    e->t = std::chrono::steady_clock::now();
    return cudaSuccess;
}
static inline cudaError_t cudaStreamWaitEvent(cudaStream_t, cudaEvent_t, unsigned)
{
    return cudaSuccess;
}
static inline cudaError_t cudaEventQuery(cudaEvent_t)
{
    return cudaSuccess;
}
static inline cudaError_t cudaEventSynchronize(cudaEvent_t)
{
    return cudaSuccess;
}
static inline cudaError_t cudaEventElapsedTime(float* ms, cudaEvent_t a, cudaEvent_t b)
{
    *ms = std::chrono::duration<float, std::milli>(b->t - a->t).count();
    return cudaSuccess;
}
template<typename F>
static inline cudaError_t cudaFuncSetAttribute(F, int, int)
{
    return cudaSuccess;
}

#endif
