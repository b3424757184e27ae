// Shared helpers for the libvfmops kernels (sm_100a only).
#pragma once

#include <cuda_runtime.h>
#include <cuda_fp16.h>
#include <stdint.h>
#include <stdio.h>
#include <stdarg.h>

#include "../../include/vfm_ops.h"

namespace vfm {

constexpr int kNumSMs = 148;  // B200

// ---- error plumbing (host) ----------------------------------------------------------------------------------
void set_error(const char* fmt, ...);
void count_launch(int n = 1);

// Scope guard around one kernel launch (or a short launch sequence): when vfm_timing_enable(1) is active it records a
// CUDA event before and after on the launching stream, tagged with the algorithmic FLOPs / bytes of the launch, so that
// bench.py can report achieved TFLOP/s or GB/s per kernel from inside the timed region.  Free when timing is off.
bool timing_enabled();
class KernelTimer {
public:
    // `tag_fmt` (optional, printf-style): a shape tag appended to the name when VFM_TIMING_DETAIL=1 (per-layer tables)
    KernelTimer(const char* name, cudaStream_t stream, double flops, double bytes, const char* tag_fmt = nullptr, ...);
    ~KernelTimer();
private:
    char name_[64]; cudaStream_t stream_; double flops_, bytes_; cudaEvent_t e0_; bool on_;
};

#define VFM_CHECK_ARG(cond, ...)             \
    do {                                         \
        if (!(cond)) {                           \
            vfm::set_error(__VA_ARGS__);         \
            return VFM_ERR_INVALID;              \
        }                                        \
    } while (0)

#define VFM_CUDA_OK(expr)                                                                  \
    do {                                                                                   \
        cudaError_t err__ = (expr);                                                        \
        if (err__ != cudaSuccess) {                                                        \
            vfm::set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(err__), __FILE__, __LINE__); \
            return VFM_ERR_CUDA;                                                           \
        }                                                                                  \
    } while (0)

inline int launch_status(const char* what) {
    cudaError_t err = cudaGetLastError();
    if (err != cudaSuccess) {
        set_error("launch of %s failed: %s", what, cudaGetErrorString(err));
        return VFM_ERR_CUDA;
    }
    count_launch();
    return VFM_OK;
}

// ---- dtype traits -------------------------------------------------------------------------------------------
template <class T> struct Acc { typedef float type; };
template <> struct Acc<double> { typedef double type; };

template <class T> __device__ __forceinline__ typename Acc<T>::type to_acc(T v) { return (typename Acc<T>::type)v; }
template <> __device__ __forceinline__ float to_acc<__half>(__half v) { return __half2float(v); }

template <class T, class A> __device__ __forceinline__ T from_acc(A v) { return (T)v; }
template <> __device__ __forceinline__ __half from_acc<__half, float>(float v) { return __float2half_rn(v); }

// 16-byte vector of T
template <class T> struct Vec16 {
    static constexpr int N = 16 / sizeof(T);
    union { uint4 u; T v[16 / sizeof(T)]; };
};

__device__ __forceinline__ uint4 ldg_stream(const void* p) {
    uint4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
    return r;
}
__device__ __forceinline__ void stg_stream(void* p, const uint4& v) {
    asm volatile("st.global.L1::no_allocate.v4.u32 [%0], {%1,%2,%3,%4};"
                 :: "l"(p), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}

template <class A> __device__ __forceinline__ A warp_sum(A v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// Sum over a thread block; result valid in thread 0.  `smem` must hold >= 32 values.
template <class A> __device__ __forceinline__ A block_sum(A v, A* smem) {
    v = warp_sum(v);
    int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    __syncthreads();
    if (lane == 0) smem[warp] = v;
    __syncthreads();
    int nwarps = (blockDim.x + 31) >> 5;
    v = (threadIdx.x < nwarps) ? smem[threadIdx.x] : (A)0;
    if (warp == 0) v = warp_sum(v);
    return v;
}

inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

__host__ __device__ __forceinline__ int64_t ceil_div64(int64_t a, int64_t b) { return (a + b - 1) / b; }
__host__ __device__ __forceinline__ int ceil_div(int a, int b) { return (a + b - 1) / b; }

// floor division for possibly negative numerators, positive divisor
__host__ __device__ __forceinline__ int floor_div(int a, int b) {
    int q = a / b;
    return (a % b != 0 && a < 0) ? q - 1 : q;
}

}  // namespace vfm
