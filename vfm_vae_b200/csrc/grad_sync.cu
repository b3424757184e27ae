// Gradient exchange helper: the post-all-reduce pass of the reference's sync_grads (training/training_loop.py:281-289)
//     flat = flat / world_size;  flat = flat * gain (gain != None);  nan_to_num(flat, nan=0, posinf=1e5, neginf=-1e5)
// as ONE in-place streaming pass over the persistent flat fp32 gradient buffer (the reference runs three full passes plus a
// concat before and a split/cast loop after).  HBM-bound: 8 bytes per element (read + write), 16-byte accesses, 4 vectors in
// flight per thread, grid sized in CTAs-per-SM multiples of the 148 SMs with a grid-stride loop.
#include "common.cuh"

namespace vfm {

__device__ __forceinline__ float finalize_one(float g, float inv_world, float world, int exact_div, float gain, int use_gain,
                                              float nan_v, float pos_v, float neg_v) {
    // one-box world sizes are powers of two, where the reciprocal multiply equals the reference's division bit for bit;
    // any other world size takes the true division
    g = exact_div ? g * inv_world : g / world;
    if (use_gain) g = g * gain;
    if (g != g) return nan_v;
    if (g == __int_as_float(0x7f800000)) return pos_v;
    if (g == __int_as_float(0xff800000)) return neg_v;
    return g;
}

__global__ void __launch_bounds__(256) grad_finalize_kernel(float* __restrict__ g, int64_t n, float inv_world, float world, int exact_div,
                                                            float gain, int use_gain, float nan_v, float pos_v, float neg_v) {
    const int64_t nvec = n >> 2;
    float4* g4 = reinterpret_cast<float4*>(g);
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    // 4 independent 16-byte loads in flight per thread
    for (; i + 3 * stride < nvec; i += 4 * stride) {
        float4 v[4];
#pragma unroll
        for (int k = 0; k < 4; k++) v[k] = g4[i + k * stride];
#pragma unroll
        for (int k = 0; k < 4; k++) {
            float* e = reinterpret_cast<float*>(&v[k]);
#pragma unroll
            for (int j = 0; j < 4; j++) e[j] = finalize_one(e[j], inv_world, world, exact_div, gain, use_gain, nan_v, pos_v, neg_v);
            g4[i + k * stride] = v[k];
        }
    }
    for (; i < nvec; i += stride) {
        float4 v = g4[i];
        float* e = reinterpret_cast<float*>(&v);
#pragma unroll
        for (int j = 0; j < 4; j++) e[j] = finalize_one(e[j], inv_world, world, exact_div, gain, use_gain, nan_v, pos_v, neg_v);
        g4[i] = v;
    }
    // tail (n % 4 elements)
    int64_t t = (nvec << 2) + (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t < n) g[t] = finalize_one(g[t], inv_world, world, exact_div, gain, use_gain, nan_v, pos_v, neg_v);
}

}  // namespace vfm

extern "C" int vfm_grad_finalize(const vfm_grad_finalize_params* p, void* stream_) {
    using namespace vfm;
    cudaStream_t stream = (cudaStream_t)stream_;
    VFM_CHECK_ARG(p != nullptr && p->grads != nullptr, "grad_finalize: null buffer");
    VFM_CHECK_ARG(p->numel >= 0 && p->world_size >= 1, "grad_finalize: bad numel / world_size");
    VFM_CHECK_ARG(aligned16(p->grads), "grad_finalize: the flat gradient buffer must be 16-byte aligned");
    if (p->numel == 0) return VFM_OK;
    const int w = p->world_size;
    const int exact = (w & (w - 1)) == 0;
    int64_t nvec = p->numel >> 2;
    int64_t want = ceil_div64(nvec > 0 ? nvec : 1, 256 * 4);
    int grid = (int)(want < (int64_t)kNumSMs * 8 ? (want > 0 ? want : 1) : (int64_t)kNumSMs * 8);
    KernelTimer kt("grad_finalize", stream, 0.0, 8.0 * (double)p->numel);
    grad_finalize_kernel<<<grid, 256, 0, stream>>>(p->grads, p->numel, 1.0f / (float)w, (float)w, exact, (float)p->gain, p->use_gain,
                                                   (float)p->nan, (float)p->posinf, (float)p->neginf);
    return launch_status("grad_finalize_kernel");
}
