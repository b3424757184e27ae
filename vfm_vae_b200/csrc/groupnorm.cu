// Group-norm statistics as a per-(sample, channel) affine map, for the fused residual layers of the legacy decoder.
//
// The reference normalises the input of every residual SynthesisLayer with GroupNorm32 (networks/generator.py:261-263,
// networks/utils/shared.py: nn.GroupNorm evaluated in fp32) before the modulated conv, and adds the *normalised* tensor
// back after it.  GroupNorm is x -> x * A[n,c] + B[n,c] with A = rstd[n,g] * gamma[c], B = beta[c] - mean[n,g] * A, so the
// only pass over the activations that cannot be folded into the conv's own kernels is the statistics.  This file is
// that pass: one read of x (HBM-bound), fp32 accumulation around a pivot (the group's first element) so that
// E[(x-K)^2] - E[x-K]^2 does not cancel, block reduction in double precision.  The modulated conv then applies the map
// in its operand pre-pass and, for the residual, in its epilogue (vfm_modconv_fwd_params::x_scale / x_shift).
#include "common.cuh"

namespace vfm {
namespace {

template <class T>
__global__ void __launch_bounds__(512) gn_affine_kernel(const T* __restrict__ x, const float* __restrict__ gamma, const float* __restrict__ beta,
                                                        float* __restrict__ scale, float* __restrict__ shift, int C, int64_t HW, int groups, float eps) {
    constexpr int V = 16 / (int)sizeof(T);
    const int n = blockIdx.x / groups, g = blockIdx.x - n * groups;
    const int cpg = C / groups;
    const int64_t count = (int64_t)cpg * HW;
    const T* base = x + ((int64_t)n * C + (int64_t)g * cpg) * HW;        // the group's channels are contiguous in NCHW
    const float K = to_acc(base[0]);
    float s1 = 0.f, s2 = 0.f;
    if ((reinterpret_cast<uintptr_t>(base) & 15u) == 0 && count % V == 0) {
        const int64_t nvec = count / V;
        const uint4* bv = (const uint4*)base;
        for (int64_t i = threadIdx.x; i < nvec; i += (int64_t)blockDim.x * 4) {
            uint4 u[4];
#pragma unroll
            for (int k = 0; k < 4; k++) { const int64_t j = i + (int64_t)k * blockDim.x; u[k] = (j < nvec) ? ldg_stream(bv + j) : make_uint4(0, 0, 0, 0); }
#pragma unroll
            for (int k = 0; k < 4; k++) {
                if (i + (int64_t)k * blockDim.x >= nvec) continue;
                const T* e = (const T*)&u[k];
#pragma unroll
                for (int q = 0; q < V; q++) { const float d = to_acc(e[q]) - K; s1 += d; s2 = fmaf(d, d, s2); }
            }
        }
    } else {
        for (int64_t i = threadIdx.x; i < count; i += blockDim.x) { const float d = to_acc(base[i]) - K; s1 += d; s2 = fmaf(d, d, s2); }
    }
    __shared__ double red[2][32];
    double d1 = (double)s1, d2 = (double)s2;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) { d1 += __shfl_xor_sync(0xffffffffu, d1, o); d2 += __shfl_xor_sync(0xffffffffu, d2, o); }
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (lane == 0) { red[0][warp] = d1; red[1][warp] = d2; }
    __syncthreads();
    __shared__ float s_mean, s_rstd;
    if (threadIdx.x == 0) {
        double a = 0.0, b = 0.0;
        for (int w = 0; w < (int)(blockDim.x >> 5); w++) { a += red[0][w]; b += red[1][w]; }
        const double m = a / (double)count;
        double var = b / (double)count - m * m;                            // biased variance, as nn.GroupNorm
        if (var < 0.0) var = 0.0;
        s_mean = (float)((double)K + m);
        s_rstd = (float)(1.0 / sqrt(var + (double)eps));
    }
    __syncthreads();
    for (int c = threadIdx.x; c < cpg; c += blockDim.x) {
        const int ch = g * cpg + c;
        const float a = s_rstd * (gamma ? gamma[ch] : 1.f);
        scale[(int64_t)n * C + ch] = a;
        shift[(int64_t)n * C + ch] = (beta ? beta[ch] : 0.f) - s_mean * a;
    }
}

}  // namespace
}  // namespace vfm

extern "C" int vfm_group_norm_affine(const vfm_group_norm_affine_params* p, void* stream_) {
    using namespace vfm;
    cudaStream_t stream = (cudaStream_t)stream_;
    VFM_CHECK_ARG(p != nullptr, "group_norm_affine: params is NULL");
    VFM_CHECK_ARG(p->x && p->scale && p->shift, "group_norm_affine: x, scale and shift must be non-NULL");
    VFM_CHECK_ARG(p->batch >= 1 && p->channels >= 1 && p->hw >= 1, "group_norm_affine: x is empty");
    VFM_CHECK_ARG(p->groups >= 1 && p->channels % p->groups == 0, "group_norm_affine: channels (%d) must be divisible by groups (%d)", p->channels, p->groups);
    VFM_CHECK_ARG(p->dtype == VFM_F16 || p->dtype == VFM_F32, "group_norm_affine: unsupported dtype %d", p->dtype);
    const int64_t blocks = (int64_t)p->batch * p->groups;
    VFM_CHECK_ARG(blocks <= 0x7fffffffLL, "group_norm_affine: grid too large");
    const double es = p->dtype == VFM_F16 ? 2.0 : 4.0;
    KernelTimer timer("group_norm_affine", stream, 0.0, (double)p->batch * p->channels * (double)p->hw * es + 8.0 * p->batch * p->channels, "c%dhw%lld",
                      p->channels, (long long)p->hw);
    if (p->dtype == VFM_F16)
        gn_affine_kernel<__half><<<(unsigned)blocks, 512, 0, stream>>>((const __half*)p->x, p->gamma, p->beta, p->scale, p->shift, p->channels, p->hw, p->groups, (float)p->eps);
    else
        gn_affine_kernel<float><<<(unsigned)blocks, 512, 0, stream>>>((const float*)p->x, p->gamma, p->beta, p->scale, p->shift, p->channels, p->hw, p->groups, (float)p->eps);
    return launch_status("gn_affine_kernel");
}
