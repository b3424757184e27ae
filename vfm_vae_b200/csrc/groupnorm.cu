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
                                                        float* __restrict__ scale, float* __restrict__ shift, int C, int64_t HW, int groups, float eps,
                                                        float* __restrict__ mean_out, float* __restrict__ rstd_out) {
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
        if (mean_out) { mean_out[blockIdx.x] = s_mean; rstd_out[blockIdx.x] = s_rstd; }
    }
    __syncthreads();
    for (int c = threadIdx.x; c < cpg; c += blockDim.x) {
        const int ch = g * cpg + c;
        const float a = s_rstd * (gamma ? gamma[ch] : 1.f);
        scale[(int64_t)n * C + ch] = a;
        shift[(int64_t)n * C + ch] = (beta ? beta[ch] : 0.f) - s_mean * a;
    }
}

// out[row, :] = a[row, :] * P[row] (+ b[row, :] * Q[row]) + R[row]   -- one (n, c) plane per run of blocks_per_row CTAs, 16-byte vectors, 4 in flight.
// Forward apply: NIN = 1 (x, scale, shift).  Backward: NIN = 2 (dy, x) with the coefficients of gn_bwd_reduce_kernel.
template <class T, int NIN>
__global__ void __launch_bounds__(256) rows_affine_kernel(const T* __restrict__ a, const T* __restrict__ b, const float* __restrict__ P, const float* __restrict__ Q,
                                                          const float* __restrict__ R, T* __restrict__ out, int64_t HW, int vec_per_block, int blocks_per_row) {
    constexpr int V = 16 / (int)sizeof(T);
    // 1-D grid (rows * blocks_per_row <= 2^31 - 1): gridDim.y would cap N*C at 65535
    const int64_t row = blockIdx.x / (unsigned)blocks_per_row;
    const int chunk = (int)(blockIdx.x - (unsigned)row * (unsigned)blocks_per_row);
    const float pp = P[row], qq = (NIN == 2) ? Q[row] : 0.f, rr = R[row];
    const T* ap = a + row * HW;
    const T* bp = (NIN == 2) ? b + row * HW : nullptr;
    T* op = out + row * HW;
    if ((HW % V) == 0 && ((reinterpret_cast<uintptr_t>(a) | reinterpret_cast<uintptr_t>(out) | (NIN == 2 ? reinterpret_cast<uintptr_t>(b) : 0)) & 15u) == 0) {
        const int64_t nvec = HW / V;
        const int64_t v_begin = (int64_t)chunk * vec_per_block, v_end = (v_begin + vec_per_block < nvec) ? v_begin + vec_per_block : nvec;
        for (int64_t v0 = v_begin + threadIdx.x; v0 < v_end; v0 += (int64_t)blockDim.x * 4) {
            uint4 ua[4], ub[4];
#pragma unroll
            for (int u = 0; u < 4; u++) {
                const int64_t v = v0 + (int64_t)u * blockDim.x;
                if (v < v_end) { ua[u] = ldg_stream((const uint4*)ap + v); if (NIN == 2) ub[u] = ldg_stream((const uint4*)bp + v); }
            }
#pragma unroll
            for (int u = 0; u < 4; u++) {
                const int64_t v = v0 + (int64_t)u * blockDim.x;
                if (v >= v_end) continue;
                const T* ea = (const T*)&ua[u];
                const T* eb = (const T*)&ub[u];
                struct alignas(16) { T e[V]; } o;
#pragma unroll
                for (int k = 0; k < V; k++) {
                    float r = fmaf(to_acc(ea[k]), pp, rr);
                    if (NIN == 2) r = fmaf(to_acc(eb[k]), qq, r);
                    o.e[k] = from_acc<T, float>(r);
                }
                stg_stream((uint4*)op + v, *(const uint4*)&o);
            }
        }
    } else {
        const int64_t e_begin = (int64_t)chunk * vec_per_block * V;
        const int64_t e_end = (e_begin + (int64_t)vec_per_block * V < HW) ? e_begin + (int64_t)vec_per_block * V : HW;
        for (int64_t i = e_begin + threadIdx.x; i < e_end; i += blockDim.x) {
            float r = fmaf(to_acc(ap[i]), pp, rr);
            if (NIN == 2) r = fmaf(to_acc(bp[i]), qq, r);
            op[i] = from_acc<T, float>(r);
        }
    }
}

// Backward reductions of GroupNorm, one CTA per (sample, group): per channel c1 = sum dy, c2 = sum dy * xhat (-> dbeta / dgamma
// after a sum over the batch), group sums s1 = sum_c gamma c1, s2 = sum_c gamma c2, and the coefficients of the elementwise pass
//   dx = dy * P + x * Q + R,   P = rstd * gamma,  Q = -rstd^2 * s2 / m,  R = rstd * (rstd * mean * s2 - s1) / m      (m = group size)
template <class T>
__global__ void __launch_bounds__(512) gn_bwd_reduce_kernel(const T* __restrict__ dy, const T* __restrict__ x, const float* __restrict__ gamma,
                                                            const float* __restrict__ mean, const float* __restrict__ rstd, float* __restrict__ c1_out,
                                                            float* __restrict__ c2_out, float* __restrict__ P, float* __restrict__ Q, float* __restrict__ R,
                                                            int C, int64_t HW, int groups) {
    constexpr int V = 16 / (int)sizeof(T);
    const int n = blockIdx.x / groups, g = blockIdx.x - n * groups;
    const int cpg = C / groups;
    const float mu = mean[blockIdx.x], rs = rstd[blockIdx.x];
    // The 16 warps form `teams` = min(cpg, 16) teams of TS warps; a team owns the channels c = team (mod teams) and reduces them with
    // shuffles only, so no CTA barrier sits between channels: the layers with many channels per group and few pixels (8x8 .. 32x32)
    // reduce up to 16 channels at once instead of one after the other with two barriers each.
    __shared__ float s_part[2][64][16];            // [sum dy | sum dy*(x-mu)][channel][warp of the team]; cpg <= 64 (checked on the host)
    __shared__ float s_c1[64], s_c2[64];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int nwarps = (int)(blockDim.x >> 5);
    const int teams = cpg < nwarps ? cpg : nwarps;
    const int TS = nwarps / teams;
    const int team = warp / TS, wt = warp - team * TS;
    const bool vec_ok = (HW % V) == 0 && ((reinterpret_cast<uintptr_t>(dy) | reinterpret_cast<uintptr_t>(x)) & 15u) == 0;
    if (team < teams) {
        for (int c = team; c < cpg; c += teams) {
            const int ch = g * cpg + c;
            const T* dyp = dy + ((int64_t)n * C + ch) * HW;
            const T* xp = x + ((int64_t)n * C + ch) * HW;
            float a1 = 0.f, a2 = 0.f;
            if (vec_ok) {
                const int nvec = (int)(HW / V);          // a plane has < 2^31 vectors
                const int step = TS * 32;
                for (int v0 = wt * 32 + lane; v0 < nvec; v0 += step * 2) {
                    uint4 ud[2], ux[2];
#pragma unroll
                    for (int u = 0; u < 2; u++) {
                        const int v = v0 + u * step;
                        if (v < nvec) { ud[u] = ldg_stream((const uint4*)dyp + v); ux[u] = ldg_stream((const uint4*)xp + v); }
                    }
#pragma unroll
                    for (int u = 0; u < 2; u++) {
                        if (v0 + u * step >= nvec) continue;
                        const T* ed = (const T*)&ud[u];
                        const T* ex = (const T*)&ux[u];
#pragma unroll
                        for (int k = 0; k < V; k++) { const float d = to_acc(ed[k]); a1 += d; a2 = fmaf(d, to_acc(ex[k]) - mu, a2); }
                    }
                }
            } else {
                for (int64_t i = wt * 32 + lane; i < HW; i += (int64_t)TS * 32) { const float d = to_acc(dyp[i]); a1 += d; a2 = fmaf(d, to_acc(xp[i]) - mu, a2); }
            }
            a1 = warp_sum(a1); a2 = warp_sum(a2);
            if (lane == 0) { s_part[0][c][wt] = a1; s_part[1][c][wt] = a2; }
        }
    }
    __syncthreads();
    for (int c = threadIdx.x; c < cpg; c += blockDim.x) {
        float t1 = 0.f, t2 = 0.f;
        for (int w = 0; w < TS; w++) { t1 += s_part[0][c][w]; t2 += s_part[1][c][w]; }
        s_c1[c] = t1; s_c2[c] = t2 * rs;              // sum dy,  sum dy * xhat
    }
    __syncthreads();
    float s1 = 0.f, s2 = 0.f;
    for (int c = 0; c < cpg; c++) { const float gm = gamma ? gamma[g * cpg + c] : 1.f; s1 = fmaf(gm, s_c1[c], s1); s2 = fmaf(gm, s_c2[c], s2); }
    const float inv_m = 1.f / ((float)cpg * (float)HW);
    for (int c = threadIdx.x; c < cpg; c += blockDim.x) {
        const int ch = g * cpg + c;
        const int64_t idx = (int64_t)n * C + ch;
        c1_out[idx] = s_c1[c];
        c2_out[idx] = s_c2[c];
        P[idx] = rs * (gamma ? gamma[ch] : 1.f);
        Q[idx] = -rs * rs * s2 * inv_m;
        R[idx] = rs * (rs * mu * s2 - s1) * inv_m;
    }
}

template <class T, int NIN>
int launch_rows_affine(const void* a, const void* b, const float* P, const float* Q, const float* R, void* out, int64_t rows, int64_t HW, const char* name,
                       cudaStream_t stream) {
    constexpr int V = 16 / (int)sizeof(T);
    const int64_t nvec = (HW + V - 1) / V;
    int64_t vpb = nvec;
    while (vpb > 256 * 8 && rows * ((nvec + vpb - 1) / vpb) < (int64_t)kNumSMs * 16) vpb = (vpb + 1) / 2;
    if (vpb > 256 * 16) vpb = 256 * 16;
    const int64_t nblk = (nvec + vpb - 1) / vpb;
    if (rows * nblk > 0x7fffffffLL) { set_error("%s: too many rows (%lld x %lld blocks)", name, (long long)rows, (long long)nblk); return VFM_ERR_INVALID; }
    KernelTimer timer(name, stream, 0.0, (double)rows * HW * sizeof(T) * (NIN + 1));
    rows_affine_kernel<T, NIN><<<(unsigned)(rows * nblk), 256, 0, stream>>>((const T*)a, (const T*)b, P, Q, R, (T*)out, HW, (int)vpb, (int)nblk);
    return launch_status("rows_affine_kernel");
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
        gn_affine_kernel<__half><<<(unsigned)blocks, 512, 0, stream>>>((const __half*)p->x, p->gamma, p->beta, p->scale, p->shift, p->channels, p->hw, p->groups, (float)p->eps, nullptr, nullptr);
    else
        gn_affine_kernel<float><<<(unsigned)blocks, 512, 0, stream>>>((const float*)p->x, p->gamma, p->beta, p->scale, p->shift, p->channels, p->hw, p->groups, (float)p->eps, nullptr, nullptr);
    return launch_status("gn_affine_kernel");
}

static int gn_check(const vfm_group_norm_params* p, const char* what) {
    using namespace vfm;
    VFM_CHECK_ARG(p != nullptr, "%s: params is NULL", what);
    VFM_CHECK_ARG(p->x && p->mean && p->rstd && p->scratch, "%s: x, mean, rstd and scratch must be non-NULL", what);
    VFM_CHECK_ARG(p->batch >= 1 && p->channels >= 1 && p->hw >= 1, "%s: x is empty", what);
    VFM_CHECK_ARG(p->groups >= 1 && p->channels % p->groups == 0 && p->channels / p->groups <= 64, "%s: channels (%d) must be divisible by groups (%d), <= 64 per group", what, p->channels, p->groups);
    VFM_CHECK_ARG(p->dtype == VFM_F16 || p->dtype == VFM_F32, "%s: unsupported dtype %d", what, p->dtype);
    VFM_CHECK_ARG((int64_t)p->batch * p->groups <= 0x7fffffffLL, "%s: grid too large", what);
    return VFM_OK;
}

extern "C" int vfm_group_norm_forward(const vfm_group_norm_params* p, void* stream_) {
    using namespace vfm;
    cudaStream_t stream = (cudaStream_t)stream_;
    int st = gn_check(p, "group_norm_forward"); if (st) return st;
    VFM_CHECK_ARG(p->y != nullptr, "group_norm_forward: y must be non-NULL");
    const size_t nc = (size_t)p->batch * p->channels;
    float* scale = p->scratch; float* shift = p->scratch + nc;
    const unsigned blocks = (unsigned)((int64_t)p->batch * p->groups);
    {
        KernelTimer timer("group_norm_affine", stream, 0.0, (double)nc * (double)p->hw * (p->dtype == VFM_F16 ? 2.0 : 4.0));
        if (p->dtype == VFM_F16)
            gn_affine_kernel<__half><<<blocks, 512, 0, stream>>>((const __half*)p->x, p->gamma, p->beta, scale, shift, p->channels, p->hw, p->groups, (float)p->eps, p->mean, p->rstd);
        else
            gn_affine_kernel<float><<<blocks, 512, 0, stream>>>((const float*)p->x, p->gamma, p->beta, scale, shift, p->channels, p->hw, p->groups, (float)p->eps, p->mean, p->rstd);
    }
    st = launch_status("gn_affine_kernel"); if (st) return st;
    if (p->dtype == VFM_F16) return launch_rows_affine<__half, 1>(p->x, nullptr, scale, nullptr, shift, p->y, (int64_t)nc, p->hw, "group_norm_apply", stream);
    return launch_rows_affine<float, 1>(p->x, nullptr, scale, nullptr, shift, p->y, (int64_t)nc, p->hw, "group_norm_apply", stream);
}

extern "C" int vfm_group_norm_backward(const vfm_group_norm_params* p, void* stream_) {
    using namespace vfm;
    cudaStream_t stream = (cudaStream_t)stream_;
    int st = gn_check(p, "group_norm_backward"); if (st) return st;
    VFM_CHECK_ARG(p->dy && p->dx && p->dgamma_nc && p->dbeta_nc, "group_norm_backward: dy, dx, dgamma_nc and dbeta_nc must be non-NULL");
    const size_t nc = (size_t)p->batch * p->channels;
    float* P = p->scratch; float* Q = p->scratch + nc; float* R = p->scratch + 2 * nc;
    const unsigned blocks = (unsigned)((int64_t)p->batch * p->groups);
    {
        KernelTimer timer("group_norm_bwd_reduce", stream, 0.0, 2.0 * (double)nc * (double)p->hw * (p->dtype == VFM_F16 ? 2.0 : 4.0));
        if (p->dtype == VFM_F16)
            gn_bwd_reduce_kernel<__half><<<blocks, 512, 0, stream>>>((const __half*)p->dy, (const __half*)p->x, p->gamma, p->mean, p->rstd, p->dbeta_nc, p->dgamma_nc, P, Q, R, p->channels, p->hw, p->groups);
        else
            gn_bwd_reduce_kernel<float><<<blocks, 512, 0, stream>>>((const float*)p->dy, (const float*)p->x, p->gamma, p->mean, p->rstd, p->dbeta_nc, p->dgamma_nc, P, Q, R, p->channels, p->hw, p->groups);
    }
    st = launch_status("gn_bwd_reduce_kernel"); if (st) return st;
    if (p->dtype == VFM_F16) return launch_rows_affine<__half, 2>(p->dy, p->x, P, Q, R, p->dx, (int64_t)nc, p->hw, "group_norm_bwd_apply", stream);
    return launch_rows_affine<float, 2>(p->dy, p->x, P, Q, R, p->dx, (int64_t)nc, p->hw, "group_norm_bwd_apply", stream);
}


// ---- generic row-wise helpers behind the fused layer-scaled residual of the training path (networks/generator.py:272-274) ----
//   vfm_rows_affine: out[r, :] = a[r, :] * P[r] (+ b[r, :] * Q[r]) + R[r]       rows = N*C planes of hw elements
//   vfm_rows_dot   : out[r]    = sum_i a[r, i] * b[r, i]
namespace vfm {
namespace {
template <class T>
__global__ void __launch_bounds__(256) rows_dot_kernel(const T* __restrict__ a, const T* __restrict__ b, float* __restrict__ out, int64_t HW) {
    constexpr int V = 16 / (int)sizeof(T);
    const int64_t row = blockIdx.x;
    const T* ap = a + row * HW;
    const T* bp = b + row * HW;
    float acc = 0.f;
    if ((HW % V) == 0 && ((reinterpret_cast<uintptr_t>(a) | reinterpret_cast<uintptr_t>(b)) & 15u) == 0) {
        const int64_t nvec = HW / V;
        for (int64_t v0 = threadIdx.x; v0 < nvec; v0 += (int64_t)blockDim.x * 4) {
            uint4 ua[4], ub[4];
#pragma unroll
            for (int u = 0; u < 4; u++) {
                const int64_t v = v0 + (int64_t)u * blockDim.x;
                if (v < nvec) { ua[u] = ldg_stream((const uint4*)ap + v); ub[u] = ldg_stream((const uint4*)bp + v); }
            }
#pragma unroll
            for (int u = 0; u < 4; u++) {
                if (v0 + (int64_t)u * blockDim.x >= nvec) continue;
                const T* ea = (const T*)&ua[u];
                const T* eb = (const T*)&ub[u];
#pragma unroll
                for (int k = 0; k < V; k++) acc = fmaf(to_acc(ea[k]), to_acc(eb[k]), acc);
            }
        }
    } else {
        for (int64_t i = threadIdx.x; i < HW; i += blockDim.x) acc = fmaf(to_acc(ap[i]), to_acc(bp[i]), acc);
    }
    __shared__ float red[32];
    acc = block_sum(acc, red);
    if (threadIdx.x == 0) out[row] = acc;
}
}  // namespace
}  // namespace vfm

extern "C" int vfm_rows_affine(const vfm_rows_params* p, void* stream_) {
    using namespace vfm;
    cudaStream_t stream = (cudaStream_t)stream_;
    VFM_CHECK_ARG(p && p->a && p->out && p->P && p->R, "rows_affine: a, out, P and R must be non-NULL");
    VFM_CHECK_ARG((p->b == nullptr) == (p->Q == nullptr), "rows_affine: b and Q come together");
    VFM_CHECK_ARG(p->rows >= 1 && p->hw >= 1, "rows_affine: empty tensor");
    VFM_CHECK_ARG(p->dtype == VFM_F16 || p->dtype == VFM_F32, "rows_affine: unsupported dtype %d", p->dtype);
    if (p->dtype == VFM_F16)
        return p->b ? launch_rows_affine<__half, 2>(p->a, p->b, p->P, p->Q, p->R, p->out, p->rows, p->hw, "rows_affine", stream)
                    : launch_rows_affine<__half, 1>(p->a, nullptr, p->P, nullptr, p->R, p->out, p->rows, p->hw, "rows_affine", stream);
    return p->b ? launch_rows_affine<float, 2>(p->a, p->b, p->P, p->Q, p->R, p->out, p->rows, p->hw, "rows_affine", stream)
                : launch_rows_affine<float, 1>(p->a, nullptr, p->P, nullptr, p->R, p->out, p->rows, p->hw, "rows_affine", stream);
}

extern "C" int vfm_rows_dot(const vfm_rows_params* p, void* stream_) {
    using namespace vfm;
    cudaStream_t stream = (cudaStream_t)stream_;
    VFM_CHECK_ARG(p && p->a && p->b && p->out, "rows_dot: a, b and out must be non-NULL");
    VFM_CHECK_ARG(p->rows >= 1 && p->rows <= 0x7fffffffLL && p->hw >= 1, "rows_dot: bad shape");
    VFM_CHECK_ARG(p->dtype == VFM_F16 || p->dtype == VFM_F32, "rows_dot: unsupported dtype %d", p->dtype);
    KernelTimer timer("rows_dot", stream, 0.0, 2.0 * (double)p->rows * (double)p->hw * (p->dtype == VFM_F16 ? 2.0 : 4.0));
    if (p->dtype == VFM_F16) rows_dot_kernel<__half><<<(unsigned)p->rows, 256, 0, stream>>>((const __half*)p->a, (const __half*)p->b, (float*)p->out, p->hw);
    else rows_dot_kernel<float><<<(unsigned)p->rows, 256, 0, stream>>>((const float*)p->a, (const float*)p->b, (float*)p->out, p->hw);
    return launch_status("rows_dot_kernel");
}
