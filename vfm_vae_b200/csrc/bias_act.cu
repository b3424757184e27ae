// bias_act for sm_100a: y = clamp(act(x + b) * gain), its first and second derivative, and the fused bias-gradient
// reduction.  HBM-bound elementwise kernel: 16-byte loads/stores, 4 independent vectors in flight per thread,
// no per-element integer division (one per 16-byte vector, 32-bit when the tensor allows), grid sized to the SM count.
//
// Semantics follow the reference plugin (torch_utils/ops/bias_act.cu:23-147, bias_act.cpp:32-90) and are checked
// against the reference's PyTorch path _bias_act_ref (torch_utils/ops/bias_act.py:91-120) through the oracle.
// Unlike the reference build (--use_fast_math) the transcendental paths use the accurate libdevice functions so
// that fp32 results stay within 1e-5 of the PyTorch reference.
#include "common.cuh"

namespace vfm {
namespace {

struct BiasActArgs {
    const void* x; const void* b; const void* xref; const void* yref; const void* dy; void* y; float* db;
    int grad; double alpha, gain, clamp;
    int64_t size_x, size_b, step_b;
    int64_t vec_per_block;  // contiguous span of vectors handled by one block
};

template <class S> __device__ __forceinline__ S fexp(S v);
template <> __device__ __forceinline__ float fexp<float>(float v) { return expf(v); }
template <> __device__ __forceinline__ double fexp<double>(double v) { return exp(v); }
template <class S> __device__ __forceinline__ S fexpm1(S v);
template <> __device__ __forceinline__ float fexpm1<float>(float v) { return expm1f(v); }
template <> __device__ __forceinline__ double fexpm1<double>(double v) { return expm1(v); }
template <class S> __device__ __forceinline__ S flog1p(S v);
template <> __device__ __forceinline__ float flog1p<float>(float v) { return log1pf(v); }
template <> __device__ __forceinline__ double flog1p<double>(double v) { return log1p(v); }
template <class S> __device__ __forceinline__ S ftanh(S v);
template <> __device__ __forceinline__ float ftanh<float>(float v) { return tanhf(v); }
template <> __device__ __forceinline__ double ftanh<double>(double v) { return tanh(v); }

// One element.  `x` is the tensor being transformed (activations for G=0, incoming gradient for G>=1).
template <int A, class S>
__device__ __forceinline__ S bias_act_elem(S x, S b, S xref, S yref, S dy, int G, S alpha, S gain, S clamp) {
    const S one = (S)1, two = (S)2, zero = (S)0;
    const S selu_scale = (S)1.0507009873554804934193349852946;
    const S selu_alpha = (S)1.6732632423543772848170429916717;
    S yy = (gain != zero) ? yref / gain : zero;
    S y = zero;
    if (G == 0) x += b; else xref += b;

    if (A == 1) { y = x; if (G == 2) y = zero; }
    if (A == 2) { if (G == 0) y = (x > zero) ? x : zero; else if (G == 1) y = (yy > zero) ? x : zero; }
    if (A == 3) { if (G == 0) y = (x > zero) ? x : x * alpha; else if (G == 1) y = (yy > zero) ? x : x * alpha; }
    if (A == 4) {
        if (G == 0) y = ftanh(x);
        else if (G == 1) y = x * (one - yy * yy);
        else y = x * (one - yy * yy) * (-two * yy);
    }
    if (A == 5) {
        if (G == 0) y = one / (one + fexp(-x));
        else if (G == 1) y = x * yy * (one - yy);
        else y = x * yy * (one - yy) * (one - two * yy);
    }
    if (A == 6) {
        if (G == 0) y = (x >= zero) ? x : fexpm1(x);
        else if (G == 1) y = (yy >= zero) ? x : x * (yy + one);
        else y = (yy >= zero) ? zero : x * (yy + one);
    }
    if (A == 7) {
        if (G == 0) y = (x >= zero) ? selu_scale * x : (selu_scale * selu_alpha) * fexpm1(x);
        else if (G == 1) y = (yy >= zero) ? x * selu_scale : x * (yy + selu_scale * selu_alpha);
        else y = (yy >= zero) ? zero : x * (yy + selu_scale * selu_alpha);
    }
    if (A == 8) {
        if (G == 0) y = (x > (S)20) ? x : flog1p(fexp(x));
        else if (G == 1) y = x * (one - fexp(-yy));
        else { S c = fexp(-yy); y = x * c * (one - c); }
    }
    if (A == 9) {
        if (G == 0) y = x / (one + fexp(-x));
        else {
            S c = fexp(xref), d = c + one;
            if (G == 1) y = (xref > (S)40) ? x : x * c * (xref + d) / (d * d);
            else y = (xref > (S)40) ? zero : x * c * (xref * (two - d) + two * d) / (d * d * d);
            yref = xref / (one + fexp(-xref)) * gain;   // swish keeps x, not y: rebuild y for the clamp mask
        }
    }
    y *= gain * dy;
    if (clamp >= zero) {
        if (G == 0) y = (y > -clamp && y < clamp) ? y : ((y >= zero) ? clamp : -clamp);
        else y = (yref > -clamp && yref < clamp) ? y : zero;
    }
    return y;
}

// MODE 0: one bias value per 16-byte vector (step_b % VEC == 0: NCHW and friends)
// MODE 1: VEC consecutive bias values per vector (step_b == 1, size_b % VEC == 0: channels_last, [N,C])
// MODE 2: scalar generic (VEC = 1)
template <class T, int A, int MODE>
__global__ void __launch_bounds__(256) bias_act_kernel(BiasActArgs p) {
    typedef typename Acc<T>::type S;
    constexpr int VEC = (MODE == 2) ? 1 : (int)(16 / sizeof(T));
    constexpr int UNROLL = 4;
    extern __shared__ float s_db[];   // [size_b] when db is requested

    const bool want_db = (p.db != nullptr);
    if (want_db) {
        for (int64_t i = threadIdx.x; i < p.size_b; i += blockDim.x) s_db[i] = 0.f;
        __syncthreads();
    }

    const int64_t nvec = (MODE == 2) ? p.size_x : p.size_x / VEC;
    const int64_t v_begin = (int64_t)blockIdx.x * p.vec_per_block;
    const int64_t v_end = min(v_begin + p.vec_per_block, nvec);
    const S alpha = (S)p.alpha, gain = (S)p.gain, clamp = (S)p.clamp;
    const int G = p.grad;
    const bool small = p.size_x < (int64_t)0x7fffffff;

    for (int64_t v0 = v_begin + threadIdx.x; v0 < v_end; v0 += (int64_t)blockDim.x * UNROLL) {
        struct alignas(MODE == 2 ? sizeof(T) : 16) Vec { T e[VEC]; };
        Vec xv[UNROLL], rv[UNROLL], yv[UNROLL], dv[UNROLL];
        bool act[UNROLL];
#pragma unroll
        for (int u = 0; u < UNROLL; u++) {
            int64_t v = v0 + (int64_t)u * blockDim.x;
            act[u] = v < v_end;
            if (!act[u]) continue;
            if (MODE == 2) {
                xv[u].e[0] = ((const T*)p.x)[v];
                if (p.xref) rv[u].e[0] = ((const T*)p.xref)[v];
                if (p.yref) yv[u].e[0] = ((const T*)p.yref)[v];
                if (p.dy) dv[u].e[0] = ((const T*)p.dy)[v];
            } else {
                *(uint4*)&xv[u] = ldg_stream((const uint4*)p.x + v);
                if (p.xref) *(uint4*)&rv[u] = ldg_stream((const uint4*)p.xref + v);
                if (p.yref) *(uint4*)&yv[u] = ldg_stream((const uint4*)p.yref + v);
                if (p.dy) *(uint4*)&dv[u] = ldg_stream((const uint4*)p.dy + v);
            }
        }
#pragma unroll
        for (int u = 0; u < UNROLL; u++) {
            if (!act[u]) continue;
            int64_t v = v0 + (int64_t)u * blockDim.x;
            int64_t e0 = v * VEC;
            int64_t c0 = 0;
            if (p.b || want_db) {
                if (small) c0 = (int64_t)(((uint32_t)e0 / (uint32_t)p.step_b) % (uint32_t)p.size_b);
                else c0 = (e0 / p.step_b) % p.size_b;
            }
            Vec outv;
            T* out = outv.e;
            S sum = (S)0;
#pragma unroll
            for (int k = 0; k < VEC; k++) {
                int64_t c = (MODE == 1) ? c0 + k : c0;
                S b = p.b ? to_acc(((const T*)p.b)[c]) : (S)0;
                S r = bias_act_elem<A, S>(to_acc(xv[u].e[k]), b, p.xref ? to_acc(rv[u].e[k]) : (S)0,
                                          p.yref ? to_acc(yv[u].e[k]) : (S)0, p.dy ? to_acc(dv[u].e[k]) : (S)1,
                                          G, alpha, gain, clamp);
                out[k] = from_acc<T, S>(r);
                if (want_db) {
                    if (MODE == 1) atomicAdd(&s_db[c], (float)r);
                    else sum += r;
                }
            }
            if (MODE == 2) ((T*)p.y)[v] = out[0];
            else stg_stream((uint4*)p.y + v, *(const uint4*)&outv);
            if (want_db && MODE != 1) {
                // lanes of a warp usually sit in the same (n,c) row: one shuffle reduction, one shared atomic
                unsigned mask = __activemask();
                int64_t cl = __shfl_sync(mask, c0, __ffs(mask) - 1);
                bool uniform = __all_sync(mask, cl == c0) && mask == 0xffffffffu;
                if (uniform) {
                    float s = warp_sum((float)sum);
                    if ((threadIdx.x & 31) == 0) atomicAdd(&s_db[c0], s);
                } else {
                    atomicAdd(&s_db[c0], (float)sum);
                }
            }
        }
    }

    if (want_db) {
        __syncthreads();
        for (int64_t i = threadIdx.x; i < p.size_b; i += blockDim.x) {
            float s = s_db[i];
            if (s != 0.f) atomicAdd(&p.db[i], s);
        }
    }
}

// ---- hot path: the decoder's linear / lrelu layers on NCHW tensors ----------------------------------------------------
// One (n,c) row per blockIdx.y: the bias is a block constant, there is no integer division anywhere, the derivative of
// lrelu is a select on the sign of y (no y/gain division), the bias gradient is one block reduction + one atomic.
// ~6 instructions per element, so the kernel is limited by HBM and not by instruction issue.
template <class T, int A, int G>
__global__ void __launch_bounds__(256) bias_act_rows_kernel(BiasActArgs p, int vec_per_row, int vec_per_block) {
    constexpr int VEC = (int)(16 / sizeof(T));
    constexpr int UNROLL = 4;
    __shared__ float red[32];
    struct alignas(16) Vec { T e[VEC]; };
    const int64_t row = blockIdx.y;
    const int c = (int)(row % p.size_b);
    const float bias = (p.b && G == 0) ? to_acc(((const T*)p.b)[c]) : 0.f;
    const float gain = (float)p.gain, alpha = (float)p.alpha, clamp = (float)p.clamp;
    const float gpos = gain, gneg = (A == 3) ? gain * alpha : gain;
    const bool pos_if_positive = gain > 0.f;      // sign(y) == sign(act(x)) * sign(gain)
    const int v_begin = blockIdx.x * vec_per_block, v_end = min(v_begin + vec_per_block, vec_per_row);
    const uint4* xin = (const uint4*)p.x + row * vec_per_row;
    const uint4* yin = (const uint4*)p.yref + row * vec_per_row;
    uint4* out = (uint4*)p.y + row * vec_per_row;
    float sum = 0.f;
    for (int v0 = v_begin + threadIdx.x; v0 < v_end; v0 += blockDim.x * UNROLL) {
        Vec xv[UNROLL], yv[UNROLL];
#pragma unroll
        for (int u = 0; u < UNROLL; u++) {
            const int v = v0 + u * blockDim.x;
            if (v < v_end) {
                *(uint4*)&xv[u] = ldg_stream(xin + v);
                if (G == 1 && p.yref) *(uint4*)&yv[u] = ldg_stream(yin + v);
            }
        }
#pragma unroll
        for (int u = 0; u < UNROLL; u++) {
            const int v = v0 + u * blockDim.x;
            if (v >= v_end) continue;
            Vec o;
#pragma unroll
            for (int k = 0; k < VEC; k++) {
                float x = to_acc(xv[u].e[k]);
                float r;
                if (G == 0) {
                    x += bias;
                    r = x * ((x > 0.f) ? gpos : gneg);
                    if (clamp >= 0.f) r = fminf(fmaxf(r, -clamp), clamp);
                } else {
                    const float yr = p.yref ? to_acc(yv[u].e[k]) : 0.f;
                    const bool pos = (A == 1) ? true : ((yr > 0.f) == pos_if_positive && yr != 0.f);
                    r = x * (pos ? gpos : gneg);
                    if (clamp >= 0.f && !(yr > -clamp && yr < clamp)) r = 0.f;
                    sum += r;
                }
                o.e[k] = from_acc<T, float>(r);
            }
            stg_stream(out + v, *(const uint4*)&o);
        }
    }
    if (G == 1 && p.db) {
        sum = block_sum(sum, red);
        if (threadIdx.x == 0) atomicAdd(&p.db[c], sum);
    }
}

template <class T, int A, int G>
int launch_rows(const BiasActArgs& a, cudaStream_t stream) {
    constexpr int VEC = (int)(16 / sizeof(T));
    const int vec_per_row = (int)(a.step_b / VEC);
    const int64_t rows = a.size_x / a.step_b;
    int vec_per_block = vec_per_row;
    // split long rows so that there are enough blocks; keep >= 4 vectors per thread when possible
    while (vec_per_block > 256 * 8 && rows * ceil_div(vec_per_row, vec_per_block) < (int64_t)kNumSMs * 16) vec_per_block = ceil_div(vec_per_block, 2);
    if (vec_per_block > 256 * 16) vec_per_block = 256 * 16;
    dim3 grid(ceil_div(vec_per_row, vec_per_block), (unsigned)rows);
    int ntens = 2 + ((G == 1 && a.yref) ? 1 : 0);
    KernelTimer timer(G == 0 ? "bias_act_fwd" : "bias_act_grad", stream, 0.0, (double)ntens * (double)a.size_x * sizeof(T) + (double)a.size_b * sizeof(T));
    bias_act_rows_kernel<T, A, G><<<grid, 256, 0, stream>>>(a, vec_per_row, vec_per_block);
    return launch_status("bias_act_rows_kernel");
}

// returns VFM_ERR_NO_KERNEL when the call does not fit the hot-path kernel
template <class T>
int try_rows(const BiasActArgs& a, int act, int mode, cudaStream_t stream) {
    constexpr int VEC = (int)(16 / sizeof(T));
    if (mode != 0 || !(act == 1 || act == 3) || a.grad > 1 || a.xref || a.dy) return VFM_ERR_NO_KERNEL;
    if (a.step_b < 256 || a.step_b % VEC != 0 || a.size_x % a.step_b != 0) return VFM_ERR_NO_KERNEL;
    if (a.size_x / a.step_b > 65535 || a.step_b / VEC > (int64_t)0x7fffffff) return VFM_ERR_NO_KERNEL;
    if (a.grad == 1 && act == 3 && !a.yref) return VFM_ERR_NO_KERNEL;
    if (act == 1) return a.grad == 0 ? launch_rows<T, 1, 0>(a, stream) : launch_rows<T, 1, 1>(a, stream);
    return a.grad == 0 ? launch_rows<T, 3, 0>(a, stream) : launch_rows<T, 3, 1>(a, stream);
}

template <class T, int A>
int launch_mode(const BiasActArgs& a0, int mode, cudaStream_t stream) {
    BiasActArgs a = a0;
    const int VEC = (mode == 2) ? 1 : (int)(16 / sizeof(T));
    const int64_t nvec = (mode == 2) ? a.size_x : a.size_x / VEC;
    const int threads = 256;
    const int64_t per_iter = (int64_t)threads * 4;
    // contiguous span per block: a multiple of one unrolled sweep, aiming at ~8 blocks per SM
    int64_t target_blocks = (int64_t)kNumSMs * 8;
    int64_t span = ceil_div64(ceil_div64(nvec, target_blocks), per_iter) * per_iter;
    if (span < per_iter) span = per_iter;
    a.vec_per_block = span;
    int64_t blocks = ceil_div64(nvec, span);
    size_t smem = a.db ? (size_t)a.size_b * sizeof(float) : 0;
    void (*kern)(BiasActArgs) = nullptr;
    if (mode == 0) kern = bias_act_kernel<T, A, 0>;
    else if (mode == 1) kern = bias_act_kernel<T, A, 1>;
    else kern = bias_act_kernel<T, A, 2>;
    if (smem > 48 * 1024) VFM_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int ntens = 2 + (a.xref ? 1 : 0) + (a.yref ? 1 : 0) + (a.dy ? 1 : 0);
    KernelTimer timer(a.grad == 0 ? "bias_act_fwd" : (a.grad == 1 ? "bias_act_grad" : "bias_act_grad2"), stream, 0.0,
                      (double)ntens * (double)a.size_x * sizeof(T) + (double)a.size_b * sizeof(T));
    kern<<<(unsigned)blocks, threads, smem, stream>>>(a);
    return launch_status("bias_act_kernel");
}

template <class T>
int launch_act(const BiasActArgs& a, int act, int mode, cudaStream_t stream) {
    switch (act) {
        case 1: return launch_mode<T, 1>(a, mode, stream);
        case 2: return launch_mode<T, 2>(a, mode, stream);
        case 3: return launch_mode<T, 3>(a, mode, stream);
        case 4: return launch_mode<T, 4>(a, mode, stream);
        case 5: return launch_mode<T, 5>(a, mode, stream);
        case 6: return launch_mode<T, 6>(a, mode, stream);
        case 7: return launch_mode<T, 7>(a, mode, stream);
        case 8: return launch_mode<T, 8>(a, mode, stream);
        case 9: return launch_mode<T, 9>(a, mode, stream);
    }
    set_error("bias_act: no kernel for activation index %d", act);
    return VFM_ERR_NO_KERNEL;
}

}  // namespace
}  // namespace vfm

extern "C" int vfm_bias_act(const vfm_bias_act_params* p, void* stream_) {
    using namespace vfm;
    cudaStream_t stream = (cudaStream_t)stream_;
    VFM_CHECK_ARG(p != nullptr, "bias_act: params is NULL");
    VFM_CHECK_ARG(p->x && p->y, "bias_act: x and y must be non-NULL");
    VFM_CHECK_ARG(p->size_x >= 0, "bias_act: negative size");
    VFM_CHECK_ARG(p->grad >= 0 && p->grad <= 2, "bias_act: grad must be 0, 1 or 2 (third-order gradients are not supported)");
    VFM_CHECK_ARG(p->dtype == VFM_F16 || p->dtype == VFM_F32 || p->dtype == VFM_F64, "bias_act: unsupported dtype %d", p->dtype);
    VFM_CHECK_ARG(!(p->b || p->db) || (p->size_b > 0 && p->step_b > 0), "bias_act: b/db given but size_b/step_b not positive");
    VFM_CHECK_ARG(!p->db || p->size_b * 4 <= 200 * 1024, "bias_act: db with more than 51200 channels is not supported");
    if (p->size_x == 0) return VFM_OK;

    BiasActArgs a;
    a.x = p->x; a.b = p->b; a.xref = p->xref; a.yref = p->yref; a.dy = p->dy; a.y = p->y; a.db = p->db;
    a.grad = p->grad; a.alpha = p->alpha; a.gain = p->gain; a.clamp = p->clamp;
    a.size_x = p->size_x;
    a.size_b = (p->b || p->db) ? p->size_b : 1;
    a.step_b = (p->b || p->db) ? p->step_b : 1;
    a.vec_per_block = 0;

    const int esize = (p->dtype == VFM_F16) ? 2 : (p->dtype == VFM_F32 ? 4 : 8);
    const int vec = 16 / esize;
    bool al = aligned16(p->x) && aligned16(p->y) && (!p->xref || aligned16(p->xref)) && (!p->yref || aligned16(p->yref)) &&
              (!p->dy || aligned16(p->dy)) && (p->size_x % vec == 0);
    int mode = 2;
    if (al) {
        if (!(p->b || p->db)) mode = 0;
        else if (a.step_b % vec == 0) mode = 0;
        else if (a.step_b == 1 && a.size_b % vec == 0) mode = 1;
    }
    if (p->dtype != VFM_F64) {
        int st = (p->dtype == VFM_F16) ? try_rows<__half>(a, p->act, mode, stream) : try_rows<float>(a, p->act, mode, stream);
        if (st != VFM_ERR_NO_KERNEL) return st;
    }
    switch (p->dtype) {
        case VFM_F16: return launch_act<__half>(a, p->act, mode, stream);
        case VFM_F32: return launch_act<float>(a, p->act, mode, stream);
        default:      return launch_act<double>(a, p->act, mode, stream);
    }
}
