// 1x1 modulated conv with a handful of output channels (the decoder's ToRGB layers: C -> 3, demodulate off,
// networks/generator.py:284-312).  With 3 outputs there is no GEMM to speak of: the op reads x once and writes 3 planes,
// so it is HBM-bound and handled by streaming SIMT kernels instead of the tensor-core path.
//
//   forward : y[n,o,p]  = osc[n,o] * sum_c W[o,c] * isc[n,c] * x[n,c,p] (+ noise)
//   backward: dx[n,c,p] = isc[n,c] * sum_o W[o,c] * osc[n,o] * dy[n,o,p]
//             R[n,o,c]  = sum_p dy[n,o,p] * x[n,c,p]                      (one pass over x)
//             dW[o,c]   = sum_n osc[n,o] * isc[n,c] * R[n,o,c]
//             dsum[n,c] = sum_o W[o,c] * osc[n,o] * R[n,o,c]               (= sum_p x * dxpre, feeds dstyles)
#include "modconv_common.cuh"

namespace vfm {
namespace modconv {

namespace {

constexpr int OMAX = 4;

// Forward.  A CTA of 256 threads = PT pixel threads (16 bytes of one plane each) x S = 256/PT channel slices: slice j accumulates the
// channels c = j (mod S), the slices are summed through shared memory.  S = 1 for the large images (enough CTAs and a long channel
// loop per thread); the 8x8 .. 64x64 layers have few pixels and many channels -- with one thread per pixel group they were a chain of
// C dependent-latency loads on a handful of warps (0.1 - 0.2 ms for a few MB) -- so the host picks S up to 16 there.
template <class T>
__global__ void __launch_bounds__(256) pw_fwd_kernel(const T* __restrict__ x, const float* __restrict__ w, const float* __restrict__ isc,
                                                     const float* __restrict__ osc, const float* __restrict__ add, int64_t add_sn,
                                                     T* __restrict__ y, int C, int O, int HW, int PT, Epilogue ep) {
    constexpr int VEC = (int)(16 / sizeof(T));
    extern __shared__ float s_w[];     // [C][OMAX]: W[o,c] * isc[n,c]; then [S][OMAX][PT][VEC] partial sums (S > 1)
    const int n = blockIdx.y;
    for (int i = threadIdx.x; i < C * OMAX; i += blockDim.x) {
        int c = i / OMAX, o = i - c * OMAX;
        s_w[i] = (o < O) ? w[(size_t)o * C + c] * isc[(size_t)n * C + c] : 0.f;
    }
    __syncthreads();
    const int S = 256 / PT;
    const int pg = threadIdx.x % PT, slice = threadIdx.x / PT;
    const int p0 = (blockIdx.x * PT + pg) * VEC;
    const bool live = p0 < HW;
    float acc[OMAX][VEC];
#pragma unroll
    for (int o = 0; o < OMAX; o++)
#pragma unroll
        for (int v = 0; v < VEC; v++) acc[o][v] = 0.f;
    struct alignas(16) Vec { T e[VEC]; };
    if (live) {
        const T* xp = x + (size_t)n * C * HW + p0;
#pragma unroll 4
        for (int c = slice; c < C; c += S) {
            Vec xv;
            *(uint4*)&xv = ldg_stream((const uint4*)(xp + (size_t)c * HW));
            const float4 wv = *(const float4*)&s_w[c * OMAX];
#pragma unroll
            for (int v = 0; v < VEC; v++) {
                const float xf = to_acc(xv.e[v]);
                acc[0][v] += wv.x * xf; acc[1][v] += wv.y * xf; acc[2][v] += wv.z * xf; acc[3][v] += wv.w * xf;
            }
        }
    }
    if (S > 1) {
        float* red = s_w + (size_t)C * OMAX;           // [S][OMAX][PT][VEC]
#pragma unroll
        for (int o = 0; o < OMAX; o++)
#pragma unroll
            for (int v = 0; v < VEC; v += 4)
                *(float4*)&red[(((size_t)slice * OMAX + o) * PT + pg) * VEC + v] = make_float4(acc[o][v], acc[o][v + 1], acc[o][v + 2], acc[o][v + 3]);
        __syncthreads();
    }
    if (!live) return;
    // output plane o is finished by slice o mod S (all planes by the only slice when S == 1)
    for (int o = (S > 1 ? slice : 0); o < O; o += S) {
        float r[VEC];
        if (S > 1) {
            const float* red = s_w + (size_t)C * OMAX;
#pragma unroll
            for (int v = 0; v < VEC; v++) r[v] = 0.f;
            for (int j = 0; j < S; j++)
#pragma unroll
                for (int v = 0; v < VEC; v += 4) {
                    const float4 t = *(const float4*)&red[(((size_t)j * OMAX + o) * PT + pg) * VEC + v];
                    r[v] += t.x; r[v + 1] += t.y; r[v + 2] += t.z; r[v + 3] += t.w;
                }
        } else {
#pragma unroll
            for (int v = 0; v < VEC; v++) r[v] = o == 0 ? acc[0][v] : o == 1 ? acc[1][v] : o == 2 ? acc[2][v] : acc[3][v];
        }
        const float sc = osc[(size_t)n * O + o];
        Vec out;
#pragma unroll
        for (int v = 0; v < VEC; v++) {
            float t = r[v] * sc;
            if (add) t += add[(size_t)n * add_sn + p0 + v];
            if (ep.enable) t = apply_epilogue<T>(ep, t, o, ((size_t)n * O + o) * HW + p0 + v);
            out.e[v] = from_acc<T, float>(t);
        }
        *(uint4*)(y + ((size_t)n * O + o) * HW + p0) = *(const uint4*)&out;
    }
}

template <class T>
__global__ void __launch_bounds__(256) pw_dgrad_kernel(const T* __restrict__ dy, const float* __restrict__ w, const float* __restrict__ isc,
                                                       const float* __restrict__ osc, T* __restrict__ dx, int C, int O, int HW) {
    constexpr int VEC = (int)(16 / sizeof(T));
    extern __shared__ float s_w[];     // [C][OMAX]: W[o,c] * osc[n,o] * isc[n,c]
    const int n = blockIdx.y;
    for (int i = threadIdx.x; i < C * OMAX; i += blockDim.x) {
        int c = i / OMAX, o = i - c * OMAX;
        s_w[i] = (o < O) ? w[(size_t)o * C + c] * osc[(size_t)n * O + o] * isc[(size_t)n * C + c] : 0.f;
    }
    __syncthreads();
    const int p0 = (blockIdx.x * blockDim.x + threadIdx.x) * VEC;
    if (p0 >= HW) return;
    struct alignas(16) Vec { T e[VEC]; };
    float dv[OMAX][VEC];
#pragma unroll
    for (int o = 0; o < OMAX; o++) {
        Vec t;
        if (o < O) *(uint4*)&t = ldg_stream((const uint4*)(dy + ((size_t)n * O + o) * HW + p0));
#pragma unroll
        for (int v = 0; v < VEC; v++) dv[o][v] = (o < O) ? to_acc(t.e[v]) : 0.f;
    }
    T* dxp = dx + (size_t)n * C * HW + p0;
#pragma unroll 4
    for (int c = 0; c < C; c++) {
        const float4 wv = *(const float4*)&s_w[c * OMAX];
        Vec out;
#pragma unroll
        for (int v = 0; v < VEC; v++) out.e[v] = from_acc<T, float>(wv.x * dv[0][v] + wv.y * dv[1][v] + wv.z * dv[2][v] + wv.w * dv[3][v]);
        stg_stream((uint4*)(dxp + (size_t)c * HW), *(const uint4*)&out);
    }
}

// R[n,o,c] = sum_p dy[n,o,p] * x[n,c,p]; one block per (c, n)
template <class T>
__global__ void __launch_bounds__(256) pw_corr_kernel(const T* __restrict__ dy, const T* __restrict__ x, float* __restrict__ R, int C, int O, int HW) {
    constexpr int VEC = (int)(16 / sizeof(T));
    __shared__ float red[32];
    const int c = blockIdx.x, n = blockIdx.y;
    const T* xp = x + ((size_t)n * C + c) * HW;
    const T* dyp = dy + (size_t)n * O * HW;
    struct alignas(16) Vec { T e[VEC]; };
    float acc[OMAX] = {0.f, 0.f, 0.f, 0.f};
    for (int p0 = threadIdx.x * VEC; p0 < HW; p0 += blockDim.x * VEC) {
        Vec xv;
        *(uint4*)&xv = ldg_stream((const uint4*)(xp + p0));
#pragma unroll
        for (int o = 0; o < OMAX; o++) {
            if (o >= O) break;
            Vec dv;
            *(uint4*)&dv = *(const uint4*)(dyp + (size_t)o * HW + p0);     // re-read per channel block: stays in L2
#pragma unroll
            for (int v = 0; v < VEC; v++) acc[o] += to_acc(xv.e[v]) * to_acc(dv.e[v]);
        }
    }
    for (int o = 0; o < O; o++) {
        float s = block_sum(acc[o], red);
        if (threadIdx.x == 0) R[((size_t)n * O + o) * C + c] = s;
    }
}

// small images (HW <= 2 vectors per lane): one WARP per (c, n), 8 channels per CTA, shuffle reduction only
template <class T>
__global__ void __launch_bounds__(256) pw_corr_warp_kernel(const T* __restrict__ dy, const T* __restrict__ x, float* __restrict__ R, int C, int O, int HW) {
    constexpr int VEC = (int)(16 / sizeof(T));
    const int lane = threadIdx.x & 31, c = blockIdx.x * 8 + (threadIdx.x >> 5), n = blockIdx.y;
    if (c >= C) return;
    const T* xp = x + ((size_t)n * C + c) * HW;
    const T* dyp = dy + (size_t)n * O * HW;
    struct alignas(16) Vec { T e[VEC]; };
    float acc[OMAX] = {0.f, 0.f, 0.f, 0.f};
    for (int p0 = lane * VEC; p0 < HW; p0 += 32 * VEC) {
        Vec xv;
        *(uint4*)&xv = ldg_stream((const uint4*)(xp + p0));
#pragma unroll
        for (int o = 0; o < OMAX; o++) {
            if (o >= O) break;
            Vec dv;
            *(uint4*)&dv = *(const uint4*)(dyp + (size_t)o * HW + p0);
#pragma unroll
            for (int v = 0; v < VEC; v++) acc[o] += to_acc(xv.e[v]) * to_acc(dv.e[v]);
        }
    }
#pragma unroll
    for (int o = 0; o < OMAX; o++) {
#pragma unroll
        for (int m = 16; m > 0; m >>= 1) acc[o] += __shfl_xor_sync(0xffffffffu, acc[o], m);
        if (lane == 0 && o < O) R[((size_t)n * O + o) * C + c] = acc[o];
    }
}

// dW[o,c] += sum_n osc*isc*R ; dsum[n,c] = sum_o W*osc*R
__global__ void pw_finish_kernel(const float* __restrict__ R, const float* __restrict__ w, const float* __restrict__ isc, const float* __restrict__ osc,
                                 float* dweight, float* dsum, int N, int C, int O) {
    int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (dweight && idx < O * C) {
        int o = idx / C, c = idx - o * C;
        float s = 0.f;
        for (int n = 0; n < N; n++) s += osc[(size_t)n * O + o] * isc[(size_t)n * C + c] * R[((size_t)n * O + o) * C + c];
        dweight[idx] += s;
    }
    if (dsum && idx < N * C) {
        int n = idx / C, c = idx - n * C;
        float s = 0.f;
        for (int o = 0; o < O; o++) s += w[(size_t)o * C + c] * osc[(size_t)n * O + o] * R[((size_t)n * O + o) * C + c];
        dsum[idx] += s;
    }
}

}  // namespace

bool pw_supported(const vfm_modconv_desc& d) {
    if (d.dtype != VFM_F16 && d.dtype != VFM_F32) return false;
    if (d.kh != 1 || d.kw != 1 || d.up != 1 || d.padding != 0) return false;
    if (d.out_channels > OMAX || d.batch > 65535) return false;
    const int vec = d.dtype == VFM_F16 ? 8 : 4;
    if ((d.in_h * d.in_w) % vec != 0) return false;
    return (size_t)d.in_channels * OMAX * sizeof(float) <= 48 * 1024;
}

size_t pw_workspace_bytes(const vfm_modconv_desc& d, int direction) {
    return direction == 1 ? (size_t)d.batch * d.out_channels * d.in_channels * sizeof(float) + 512 : 0;
}

template <class T>
static int pw_forward_t(const vfm_modconv_desc& d, const void* x, const float* weight, const Coefs& k, void* y, const float* noise, int64_t noise_sn, const Epilogue& ep, cudaStream_t stream) {
    constexpr int VEC = (int)(16 / sizeof(T));
    const int HW = d.in_h * d.in_w;
    // pixel threads per CTA: halve (= double the channel slices) while the grid has fewer than 4 CTAs per SM
    int PT = 256;
    while (PT > 16 && (PT * VEC >= 2 * HW || (int64_t)ceil_div(HW, PT * VEC) * d.batch < 4 * 148)) PT >>= 1;
    const int S = 256 / PT;
    dim3 grid(ceil_div(HW, PT * VEC), d.batch);
    const size_t smem = (size_t)d.in_channels * OMAX * sizeof(float) + (S > 1 ? (size_t)256 * OMAX * VEC * sizeof(float) : 0);
    KernelTimer timer("modconv_pointwise_fwd", stream, 2.0 * d.batch * HW * (double)d.out_channels * d.in_channels,
                      (double)d.batch * HW * (d.in_channels + d.out_channels) * sizeof(T));
    if (smem > 48 * 1024) cudaFuncSetAttribute(pw_fwd_kernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    pw_fwd_kernel<T><<<grid, 256, smem, stream>>>((const T*)x, weight, k.iscale, k.oscale, noise, noise_sn, (T*)y,
                                                  d.in_channels, d.out_channels, HW, PT, ep);
    return launch_status("modconv pw_fwd_kernel");
}

int pw_stage1_forward(const vfm_modconv_desc& d, const void* x, const float* weight, const Coefs& k, void* y, const float* noise, int64_t noise_sn, const Epilogue& ep, cudaStream_t stream) {
    return d.dtype == VFM_F16 ? pw_forward_t<__half>(d, x, weight, k, y, noise, noise_sn, ep, stream) : pw_forward_t<float>(d, x, weight, k, y, noise, noise_sn, ep, stream);
}

template <class T>
static int pw_backward_t(const vfm_modconv_desc& d, const void* dy, const void* x, const float* weight, const Coefs& k, void* dx, float* dsum, float* dweight,
                         float* R, cudaStream_t stream) {
    constexpr int VEC = (int)(16 / sizeof(T));
    const int HW = d.in_h * d.in_w, N = d.batch, C = d.in_channels, O = d.out_channels;
    int st;
    if (dx) {
        dim3 grid(ceil_div(HW, 256 * VEC), N);
        KernelTimer timer("modconv_pointwise_dgrad", stream, 2.0 * N * HW * (double)O * C, (double)N * HW * (C + O) * sizeof(T));
        pw_dgrad_kernel<T><<<grid, 256, (size_t)C * OMAX * sizeof(float), stream>>>((const T*)dy, weight, k.iscale, k.oscale, (T*)dx, C, O, HW);
        st = launch_status("modconv pw_dgrad_kernel"); if (st) return st;
    }
    if (dsum || dweight) {
        {
            KernelTimer timer("modconv_pointwise_corr", stream, 2.0 * N * HW * (double)O * C, (double)N * HW * (C + O) * sizeof(T));
            if (HW <= 64 * VEC) pw_corr_warp_kernel<T><<<dim3(ceil_div(C, 8), N), 256, 0, stream>>>((const T*)dy, (const T*)x, R, C, O, HW);
            else pw_corr_kernel<T><<<dim3(C, N), 256, 0, stream>>>((const T*)dy, (const T*)x, R, C, O, HW);
            st = launch_status("modconv pw_corr_kernel"); if (st) return st;
        }
        int total = max(O * C, N * C);
        pw_finish_kernel<<<ceil_div(total, 256), 256, 0, stream>>>(R, weight, k.iscale, k.oscale, dweight, dsum, N, C, O);
        st = launch_status("modconv pw_finish_kernel"); if (st) return st;
    }
    return VFM_OK;
}

int pw_stage1_backward(const vfm_modconv_desc& d, const void* dy, const void* x, const float* weight, const Coefs& k, void* dx, float* dsum, float* dweight,
                       void* ws, size_t ws_bytes, cudaStream_t stream) {
    Carver cv(ws, ws_bytes);
    float* R = cv.take<float>((size_t)d.batch * d.out_channels * d.in_channels);
    if (!cv.ok()) { set_error("modulated_conv2d backward: pointwise workspace too small"); return VFM_ERR_WORKSPACE; }
    return d.dtype == VFM_F16 ? pw_backward_t<__half>(d, dy, x, weight, k, dx, dsum, dweight, R, stream)
                              : pw_backward_t<float>(d, dy, x, weight, k, dx, dsum, dweight, R, stream);
}

}  // namespace modconv
}  // namespace vfm
