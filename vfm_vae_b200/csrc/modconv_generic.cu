// Modulated conv2d: coefficient kernels + the generic SIMT implicit-conv path (any dtype incl. fp64, any channel count).
//
// The generic path is the always-available CUDA implementation: it covers every parameter combination the reference's
// modulated_conv2d accepts on the decoder path (k odd <= 5, up in {1,2}, demodulate on/off, both flip modes, all noise
// modes) and is what the tcgen05 path (modconv_tc.cu) is cross-checked against on the device.  Hot shapes (fp16/fp32,
// channel counts that are multiples of 64) are routed to the tensor-core path by vfm_modconv_forward/backward.
//
// Math (networks/generator.py:46-103, conv2d_resample.py:46-141), with s' = c[n]*s, W' = a[o]*W the fp16
// pre-normalised operands (a = c = 1 otherwise):
//   d[n,o] = rsqrt(sum_i wsq[o,i] * s'[n,i]^2 + 1e-8),  wsq[o,i] = a[o]^2 sum_k W[o,i,k]^2
//   y[n,o] = (d*a)[n,o] * conv(x[n,i] * s'[n,i], W) + noise
// Backward:
//   dxpre = conv^T_W((d*a) * dy);  dx = s' * dxpre;  dsum[n,i] = sum_p x * dxpre
//   M[o,i,k] = sum_{n,p} ((d*a)*dy)[n,o,p] * (s'*x)[n,i,p+k]
//   g[n,o] = sum_p dy * (y - noise) / d;  h = g * d^3
//   dW[o,i,k] = M - a[o]^2 W[o,i,k] sum_n h[n,o] s'[n,i]^2
//   ds[n,i]   = c[n] * (dsum[n,i] - s'[n,i] sum_o h[n,o] wsq[o,i])
#include "modconv_common.cuh"

namespace vfm {
namespace modconv {

// ---------------------------------------------------------------------------------------------- coefficient kernels
__global__ void wprep_kernel(const float* __restrict__ w, int O, int I, int KK, int normalize, float* a, float* wsq) {
    __shared__ float red[32];
    __shared__ float s_a;
    int o = blockIdx.x;
    const float* wo = w + (size_t)o * I * KK;
    float av = 1.f;
    if (normalize) {
        float m = 0.f;
        for (int i = threadIdx.x; i < I * KK; i += blockDim.x) m = fmaxf(m, fabsf(wo[i]));
#pragma unroll
        for (int s = 16; s > 0; s >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, s));
        if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = m;
        __syncthreads();
        if (threadIdx.x == 0) {
            float mm = 0.f;
            for (int k = 0; k < (int)(blockDim.x + 31) / 32; k++) mm = fmaxf(mm, red[k]);
            s_a = 1.0f / (sqrtf((float)(I * KK)) * mm);
        }
        __syncthreads();
        av = s_a;
    }
    if (threadIdx.x == 0) a[o] = av;
    for (int i = threadIdx.x; i < I; i += blockDim.x) {
        float s = 0.f;
        for (int k = 0; k < KK; k++) { float v = wo[i * KK + k] * av; s += v * v; }
        wsq[(size_t)o * I + i] = s;
    }
}

__global__ void sprep_kernel(const float* __restrict__ styles, int N, int I, int normalize, float* c, float* iscale) {
    __shared__ float red[32];
    __shared__ float s_c;
    int n = blockIdx.x;
    const float* sn = styles + (size_t)n * I;
    float cv = 1.f;
    if (normalize) {
        float m = 0.f;
        for (int i = threadIdx.x; i < I; i += blockDim.x) m = fmaxf(m, fabsf(sn[i]));
#pragma unroll
        for (int s = 16; s > 0; s >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, s));
        if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = m;
        __syncthreads();
        if (threadIdx.x == 0) {
            float mm = 0.f;
            for (int k = 0; k < (int)(blockDim.x + 31) / 32; k++) mm = fmaxf(mm, red[k]);
            s_c = 1.0f / mm;
        }
        __syncthreads();
        cv = s_c;
    }
    if (threadIdx.x == 0) c[n] = cv;
    for (int i = threadIdx.x; i < I; i += blockDim.x) iscale[(size_t)n * I + i] = sn[i] * cv;
}

// one warp per (n,o)
__global__ void dcoef_kernel(const float* __restrict__ wsq, const float* __restrict__ iscale, const float* __restrict__ a,
                             int N, int O, int I, int demodulate, const float* dcoefs_in, float* dcoefs_out, float* oscale) {
    int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (warp >= N * O) return;
    int n = warp / O, o = warp - n * O;
    float d = 1.f;
    if (dcoefs_in) d = dcoefs_in[warp];
    else if (demodulate) {
        float s = 0.f;
        for (int i = lane; i < I; i += 32) { float t = iscale[(size_t)n * I + i]; s += wsq[(size_t)o * I + i] * t * t; }
        s = warp_sum(s);
        d = rsqrtf(s + 1e-8f);
        // one Newton step: rsqrtf is ~2 ulp, the reference uses a correctly rounded sqrt + divide
        d = d * (1.5f - 0.5f * (s + 1e-8f) * d * d);
    }
    if (lane == 0) {
        if (dcoefs_out) dcoefs_out[warp] = d;
        oscale[warp] = d * a[o];
    }
}

int compute_coefs(const vfm_modconv_desc& d, const float* weight, const float* styles, const Coefs& k,
                  float* dcoefs_out, const float* dcoefs_in, cudaStream_t stream) {
    int normalize = (d.dtype == VFM_F16 && d.demodulate) ? 1 : 0;
    wprep_kernel<<<d.out_channels, 256, 0, stream>>>(weight, d.out_channels, d.in_channels, d.kh * d.kw, normalize, k.a, k.wsq);
    int st = launch_status("modconv wprep_kernel"); if (st) return st;
    sprep_kernel<<<d.batch, 256, 0, stream>>>(styles, d.batch, d.in_channels, normalize, k.c, k.iscale);
    st = launch_status("modconv sprep_kernel"); if (st) return st;
    int warps = d.batch * d.out_channels;
    dcoef_kernel<<<ceil_div(warps, 8), 256, 0, stream>>>(k.wsq, k.iscale, k.a, d.batch, d.out_channels, d.in_channels,
                                                       d.demodulate, dcoefs_in, dcoefs_out, k.oscale);
    return launch_status("modconv dcoef_kernel");
}

// ------------------------------------------------------------------------------------------------ generic conv kernel

constexpr int CBM = 64, CBN = 64, CBK = 8;

template <class T>
__global__ void __launch_bounds__(256) conv_generic_kernel(ConvArgs p) {
    typedef typename Acc<T>::type S;
    __shared__ S sA[CBK][CBM + 4];   // [ci][pixel]
    __shared__ S sB[CBK][CBN + 4];   // [ci][co]
    const int tid = threadIdx.x;
    const int tx = tid & 15, ty = tid >> 4;          // tx -> pixels, ty -> channels
    const int HWo = p.Hout * p.Wout;
    const int pix0 = blockIdx.x * CBM, co0 = blockIdx.y * CBN, n = blockIdx.z;
    const T* in_n = (const T*)p.in + (size_t)n * p.Cin * p.Hin * p.Win;

    // per-thread gather assignment for A: 64 pixels x 8 ci = 512 elements, 2 per thread
    const int a_pix = tid & 63, a_ci0 = tid >> 6;    // ci = a_ci0, a_ci0 + 4
    const int gp = pix0 + a_pix;
    const int a_oy = (gp < HWo) ? gp / p.Wout : 0, a_ox = (gp < HWo) ? gp - a_oy * p.Wout : 0;
    // B: 64 co x 8 ci = 512 elements, 2 per thread
    const int b_co = tid & 63, b_ci0 = tid >> 6;

    S acc[4][4];   // [channel][pixel]
#pragma unroll
    for (int i = 0; i < 4; i++)
#pragma unroll
        for (int j = 0; j < 4; j++) acc[i][j] = (S)0;

    for (int t = 0; t < p.taps.ntaps; t++) {
        int ny = a_oy * p.sn + p.taps.off_y[t], nx = a_ox * p.sn + p.taps.off_x[t];
        bool valid = (gp < HWo) && (ny % p.sd == 0) && (nx % p.sd == 0);
        int iy = ny / p.sd, ix = nx / p.sd;
        valid = valid && iy >= 0 && iy < p.Hin && ix >= 0 && ix < p.Win;
        const int wi = p.taps.widx[t];
        for (int c0 = 0; c0 < p.Cin; c0 += CBK) {
#pragma unroll
            for (int r = 0; r < 2; r++) {
                int ci = c0 + a_ci0 + r * 4;
                S v = (S)0;
                if (valid && ci < p.Cin) {
                    v = to_acc(in_n[((size_t)ci * p.Hin + iy) * p.Win + ix]);
                    if (p.in_scale) v *= (S)p.in_scale[(size_t)n * p.Cin + ci];
                }
                sA[a_ci0 + r * 4][a_pix] = v;
                int cib = c0 + b_ci0 + r * 4, co = co0 + b_co;
                S wv = (S)0;
                if (cib < p.Cin && co < p.Cout) wv = (S)p.w[(size_t)co * p.w_s_co + (size_t)cib * p.w_s_ci + wi];
                sB[b_ci0 + r * 4][b_co] = wv;
            }
            __syncthreads();
#pragma unroll
            for (int k = 0; k < CBK; k++) {
                S av[4], bv[4];
#pragma unroll
                for (int j = 0; j < 4; j++) av[j] = sA[k][tx * 4 + j];
#pragma unroll
                for (int i = 0; i < 4; i++) bv[i] = sB[k][ty * 4 + i];
#pragma unroll
                for (int i = 0; i < 4; i++)
#pragma unroll
                    for (int j = 0; j < 4; j++) acc[i][j] += bv[i] * av[j];
            }
            __syncthreads();
        }
    }

    // epilogue
#pragma unroll
    for (int i = 0; i < 4; i++) {
        int co = co0 + ty * 4 + i;
        S part = (S)0;
        if (co < p.Cout) {
            S osc = p.out_scale ? (S)p.out_scale[(size_t)n * p.Cout + co] : (S)1;
#pragma unroll
            for (int j = 0; j < 4; j++) {
                int pix = pix0 + tx * 4 + j;
                if (pix >= HWo) continue;
                size_t oidx = ((size_t)n * p.Cout + co) * HWo + pix;
                S v = acc[i][j];
                if (p.aux) part += to_acc(((const T*)p.aux)[oidx]) * v;
                v *= osc;
                if (p.add) { int oy = pix / p.Wout, ox = pix - oy * p.Wout; v += (S)p.add[(size_t)n * p.add_sn + (size_t)oy * p.add_sh + ox]; }
                ((T*)p.out)[oidx] = from_acc<T, S>(v);
            }
        }
        if (p.aux_sum) {
            // reduce over the 16 pixel-threads (tx) that share this channel: lanes tx = 0..15 within a half warp
#pragma unroll
            for (int s = 8; s > 0; s >>= 1) part += __shfl_xor_sync(0xffffffffu, part, s);
            if (tx == 0 && co < p.Cout) atomicAdd(&p.aux_sum[(size_t)n * p.Cout + co], (float)part);
        }
    }
}

template <class T>
int launch_conv(const ConvArgs& a, cudaStream_t stream) {
    dim3 grid(ceil_div(a.Hout * a.Wout, CBM), ceil_div(a.Cout, CBN), a.N);
    KernelTimer timer(a.aux_sum ? "modconv_generic_dgrad" : "modconv_generic_conv", stream,
                      2.0 * a.N * a.Hout * a.Wout * a.Cout * (double)a.Cin * a.taps.ntaps / (a.sd * a.sd),
                      ((double)a.N * a.Cin * a.Hin * a.Win + (double)a.N * a.Cout * a.Hout * a.Wout) * sizeof(T));
    conv_generic_kernel<T><<<grid, 256, 0, stream>>>(a);
    return launch_status("modconv conv_generic_kernel");
}

int run_conv(int dtype, const ConvArgs& a, cudaStream_t stream) {
    if (a.N > 65535 || ceil_div(a.Cout, CBN) > 65535) { set_error("modulated_conv2d: batch/channels too large for the generic kernel"); return VFM_ERR_INVALID; }
    switch (dtype) {
        case VFM_F16: return launch_conv<__half>(a, stream);
        case VFM_F32: return launch_conv<float>(a, stream);
        default:      return launch_conv<double>(a, stream);
    }
}

// ------------------------------------------------------------------------------------------------ generic wgrad kernel
// dW[co,ci,tap] (+)= sum_{n,p} (dy[n,co,p]*oscale[n,co]) * (x[n,ci,pos(p,tap)]*iscale[n,ci])

constexpr int WBM = 64, WBN = 64, WBK = 32;

template <class T>
__global__ void __launch_bounds__(256) wgrad_generic_kernel(WgradArgs p) {
    typedef typename Acc<T>::type S;
    __shared__ S sD[WBK][WBM + 4];   // [pixel][co]
    __shared__ S sX[WBK][WBN + 4];   // [pixel][ci]
    const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;   // tx -> ci, ty -> co
    const int co0 = blockIdx.x * WBM, ci0 = blockIdx.y * WBN;
    int z = blockIdx.z;
    const int t = z % p.taps.ntaps; z /= p.taps.ntaps;
    const int chunk = z;
    const int HWd = p.Hd * p.Wd;
    const int64_t total = (int64_t)p.N * HWd;
    const int64_t q_begin = (int64_t)chunk * p.chunk_pix, q_end = min(q_begin + p.chunk_pix, total);
    const int oy_off = p.taps.off_y[t], ox_off = p.taps.off_x[t];

    S acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; i++)
#pragma unroll
        for (int j = 0; j < 4; j++) acc[i][j] = (S)0;

    // loader assignment: 32 pixels x 64 channels = 2048 elements, 8 per thread; pixel fastest for coalescing
    const int l_pix = tid & 31, l_ch0 = tid >> 5;   // channels l_ch0 + 8*r
    for (int64_t q0 = q_begin; q0 < q_end; q0 += WBK) {
        int64_t q = q0 + l_pix;
        bool qv = q < q_end;
        int n = 0, py = 0, px = 0;
        if (qv) { n = (int)(q / HWd); int r = (int)(q - (int64_t)n * HWd); py = r / p.Wd; px = r - py * p.Wd; }
        int ny = py * p.sn + oy_off, nx = px * p.sn + ox_off;
        bool xv = qv && (ny % p.sd == 0) && (nx % p.sd == 0);
        int iy = ny / p.sd, ix = nx / p.sd;
        xv = xv && iy >= 0 && iy < p.Hx && ix >= 0 && ix < p.Wx;
#pragma unroll
        for (int r = 0; r < 8; r++) {
            int ch = l_ch0 + r * 8;
            int co = co0 + ch, ci = ci0 + ch;
            S dv = (S)0, xvv = (S)0;
            if (qv && co < p.Co) {
                dv = to_acc(((const T*)p.dy)[((size_t)n * p.Co + co) * HWd + (size_t)py * p.Wd + px]);
                if (p.oscale) dv *= (S)p.oscale[(size_t)n * p.Co + co];
            }
            if (xv && ci < p.Ci) {
                xvv = to_acc(((const T*)p.x)[(((size_t)n * p.Ci + ci) * p.Hx + iy) * p.Wx + ix]);
                if (p.iscale) xvv *= (S)p.iscale[(size_t)n * p.Ci + ci];
            }
            sD[l_pix][ch] = dv;
            sX[l_pix][ch] = xvv;
        }
        __syncthreads();
#pragma unroll 8
        for (int k = 0; k < WBK; k++) {
            S dvv[4], xv4[4];
#pragma unroll
            for (int i = 0; i < 4; i++) dvv[i] = sD[k][ty * 4 + i];
#pragma unroll
            for (int j = 0; j < 4; j++) xv4[j] = sX[k][tx * 4 + j];
#pragma unroll
            for (int i = 0; i < 4; i++)
#pragma unroll
                for (int j = 0; j < 4; j++) acc[i][j] += dvv[i] * xv4[j];
        }
        __syncthreads();
    }
    const int wi = p.taps.widx[t];
#pragma unroll
    for (int i = 0; i < 4; i++)
#pragma unroll
        for (int j = 0; j < 4; j++) {
            int co = co0 + ty * 4 + i, ci = ci0 + tx * 4 + j;
            if (co < p.Co && ci < p.Ci) atomicAdd(&p.dw[(size_t)co * p.s_co + (size_t)ci * p.s_ci + wi], (float)acc[i][j]);
        }
}

int run_wgrad(int dtype, WgradArgs a, cudaStream_t stream) {
    int64_t total = (int64_t)a.N * a.Hd * a.Wd;
    int tiles = ceil_div(a.Co, WBM) * ceil_div(a.Ci, WBN) * a.taps.ntaps;
    int want = max(1, (kNumSMs * 4) / max(tiles, 1));
    int64_t chunk_pix = ceil_div64(ceil_div64(total, want), WBK) * WBK;
    a.chunk_pix = (int)(chunk_pix < (int64_t)(1 << 30) ? chunk_pix : (int64_t)(1 << 30));
    a.chunks = (int)ceil_div64(total, a.chunk_pix);
    dim3 grid(ceil_div(a.Co, WBM), ceil_div(a.Ci, WBN), a.taps.ntaps * a.chunks);
    if (grid.z > 65535 || grid.y > 65535) { set_error("modulated_conv2d: wgrad grid too large"); return VFM_ERR_INVALID; }
    KernelTimer timer("modconv_generic_wgrad", stream, 2.0 * a.N * a.Hd * a.Wd * a.Co * (double)a.Ci * a.taps.ntaps / (a.sd * a.sd), 0.0);
    switch (dtype) {
        case VFM_F16: wgrad_generic_kernel<__half><<<grid, 256, 0, stream>>>(a); break;
        case VFM_F32: wgrad_generic_kernel<float><<<grid, 256, 0, stream>>>(a); break;
        default:      wgrad_generic_kernel<double><<<grid, 256, 0, stream>>>(a); break;
    }
    return launch_status("modconv wgrad_generic_kernel");
}

// ------------------------------------------------------------------------------------------------ reductions
// g[n,o] = sum_p dy * (y - noise) / d      (one block per (n,o) plane)
template <class T>
__global__ void __launch_bounds__(256) gsum_kernel(const T* __restrict__ dy, const T* __restrict__ y, const float* __restrict__ noise,
                                                   int64_t noise_sn, const float* __restrict__ dcoefs, int O, int HW, float* g) {
    typedef typename Acc<T>::type S;
    __shared__ S red[32];
    int plane = blockIdx.x, n = plane / O;
    const T* dyp = dy + (size_t)plane * HW;
    const T* yp = y + (size_t)plane * HW;
    const float* np_ = noise ? noise + (size_t)n * noise_sn : nullptr;
    S s = (S)0;
    for (int i = threadIdx.x; i < HW; i += blockDim.x) {
        S yv = to_acc(yp[i]);
        if (np_) yv -= (S)np_[i];
        s += to_acc(dyp[i]) * yv;
    }
    s = block_sum(s, red);
    if (threadIdx.x == 0) g[plane] = (float)(s / (S)dcoefs[plane]);
}

// dnoise: HW mode: [H,W] += sum_{n,o} dy ; N1HW mode: [N,1,H,W] += sum_o dy.  grid (pix/256, splits, N or 1)
template <class T>
__global__ void __launch_bounds__(256) dnoise_kernel(const T* __restrict__ dy, int N, int O, int HW, int per_sample, int planes_per_split, float* dnoise) {
    typedef typename Acc<T>::type S;
    int pix = blockIdx.x * blockDim.x + threadIdx.x;
    if (pix >= HW) return;
    int total_planes = per_sample ? O : N * O;
    int p0 = blockIdx.y * planes_per_split, p1 = min(p0 + planes_per_split, total_planes);
    const T* base = dy + (per_sample ? (size_t)blockIdx.z * O * HW : 0);
    S s = (S)0;
    for (int pl = p0; pl < p1; pl++) s += to_acc(base[(size_t)pl * HW + pix]);
    atomicAdd(&dnoise[(per_sample ? (size_t)blockIdx.z * HW : 0) + pix], (float)s);
}

// Both reductions of the backward that read the whole dy in ONE pass (16-byte loads, 8 vectors in flight per thread):
//   g[n,o] += (1/d[n,o]) * sum_p dy*(y - noise)          dnoise[(n,)p] += sum_o dy
// A CTA owns one sample and a chunk of 1024 pixels and walks over the channels: a thread keeps its 8 pixels' dnoise sums
// in registers, the per-channel partials of g are reduced by warp shuffles into a shared-memory array and leave as
// one atomic per (CTA, channel).  Requires HW % 8 == 0 and 16-byte aligned tensors; otherwise the two simple kernels above run.
template <class T>
__global__ void __launch_bounds__(128) gsum_dnoise_kernel(const T* __restrict__ dy, const T* __restrict__ y, const float* __restrict__ noise, int64_t noise_sn,
                                                          const float* __restrict__ dcoefs, int O, int HW, float* g, float* dnoise, int64_t dnoise_sn, int o_per_slice) {
    constexpr int V = 16 / (int)sizeof(T);              // elements per vector: a thread covers 8 pixels = 8/V vectors
    constexpr int NV = 8 / V;
    extern __shared__ float s_g[];                       // [O]
    const int n = blockIdx.y;
    const int p0 = blockIdx.x * 1024 + threadIdx.x * 8;
    const bool live = p0 < HW;
    const int o_begin = blockIdx.z * o_per_slice, o_end = min(O, o_begin + o_per_slice);     // this CTA's channel slice (a multiple of 8 channels)
    for (int o = threadIdx.x; o < O; o += blockDim.x) s_g[o] = 0.f;
    __syncthreads();
    float nz[8], dn[8];
#pragma unroll
    for (int i = 0; i < 8; i++) { nz[i] = (noise && live) ? noise[(size_t)n * noise_sn + p0 + i] : 0.f; dn[i] = 0.f; }
    const T* dyp = dy + (size_t)n * O * HW + p0;
    const T* yp = y ? y + (size_t)n * O * HW + p0 : nullptr;
    const int lane = threadIdx.x & 31;
    for (int o0 = o_begin; o0 < o_end; o0 += 4) {
        uint4 a[4][NV], b[4][NV];
#pragma unroll
        for (int k = 0; k < 4; k++)
#pragma unroll
            for (int q = 0; q < NV; q++) {
                const bool ok = live && o0 + k < o_end;
                a[k][q] = ok ? ldg_stream(dyp + (size_t)(o0 + k) * HW + q * V) : make_uint4(0, 0, 0, 0);
                b[k][q] = (ok && yp) ? ldg_stream(yp + (size_t)(o0 + k) * HW + q * V) : make_uint4(0, 0, 0, 0);
            }
#pragma unroll
        for (int k = 0; k < 4; k++) {
            float part = 0.f;
#pragma unroll
            for (int q = 0; q < NV; q++) {
                const T* av = (const T*)&a[k][q];
                const T* bv = (const T*)&b[k][q];
#pragma unroll
                for (int e = 0; e < V; e++) {
                    const float d = to_acc(av[e]);
                    dn[q * V + e] += d;
                    part = fmaf(d, to_acc(bv[e]) - nz[q * V + e], part);
                }
            }
            if (g) {
                part = warp_sum(part);
                if (lane == 0 && o0 + k < o_end) atomicAdd(&s_g[o0 + k], part);
            }
        }
    }
    if (dnoise && live) {
#pragma unroll
        for (int i = 0; i < 8; i++) atomicAdd(&dnoise[(size_t)n * dnoise_sn + p0 + i], dn[i]);
    }
    if (g) {
        __syncthreads();
        for (int o = o_begin + threadIdx.x; o < o_end; o += blockDim.x) atomicAdd(&g[(size_t)n * O + o], s_g[o] / dcoefs[(size_t)n * O + o]);
    }
}

// Training with the fused layer epilogue (y_act = clamp(gain * lrelu(conv * d + noise + b), +-clamp) written by the forward, nothing else
// kept): the backward of bias_act (torch_utils/ops/bias_act.py:158-179) folded into the same single pass over the incoming gradient:
//   dz       = dy_act * gain * (y_act > 0 ? 1 : alpha) * [|y_act| < clamp]           -> written out (the gradient w.r.t. the conv output)
//   g[n,o]  += (1/d) * sum_p dz * (y_act / (gain * slope) - b[o] - noise)             (the pre-activation value recovered from y_act; where
//                                                                                       the clamp saturated, dz = 0 and the term vanishes)
//   gz[n,o] += sum_p dz   (bias gradient per sample)          dnoise[(n,)p] += sum_o dz
// instead of bias_act_grad (2 reads + 1 write) followed by gsum_dnoise (2 reads): 2 reads + 1 write in total.
template <class T>
__global__ void __launch_bounds__(128) act_grad_gsum_dnoise_kernel(const T* __restrict__ dy, const T* __restrict__ y, const T* __restrict__ bias,
                                                                   const float* __restrict__ noise, int64_t noise_sn, const float* __restrict__ dcoefs, int O, int HW,
                                                                   float gain, float alpha, float clamp, int act, T* __restrict__ dz, float* g, float* gz,
                                                                   float* dnoise, int64_t dnoise_sn, int o_per_slice) {
    constexpr int V = 16 / (int)sizeof(T);
    constexpr int NV = 8 / V;
    extern __shared__ float s_red[];                     // [2][O]: g and gz partials of this CTA
    float* s_g = s_red;
    float* s_z = s_red + O;
    const int n = blockIdx.y;
    const int p0 = blockIdx.x * 1024 + threadIdx.x * 8;
    const bool live = p0 < HW;
    const int o_begin = blockIdx.z * o_per_slice, o_end = min(O, o_begin + o_per_slice);     // this CTA's channel slice (a multiple of 8 channels)
    for (int o = threadIdx.x; o < 2 * O; o += blockDim.x) s_red[o] = 0.f;
    __syncthreads();
    float nz[8], dn[8];
#pragma unroll
    for (int i = 0; i < 8; i++) { nz[i] = (noise && live) ? noise[(size_t)n * noise_sn + p0 + i] : 0.f; dn[i] = 0.f; }
    const size_t base = (size_t)n * O * HW + p0;
    const int lane = threadIdx.x & 31;
    const float g_pos = gain, g_neg = (act == 3) ? gain * alpha : gain;
    const float r_pos = 1.f / g_pos, r_neg = 1.f / g_neg;
    // software pipeline over groups of 4 channels: the loads of group k+1 are issued before group k is consumed
    uint4 a[2][4][NV], b[2][4][NV];
    auto fetch = [&](int o0, int slot) {
#pragma unroll
        for (int k = 0; k < 4; k++)
#pragma unroll
            for (int q = 0; q < NV; q++) {
                const bool ok = live && o0 + k < o_end;
                a[slot][k][q] = ok ? ldg_stream(dy + base + (size_t)(o0 + k) * HW + q * V) : make_uint4(0, 0, 0, 0);
                b[slot][k][q] = ok ? ldg_stream(y + base + (size_t)(o0 + k) * HW + q * V) : make_uint4(0, 0, 0, 0);
            }
    };
    auto consume = [&](int o0, int slot) {
        float red[8];                                        // [part of channel 0..3 | zsum of channel 0..3]
#pragma unroll
        for (int k = 0; k < 4; k++) {
            const float bo = (bias && o0 + k < o_end) ? to_acc(bias[o0 + k]) : 0.f;
            float part = 0.f, zsum = 0.f;
#pragma unroll
            for (int q = 0; q < NV; q++) {
                const T* av = (const T*)&a[slot][k][q];
                const T* bv = (const T*)&b[slot][k][q];
                struct alignas(16) { T e[V]; } out;
#pragma unroll
                for (int e = 0; e < V; e++) {
                    const float ya = to_acc(bv[e]);
                    const bool pos = ya > 0.f;
                    float d = to_acc(av[e]) * (pos ? g_pos : g_neg);
                    if (clamp >= 0.f && !(fabsf(ya) < clamp)) d = 0.f;
                    const T dq = from_acc<T, float>(d);
                    out.e[e] = dq;
                    d = to_acc(dq);                              // downstream kernels see the rounded value: keep the reductions consistent with it
                    dn[q * V + e] += d;
                    zsum += d;
                    part = fmaf(d, ya * (pos ? r_pos : r_neg) - bo - nz[q * V + e], part);
                }
                if (live && o0 + k < o_end) stg_stream(dz + base + (size_t)(o0 + k) * HW + q * V, *(const uint4*)&out);
            }
            red[k] = part; red[4 + k] = zsum;
        }
        // 8 sums over the warp with 9 shuffles instead of 40: halve the value set while halving the lane set (bits 16, 8, 4), then two
        // plain butterfly steps; lane L ends up with the total of value index ((L>>4)&1)*4 + ((L>>3)&1)*2 + ((L>>2)&1)
        float v4[4], v2[2], v1;
        {
            const bool hi = lane & 16;
#pragma unroll
            for (int i = 0; i < 4; i++) { const float send = hi ? red[i] : red[4 + i]; const float keep = hi ? red[4 + i] : red[i]; v4[i] = keep + __shfl_xor_sync(0xffffffffu, send, 16); }
        }
        {
            const bool hi = lane & 8;
#pragma unroll
            for (int i = 0; i < 2; i++) { const float send = hi ? v4[i] : v4[2 + i]; const float keep = hi ? v4[2 + i] : v4[i]; v2[i] = keep + __shfl_xor_sync(0xffffffffu, send, 8); }
        }
        {
            const bool hi = lane & 4;
            const float send = hi ? v2[0] : v2[1]; const float keep = hi ? v2[1] : v2[0];
            v1 = keep + __shfl_xor_sync(0xffffffffu, send, 4);
        }
        v1 += __shfl_xor_sync(0xffffffffu, v1, 2);
        v1 += __shfl_xor_sync(0xffffffffu, v1, 1);
        if ((lane & 3) == 0) {
            const int k = ((lane >> 3) & 1) * 2 + ((lane >> 2) & 1);
            if (o0 + k < o_end) atomicAdd(((lane & 16) ? s_z : s_g) + o0 + k, v1);
        }
    };
    fetch(o_begin, 0);
    for (int o0 = o_begin; o0 < o_end; o0 += 8) {
        if (o0 + 4 < o_end) fetch(o0 + 4, 1);
        consume(o0, 0);
        if (o0 + 8 < o_end) fetch(o0 + 8, 0);
        if (o0 + 4 < o_end) consume(o0 + 4, 1);
    }
    if (dnoise && live) {
#pragma unroll
        for (int i = 0; i < 8; i++) atomicAdd(&dnoise[(size_t)n * dnoise_sn + p0 + i], dn[i]);
    }
    __syncthreads();
    for (int o = o_begin + threadIdx.x; o < o_end; o += blockDim.x) {
        if (g) atomicAdd(&g[(size_t)n * O + o], s_g[o] / dcoefs[(size_t)n * O + o]);
        if (gz) atomicAdd(&gz[(size_t)n * O + o], s_z[o]);
    }
}

// dW[o,i,k] = M[o,i,k] - a[o]^2 W[o,i,k] sum_n h[n,o] s'[n,i]^2,  h = g d^3          (thread per (o,i))
__global__ void dw_fix_kernel(float* dw, const float* __restrict__ w, const float* __restrict__ a, const float* __restrict__ g,
                              const float* __restrict__ dcoefs, const float* __restrict__ iscale, int N, int O, int I, int KK) {
    int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= O * I) return;
    int o = idx / I, i = idx - o * I;
    float t = 0.f;
    for (int n = 0; n < N; n++) {
        float d = dcoefs[n * O + o], s = iscale[n * I + i];
        t += g[n * O + o] * d * d * d * s * s;
    }
    float aa = a[o] * a[o];
    for (int k = 0; k < KK; k++) dw[(size_t)idx * KK + k] -= aa * w[(size_t)idx * KK + k] * t;
}

// ds[n,i] = c[n] * (dsum[n,i] - s'[n,i] sum_o h[n,o] wsq[o,i])
// CTA = 32 input channels of one sample x 8 slices of the output channels (summed through shared memory): a thread per (n,i) walking
// all O channels alone was a chain of O dependent-latency rounds on 128 CTAs (46 us per call in the ncu launch list).
__global__ void __launch_bounds__(256) ds_fix_kernel(float* ds, const float* __restrict__ dsum, const float* __restrict__ c, const float* __restrict__ g,
                                                     const float* __restrict__ dcoefs, const float* __restrict__ iscale, const float* __restrict__ wsq,
                                                     int N, int O, int I, int demodulate) {
    __shared__ float part[8][32];
    const int il = threadIdx.x & 31, sl = threadIdx.x >> 5;
    const int bpn = (I + 31) / 32;
    const int n = blockIdx.x / bpn, i = (blockIdx.x - n * bpn) * 32 + il;
    float t = 0.f;
    if (demodulate && i < I) {
#pragma unroll 4
        for (int o = sl; o < O; o += 8) { const float d = dcoefs[n * O + o]; t += g[n * O + o] * d * d * d * wsq[(size_t)o * I + i]; }
    }
    part[sl][il] = t;
    __syncthreads();
    if (sl == 0 && i < I) {
        t = 0.f;
#pragma unroll
        for (int k = 0; k < 8; k++) t += part[k][il];
        const int idx = n * I + i;
        ds[idx] = c[n] * (dsum[idx] - iscale[idx] * t);
    }
}

template <class T>
int launch_gsum(const void* dy, const void* y, const float* noise, int64_t noise_sn, const float* dcoefs, int planes, int O, int HW, float* g, cudaStream_t stream) {
    gsum_kernel<T><<<planes, 256, 0, stream>>>((const T*)dy, (const T*)y, noise, noise_sn, dcoefs, O, HW, g);
    return launch_status("modconv gsum_kernel");
}
int run_gsum(int dtype, const void* dy, const void* y, const float* noise, int64_t noise_sn, const float* dcoefs, int planes, int O, int HW, float* g, cudaStream_t stream) {
    switch (dtype) {
        case VFM_F16: return launch_gsum<__half>(dy, y, noise, noise_sn, dcoefs, planes, O, HW, g, stream);
        case VFM_F32: return launch_gsum<float>(dy, y, noise, noise_sn, dcoefs, planes, O, HW, g, stream);
        default:      return launch_gsum<double>(dy, y, noise, noise_sn, dcoefs, planes, O, HW, g, stream);
    }
}
int run_dnoise(int dtype, const void* dy, int N, int O, int HW, int per_sample, float* dnoise, cudaStream_t stream) {
    int total_planes = per_sample ? O : N * O;
    int splits = min(total_planes, max(1, (kNumSMs * 8) / max(1, ceil_div(HW, 256))));
    int pps = ceil_div(total_planes, splits);
    splits = ceil_div(total_planes, pps);
    dim3 grid(ceil_div(HW, 256), splits, per_sample ? N : 1);
    switch (dtype) {
        case VFM_F16: dnoise_kernel<__half><<<grid, 256, 0, stream>>>((const __half*)dy, N, O, HW, per_sample, pps, dnoise); break;
        case VFM_F32: dnoise_kernel<float><<<grid, 256, 0, stream>>>((const float*)dy, N, O, HW, per_sample, pps, dnoise); break;
        default:      dnoise_kernel<double><<<grid, 256, 0, stream>>>((const double*)dy, N, O, HW, per_sample, pps, dnoise); break;
    }
    return launch_status("modconv dnoise_kernel");
}
// g (zero-initialised by this function) and/or dnoise (zero-initialised by the caller) in one pass over dy; returns
// VFM_ERR_NO_KERNEL when the tensors do not allow the vector kernel (the caller then uses run_gsum / run_dnoise)
// channel slices (blockIdx.z) of the one-pass reduction kernels: the grid of (1024-pixel chunk, sample) CTAs of 4 warps is too small to
// keep HBM busy for the 64x64 / 128x128 layers (256 / 1024 CTAs: 2.8 TB/s in ncu at 512 channels, 64x64), so the channels are cut
// into slices of a multiple of 8 until there are >= 16 warps per SM; a slice adds its own partial dnoise sums (atomics, as before)
static int reduce_slices(int O, int HW, int N, int& o_per_slice) {
    int slices = 1;
    const long long ctas = (long long)ceil_div(HW, 1024) * N;
    while (slices < 8 && ctas * slices * 4 < (long long)kNumSMs * 16 && O / (slices * 2) >= 32) slices *= 2;
    o_per_slice = ceil_div(ceil_div(O, slices), 8) * 8;
    return ceil_div(O, o_per_slice);
}
int run_gsum_dnoise(int dtype, const void* dy, const void* y, const float* noise, int64_t noise_sn, const float* dcoefs, int N, int O, int HW,
                    float* g, float* dnoise, int dnoise_per_sample, cudaStream_t stream) {
    if ((dtype != VFM_F16 && dtype != VFM_F32) || HW % 8 != 0 || HW < 2048 || !aligned16(dy) || (y && !aligned16(y)) || (size_t)O * sizeof(float) > 48 * 1024) return VFM_ERR_NO_KERNEL;
    if (g) VFM_CUDA_OK(cudaMemsetAsync(g, 0, sizeof(float) * (size_t)N * O, stream));
    int ops = O;
    dim3 grid(ceil_div(HW, 1024), N, reduce_slices(O, HW, N, ops));
    const double es = dtype == VFM_F16 ? 2.0 : 4.0;
    KernelTimer timer("modconv_gsum_dnoise", stream, 0.0, (double)N * O * HW * es * (g ? 2.0 : 1.0), "o%dhw%d", O, HW);
    const int64_t dsn = dnoise_per_sample ? HW : 0;
    if (dtype == VFM_F16)
        gsum_dnoise_kernel<__half><<<grid, 128, O * sizeof(float), stream>>>((const __half*)dy, g ? (const __half*)y : nullptr, g ? noise : nullptr, noise_sn, dcoefs, O, HW, g, dnoise, dsn, ops);
    else
        gsum_dnoise_kernel<float><<<grid, 128, O * sizeof(float), stream>>>((const float*)dy, g ? (const float*)y : nullptr, g ? noise : nullptr, noise_sn, dcoefs, O, HW, g, dnoise, dsn, ops);
    return launch_status("modconv gsum_dnoise_kernel");
}
// fused backward of the layer epilogue (see act_grad_gsum_dnoise_kernel); g / gz are zero-initialised here, dnoise by the caller
bool act_grad_fused_supported(int dtype, int O, int HW, const void* dy, const void* y, const void* dz) {
    return (dtype == VFM_F16 || dtype == VFM_F32) && HW % 8 == 0 && HW >= 2048 && aligned16(dy) && aligned16(y) && aligned16(dz) && (size_t)O * 2 * sizeof(float) <= 48 * 1024;
}
int run_act_grad_gsum_dnoise(int dtype, const void* dy, const void* y, const void* bias, const float* noise, int64_t noise_sn, const float* dcoefs, int N, int O, int HW,
                             float gain, float alpha, float clamp, int act, void* dz, float* g, float* gz, float* dnoise, int dnoise_per_sample, cudaStream_t stream) {
    if (!act_grad_fused_supported(dtype, O, HW, dy, y, dz)) { set_error("modulated_conv2d backward: the fused epilogue gradient needs >= 2048 output pixels (a multiple of 8) and 16-byte aligned tensors"); return VFM_ERR_NO_KERNEL; }
    if (g) VFM_CUDA_OK(cudaMemsetAsync(g, 0, sizeof(float) * (size_t)N * O, stream));
    if (gz) VFM_CUDA_OK(cudaMemsetAsync(gz, 0, sizeof(float) * (size_t)N * O, stream));
    int ops = O;
    dim3 grid(ceil_div(HW, 1024), N, reduce_slices(O, HW, N, ops));
    const double es = dtype == VFM_F16 ? 2.0 : 4.0;
    KernelTimer timer("modconv_act_grad_gsum", stream, 0.0, (double)N * O * HW * es * 3.0, "o%dhw%d", O, HW);
    const int64_t dsn = dnoise_per_sample ? HW : 0;
    const size_t smem = (size_t)O * 2 * sizeof(float);
    if (dtype == VFM_F16)
        act_grad_gsum_dnoise_kernel<__half><<<grid, 128, smem, stream>>>((const __half*)dy, (const __half*)y, (const __half*)bias, noise, noise_sn, dcoefs, O, HW, gain, alpha, clamp, act,
                                                                         (__half*)dz, g, gz, dnoise, dsn, ops);
    else
        act_grad_gsum_dnoise_kernel<float><<<grid, 128, smem, stream>>>((const float*)dy, (const float*)y, (const float*)bias, noise, noise_sn, dcoefs, O, HW, gain, alpha, clamp, act,
                                                                        (float*)dz, g, gz, dnoise, dsn, ops);
    return launch_status("modconv act_grad_gsum_dnoise_kernel");
}
int run_dw_fix(float* dw, const float* w, const float* a, const float* g, const float* dcoefs, const float* iscale, int N, int O, int I, int KK, cudaStream_t stream) {
    dw_fix_kernel<<<ceil_div(O * I, 256), 256, 0, stream>>>(dw, w, a, g, dcoefs, iscale, N, O, I, KK);
    return launch_status("modconv dw_fix_kernel");
}
int run_ds_fix(float* ds, const float* dsum, const float* c, const float* g, const float* dcoefs, const float* iscale, const float* wsq, int N, int O, int I, int demod, cudaStream_t stream) {
    ds_fix_kernel<<<N * ceil_div(I, 32), 256, 0, stream>>>(ds, dsum, c, g, dcoefs, iscale, wsq, N, O, I, demod);
    return launch_status("modconv ds_fix_kernel");
}

}  // namespace modconv
}  // namespace vfm
