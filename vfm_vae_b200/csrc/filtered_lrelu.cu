// filtered_lrelu for sm_100a:  y = down_fd( clamp( lrelu( up_fu(x + b) * up^2 * gain ) ) )  in ONE kernel.
//
// A CTA owns one output tile of one (n,c) plane and keeps every intermediate in shared memory:
//   input halo tile (+bias, zero outside the image)  ->  horizontal up-FIR  ->  vertical up-FIR + gain/lrelu/clamp
//   (+ 2-bit sign write or read)  ->  horizontal down-FIR  ->  vertical down-FIR  ->  global store.
// Only x is read from and y (+ the packed sign tensor) written to HBM.  Separable (1-D) and full (2-D) filters are
// both handled for fu and fd independently; the polyphase structure (only taps that hit non-zero samples of the
// zero-inserted signal) is exploited in both up passes.  Filters travel as kernel-visible global pointers and are
// staged to shared memory per CTA -- there is no __constant__/global scratch state, so the op is stream-safe (the
// reference is not: torch_utils/ops/filtered_lrelu.cu:77-78, filtered_lrelu.py:215-216).
//
// Semantics: torch_utils/ops/filtered_lrelu.py:121-153 (_filtered_lrelu_ref) via the oracle; sign tensor layout and
// read/write rules follow torch_utils/ops/filtered_lrelu.cpp:80-131 and filtered_lrelu.cu:484-579,1105-1205.
#include "common.cuh"

namespace vfm {
namespace {

constexpr int kMaxTaps = 32;

struct FlrArgs {
    const void* x; void* y; const void* b; uint8_t* s; const float* fu; const float* fd;
    int up, down;
    int fu_w, fu_h, fd_w, fd_h;          // *_h == 0 -> separable
    int64_t fu_sw, fu_sh, fd_sw, fd_sh;
    int pad_x0, pad_y0;
    float gain, slope, clamp;
    int flip;
    int x_w, x_h, channels, batch;
    int64_t xsw, xsh, xsc, xsn;
    int y_w, y_h;
    int64_t ysw, ysh, ysc, ysn;
    int64_t b_stride;
    int s_w_bytes, s_h, s_ofs_x, s_ofs_y, s_w_active;
    // tiling (host-computed)
    int tow, toh;            // output tile
    int tuw, tuh;            // upsampled tile (tuw multiple of 4)
    int tiw, tih;            // input tile
    int tiles_x, tiles_y;
};

// SIGN: 0 = none, 1 = write, 2 = read
template <class T, int SIGN>
__global__ void __launch_bounds__(256) filtered_lrelu_kernel(FlrArgs p) {
    extern __shared__ __align__(16) float smem[];
    const int fuh = p.fu_h ? p.fu_h : p.fu_w, fdh = p.fd_h ? p.fd_h : p.fd_w;   // vertical tap counts
    const bool fu_sep = (p.fu_h == 0), fd_sep = (p.fd_h == 0);
    // shared memory carve-up
    float* s_fu = smem;                                    // separable: [fu_w]; full: [fuh*fu_w]
    float* s_fd = s_fu + (fu_sep ? p.fu_w : fuh * p.fu_w);
    float* s_in = s_fd + (fd_sep ? p.fd_w : fdh * p.fd_w);
    s_in = (float*)(((uintptr_t)s_in + 15) & ~(uintptr_t)15);
    float* s_ux = s_in + p.tih * p.tiw;                    // [tih][tuw]   (separable fu only)
    float* s_u = s_ux + (fu_sep ? p.tih * p.tuw : 0);      // [tuh][tuw]
    float* s_dx = s_u + p.tuh * p.tuw;                     // [tuh][tow]   (separable fd only)

    const int tid = threadIdx.x, nthr = blockDim.x;
    int64_t bid = blockIdx.x;
    const int tile_x = (int)(bid % p.tiles_x); bid /= p.tiles_x;
    const int tile_y = (int)(bid % p.tiles_y); bid /= p.tiles_y;
    const int c = (int)(bid % p.channels), n = (int)(bid / p.channels);
    const int64_t plane = (int64_t)n * p.channels + c;

    // ---- taps as correlation taps (flip == 0 means true convolution -> reverse) ----
    if (fu_sep) { for (int i = tid; i < p.fu_w; i += nthr) s_fu[i] = p.fu[(p.flip ? i : p.fu_w - 1 - i) * p.fu_sw]; }
    else {
        for (int i = tid; i < fuh * p.fu_w; i += nthr) {
            int ty = i / p.fu_w, tx = i - ty * p.fu_w;
            s_fu[i] = p.fu[(p.flip ? ty : fuh - 1 - ty) * p.fu_sh + (p.flip ? tx : p.fu_w - 1 - tx) * p.fu_sw];
        }
    }
    if (fd_sep) { for (int i = tid; i < p.fd_w; i += nthr) s_fd[i] = p.fd[(p.flip ? i : p.fd_w - 1 - i) * p.fd_sw]; }
    else {
        for (int i = tid; i < fdh * p.fd_w; i += nthr) {
            int ty = i / p.fd_w, tx = i - ty * p.fd_w;
            s_fd[i] = p.fd[(p.flip ? ty : fdh - 1 - ty) * p.fd_sh + (p.flip ? tx : p.fd_w - 1 - tx) * p.fd_sw];
        }
    }

    // ---- tile geometry ----
    const int ox0 = tile_x * p.tow, oy0 = tile_y * p.toh;       // first output of the tile
    const int ux0 = ox0 * p.down, uy0 = oy0 * p.down;           // first intermediate sample of the tile
    const int ix0 = floor_div(ux0 - p.pad_x0, p.up);            // first input sample (may be negative)
    const int iy0 = floor_div(uy0 - p.pad_y0, p.up);
    // note: intermediate u[ux] = sum_t g[t] * xup[ux - pad0 + t], xup[j] = x[j/up] if j % up == 0

    // ---- stage 1: input tile + bias ----
    {
        const float bias = to_acc(*(const T*)((const char*)p.b + (int64_t)c * p.b_stride * (int64_t)sizeof(T)));
        const T* xp = (const T*)p.x + (int64_t)n * p.xsn + (int64_t)c * p.xsc;
        for (int i = tid; i < p.tih * p.tiw; i += nthr) {
            int ry = i / p.tiw, rx = i - ry * p.tiw;
            int ix = ix0 + rx, iy = iy0 + ry;
            float v = 0.f;
            if (ix >= 0 && ix < p.x_w && iy >= 0 && iy < p.x_h) v = to_acc(xp[(int64_t)iy * p.xsh + (int64_t)ix * p.xsw]) + bias;
            s_in[i] = v;
        }
    }
    __syncthreads();

    const float act_gain = (float)(p.up * p.up) * p.gain;

    // ---- activation + signs for one intermediate sample ----
    auto activate = [&](float v, int rux, int ruy, uint32_t& code) -> float {
        v *= act_gain;
        code = 0;
        if (SIGN == 2) {
            int sx = ux0 + rux + p.s_ofs_x, sy = uy0 + ruy + p.s_ofs_y;
            if (sx >= 0 && sx < p.s_w_active && sy >= 0 && sy < p.s_h) {
                uint32_t sb = p.s[(plane * p.s_h + sy) * (int64_t)p.s_w_bytes + (sx >> 2)];
                sb >>= (sx & 3) << 1;
                if (sb & 1) v *= p.slope;
                if (sb & 2) v = 0.f;
            }
        } else {
            if (v < 0.f) { v *= p.slope; code = 1; }
            if (fabsf(v) > p.clamp) { v = copysignf(p.clamp, v); code = 2; }
        }
        return v;
    };

    // ---- stage 2: up-FIR ----
    if (fu_sep) {
        // horizontal: s_ux[ry][rux] for every input row of the tile
        for (int i = tid; i < p.tih * p.tuw; i += nthr) {
            int ry = i / p.tuw, rux = i - ry * p.tuw;
            int j0 = ux0 + rux - p.pad_x0;                       // zero-inserted coordinate of tap 0
            int t0 = ((-j0) % p.up + p.up) % p.up;
            int rx = (j0 + t0) / p.up - ix0;                     // exact division
            const float* row = s_in + ry * p.tiw;
            float acc = 0.f;
            for (int t = t0; t < p.fu_w; t += p.up, rx++) acc += s_fu[t] * row[rx];
            s_ux[i] = acc;
        }
        __syncthreads();
    }
    // vertical (separable) or full 2-D, then activation.  Threads own groups of 4 consecutive columns so that the
    // 2-bit sign codes of one byte are produced by a single thread.
    {
        const int groups_x = p.tuw >> 2;
        for (int i = tid; i < p.tuh * groups_x; i += nthr) {
            int ruy = i / groups_x, gx = i - ruy * groups_x;
            int jy0 = uy0 + ruy - p.pad_y0;
            int ty0 = ((-jy0) % p.up + p.up) % p.up;
            int ry0 = (jy0 + ty0) / p.up - iy0;
            uint32_t packed = 0;
#pragma unroll
            for (int k = 0; k < 4; k++) {
                int rux = gx * 4 + k;
                float acc = 0.f;
                if (fu_sep) {
                    int ry = ry0;
                    for (int t = ty0; t < fuh; t += p.up, ry++) acc += s_fu[t] * s_ux[ry * p.tuw + rux];
                } else {
                    int j0 = ux0 + rux - p.pad_x0;
                    int t0 = ((-j0) % p.up + p.up) % p.up;
                    int rx0 = (j0 + t0) / p.up - ix0;
                    int ry = ry0;
                    for (int ty = ty0; ty < fuh; ty += p.up, ry++) {
                        int rx = rx0;
                        for (int tx = t0; tx < p.fu_w; tx += p.up, rx++) acc += s_fu[ty * p.fu_w + tx] * s_in[ry * p.tiw + rx];
                    }
                }
                uint32_t code;
                s_u[ruy * p.tuw + rux] = activate(acc, rux, ruy, code);
                packed |= code << (2 * k);
            }
            if (SIGN == 1) {
                int sx = ux0 + gx * 4, sy = uy0 + ruy;           // write mode: sign offset is zero
                if (sx < p.s_w_active && sy < p.s_h) p.s[(plane * p.s_h + sy) * (int64_t)p.s_w_bytes + (sx >> 2)] = (uint8_t)packed;
            }
        }
    }
    __syncthreads();

    // ---- stage 3: down-FIR ----
    T* yp = (T*)p.y + (int64_t)n * p.ysn + (int64_t)c * p.ysc;
    if (fd_sep) {
        for (int i = tid; i < p.tuh * p.tow; i += nthr) {
            int ruy = i / p.tow, rox = i - ruy * p.tow;
            const float* row = s_u + ruy * p.tuw + rox * p.down;
            float acc = 0.f;
            for (int t = 0; t < p.fd_w; t++) acc += s_fd[t] * row[t];
            s_dx[i] = acc;
        }
        __syncthreads();
        for (int i = tid; i < p.toh * p.tow; i += nthr) {
            int roy = i / p.tow, rox = i - roy * p.tow;
            int ox = ox0 + rox, oy = oy0 + roy;
            if (ox >= p.y_w || oy >= p.y_h) continue;
            const float* col = s_dx + (roy * p.down) * p.tow + rox;
            float acc = 0.f;
            for (int t = 0; t < fdh; t++) acc += s_fd[t] * col[t * p.tow];
            yp[(int64_t)oy * p.ysh + (int64_t)ox * p.ysw] = from_acc<T, float>(acc);
        }
    } else {
        for (int i = tid; i < p.toh * p.tow; i += nthr) {
            int roy = i / p.tow, rox = i - roy * p.tow;
            int ox = ox0 + rox, oy = oy0 + roy;
            if (ox >= p.y_w || oy >= p.y_h) continue;
            const float* base = s_u + (roy * p.down) * p.tuw + rox * p.down;
            float acc = 0.f;
            for (int ty = 0; ty < fdh; ty++)
                for (int tx = 0; tx < p.fd_w; tx++) acc += s_fd[ty * p.fd_w + tx] * base[ty * p.tuw + tx];
            yp[(int64_t)oy * p.ysh + (int64_t)ox * p.ysw] = from_acc<T, float>(acc);
        }
    }
}

// ------------------------------------------------------------------------------------------------------------
struct FlrActArgs {
    void* x; uint8_t* s;
    double gain, slope, clamp;
    int x_w, x_h, channels, batch;
    int64_t xsw, xsh, xsc, xsn;
    int s_w, s_h, s_ofs_x, s_ofs_y;
    int groups_x;    // number of 4-element groups per row covered by the launch
    int rows;        // rows covered by the launch
};

// One thread per group of 4 consecutive elements of a row, so each sign byte has a single writer.
template <class T, int SIGN>
__global__ void __launch_bounds__(256) filtered_lrelu_act_kernel(FlrActArgs p, int64_t total) {
    typedef typename Acc<T>::type S;
    for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (int64_t)gridDim.x * blockDim.x) {
        int gx = (int)(idx % p.groups_x);
        int64_t r = idx / p.groups_x;
        int y = (int)(r % p.rows);
        int64_t plane = r / p.rows;
        int c = (int)(plane % p.channels), n = (int)(plane / p.channels);
        T* row = (T*)p.x + (int64_t)n * p.xsn + (int64_t)c * p.xsc + (int64_t)y * p.xsh;
        uint32_t packed = 0;
#pragma unroll
        for (int k = 0; k < 4; k++) {
            int x = gx * 4 + k;
            if (x >= p.x_w || y >= p.x_h) continue;
            S v = to_acc(row[(int64_t)x * p.xsw]) * (S)p.gain;
            if (SIGN == 2) {
                int sx = x + p.s_ofs_x, sy = y + p.s_ofs_y;
                if (sx >= 0 && sx < p.s_w && sy >= 0 && sy < p.s_h) {
                    uint32_t sb = p.s[(plane * p.s_h + sy) * (int64_t)(p.s_w >> 2) + (sx >> 2)];
                    sb >>= (sx & 3) << 1;
                    if (sb & 1) v *= (S)p.slope;
                    if (sb & 2) v = (S)0;
                }
            } else {
                uint32_t code = 0;
                if (v < (S)0) { v *= (S)p.slope; code = 1; }
                if (fabs((double)v) > (double)p.clamp) { v = (v < (S)0) ? -(S)p.clamp : (S)p.clamp; code = 2; }
                packed |= code << (2 * k);
            }
            row[(int64_t)x * p.xsw] = from_acc<T, S>(v);
        }
        if (SIGN == 1 && gx * 4 < p.s_w && y < p.s_h)
            p.s[(plane * p.s_h + y) * (int64_t)(p.s_w >> 2) + gx] = (uint8_t)packed;
    }
}

template <class T>
int launch_fused(FlrArgs a, int sign, size_t smem, cudaStream_t stream) {
    void (*kern)(FlrArgs) = (sign == 1) ? filtered_lrelu_kernel<T, 1> : (sign == 2) ? filtered_lrelu_kernel<T, 2> : filtered_lrelu_kernel<T, 0>;
    if (smem > 48 * 1024) VFM_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int64_t blocks = (int64_t)a.tiles_x * a.tiles_y * a.channels * a.batch;
    if (blocks > 0x7fffffffLL) { set_error("filtered_lrelu: grid too large"); return VFM_ERR_INVALID; }
    double planes = (double)a.channels * a.batch;
    KernelTimer timer("filtered_lrelu_fused", stream, 0.0,
                      ((double)a.x_w * a.x_h + (double)a.y_w * a.y_h) * planes * sizeof(T) + (double)a.channels * sizeof(T) +
                      (sign ? planes * a.s_h * a.s_w_bytes : 0.0));
    kern<<<(unsigned)blocks, 256, smem, stream>>>(a);
    return launch_status("filtered_lrelu_kernel");
}

template <class T>
int launch_act(FlrActArgs a, int sign, cudaStream_t stream) {
    int w = (sign == 1) ? max(a.x_w, a.s_w) : a.x_w;
    a.groups_x = (w + 3) >> 2;
    a.rows = (sign == 1) ? max(a.x_h, a.s_h) : a.x_h;
    int64_t total = (int64_t)a.groups_x * a.rows * a.channels * a.batch;
    int64_t blocks = ceil_div64(total, 256);
    int64_t cap = (int64_t)kNumSMs * 32;
    if (blocks > cap) blocks = cap;
    if (blocks < 1) blocks = 1;
    void (*kern)(FlrActArgs, int64_t) = (sign == 1) ? filtered_lrelu_act_kernel<T, 1> : (sign == 2) ? filtered_lrelu_act_kernel<T, 2> : filtered_lrelu_act_kernel<T, 0>;
    kern<<<(unsigned)blocks, 256, 0, stream>>>(a, total);
    return launch_status("filtered_lrelu_act_kernel");
}

}  // namespace
}  // namespace vfm

extern "C" int vfm_filtered_lrelu(const vfm_filtered_lrelu_params* p, void* stream_) {
    using namespace vfm;
    cudaStream_t stream = (cudaStream_t)stream_;
    VFM_CHECK_ARG(p != nullptr, "filtered_lrelu: params is NULL");
    VFM_CHECK_ARG(p->x && p->y && p->b && p->fu && p->fd, "filtered_lrelu: x, y, b, fu, fd must be non-NULL");
    VFM_CHECK_ARG(p->up >= 1 && p->down >= 1, "filtered_lrelu: up and down must be at least 1");
    VFM_CHECK_ARG(p->x_w >= 1 && p->x_h >= 1 && p->channels >= 1 && p->batch >= 1, "filtered_lrelu: x is empty");
    VFM_CHECK_ARG(p->y_w >= 1 && p->y_h >= 1, "filtered_lrelu: output must be at least 1x1");
    VFM_CHECK_ARG(p->fu_w >= 1 && p->fd_w >= 1 && p->fu_h >= 0 && p->fd_h >= 0, "filtered_lrelu: fu/fd is empty");
    VFM_CHECK_ARG(!(p->write_signs && p->read_signs), "filtered_lrelu: cannot both read and write signs");
    VFM_CHECK_ARG(!(p->write_signs || p->read_signs) || p->s, "filtered_lrelu: sign tensor missing");
    // outside the fused kernel's envelope -> tell the caller to compose (same contract as the reference's rc = -1)
    const int fuh = p->fu_h ? p->fu_h : p->fu_w, fdh = p->fd_h ? p->fd_h : p->fd_w;
    if (p->dtype != VFM_F16 && p->dtype != VFM_F32) return VFM_ERR_NO_KERNEL;
    if (!(p->up == 1 || p->up == 2 || p->up == 4) || !(p->down == 1 || p->down == 2 || p->down == 4)) return VFM_ERR_NO_KERNEL;
    if (p->fu_w > kMaxTaps || fuh > kMaxTaps || p->fd_w > kMaxTaps || fdh > kMaxTaps) return VFM_ERR_NO_KERNEL;

    FlrArgs a;
    a.x = p->x; a.y = p->y; a.b = p->b; a.s = p->s; a.fu = p->fu; a.fd = p->fd;
    a.up = p->up; a.down = p->down;
    a.fu_w = p->fu_w; a.fu_h = p->fu_h; a.fd_w = p->fd_w; a.fd_h = p->fd_h;
    a.fu_sw = p->fu_stride_w; a.fu_sh = p->fu_stride_h; a.fd_sw = p->fd_stride_w; a.fd_sh = p->fd_stride_h;
    a.pad_x0 = p->pad_x0; a.pad_y0 = p->pad_y0;
    a.gain = p->gain; a.slope = p->slope; a.clamp = p->clamp; a.flip = p->flip ? 1 : 0;
    a.x_w = p->x_w; a.x_h = p->x_h; a.channels = p->channels; a.batch = p->batch;
    a.xsw = p->x_stride_w; a.xsh = p->x_stride_h; a.xsc = p->x_stride_c; a.xsn = p->x_stride_n;
    a.y_w = p->y_w; a.y_h = p->y_h;
    a.ysw = p->y_stride_w; a.ysh = p->y_stride_h; a.ysc = p->y_stride_c; a.ysn = p->y_stride_n;
    a.b_stride = p->b_stride;
    a.s_w_bytes = p->s_w_bytes; a.s_h = p->s_h; a.s_ofs_x = p->s_ofs_x; a.s_ofs_y = p->s_ofs_y; a.s_w_active = p->s_w_active;

    // pick the largest square-ish output tile whose intermediates fit in ~100 KB (2 CTAs per SM)
    const size_t budget = 100 * 1024;
    size_t smem = 0;
    int tow = 0, toh = 0;
    const int cand[][2] = {{64, 32}, {32, 32}, {32, 16}, {16, 16}, {16, 8}, {8, 8}, {4, 4}};
    for (auto& cd : cand) {
        tow = cd[0]; toh = cd[1];
        if (tow > ((p->y_w + 3) & ~3) * 2 && tow > 4) continue;   // do not waste a big tile on a small image
        a.tow = tow; a.toh = toh;
        a.tuw = (((tow - 1) * p->down + p->fd_w) + 3) & ~3;
        a.tuh = (toh - 1) * p->down + fdh;
        a.tiw = (a.tuw + p->fu_w - 1 + p->up - 1) / p->up + 1;
        a.tih = (a.tuh + fuh - 1 + p->up - 1) / p->up + 1;
        size_t fl = (size_t)(p->fu_h ? fuh * p->fu_w : p->fu_w) + (size_t)(p->fd_h ? fdh * p->fd_w : p->fd_w) + 4;
        fl += (size_t)a.tih * a.tiw + (p->fu_h == 0 ? (size_t)a.tih * a.tuw : 0) + (size_t)a.tuh * a.tuw +
              (p->fd_h == 0 ? (size_t)a.tuh * a.tow : 0);
        smem = fl * sizeof(float);
        if (smem <= budget) break;
        smem = 0;
    }
    if (smem == 0) return VFM_ERR_NO_KERNEL;
    a.tiles_x = ceil_div(p->y_w, a.tow);
    a.tiles_y = ceil_div(p->y_h, a.toh);
    int sign = p->write_signs ? 1 : (p->read_signs ? 2 : 0);
    if (p->dtype == VFM_F16) return launch_fused<__half>(a, sign, smem, stream);
    return launch_fused<float>(a, sign, smem, stream);
}

extern "C" int vfm_filtered_lrelu_act(const vfm_filtered_lrelu_act_params* p, void* stream_) {
    using namespace vfm;
    cudaStream_t stream = (cudaStream_t)stream_;
    VFM_CHECK_ARG(p != nullptr, "filtered_lrelu_act: params is NULL");
    VFM_CHECK_ARG(p->x != nullptr, "filtered_lrelu_act: x must be non-NULL");
    VFM_CHECK_ARG(p->x_w >= 1 && p->x_h >= 1 && p->channels >= 1 && p->batch >= 1, "filtered_lrelu_act: x is empty");
    VFM_CHECK_ARG(!(p->write_signs && p->read_signs), "filtered_lrelu_act: cannot both read and write signs");
    VFM_CHECK_ARG(!(p->write_signs || p->read_signs) || (p->s && (p->s_w & 3) == 0), "filtered_lrelu_act: bad sign tensor");
    VFM_CHECK_ARG(p->dtype == VFM_F16 || p->dtype == VFM_F32 || p->dtype == VFM_F64, "filtered_lrelu_act: unsupported dtype");
    FlrActArgs a;
    a.x = p->x; a.s = p->s; a.gain = p->gain; a.slope = p->slope; a.clamp = p->clamp;
    a.x_w = p->x_w; a.x_h = p->x_h; a.channels = p->channels; a.batch = p->batch;
    a.xsw = p->x_stride_w; a.xsh = p->x_stride_h; a.xsc = p->x_stride_c; a.xsn = p->x_stride_n;
    a.s_w = p->s_w; a.s_h = p->s_h; a.s_ofs_x = p->s_ofs_x; a.s_ofs_y = p->s_ofs_y;
    a.groups_x = a.rows = 0;
    int sign = p->write_signs ? 1 : (p->read_signs ? 2 : 0);
    switch (p->dtype) {
        case VFM_F16: return launch_act<__half>(a, sign, stream);
        case VFM_F32: return launch_act<float>(a, sign, stream);
        default:      return launch_act<double>(a, sign, stream);
    }
}
