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
    float* ysum;             // optional [channels] fp32: += sum of the (rounded) outputs of each plane: the fused bias gradient of the backward pass
    // tiling (host-computed)
    int tow, toh;            // output tile
    int tuw, tuh;            // upsampled tile (tuw multiple of 4)
    int tiw, tih;            // input tile
    int tiles_x, tiles_y;
};

// sum of a per-thread partial over the 256-thread CTA -> one atomicAdd into ysum[c] (fused db = dx.sum([0,2,3]) of the backward pass,
// reference filtered_lrelu.py:266).  Called by every thread of the CTA after its last store.
__device__ __forceinline__ void flr_accumulate_ysum(float* ysum, int c, float part) {
    __shared__ float s_part[8];
    part = warp_sum(part);
    __syncthreads();
    if ((threadIdx.x & 31) == 0) s_part[threadIdx.x >> 5] = part;
    __syncthreads();
    if (threadIdx.x == 0) {
        float t = 0.f;
#pragma unroll
        for (int w = 0; w < 8; w++) t += s_part[w];
        atomicAdd(&ysum[c], t);
    }
}

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
    float ysum_part = 0.f;                 // this thread's share of sum(y) over the plane (p.ysum)
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
            const T q = from_acc<T, float>(acc);
            yp[(int64_t)oy * p.ysh + (int64_t)ox * p.ysw] = q;
            ysum_part += to_acc(q);
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
            const T q = from_acc<T, float>(acc);
            yp[(int64_t)oy * p.ysh + (int64_t)ox * p.ysw] = q;
            ysum_part += to_acc(q);
        }
    }
    if (p.ysum) flr_accumulate_ysum(p.ysum, c, ysum_part);
}

// ------------------------------------------------------------------------------------------------------------
// Specialised fused kernel for separable fu and fd with compile-time up/down factors and <= 6 taps per polyphase
// (fu_w <= 6*UP, fd_w <= 6*DOWN: the StyleGAN3 family, e.g. 12-tap up2/down2).  Same five stages as the generic kernel,
// but every FIR stage is register-blocked: a thread loads a short run of samples once (128-bit shared-memory loads where
// the layout allows), keeps the taps in registers and produces 8 (or 8x4) outputs from them, so the inner loops are pure
// FFMA with compile-time polyphase indexing.  The phase alignment of the zero-inserted signal is a run-time property
// of the padding; it is absorbed by starting the thread blocks of the up passes at column/row (c - UP), c = (pad0 -
// tile origin) mod UP, which makes the phase of a thread's i-th output a compile-time function of i.
template <int UP, int DOWN> struct SepCfg {
    static constexpr int TU = 6;                                   // taps per polyphase of the up filter
    static constexpr int FU = TU * UP, FD = 6 * DOWN;              // padded tap counts
    static constexpr int TO = (DOWN == 4) ? 16 : 32;               // output tile (square)
    static constexpr int TUA = (TO - 1) * DOWN + FD;               // intermediate samples needed per dimension
    static constexpr int RD4 = (7 * DOWN + FD + 3) / 4;            // float4 loads of one 8-output run of the horizontal down pass
    static constexpr int TUNEED = ((TO / 8 - 1) * 8 * DOWN + RD4 * 4 > TUA) ? (TO / 8 - 1) * 8 * DOWN + RD4 * 4 : TUA;
    static constexpr int TUR = (TUNEED + 3) & ~3;
    static constexpr int TUP = (TUR % 8 == 0) ? TUR + 4 : TUR;     // row pitch of the intermediates: == 4 (mod 8) -> conflict-free float4 columns
    static constexpr int NB = (TUP + UP + 7) / 8;                  // 8-sample thread blocks of the up passes: every column < TUP is
                                                                   // a real intermediate sample (its sign byte may be shared with the next tile)
    static constexpr int NIN = (UP == 1) ? 8 + TU - 1 : 8 / UP + TU;   // input samples feeding one 8-sample block
    static constexpr int TI = (8 * NB + FU) / UP + 2;              // input tile per dimension
    static constexpr int TIP = TI | 1;                             // odd pitch
    static constexpr int TOP = TO + 4;                             // pitch of the horizontally down-filtered rows (== 4 mod 8)
    static constexpr int r4(int v) { return (v + 3) & ~3; }
    static constexpr int OFF_FD = r4(FU), OFF_IN = OFF_FD + r4(FD), OFF_UX = OFF_IN + r4(TI * TIP), OFF_U = OFF_UX + TI * TUP,
                         OFF_DX = OFF_U + TUA * TUP, smem_floats = OFF_DX + TUA * TOP;
};

// out[i] (i < 8) = sum_k g[phase(i) + UP*k] * in[base(i) + k]: eight consecutive samples of an up-FIR whose first sample
// is phase-aligned (zero-inserted coordinate of tap 0 is a multiple of UP)
template <int UP, int TU, class V> __device__ __forceinline__ void up8(const float* g, const V* in, V* out) {
#pragma unroll
    for (int i = 0; i < 8; i++) {
        constexpr int dummy = 0; (void)dummy;
        const int t0 = (UP - i % UP) % UP, base = (i + UP - 1) / UP;
        V acc = in[base] * g[t0];
#pragma unroll
        for (int k = 1; k < TU; k++) acc = acc + in[base + k] * g[t0 + UP * k];
        out[i] = acc;
    }
}
struct F4 { float x, y, z, w; };
__device__ __forceinline__ F4 operator*(const F4& a, float s) { return F4{a.x * s, a.y * s, a.z * s, a.w * s}; }
__device__ __forceinline__ F4 operator+(const F4& a, const F4& b) { return F4{a.x + b.x, a.y + b.y, a.z + b.z, a.w + b.w}; }

template <class T, int UP, int DOWN, int SIGN>
__global__ void __launch_bounds__(256) flr_sep_kernel(FlrArgs p) {
    typedef SepCfg<UP, DOWN> C;
    extern __shared__ __align__(16) float smem[];
    float* s_fu = smem;                       // [FU]
    float* s_fd = smem + C::OFF_FD;           // [FD]
    float* s_in = smem + C::OFF_IN;           // [TI][TIP]
    float* s_ux = smem + C::OFF_UX;           // [TI][TUP]   (16-byte aligned, like s_u and s_dx)
    float* s_u = smem + C::OFF_U;             // [TUA][TUP]
    float* s_dx = smem + C::OFF_DX;           // [TUA][TOP]

    const int tid = threadIdx.x;
    int64_t bid = blockIdx.x;
    const int tile_x = (int)(bid % p.tiles_x); bid /= p.tiles_x;
    const int tile_y = (int)(bid % p.tiles_y); bid /= p.tiles_y;
    const int c = (int)(bid % p.channels), n = (int)(bid / p.channels);
    float ysum_part = 0.f;                 // this thread's share of sum(y) over the plane (p.ysum)
    const int64_t plane = (int64_t)n * p.channels + c;

    // correlation taps (flip == 0 means true convolution -> reversed), zero-extended to the padded counts
    for (int i = tid; i < C::FU; i += 256) s_fu[i] = (i < p.fu_w) ? p.fu[(p.flip ? i : p.fu_w - 1 - i) * p.fu_sw] : 0.f;
    for (int i = tid; i < C::FD; i += 256) s_fd[i] = (i < p.fd_w) ? p.fd[(p.flip ? i : p.fd_w - 1 - i) * p.fd_sw] : 0.f;

    const int ox0 = tile_x * C::TO, oy0 = tile_y * C::TO;
    const int ux0 = ox0 * DOWN, uy0 = oy0 * DOWN;                       // first intermediate sample of the tile
    const int cx = ((p.pad_x0 - ux0) % UP + UP) % UP, cy = ((p.pad_y0 - uy0) % UP + UP) % UP;
    const int sx0 = cx - UP, sy0 = cy - UP;                             // first column / row of the up-pass thread blocks (negative)
    const int ix0 = (ux0 + sx0 - p.pad_x0) / UP, iy0 = (uy0 + sy0 - p.pad_y0) / UP;   // exact divisions: first input sample of the tile

    // ---- stage 1: input tile + bias (zero outside the image) ----
    {
        const float bias = to_acc(*(const T*)((const char*)p.b + (int64_t)c * p.b_stride * (int64_t)sizeof(T)));
        const T* xp = (const T*)p.x + (int64_t)n * p.xsn + (int64_t)c * p.xsc;
        for (int i = tid; i < C::TI * C::TI; i += 256) {
            const int ry = i / C::TI, rx = i - ry * C::TI;
            const int ix = ix0 + rx, iy = iy0 + ry;
            float v = 0.f;
            if (ix >= 0 && ix < p.x_w && iy >= 0 && iy < p.x_h) v = to_acc(xp[(int64_t)iy * p.xsh + (int64_t)ix * p.xsw]) + bias;
            s_in[ry * C::TIP + rx] = v;
        }
    }
    __syncthreads();

    // ---- stage 2: horizontal up-FIR.  item = (input row ry, block b): columns 8b + sx0 .. +7 ----
    {
        float g[C::FU];
#pragma unroll
        for (int i = 0; i < C::FU; i++) g[i] = s_fu[i];
        for (int it = tid; it < C::TI * C::NB; it += 256) {
            const int b = it / C::TI, ry = it - b * C::TI;             // lanes along rows: odd pitch -> conflict-free
            const float* src = s_in + ry * C::TIP + (8 * b) / UP;       // (ux0 + 8b + sx0 - pad)/UP - ix0 == 8b/UP
            float in[C::NIN], out[8];
#pragma unroll
            for (int k = 0; k < C::NIN; k++) in[k] = src[k];
            up8<UP, C::TU, float>(g, in, out);
            float* dst = s_ux + ry * C::TUP + 8 * b + sx0;
#pragma unroll
            for (int i = 0; i < 8; i++) { const int rux = 8 * b + sx0 + i; if (rux >= 0 && rux < C::TUP) dst[i] = out[i]; }
        }
    }
    __syncthreads();

    // ---- stage 3: vertical up-FIR + gain / lrelu / clamp (+ signs).  item = (4-column group, block rb): rows 8rb + sy0 .. +7 ----
    {
        float g[C::FU];
#pragma unroll
        for (int i = 0; i < C::FU; i++) g[i] = s_fu[i];
        const float act_gain = (float)(UP * UP) * p.gain;
        constexpr int NG = C::TUP / 4;
        for (int it = tid; it < NG * C::NB; it += 256) {
            const int rb = it / NG, g4 = it - rb * NG;                 // lanes along column groups: consecutive float4
            const float* src = s_ux + ((8 * rb) / UP) * C::TUP + 4 * g4;
            F4 in[C::NIN], out[8];
#pragma unroll
            for (int k = 0; k < C::NIN; k++) { const float4 t = *(const float4*)(src + k * C::TUP); in[k] = F4{t.x, t.y, t.z, t.w}; }
            up8<UP, C::TU, F4>(g, in, out);
#pragma unroll
            for (int i = 0; i < 8; i++) {
                const int ruy = 8 * rb + sy0 + i;
                if (ruy < 0 || ruy >= C::TUA) continue;
                float v[4] = {out[i].x * act_gain, out[i].y * act_gain, out[i].z * act_gain, out[i].w * act_gain};
                const int sy = uy0 + ruy + (SIGN == 2 ? p.s_ofs_y : 0);
                if (SIGN == 2) {
                    if (sy >= 0 && sy < p.s_h) {
                        const uint8_t* srow = p.s + (plane * p.s_h + sy) * (int64_t)p.s_w_bytes;
#pragma unroll
                        for (int k = 0; k < 4; k++) {
                            const int sx = ux0 + 4 * g4 + k + p.s_ofs_x;
                            if (sx >= 0 && sx < p.s_w_active) {
                                const uint32_t sb = (uint32_t)srow[sx >> 2] >> ((sx & 3) << 1);
                                if (sb & 1) v[k] *= p.slope;
                                if (sb & 2) v[k] = 0.f;
                            }
                        }
                    }
                } else {
                    uint32_t packed = 0;
#pragma unroll
                    for (int k = 0; k < 4; k++) {
                        uint32_t code = 0;
                        if (v[k] < 0.f) { v[k] *= p.slope; code = 1; }
                        if (fabsf(v[k]) > p.clamp) { v[k] = copysignf(p.clamp, v[k]); code = 2; }
                        packed |= code << (2 * k);
                    }
                    if (SIGN == 1) {
                        const int sx = ux0 + 4 * g4;                     // write mode: sign offset is zero, ux0 is a multiple of 4
                        if (sx < p.s_w_active && sy < p.s_h) p.s[(plane * p.s_h + sy) * (int64_t)p.s_w_bytes + (sx >> 2)] = (uint8_t)packed;
                    }
                }
                *(float4*)(s_u + ruy * C::TUP + 4 * g4) = make_float4(v[0], v[1], v[2], v[3]);
            }
        }
    }
    __syncthreads();

    // ---- stage 4: horizontal down-FIR.  item = (intermediate row, q): outputs 8q .. 8q+7 ----
    {
        float h[C::FD];
#pragma unroll
        for (int i = 0; i < C::FD; i++) h[i] = s_fd[i];
        constexpr int NQ = C::TO / 8;
        for (int it = tid; it < C::TUA * NQ; it += 256) {
            const int q = it / C::TUA, ruy = it - q * C::TUA;          // lanes along rows: pitch == 4 (mod 8) -> conflict-free float4
            const float* src = s_u + ruy * C::TUP + 8 * q * DOWN;
            float in[C::RD4 * 4];
#pragma unroll
            for (int k = 0; k < C::RD4; k++) { const float4 t = *(const float4*)(src + 4 * k); in[4 * k] = t.x; in[4 * k + 1] = t.y; in[4 * k + 2] = t.z; in[4 * k + 3] = t.w; }
            float out[8];
#pragma unroll
            for (int i = 0; i < 8; i++) {
                float acc = in[i * DOWN] * h[0];
#pragma unroll
                for (int t = 1; t < C::FD; t++) acc = fmaf(in[i * DOWN + t], h[t], acc);
                out[i] = acc;
            }
            float* dst = s_dx + ruy * C::TOP + 8 * q;
            *(float4*)dst = make_float4(out[0], out[1], out[2], out[3]);
            *(float4*)(dst + 4) = make_float4(out[4], out[5], out[6], out[7]);
        }
    }
    __syncthreads();

    // ---- stage 5: vertical down-FIR + store.  item = (output row pair, 4-column group) ----
    {
        float h[C::FD];
#pragma unroll
        for (int i = 0; i < C::FD; i++) h[i] = s_fd[i];
        T* yp = (T*)p.y + (int64_t)n * p.ysn + (int64_t)c * p.ysc;
        constexpr int NG = C::TO / 4, NR = C::TO / 2;
        for (int it = tid; it < NG * NR; it += 256) {
            const int rp = it / NG, g4 = it - rp * NG;
            const float* src = s_dx + (2 * rp * DOWN) * C::TOP + 4 * g4;
            F4 acc[2] = {F4{0.f, 0.f, 0.f, 0.f}, F4{0.f, 0.f, 0.f, 0.f}};
#pragma unroll
            for (int k = 0; k < DOWN + C::FD; k++) {
                const float4 t = *(const float4*)(src + k * C::TOP);
                const F4 v = F4{t.x, t.y, t.z, t.w};
                if (k < C::FD) acc[0] = acc[0] + v * h[k];
                if (k >= DOWN) acc[1] = acc[1] + v * h[k - DOWN];
            }
#pragma unroll
            for (int r = 0; r < 2; r++) {
                const int oy = oy0 + 2 * rp + r;
                if (oy >= p.y_h) continue;
                const float o[4] = {acc[r].x, acc[r].y, acc[r].z, acc[r].w};
#pragma unroll
                for (int k = 0; k < 4; k++) {
                    const int ox = ox0 + 4 * g4 + k;
                    if (ox < p.y_w) { const T q = from_acc<T, float>(o[k]); yp[(int64_t)oy * p.ysh + (int64_t)ox * p.ysw] = q; ysum_part += to_acc(q); }
                }
            }
        }
    }
    if (p.ysum) flr_accumulate_ysum(p.ysum, c, ysum_part);
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

template <class T, int UP, int DOWN>
int launch_sep(FlrArgs a, int sign, cudaStream_t stream) {
    typedef SepCfg<UP, DOWN> C;
    void (*kern)(FlrArgs) = (sign == 1) ? flr_sep_kernel<T, UP, DOWN, 1> : (sign == 2) ? flr_sep_kernel<T, UP, DOWN, 2> : flr_sep_kernel<T, UP, DOWN, 0>;
    const size_t smem = ((size_t)C::smem_floats + 8) * sizeof(float);
    VFM_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    a.tow = a.toh = C::TO;
    a.tiles_x = ceil_div(a.y_w, C::TO);
    a.tiles_y = ceil_div(a.y_h, C::TO);
    int64_t blocks = (int64_t)a.tiles_x * a.tiles_y * a.channels * a.batch;
    if (blocks > 0x7fffffffLL) { set_error("filtered_lrelu: grid too large"); return VFM_ERR_INVALID; }
    double planes = (double)a.channels * a.batch;
    KernelTimer timer("filtered_lrelu_sep", stream, 0.0,
                      ((double)a.x_w * a.x_h + (double)a.y_w * a.y_h) * planes * sizeof(T) + (double)a.channels * sizeof(T) +
                      (sign ? planes * a.s_h * a.s_w_bytes : 0.0), "u%dd%dw%d", UP, DOWN, a.y_w);
    kern<<<(unsigned)blocks, 256, smem, stream>>>(a);
    return launch_status("flr_sep_kernel");
}

// returns VFM_ERR_NO_KERNEL when the specialised kernel does not cover the configuration
template <class T>
int try_sep(const FlrArgs& a, int sign, cudaStream_t stream) {
    if (a.fu_h != 0 || a.fd_h != 0) return VFM_ERR_NO_KERNEL;                       // both filters separable
    if (a.fu_w > 6 * a.up || a.fd_w > 6 * a.down) return VFM_ERR_NO_KERNEL;         // <= 6 taps per polyphase
    if (a.up == 2 && a.down == 2) return launch_sep<T, 2, 2>(a, sign, stream);
    if (a.up == 4 && a.down == 2) return launch_sep<T, 4, 2>(a, sign, stream);
    if (a.up == 2 && a.down == 4) return launch_sep<T, 2, 4>(a, sign, stream);
    if (a.up == 2 && a.down == 1) return launch_sep<T, 2, 1>(a, sign, stream);
    if (a.up == 1 && a.down == 2) return launch_sep<T, 1, 2>(a, sign, stream);
    return VFM_ERR_NO_KERNEL;
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
    a.ysum = p->y_sum;

    {
        const int sign = p->write_signs ? 1 : (p->read_signs ? 2 : 0);
        a.tow = a.toh = a.tuw = a.tuh = a.tiw = a.tih = a.tiles_x = a.tiles_y = 0;
        const int st = (p->dtype == VFM_F16) ? try_sep<__half>(a, sign, stream) : try_sep<float>(a, sign, stream);
        if (st != VFM_ERR_NO_KERNEL) return st;
    }
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
