// upfirdn2d for sm_100a: zero-insert -> pad/crop -> FIR -> decimate per (n,c) plane.
//
// Two kernels:
//   * upfirdn2d_tiled<T,UP,DOWN,FW,FH>: the HBM-bound workhorse for W-contiguous tensors.  A CTA stages the input
//     halo tile in shared memory with coalesced loads (zero-filled outside the image, which is what implements both
//     the zero padding and negative-padding crops), every thread then produces a 4x4 register block of outputs from a
//     register patch with the (flipped, gain-scaled) taps held in registers.  All polyphase index arithmetic is
//     resolved at compile time by aligning the tile grid to the up-sampling phase, so the inner loop is pure FFMA:
//     no per-tap shared-memory filter reads (the reference kernel does two LDS per FMA,
//     torch_utils/ops/upfirdn2d.cu:189-193).
//   * upfirdn2d_generic<T>: one thread per output, runtime parameters, arbitrary strides (channels_last, separable
//     passes, odd up/down combinations).  Correct for everything; used when no tiled specialisation applies.
//
// Semantics: torch_utils/ops/upfirdn2d.py:167-211 (_upfirdn2d_ref) via the oracle; kernel parameter meaning
// follows torch_utils/ops/upfirdn2d.cpp:16-98.
#include "common.cuh"

namespace vfm {
namespace {

struct UpfirdnArgs {
    const void* x; const float* f; void* y; const float* add;
    int upx, upy, downx, downy, padx0, pady0, flip;
    double gain;
    int in_w, in_h, channels, batch;
    int64_t isw, ish, isc, isn;
    int fw, fh; int64_t fsw, fsh;
    int out_w, out_h;
    int64_t osw, osh, osc, osn;
    int64_t add_sh, add_sn;
    // tiled only
    int xstart, ystart, tiles_x, tiles_y;
    int vec_ok;   // input rows can be staged with 16-byte loads
    int ep_enable, ep_act; float ep_alpha, ep_gain, ep_clamp; const void* ep_bias;
    int pad_mode;    // 0 = zero padding, 1 = replicate (clamp to edge; streaming blur kernel only)
    int64_t fsc;     // per-channel stride of f (0 = one shared filter; != 0: depthwise conv with learned taps)
};

template <class T, class S> __device__ __forceinline__ S ep_apply(const UpfirdnArgs& p, S v, int c) {
    if (p.ep_bias) v += to_acc(((const T*)p.ep_bias)[c]);
    if (p.ep_act == 3) v = (v > (S)0) ? v : v * (S)p.ep_alpha;
    v *= (S)p.ep_gain;
    if (p.ep_clamp >= 0.f) v = (v > (S)p.ep_clamp) ? (S)p.ep_clamp : ((v < -(S)p.ep_clamp) ? -(S)p.ep_clamp : v);
    return v;
}

// ------------------------------------------------------------------------------------------------------------
template <class T>
__global__ void __launch_bounds__(256) upfirdn2d_generic(UpfirdnArgs p, int64_t total, int c_fastest) {
    typedef typename Acc<T>::type S;
    const float gain32 = (float)p.gain;   // taps are scaled in fp32, as the reference scales its fp32 filter tensor
    for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (int64_t)gridDim.x * blockDim.x) {
        int ox, oy, c, n;
        int64_t r = idx;
        if (c_fastest) {
            c = (int)(r % p.channels); r /= p.channels;
            ox = (int)(r % p.out_w); r /= p.out_w;
            oy = (int)(r % p.out_h); n = (int)(r / p.out_h);
        } else {
            ox = (int)(r % p.out_w); r /= p.out_w;
            oy = (int)(r % p.out_h); r /= p.out_h;
            c = (int)(r % p.channels); n = (int)(r / p.channels);
        }
        // position of tap 0 in the zero-inserted (unpadded) signal
        int ux = ox * p.downx - p.padx0, uy = oy * p.downy - p.pady0;
        int tx0 = ((-ux) % p.upx + p.upx) % p.upx, ty0 = ((-uy) % p.upy + p.upy) % p.upy;
        const T* xp = (const T*)p.x + (int64_t)n * p.isn + (int64_t)c * p.isc;
        S acc = (S)0;
        for (int ty = ty0; ty < p.fh; ty += p.upy) {
            int iy = (uy + ty) / p.upy;   // exact
            if (iy < 0 || iy >= p.in_h) continue;
            int fy = p.flip ? ty : p.fh - 1 - ty;
            for (int tx = tx0; tx < p.fw; tx += p.upx) {
                int ix = (ux + tx) / p.upx;
                if (ix < 0 || ix >= p.in_w) continue;
                int fx = p.flip ? tx : p.fw - 1 - tx;
                acc += to_acc(xp[(int64_t)iy * p.ish + (int64_t)ix * p.isw]) * (S)(__ldg(&p.f[fy * p.fsh + fx * p.fsw]) * gain32);
            }
        }
        if (p.add) acc += (S)p.add[(int64_t)n * p.add_sn + (int64_t)oy * p.add_sh + ox];
        if (p.ep_enable) acc = ep_apply<T, S>(p, acc, c);
        ((T*)p.y)[(int64_t)n * p.osn + (int64_t)c * p.osc + (int64_t)oy * p.osh + (int64_t)ox * p.osw] = from_acc<T, S>(acc);
    }
}

// ------------------------------------------------------------------------------------------------------------
// Tiled kernel.  Requires: in/out W-contiguous (stride_w == 1), (UP == 1 || DOWN == 1).
//
// Shared-memory layout: one row per input row; inside a row the columns are stored *swizzled* so that the register-patch
// reads of a warp (lane l reads column G*l + t, G = OX*DOWN/UP input columns per thread) hit 32 different banks:
//     column c = G*a + b  (0 <= b < G)   ->   slot  b*A + a      (A = columns per residue class)
// The tile origin in shared memory is aligned to 16 bytes of the *global* row, so staging is one 16-byte global load per
// thread (when the tensor's strides allow it; scalar loads otherwise) followed by V scalar shared stores (<= 2-way conflict).
template <int UP, int DOWN, int FW, int FH>
struct TileCfg {
    static constexpr int OX = 4, OY = (DOWN == 1) ? 4 : 2;   // outputs per thread
    static constexpr int TX = 32, TY = 8;            // threads
    static constexpr int TW = TX * OX, TH = TY * OY; // output tile
    static constexpr int G = OX * DOWN / UP;         // input columns between neighbouring lanes (4, 2, 1 or 8)
    static constexpr int PW = ((OX - 1) * DOWN + FW - 1) / UP + 1;   // register patch
    static constexpr int PH = ((OY - 1) * DOWN + FH - 1) / UP + 1;
    static constexpr int IW = ((TW - 1) * DOWN + FW - 1) / UP + 1;   // input tile (exact)
    static constexpr int IH = ((TH - 1) * DOWN + FH - 1) / UP + 1;
    static constexpr int MAXSHIFT = 8;                                // alignment slack in columns (16 bytes of fp16)
    static constexpr int NCOL = IW + MAXSHIFT;                        // columns staged per row
    static constexpr int A = (NCOL + 8 + G - 1) / G;                  // columns per residue class (incl. vector overrun)
    static constexpr int IWP = G * A;                                 // slots per row
    __host__ __device__ static constexpr int slot(int c) { return (c % G) * A + (c / G); }
};

template <class T, int UP, int DOWN, int FW, int FH>
__global__ void __launch_bounds__(256) upfirdn2d_tiled(UpfirdnArgs p) {
    typedef typename Acc<T>::type S;
    typedef TileCfg<UP, DOWN, FW, FH> C;
    constexpr int V = (int)(16 / sizeof(T));        // elements per 16-byte global vector
    __shared__ S s_in[C::IH * C::IWP];
    __shared__ S s_f[FH * FW];

    // flat block index -> (plane, tile_y, tile_x)
    int64_t bid = blockIdx.x;
    int tile_x = (int)(bid % p.tiles_x); bid /= p.tiles_x;
    int tile_y = (int)(bid % p.tiles_y); bid /= p.tiles_y;
    int c = (int)(bid % p.channels), n = (int)(bid / p.channels);

    // taps: store as correlation taps with the gain folded in
    for (int i = threadIdx.x; i < FW * FH; i += blockDim.x) {
        int ty = i / FW, tx = i - ty * FW;
        S v = (S)0;
        if (tx < p.fw && ty < p.fh) {
            int fx = p.flip ? tx : p.fw - 1 - tx, fy = p.flip ? ty : p.fh - 1 - ty;
            v = (S)(p.f[fy * p.fsh + fx * p.fsw] * (float)p.gain);   // product in fp32: the reference scales its fp32 filter tensor
        }
        s_f[i] = v;
    }

    // tile origin in output coordinates; (ox0*DOWN - pad0) is a multiple of UP by construction of xstart/ystart
    const int ox_t = p.xstart + tile_x * C::TW, oy_t = p.ystart + tile_y * C::TH;
    const int ix_t = (ox_t * DOWN - p.padx0) / UP, iy_t = (oy_t * DOWN - p.pady0) / UP;   // exact, may be negative
    const int ix_al = floor_div(ix_t, V) * V;        // 16-byte aligned column the staged tile starts at
    const int shift = ix_t - ix_al;                  // 0 .. V-1, uniform over the grid (TW*DOWN/UP is a multiple of 8)
    const T* xp = (const T*)p.x + (int64_t)n * p.isn + (int64_t)c * p.isc;
    {
        constexpr int NV = (C::NCOL + V - 1) / V;    // vectors per row
        for (int i = threadIdx.x; i < C::IH * NV; i += blockDim.x) {
            const int ry = i / NV, j = i - ry * NV;
            const int iy = iy_t + ry, ix0 = ix_al + j * V;
            struct alignas(16) Vec { T e[V]; } v;
            const bool row_ok = (iy >= 0 && iy < p.in_h);
            if (row_ok && p.vec_ok && ix0 >= 0 && ix0 + V <= p.in_w) {
                *(uint4*)&v = *(const uint4*)(xp + (int64_t)iy * p.ish + ix0);
            } else {
#pragma unroll
                for (int k = 0; k < V; k++) {
                    const int ix = ix0 + k;
                    v.e[k] = (row_ok && ix >= 0 && ix < p.in_w) ? xp[(int64_t)iy * p.ish + ix] : from_acc<T, S>((S)0);
                }
            }
            S* dst = s_in + ry * C::IWP;
#pragma unroll
            for (int k = 0; k < V; k++) dst[C::slot(j * V + k)] = to_acc(v.e[k]);
        }
    }
    __syncthreads();

    S f[FH][FW];
#pragma unroll
    for (int ty = 0; ty < FH; ty++)
#pragma unroll
        for (int tx = 0; tx < FW; tx++) f[ty][tx] = s_f[ty * FW + tx];

    const int lx = threadIdx.x & 31, ly = threadIdx.x >> 5;
    // thread's first output (tile-relative) and first patch element (tile-relative, input coords)
    const int jx0 = lx * C::OX, jy0 = ly * C::OY;
    const int py0 = jy0 * DOWN / UP;                 // exact: OY multiple of UP
    // slot of patch column q for this thread: column = G*lx + shift + q
    int pslot[C::PW];
#pragma unroll
    for (int q = 0; q < C::PW; q++) pslot[q] = C::slot(C::G * lx + shift + q);

    S acc[C::OY][C::OX];
#pragma unroll
    for (int a = 0; a < C::OY; a++)
#pragma unroll
        for (int b = 0; b < C::OX; b++) acc[a][b] = (S)0;

#pragma unroll
    for (int r = 0; r < C::PH; r++) {
        S row[C::PW];
        const S* src = &s_in[(py0 + r) * C::IWP];
#pragma unroll
        for (int q = 0; q < C::PW; q++) row[q] = src[pslot[q]];
#pragma unroll
        for (int jy = 0; jy < C::OY; jy++) {
#pragma unroll
            for (int ty = 0; ty < FH; ty++) {
                if ((jy * DOWN + ty) % UP != 0 || (jy * DOWN + ty) / UP != r) continue;
#pragma unroll
                for (int jx = 0; jx < C::OX; jx++) {
#pragma unroll
                    for (int tx = 0; tx < FW; tx++) {
                        if ((jx * DOWN + tx) % UP != 0) continue;
                        acc[jy][jx] += f[ty][tx] * row[(jx * DOWN + tx) / UP];
                    }
                }
            }
        }
    }

    T* yp = (T*)p.y + (int64_t)n * p.osn + (int64_t)c * p.osc;
    const int ox0 = ox_t + jx0, oy0 = oy_t + jy0;
#pragma unroll
    for (int jy = 0; jy < C::OY; jy++) {
        int oy = oy0 + jy;
        if (oy < 0 || oy >= p.out_h) continue;
        T* rowp = yp + (int64_t)oy * p.osh;
        struct alignas(16) OutV { T e[C::OX]; } outv;
        T* out = outv.e;
#pragma unroll
        for (int jx = 0; jx < C::OX; jx++) {
            S v = acc[jy][jx];
            int ox = ox0 + jx;
            if (p.add && ox >= 0 && ox < p.out_w) v += (S)p.add[(int64_t)n * p.add_sn + (int64_t)oy * p.add_sh + ox];
            if (p.ep_enable) v = ep_apply<T, S>(p, v, c);
            out[jx] = from_acc<T, S>(v);
        }
        bool full = ox0 >= 0 && ox0 + C::OX <= p.out_w;
        constexpr int BYTES = C::OX * (int)sizeof(T);
        if (full && ((reinterpret_cast<uintptr_t>(rowp + ox0) & (BYTES - 1)) == 0)) {
            if (BYTES == 8) *(uint2*)(rowp + ox0) = *(const uint2*)out;
            else if (BYTES == 16) *(uint4*)(rowp + ox0) = *(const uint4*)out;
            else { *(uint4*)(rowp + ox0) = ((const uint4*)out)[0]; *(uint4*)(rowp + ox0 + 2) = ((const uint4*)out)[1]; }
        } else {
#pragma unroll
            for (int jx = 0; jx < C::OX; jx++) {
                int ox = ox0 + jx;
                if (ox >= 0 && ox < p.out_w) rowp[ox] = out[jx];
            }
        }
    }
}

// ------------------------------------------------------------------------------------------------------------
// Streaming blur: up == down == 1, filter <= 4x4 (the modulated conv's post-transposed-conv FIR and its backward, which
// is where almost all upfirdn2d bytes of the decoder go).  No shared memory: a thread owns 8 consecutive output
// columns and walks down a strip of rows.  Per input row it issues ONE 16-byte (fp16) / 32-byte (fp32) load of its own
// 8 columns, takes the <= 3 halo columns on either side from the neighbouring lanes with warp shuffles (global loads only
// at strip edges), and keeps the four partially accumulated output rows in registers, so every input element is
// converted once and every output row leaves as one aligned vector store.  A rank-1 filter (the decoder's
// [1,3,3,1] x [1,3,3,1]) is detected in the kernel and evaluated separably (8 instead of 16 FMAs per output).
// Requires 16-byte aligned rows on both sides; everything else goes to the tiled / generic kernels.
template <int FT> struct BlurTaps { float f[FT][FT]; float fx[FT], fy[FT]; };

// 128- / 256-bit streaming accesses of NB bytes (16, 32 or 64)
template <int NB> __device__ __forceinline__ void ldg_words(const void* p, uint32_t* w) {
    if (NB == 16) { const uint4 u = ldg_stream(p); w[0] = u.x; w[1] = u.y; w[2] = u.z; w[3] = u.w; }
    else {
#pragma unroll
        for (int q = 0; q < NB / 32; q++)
            asm volatile("ld.global.nc.L1::no_allocate.v8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                         : "=r"(w[8 * q]), "=r"(w[8 * q + 1]), "=r"(w[8 * q + 2]), "=r"(w[8 * q + 3]), "=r"(w[8 * q + 4]), "=r"(w[8 * q + 5]), "=r"(w[8 * q + 6]), "=r"(w[8 * q + 7])
                         : "l"((const char*)p + 32 * q));
    }
}
template <int NB> __device__ __forceinline__ void stg_words(void* p, const uint32_t* w) {
    if (NB == 16) stg_stream(p, make_uint4(w[0], w[1], w[2], w[3]));
    else {
#pragma unroll
        for (int q = 0; q < NB / 32; q++)
            asm volatile("st.global.L1::no_allocate.v8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};"
                         :: "l"((char*)p + 32 * q), "r"(w[8 * q]), "r"(w[8 * q + 1]), "r"(w[8 * q + 2]), "r"(w[8 * q + 3]), "r"(w[8 * q + 4]), "r"(w[8 * q + 5]), "r"(w[8 * q + 6]), "r"(w[8 * q + 7]) : "memory");
    }
}

// NC = output columns per thread (8, or 16 for fp16 rows aligned to 32 bytes: one 256-bit load / store per row)
// FT = taps per dimension (4, or 5 for the pixel-shuffle upsampler's [1,4,6,4,1] blur); REPL = replicate (clamp-to-edge) padding
template <class T, int PX, bool SEP, int NC, int FT, bool REPL>
__device__ __forceinline__ void blur_body(const UpfirdnArgs& p, const BlurTaps<FT>& k, int CG, int strips, int strip_rows) {
    constexpr int NL = PX;                 // halo columns on the left
    constexpr int NR = FT - 1 - PX;        // halo columns on the right
    constexpr int EW = (int)(4 / sizeof(T));   // elements per 32-bit word (2 for fp16, 1 for fp32)
    constexpr int NW = NC / EW;            // words of the thread's own NC columns
    constexpr int NBYTES = NC * (int)sizeof(T);
    constexpr int WL = (NL + EW - 1) / EW, WR = (NR + EW - 1) / EW;    // halo words needed on each side
    constexpr int NE = WL + WR;            // halo words an edge lane fetches itself
    // CG = column groups per row (any number: a warp may straddle two strips, see the edge flags below)
    const int64_t gid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int cgi = (int)(gid % CG);
    int64_t rest = gid / CG;
    const int strip = (int)(rest % strips);
    const int64_t plane = rest / strips;
    const bool alive = plane < (int64_t)p.channels * p.batch;
    const int c = alive ? (int)(plane % p.channels) : 0, n = alive ? (int)(plane / p.channels) : 0;
    const int lane = threadIdx.x & 31;
    // a lane's shuffle neighbour is its column neighbour unless the lane sits at the end of the warp or of the row
    const bool edge_l = lane == 0 || cgi == 0, edge_r = lane == 31 || cgi == CG - 1;

    const int ox0 = cgi * NC;
    const int oy_begin = strip * strip_rows;
    const int oy_end = min(oy_begin + strip_rows, p.out_h);
    const bool active = alive && ox0 < p.out_w && oy_begin < oy_end;      // produces outputs
    const bool feeds = alive && oy_begin < oy_end;                         // loads input (also for its neighbours' halos)
    const bool add_vec = p.add && ((reinterpret_cast<uintptr_t>(p.add) & 15u) == 0) && (p.add_sh & 3) == 0 && (p.add_sn & 3) == 0;
    const T* xp = (const T*)p.x + (int64_t)n * p.isn + (int64_t)c * p.isc;
    T* yp = (T*)p.y + (int64_t)n * p.osn + (int64_t)c * p.osc;
    // fused layer epilogue y = clamp(gain * act(blur + add + bias)):  with gain > 0 the gain is folded into the taps (by the
    // caller of blur_body), the bias into the initial value of the accumulators, and lrelu(v) = max(v, alpha * v) for 0 <= alpha <= 1
    const float eg = p.ep_enable ? p.ep_gain : 1.f;                       // > 0 (checked on the host)
    const float acc_init = (p.ep_enable && p.ep_bias) ? to_acc(((const T*)p.ep_bias)[c]) * eg : 0.f;
    const bool lrelu = p.ep_enable && p.ep_act == 3;
    const float alpha = p.ep_alpha;

    float acc[FT][NC];
#pragma unroll
    for (int a = 0; a < FT; a++)
#pragma unroll
        for (int b = 0; b < NC; b++) acc[a][b] = acc_init;

    // input rows iy = oy - pady0 + ty  ->  rows [oy_begin - pady0, oy_end - 1 - pady0 + FT - 1]
    const int nrows = feeds ? (oy_end - oy_begin + FT - 1) : 0;
    const int nrows_warp = __reduce_max_sync(0xffffffffu, nrows);      // shuffles need the whole warp in the loop

    // Loads never wait for their data inside fetch(): a vector / word that straddles the row end is loaded whole (it is
    // aligned and contains at least one valid element, so it lies inside the allocation) and the columns >= in_w are
    // masked to zero when the words are unpacked.  keep[i] = byte mask of word i of {own, left halo, right halo}.
    uint32_t keep[NW + NE];
    bool partial = false;
    {
        auto word_mask = [&](int ix) -> uint32_t {          // word covering columns ix .. ix + EW - 1
            if (EW == 1) return (ix >= 0 && ix < p.in_w) ? 0xffffffffu : 0u;
            if (ix < 0 || ix >= p.in_w) return 0u;
            return (ix + 1 < p.in_w) ? 0xffffffffu : 0x0000ffffu;
        };
#pragma unroll
        for (int i = 0; i < NW; i++) keep[i] = word_mask(ox0 + i * EW);
#pragma unroll
        for (int i = 0; i < WL; i++) keep[NW + i] = word_mask(ox0 - (WL - i) * EW);
#pragma unroll
        for (int i = 0; i < WR; i++) keep[NW + WL + i] = word_mask(ox0 + NC + i * EW);
#pragma unroll
        for (int i = 0; i < NW + NE; i++) partial = partial || (keep[i] != 0xffffffffu);
    }
    // rows this thread may load: inside the image and not beyond the last row its strip needs
    const unsigned in_h_eff = feeds ? (unsigned)max(0, min(p.in_h, oy_end - p.pady0 + FT - 1)) : 0u;
    const int iy_last = oy_end - 1 - p.pady0 + FT - 1;                     // last input row the strip needs (REPL: rows beyond are not fetched)
    // the aligned vector(s) of the own columns that contain at least one valid element
    bool own_ok[NBYTES / 16 >= 2 && NBYTES == 16 * 2 && sizeof(T) == 4 ? 2 : 1];
    constexpr int NSEG = (sizeof(T) == 4 && NBYTES == 32) ? 2 : 1;          // fp32 rows are only 16-byte aligned: two 128-bit loads
    own_ok[0] = ox0 < p.in_w;
    if (NSEG == 2) own_ok[NSEG - 1] = ox0 + 4 < p.in_w;
    bool el_ok[WL > 0 ? WL : 1], er_ok[WR > 0 ? WR : 1];
#pragma unroll
    for (int i = 0; i < WL; i++) el_ok[i] = edge_l && keep[NW + i] != 0u;
#pragma unroll
    for (int i = 0; i < WR; i++) er_ok[i] = edge_r && keep[NW + WL + i] != 0u;
    // running fetch cursor: row iy_f at rp_f (only dereferenced when the row exists)
    int iy_f = oy_begin - p.pady0;
    const T* rp_f = xp + ox0 + (int64_t)iy_f * p.ish;
    auto fetch = [&](uint32_t* w) {
#pragma unroll
        for (int i = 0; i < NW + NE; i++) w[i] = 0u;
        // zero padding: rows outside the image stay zero;  replicate: they are the clamped row
        const bool ok = REPL ? (feeds && iy_f <= iy_last) : ((unsigned)iy_f < in_h_eff);
        const T* rp = REPL ? xp + ox0 + (int64_t)min(max(iy_f, 0), p.in_h - 1) * p.ish : rp_f;
        if (NSEG == 1) {
            if (ok && own_ok[0]) ldg_words<NBYTES>(rp, w);
        } else {
            if (ok && own_ok[0]) ldg_words<16>(rp, w);
            if (ok && own_ok[NSEG - 1]) ldg_words<16>(rp + 4, w + 4);
        }
#pragma unroll
        for (int i = 0; i < WL; i++) if (ok && el_ok[i]) w[NW + i] = __ldg((const uint32_t*)(rp - (WL - i) * EW));
#pragma unroll
        for (int i = 0; i < WR; i++) if (ok && er_ok[i]) w[NW + WL + i] = __ldg((const uint32_t*)(rp + NC + i * EW));
        rp_f += p.ish;
        iy_f++;
    };
    const bool row_end = ox0 + NC >= p.in_w;               // REPL (in_w is a multiple of NC there): the thread owns the last columns
    // in[0 .. NC+FT-2] = input columns ox0 - PX .. ox0 - PX + NC + FT - 2, from the thread's own words and its neighbours'
    auto expand = [&](uint32_t* w, float* in) {
        if (partial) {          // thread-constant; lanes whose vectors lie wholly inside the row skip the masking
#pragma unroll
            for (int i = 0; i < NW + NE; i++) w[i] &= keep[i];
        }
        uint32_t hl[WL > 0 ? WL : 1], hr[WR > 0 ? WR : 1];
#pragma unroll
        for (int i = 0; i < WL; i++) { const uint32_t t = __shfl_up_sync(0xffffffffu, w[NW - WL + i], 1); hl[i] = edge_l ? w[NW + i] : t; }
#pragma unroll
        for (int i = 0; i < WR; i++) { const uint32_t t = __shfl_down_sync(0xffffffffu, w[i], 1); hr[i] = edge_r ? w[NW + WL + i] : t; }
        float own[NC], left[WL * EW > 0 ? WL * EW : 1], right[WR * EW > 0 ? WR * EW : 1];
        if (EW == 2) {
#pragma unroll
            for (int i = 0; i < NW; i++) { const float2 t = __half22float2(*(const __half2*)&w[i]); own[2 * i] = t.x; own[2 * i + 1] = t.y; }
#pragma unroll
            for (int i = 0; i < WL; i++) { const float2 t = __half22float2(*(const __half2*)&hl[i]); left[2 * i] = t.x; left[2 * i + 1] = t.y; }
#pragma unroll
            for (int i = 0; i < WR; i++) { const float2 t = __half22float2(*(const __half2*)&hr[i]); right[2 * i] = t.x; right[2 * i + 1] = t.y; }
        } else {
#pragma unroll
            for (int i = 0; i < NW; i++) own[i] = __uint_as_float(w[i]);
#pragma unroll
            for (int i = 0; i < WL; i++) left[i] = __uint_as_float(hl[i]);
#pragma unroll
            for (int i = 0; i < WR; i++) right[i] = __uint_as_float(hr[i]);
        }
#pragma unroll
        for (int i = 0; i < NL; i++) in[i] = (REPL && cgi == 0) ? own[0] : left[WL * EW - NL + i];
#pragma unroll
        for (int i = 0; i < NC; i++) in[NL + i] = own[i];
#pragma unroll
        for (int i = 0; i < NR; i++) in[NL + NC + i] = (REPL && row_end) ? own[NC - 1] : right[i];
    };
    // the fused addend (noise) of an output row, fetched two rows before it is needed (running cursor oy_a / ap)
    int oy_a = oy_begin;
    const float* ap = p.add ? p.add + (int64_t)n * p.add_sn + (int64_t)oy_begin * p.add_sh + ox0 : nullptr;
    const bool full_store = ox0 + NC <= p.out_w;
    auto fetch_add = [&](float* a) {
#pragma unroll
        for (int i = 0; i < NC; i++) a[i] = 0.f;
        if (active && oy_a < oy_end) {
            if (add_vec && full_store) {
#pragma unroll
                for (int q = 0; q < NC / 4; q++) { const float4 t = __ldg((const float4*)ap + q); a[4 * q] = t.x; a[4 * q + 1] = t.y; a[4 * q + 2] = t.z; a[4 * q + 3] = t.w; }
            } else {
#pragma unroll
                for (int i = 0; i < NC; i++) if (ox0 + i < p.out_w) a[i] = __ldg(ap + i);
            }
        }
        ap += p.add_sh;
        oy_a++;
    };
    const bool has_clamp = p.ep_enable && p.ep_clamp >= 0.f;
    const unsigned emit_rows = active ? (unsigned)(oy_end - oy_begin) : 0u;      // output rows this thread stores
    int orow = -(FT - 1);                                                         // output row (strip-relative) completed by the current input row
    T* op = yp + ox0 + (int64_t)(oy_begin - (FT - 1)) * p.osh;                           // its address (dereferenced only for 0 <= orow < emit_rows)
    auto emit = [&](float* o, const float* addv) {
        if ((unsigned)orow < emit_rows) {
            if (p.add) {
#pragma unroll
                for (int i = 0; i < NC; i++) o[i] = fmaf(addv[i], eg, o[i]);
            }
            if (lrelu) {
#pragma unroll
                for (int i = 0; i < NC; i++) o[i] = fmaxf(o[i], o[i] * alpha);
            }
            if (has_clamp) {
#pragma unroll
                for (int i = 0; i < NC; i++) o[i] = fminf(fmaxf(o[i], -p.ep_clamp), p.ep_clamp);
            }
            if (full_store) {
                uint32_t w[NW];
                if (EW == 2) {
#pragma unroll
                    for (int i = 0; i < NW; i++) { const __half2 h = __floats2half2_rn(o[2 * i], o[2 * i + 1]); w[i] = *(const uint32_t*)&h; }
                } else {
#pragma unroll
                    for (int i = 0; i < NW; i++) w[i] = __float_as_uint(o[i]);
                }
                if (NSEG == 1) stg_words<NBYTES>(op, w);
                else { stg_words<16>(op, w); stg_words<16>(op + 4, w + 4); }
            } else {
#pragma unroll
                for (int i = 0; i < NC; i++) if (ox0 + i < p.out_w) op[i] = from_acc<T, float>(o[i]);
            }
        }
        op += p.osh;
        orow++;
    };

    // input row r (0-based inside the strip) feeds output rows r - ty (ty < FT); output row r - (FT-1) is complete after it.
    // A ring of FT row vectors is kept in flight (a slot is refilled with row r + FT as soon as row r has been unpacked)
    // and the addend of an output row is fetched two rows (FT == 4; otherwise one row) before it is needed.
    uint32_t wq[FT][NW + NE];
    constexpr int AD = (FT == 4) ? 2 : 1;            // addend rows in flight (other tap counts: one row ahead, the addend is L2-resident noise)
    float addq[AD][NC];
    const bool use_add = p.add != nullptr;
#pragma unroll
    for (int rr = 0; rr < FT; rr++) fetch(wq[rr]);
    if (use_add) { fetch_add(addq[AD - 1]); if (AD == 2) fetch_add(addq[0]); }
#pragma unroll 1
    for (int r0 = 0; r0 < nrows_warp; r0 += FT) {
#pragma unroll
        for (int rr = 0; rr < FT; rr++) {
            float in[NC + FT - 1];
            expand(wq[rr], in);
            fetch(wq[rr]);
            if (SEP) {
                float h[NC];
#pragma unroll
                for (int i = 0; i < NC; i++) {
                    float t = k.fx[0] * in[i];
#pragma unroll
                    for (int tx = 1; tx < FT; tx++) t = fmaf(k.fx[tx], in[i + tx], t);
                    h[i] = t;
                }
#pragma unroll
                for (int ty = 0; ty < FT; ty++)
#pragma unroll
                    for (int i = 0; i < NC; i++) acc[(rr - ty + FT) % FT][i] = fmaf(k.fy[ty], h[i], acc[(rr - ty + FT) % FT][i]);
            } else {
#pragma unroll
                for (int ty = 0; ty < FT; ty++)
#pragma unroll
                    for (int tx = 0; tx < FT; tx++)
#pragma unroll
                        for (int i = 0; i < NC; i++) acc[(rr - ty + FT) % FT][i] = fmaf(k.f[ty][tx], in[i + tx], acc[(rr - ty + FT) % FT][i]);
            }
            // output row r - (FT-1) used accumulator slot (rr + 1) % FT
            const bool emitted = orow >= 0;
            emit(acc[(rr + 1) % FT], addq[rr & (AD - 1)]);
            if (use_add && emitted) fetch_add(addq[rr & (AD - 1)]);
#pragma unroll
            for (int i = 0; i < NC; i++) acc[(rr + 1) % FT][i] = acc_init;
        }
    }
}

template <int FT> __device__ __forceinline__ bool blur_rank1(BlurTaps<FT>& k);

template <int FT>
__device__ __forceinline__ bool blur_taps(const UpfirdnArgs& p, BlurTaps<FT>& k) {
    // correlation taps with the gain folded in (fp32 product, like the reference's scaled filter tensor); missing taps = 0
#pragma unroll
    for (int ty = 0; ty < FT; ty++)
#pragma unroll
        for (int tx = 0; tx < FT; tx++) {
            float v = 0.f;
            if (tx < p.fw && ty < p.fh) {
                const int fx = p.flip ? tx : p.fw - 1 - tx, fy = p.flip ? ty : p.fh - 1 - ty;
                v = __ldg(&p.f[fy * p.fsh + fx * p.fsw]) * (float)p.gain;
                if (p.ep_enable) v *= p.ep_gain;          // fused epilogue gain (positive), see blur_body
            }
            k.f[ty][tx] = v;
        }
    return blur_rank1<FT>(k);
}

// rank-1 test: f == fy (x) fx with fx = pivot row / pivot, fy = pivot column; fills k.fx / k.fy
template <int FT>
__device__ __forceinline__ bool blur_rank1(BlurTaps<FT>& k) {
    int pi = 0, pj = 0; float best = -1.f;
#pragma unroll
    for (int ty = 0; ty < FT; ty++)
#pragma unroll
        for (int tx = 0; tx < FT; tx++) if (fabsf(k.f[ty][tx]) > best) { best = fabsf(k.f[ty][tx]); pi = ty; pj = tx; }
    float piv = 1.f, prow[FT], pcol[FT];
#pragma unroll
    for (int ty = 0; ty < FT; ty++)
#pragma unroll
        for (int tx = 0; tx < FT; tx++) {
            if (ty == pi && tx == pj) piv = k.f[ty][tx];
            if (ty == pi) prow[tx] = k.f[ty][tx];
            if (tx == pj) pcol[ty] = k.f[ty][tx];
        }
    const float inv = (best > 0.f) ? 1.f / piv : 0.f;
    bool sep = best > 0.f;
#pragma unroll
    for (int i = 0; i < FT; i++) { k.fx[i] = prow[i] * inv; k.fy[i] = pcol[i]; }
#pragma unroll
    for (int ty = 0; ty < FT; ty++)
#pragma unroll
        for (int tx = 0; tx < FT; tx++) sep = sep && (fabsf(k.f[ty][tx] - k.fy[ty] * k.fx[tx]) <= 1e-6f * best);
    return sep;
}

template <class T, int PX>
__global__ void __launch_bounds__(128, 4) upfirdn2d_blur(UpfirdnArgs p, int cg, int strips, int strip_rows) {
    BlurTaps<4> k;
    if (blur_taps<4>(p, k)) blur_body<T, PX, true, 8, 4, false>(p, k, cg, strips, strip_rows);
    else blur_body<T, PX, false, 8, 4, false>(p, k, cg, strips, strip_rows);
}
// 16 columns per thread (fp16 rows aligned to 32 bytes): half the per-row bookkeeping per output
template <int PX>
__global__ void __launch_bounds__(128, 3) upfirdn2d_blur16(UpfirdnArgs p, int cg, int strips, int strip_rows) {
    BlurTaps<4> k;
    if (blur_taps<4>(p, k)) blur_body<__half, PX, true, 16, 4, false>(p, k, cg, strips, strip_rows);
    else blur_body<__half, PX, false, 16, 4, false>(p, k, cg, strips, strip_rows);
}
// replicate-padded blur (the fixed binomial blur behind the pixel-shuffle upsampler, networks/utils/convnext_utils.py:250-255):
// <= 4 taps with left pad PX4 in the FT = 4 body, 5 taps (pad 2) in the FT = 5 body; separable filters only take the fast path
template <class T, int FT, int PX>
__global__ void __launch_bounds__(128, 3) upfirdn2d_blur_repl(UpfirdnArgs p, int cg, int strips, int strip_rows) {
    BlurTaps<FT> k;
    if (blur_taps<FT>(p, k)) blur_body<T, PX, true, 8, FT, true>(p, k, cg, strips, strip_rows);
    else blur_body<T, PX, false, 8, FT, true>(p, k, cg, strips, strip_rows);
}

// Depthwise k x k conv (k = 3, 5, 7; stride 1, "same" zero padding) with learned per-channel taps + bias: the dwconv of the ConvNeXt
// synthesis layers (networks/utils/convnext_utils.py:99,128).  It is the same streaming kernel -- a k-row ring of partially
// accumulated output rows in registers -- with the taps of the thread's channel in registers (49 FMAs per output: FMA-issue
// bound, ~1.5 ms for [64,128,256,256] fp16 against ~20 ms of the stock depthwise kernel).
template <class T, int FT, int PX>
__global__ void __launch_bounds__(128, 2) upfirdn2d_dw(UpfirdnArgs p, int cg, int strips, int strip_rows) {
    const int64_t gid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t plane = (gid / cg) / strips;
    const int c = (plane < (int64_t)p.channels * p.batch) ? (int)(plane % p.channels) : 0;
    const float* fp = p.f + (int64_t)c * p.fsc;
    BlurTaps<FT> k;
#pragma unroll
    for (int ty = 0; ty < FT; ty++)
#pragma unroll
        for (int tx = 0; tx < FT; tx++) {
            float v = 0.f;
            if (tx < p.fw && ty < p.fh) {
                const int fx = p.flip ? tx : p.fw - 1 - tx, fy = p.flip ? ty : p.fh - 1 - ty;
                v = __ldg(&fp[fy * p.fsh + fx * p.fsw]) * (float)p.gain;
                if (p.ep_enable) v *= p.ep_gain;
            }
            k.f[ty][tx] = v;
        }
    // a rank-1 5x5 filter (the fixed binomial blur whose data gradient runs through this kernel: _ReplicateBlur.backward) takes the separable
    // body, 10 instead of 25 FMAs per output; learned depthwise filters are not rank-1 and keep the dense one
    if (FT == 5 && blur_rank1<FT>(k)) blur_body<T, PX, true, 8, FT, false>(p, k, cg, strips, strip_rows);
    else blur_body<T, PX, false, 8, FT, false>(p, k, cg, strips, strip_rows);
}

template <class T>
int launch_dw(const UpfirdnArgs& a, cudaStream_t stream) {
    const int groups = ceil_div(a.out_w, 8);
    const int64_t planes = (int64_t)a.channels * a.batch;
    int strip_rows = 64;
    while (strip_rows > 8 && planes * ceil_div(a.out_h, strip_rows) * groups < (int64_t)kNumSMs * 1024) strip_rows >>= 1;
    const int strips = ceil_div(a.out_h, strip_rows);
    const int64_t blocks = ceil_div64(planes * strips * groups, 128);
    if (blocks > 0x7fffffffLL) { set_error("upfirdn2d: grid too large"); return VFM_ERR_INVALID; }
    KernelTimer timer("depthwise_conv", stream, 0.0, ((double)a.in_w * a.in_h + (double)a.out_w * a.out_h) * a.channels * a.batch * sizeof(T),
                      "k%dw%dc%d", a.fw, a.out_w, a.channels);
    if (a.fw == 7) upfirdn2d_dw<T, 7, 3><<<(unsigned)blocks, 128, 0, stream>>>(a, groups, strips, strip_rows);
    else if (a.fw == 5) upfirdn2d_dw<T, 5, 2><<<(unsigned)blocks, 128, 0, stream>>>(a, groups, strips, strip_rows);
    else upfirdn2d_dw<T, 3, 1><<<(unsigned)blocks, 128, 0, stream>>>(a, groups, strips, strip_rows);
    return launch_status("upfirdn2d_dw");
}

template <class T>
int launch_blur(const UpfirdnArgs& a, cudaStream_t stream) {
    // 16 columns per thread when the rows are fp16, 32-byte aligned and wide enough to keep a warp busy
    bool wide = false;
    if (sizeof(T) == 2 && !a.pad_mode) {
        auto al32 = [](const void* q) { return (reinterpret_cast<uintptr_t>(q) & 31u) == 0; };
        wide = al32(a.x) && al32(a.y) && (a.ish % 16) == 0 && (a.isc % 16) == 0 && (a.isn % 16) == 0 &&
               (a.osh % 16) == 0 && (a.osc % 16) == 0 && (a.osn % 16) == 0 && a.out_w >= 256;   // measured: pays off only for full-warp rows
    }
    const int nc = wide ? 16 : 8;
    const int groups = ceil_div(a.out_w, nc);
    // strips: enough threads to fill the machine, few enough that the 3-row halo stays cheap
    const int64_t planes = (int64_t)a.channels * a.batch;
    int strip_rows = 64;
    while (strip_rows > 8 && planes * ceil_div(a.out_h, strip_rows) * groups < (int64_t)kNumSMs * 2048) strip_rows >>= 1;
    const int strips = ceil_div(a.out_h, strip_rows);
    const int64_t threads = planes * strips * groups;
    const int64_t blocks = ceil_div64(threads, 128);
    if (blocks > 0x7fffffffLL) { set_error("upfirdn2d: grid too large"); return VFM_ERR_INVALID; }
    KernelTimer timer(a.pad_mode ? "upfirdn2d_blur_repl" : "upfirdn2d_blur", stream, 0.0,
                      ((double)a.in_w * a.in_h + (double)a.out_w * a.out_h) * a.channels * a.batch * sizeof(T) + (double)a.fw * a.fh * 4,
                      "w%dc%d", a.out_w, a.channels);
    if (a.pad_mode) {
        const int ft = (a.fw > 4 || a.fh > 4) ? 5 : 4;
        if (ft == 5 && a.padx0 == 2) upfirdn2d_blur_repl<T, 5, 2><<<(unsigned)blocks, 128, 0, stream>>>(a, groups, strips, strip_rows);
        else if (ft == 4 && a.padx0 == 1) upfirdn2d_blur_repl<T, 4, 1><<<(unsigned)blocks, 128, 0, stream>>>(a, groups, strips, strip_rows);
        else { set_error("upfirdn2d: replicate padding supports 3/4 taps with left pad 1 and 5 taps with left pad 2"); return VFM_ERR_NO_KERNEL; }
        return launch_status("upfirdn2d_blur_repl");
    }
    if (wide) {
        switch (a.padx0) {
            case 0: upfirdn2d_blur16<0><<<(unsigned)blocks, 128, 0, stream>>>(a, groups, strips, strip_rows); break;
            case 1: upfirdn2d_blur16<1><<<(unsigned)blocks, 128, 0, stream>>>(a, groups, strips, strip_rows); break;
            case 2: upfirdn2d_blur16<2><<<(unsigned)blocks, 128, 0, stream>>>(a, groups, strips, strip_rows); break;
            default: upfirdn2d_blur16<3><<<(unsigned)blocks, 128, 0, stream>>>(a, groups, strips, strip_rows); break;
        }
        return launch_status("upfirdn2d_blur16");
    }
    switch (a.padx0) {
        case 0: upfirdn2d_blur<T, 0><<<(unsigned)blocks, 128, 0, stream>>>(a, groups, strips, strip_rows); break;
        case 1: upfirdn2d_blur<T, 1><<<(unsigned)blocks, 128, 0, stream>>>(a, groups, strips, strip_rows); break;
        case 2: upfirdn2d_blur<T, 2><<<(unsigned)blocks, 128, 0, stream>>>(a, groups, strips, strip_rows); break;
        default: upfirdn2d_blur<T, 3><<<(unsigned)blocks, 128, 0, stream>>>(a, groups, strips, strip_rows); break;
    }
    return launch_status("upfirdn2d_blur");
}

template <class T, int UP, int DOWN, int FW, int FH>
constexpr bool tiled_fits() {
    typedef TileCfg<UP, DOWN, FW, FH> C;
    return sizeof(typename Acc<T>::type) * (C::IH * C::IWP + FW * FH) <= 48 * 1024;
}

template <class T, int UP, int DOWN, int FW, int FH>
int launch_tiled(UpfirdnArgs a, cudaStream_t stream) {
    typedef TileCfg<UP, DOWN, FW, FH> C;
    static_assert(tiled_fits<T, UP, DOWN, FW, FH>(), "static shared memory overflow");
    // align the tile grid to the up-sampling phase: (xstart*DOWN - pad0) % UP == 0, xstart in (-UP, 0]
    auto start = [](int pad0) { int m = ((pad0 % UP) + UP) % UP; return (UP == 1 || m == 0) ? 0 : m - UP; };
    a.xstart = (DOWN == 1) ? start(a.padx0) : 0;
    a.ystart = (DOWN == 1) ? start(a.pady0) : 0;
    a.tiles_x = ceil_div(a.out_w - a.xstart, C::TW);
    a.tiles_y = ceil_div(a.out_h - a.ystart, C::TH);
    int64_t blocks = (int64_t)a.tiles_x * a.tiles_y * a.channels * a.batch;
    if (blocks > 0x7fffffffLL) { set_error("upfirdn2d: grid too large"); return VFM_ERR_INVALID; }
    KernelTimer timer("upfirdn2d_tiled", stream, 0.0,
                      ((double)a.in_w * a.in_h + (double)a.out_w * a.out_h) * a.channels * a.batch * sizeof(T) + (double)a.fw * a.fh * 4);
    upfirdn2d_tiled<T, UP, DOWN, FW, FH><<<(unsigned)blocks, 256, 0, stream>>>(a);
    return launch_status("upfirdn2d_tiled");
}

template <class T>
int launch(UpfirdnArgs a, cudaStream_t stream) {
    bool wcontig = (a.isw == 1 && a.osw == 1);
    bool sym = (a.upx == a.upy && a.downx == a.downy);
    if constexpr (sizeof(T) <= 4) {
        // streaming blur: up = down = 1, <= 4x4 taps, 16-byte aligned rows, and the left padding inside the halo it handles
        const int es = (int)sizeof(T);
        const bool rows16 = a.vec_ok && aligned16(a.y) && (a.osh * es) % 16 == 0 && (a.osc * es) % 16 == 0 && (a.osn * es) % 16 == 0;
        if (a.fsc != 0) {
            // depthwise conv with per-channel taps: k in {3, 5, 7}, same-size output
            if (wcontig && a.upx == 1 && a.upy == 1 && a.downx == 1 && a.downy == 1 && a.fw == a.fh && (a.fw == 3 || a.fw == 5 || a.fw == 7) && rows16 &&
                !a.pad_mode && a.padx0 == a.fw / 2 && a.pady0 == a.fh / 2 && a.out_w == a.in_w && a.out_h == a.in_h)
                return launch_dw<T>(a, stream);
            set_error("upfirdn2d: per-channel filters are only implemented for same-size 3x3 / 5x5 / 7x7 depthwise convs of 16-byte aligned rows");
            return VFM_ERR_NO_KERNEL;
        }
        if (a.pad_mode) {
            // replicate padding: same-size output, rows of whole 8-column groups
            if (wcontig && a.upx == 1 && a.upy == 1 && a.downx == 1 && a.downy == 1 && a.fw <= 5 && a.fh <= 5 && rows16 && !a.add && !a.ep_enable &&
                a.out_w == a.in_w && a.out_h == a.in_h && a.in_w % 8 == 0)
                return launch_blur<T>(a, stream);
            set_error("upfirdn2d: replicate padding is only implemented for same-size blurs of 16-byte aligned rows");
            return VFM_ERR_NO_KERNEL;
        }
        if (wcontig && a.upx == 1 && a.upy == 1 && a.downx == 1 && a.downy == 1 && a.fw <= 4 && a.fh <= 4 && rows16 &&
            a.padx0 >= 0 && a.padx0 <= 3 && (!a.add || a.add_sh >= a.out_w) &&
            (!a.ep_enable || (a.ep_gain > 0.f && (a.ep_act != 3 || (a.ep_alpha >= 0.f && a.ep_alpha <= 1.f)))))
            return launch_blur<T>(a, stream);
    }
    if (a.pad_mode || a.fsc != 0) { set_error("upfirdn2d: replicate padding / per-channel filters need fp16/fp32"); return VFM_ERR_NO_KERNEL; }
    if (wcontig && sym && a.out_w >= 32 && a.out_h >= 8) {
        int up = a.upx, down = a.downx;
        if (up == 1 && down == 1 && a.fw <= 4 && a.fh <= 4) return launch_tiled<T, 1, 1, 4, 4>(a, stream);
        if (up == 2 && down == 1 && a.fw <= 4 && a.fh <= 4) return launch_tiled<T, 2, 1, 4, 4>(a, stream);
        if constexpr (tiled_fits<T, 1, 2, 4, 4>()) {
            if (up == 1 && down == 2 && a.fw <= 4 && a.fh <= 4) return launch_tiled<T, 1, 2, 4, 4>(a, stream);
        }
    }
    int64_t total = (int64_t)a.out_w * a.out_h * a.channels * a.batch;
    int c_fastest = (a.osc == 1 && a.channels > 1) ? 1 : 0;
    int64_t blocks = ceil_div64(total, 256);
    int64_t cap = (int64_t)kNumSMs * 32;
    if (blocks > cap) blocks = cap;
    KernelTimer timer("upfirdn2d_generic", stream, 0.0,
                      ((double)a.in_w * a.in_h + (double)a.out_w * a.out_h) * a.channels * a.batch * sizeof(T) + (double)a.fw * a.fh * 4);
    upfirdn2d_generic<T><<<(unsigned)blocks, 256, 0, stream>>>(a, total, c_fastest);
    return launch_status("upfirdn2d_generic");
}

}  // namespace
}  // namespace vfm

extern "C" int vfm_upfirdn2d(const vfm_upfirdn2d_params* p, void* stream_) {
    using namespace vfm;
    cudaStream_t stream = (cudaStream_t)stream_;
    VFM_CHECK_ARG(p != nullptr, "upfirdn2d: params is NULL");
    VFM_CHECK_ARG(p->x && p->y && p->f, "upfirdn2d: x, f and y must be non-NULL");
    VFM_CHECK_ARG(p->upx >= 1 && p->upy >= 1, "upfirdn2d: upsampling factor must be at least 1");
    VFM_CHECK_ARG(p->downx >= 1 && p->downy >= 1, "upfirdn2d: downsampling factor must be at least 1");
    VFM_CHECK_ARG(p->fw >= 1 && p->fh >= 1, "upfirdn2d: f must be at least 1x1");
    VFM_CHECK_ARG(p->in_w >= 1 && p->in_h >= 1 && p->channels >= 1 && p->batch >= 1, "upfirdn2d: x has zero size");
    VFM_CHECK_ARG(p->out_w >= 1 && p->out_h >= 1, "upfirdn2d: output must be at least 1x1");
    VFM_CHECK_ARG(p->dtype == VFM_F16 || p->dtype == VFM_F32 || p->dtype == VFM_F64, "upfirdn2d: unsupported dtype %d", p->dtype);
    UpfirdnArgs a;
    a.x = p->x; a.f = p->f; a.y = p->y; a.add = p->add;
    a.upx = p->upx; a.upy = p->upy; a.downx = p->downx; a.downy = p->downy;
    a.padx0 = p->padx0; a.pady0 = p->pady0; a.flip = p->flip ? 1 : 0; a.gain = p->gain;
    a.in_w = p->in_w; a.in_h = p->in_h; a.channels = p->channels; a.batch = p->batch;
    a.isw = p->in_stride_w; a.ish = p->in_stride_h; a.isc = p->in_stride_c; a.isn = p->in_stride_n;
    a.fw = p->fw; a.fh = p->fh; a.fsw = p->f_stride_w; a.fsh = p->f_stride_h;
    a.out_w = p->out_w; a.out_h = p->out_h;
    a.osw = p->out_stride_w; a.osh = p->out_stride_h; a.osc = p->out_stride_c; a.osn = p->out_stride_n;
    a.add_sh = p->add_stride_h; a.add_sn = p->add_stride_n;
    a.ep_enable = p->ep_enable; a.ep_act = p->ep_act; a.ep_alpha = (float)p->ep_alpha; a.ep_gain = (float)p->ep_gain; a.ep_clamp = (float)p->ep_clamp; a.ep_bias = p->ep_bias;
    VFM_CHECK_ARG(!p->ep_enable || p->ep_act == 1 || p->ep_act == 3, "upfirdn2d: fused epilogue supports linear and lrelu only");
    a.pad_mode = p->pad_mode;
    a.fsc = p->f_stride_c;
    a.xstart = a.ystart = a.tiles_x = a.tiles_y = 0;
    {
        const int64_t es = (p->dtype == VFM_F16) ? 2 : (p->dtype == VFM_F32 ? 4 : 8);
        a.vec_ok = aligned16(p->x) && (a.ish * es) % 16 == 0 && (a.isc * es) % 16 == 0 && (a.isn * es) % 16 == 0;
    }
    switch (p->dtype) {
        case VFM_F16: return launch<__half>(a, stream);
        case VFM_F32: return launch<float>(a, stream);
        default:      return launch<double>(a, stream);
    }
}


// ---------------------------------------------------------------------------------------------------------------------------
// PixelShuffle(2): y[n, c, 2h + i, 2w + j] = x[n, 4c + 2i + j, h, w]  (the upsampling step of SeparableUpsampleWithFixedBlur,
// networks/utils/convnext_utils.py:197-257).  One thread = 8 output pixels of one row: 4 + 4 inputs of the two source planes of
// that row parity, interleaved in registers -- every access is a full 8/16-byte vector, so the copy runs at the HBM rate (the
// generic 6-d permute copy it replaces reaches 1.3 TB/s).
namespace vfm {
namespace {

template <class T, bool INV>
__global__ void __launch_bounds__(256) pixel_shuffle2_kernel(const T* __restrict__ x, T* __restrict__ y, int64_t planes, int H, int W) {
    // INV = false: x = the 4-plane tensor [planes*4, H, W], y = the shuffled one [planes, 2H, 2W];  INV = true: the other way round
    // (PixelUnshuffle(2) = the backward of the shuffle)
    constexpr int NB = 4 * (int)sizeof(T);                 // bytes of the 4 elements a thread moves per source plane
    const int gpr = W / 4;                                 // groups of 8 shuffled pixels per shuffled row
    const int64_t total = planes * (2 * H) * gpr;
    for (int64_t gid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; gid < total; gid += (int64_t)gridDim.x * blockDim.x) {
        const int g = (int)(gid % gpr);
        const int64_t r = gid / gpr;
        const int oy = (int)(r % (2 * H));
        const int64_t pl = r / (2 * H);                    // n * C + c
        const int64_t o0 = ((pl * 4 + (oy & 1) * 2) * H + (oy >> 1)) * (int64_t)W + g * 4;      // plane 2i, then plane 2i + 1 at + H*W
        const int64_t o1 = o0 + (int64_t)H * W;
        const int64_t os = (pl * (2 * H) + oy) * (int64_t)(2 * W) + g * 8;
        uint32_t a[NB / 4], b[NB / 4], o[NB / 2];
        if (!INV) {
            if (NB == 8) {
                const uint2 ua = __ldg((const uint2*)(x + o0)), ub = __ldg((const uint2*)(x + o1));
                a[0] = ua.x; a[1] = ua.y; b[0] = ub.x; b[1] = ub.y;
            } else {
                ldg_words<16>(x + o0, a);
                ldg_words<16>(x + o1, b);
            }
            if (sizeof(T) == 2) {
#pragma unroll
                for (int i = 0; i < NB / 4; i++) {
                    o[2 * i] = __byte_perm(a[i], b[i], 0x5410);          // a.lo, b.lo
                    o[2 * i + 1] = __byte_perm(a[i], b[i], 0x7632);      // a.hi, b.hi
                }
            } else {
#pragma unroll
                for (int i = 0; i < NB / 4; i++) { o[2 * i] = a[i]; o[2 * i + 1] = b[i]; }
            }
            if (sizeof(T) == 2) stg_words<16>(y + os, o);
            else { stg_words<16>(y + os, o); stg_words<16>(y + os + 4, o + 4); }
        } else {
            if (sizeof(T) == 2) ldg_words<16>(x + os, o);
            else { ldg_words<16>(x + os, o); ldg_words<16>(x + os + 4, o + 4); }
            if (sizeof(T) == 2) {
#pragma unroll
                for (int i = 0; i < NB / 4; i++) {
                    a[i] = __byte_perm(o[2 * i], o[2 * i + 1], 0x5410);  // even pixels
                    b[i] = __byte_perm(o[2 * i], o[2 * i + 1], 0x7632);  // odd pixels
                }
            } else {
#pragma unroll
                for (int i = 0; i < NB / 4; i++) { a[i] = o[2 * i]; b[i] = o[2 * i + 1]; }
            }
            if (NB == 8) {
                *(uint2*)(y + o0) = make_uint2(a[0], a[1]);
                *(uint2*)(y + o1) = make_uint2(b[0], b[1]);
            } else {
                stg_words<16>(y + o0, a);
                stg_words<16>(y + o1, b);
            }
        }
    }
}

// ---------------------------------------------------------------------------------------------------------------------------
// Weight / bias gradient of the depthwise k x k conv:  dweight[c,ty,tx] = sum_{n,y,x} dy[n,c,y,x] * x[n,c,y+ty-P,x+tx-P],
// dbias[c] = sum dy.  A thread owns 8 columns of a strip of rows of one (n, c) plane: per row one vector of dy and one of x
// (requested one row ahead) plus the halo columns (L1 hits: the neighbouring threads' vectors), the k most recent x rows
// expanded in registers, k*k running sums; block reduction by shuffles + shared memory, one atomicAdd per (block, tap).
template <class T, int K>
__global__ void __launch_bounds__(128) dw_wgrad_kernel(const T* __restrict__ x, const T* __restrict__ dy, float* __restrict__ dwt, float* __restrict__ db,
                                                        int N, int C, int H, int W, int strips, int strip_rows) {
    constexpr int P = K / 2;
    constexpr int EW = (int)(4 / sizeof(T));
    constexpr int NW = 8 / EW;                              // words of 8 elements
    // work items of a channel = (sample, strip of rows, group of 8 columns), flattened over the blocks of the channel so that small
    // images still fill the CTAs (the reduction below only needs all threads of a CTA to share the channel)
    const int c = blockIdx.y;
    const int cg = W >> 3;                                  // column groups per row (W % 8 == 0)
    const int64_t item = (int64_t)blockIdx.x * 128 + threadIdx.x;
    const bool active = item < (int64_t)N * strips * cg;
    const int g = (int)(item % cg);
    const int strip = (int)((item / cg) % strips);
    const int n = active ? (int)(item / ((int64_t)cg * strips)) : 0;
    const int y0 = strip * strip_rows;
    const int y1 = min(y0 + strip_rows, H);
    const int x0 = g * 8;
    const T* xp = x + ((int64_t)n * C + c) * H * (int64_t)W;
    const T* dp = dy + ((int64_t)n * C + c) * H * (int64_t)W;

    float acc[K][K], accb = 0.f;
#pragma unroll
    for (int a = 0; a < K; a++)
#pragma unroll
        for (int b = 0; b < K; b++) acc[a][b] = 0.f;

    struct Raw { uint32_t w[NW]; T hl[P], hr[P]; };
    auto fetch_x = [&](int iy, Raw& r) {
#pragma unroll
        for (int i = 0; i < NW; i++) r.w[i] = 0u;
#pragma unroll
        for (int i = 0; i < P; i++) { r.hl[i] = from_acc<T, float>(0.f); r.hr[i] = from_acc<T, float>(0.f); }
        if (iy >= 0 && iy < H) {
            const T* rp = xp + (int64_t)iy * W + x0;
            if (sizeof(T) == 2) ldg_words<16>(rp, r.w);
            else { ldg_words<16>(rp, r.w); ldg_words<16>(rp + 4, r.w + 4); }
#pragma unroll
            for (int i = 0; i < P; i++) {
                if (x0 - P + i >= 0) r.hl[i] = __ldg(rp - P + i);
                if (x0 + 8 + i < W) r.hr[i] = __ldg(rp + 8 + i);
            }
        }
    };
    auto expand8 = [&](const uint32_t* w, float* o) {
        if (EW == 2) {
#pragma unroll
            for (int i = 0; i < NW; i++) { const float2 t = __half22float2(*(const __half2*)&w[i]); o[2 * i] = t.x; o[2 * i + 1] = t.y; }
        } else {
#pragma unroll
            for (int i = 0; i < NW; i++) o[i] = __uint_as_float(w[i]);
        }
    };
    auto expand_x = [&](const Raw& r, float* o) {          // o[j] = x[iy][x0 - P + j], j < 8 + 2P
#pragma unroll
        for (int i = 0; i < P; i++) { o[i] = to_acc(r.hl[i]); o[P + 8 + i] = to_acc(r.hr[i]); }
        expand8(r.w, o + P);
    };

    if (active) {
        float ring[K][8 + K - 1];
        // rows y0 - P .. y0 + P - 1 -> slots 0 .. K-2;  row y + P goes to slot (y - y0 + K - 1) % K
#pragma unroll
        for (int j = 0; j < K - 1; j++) { Raw r; fetch_x(y0 - P + j, r); expand_x(r, ring[j]); }
        Raw nx;
        uint32_t nd[NW];
        fetch_x(y0 + P, nx);
        if (sizeof(T) == 2) ldg_words<16>(dp + (int64_t)y0 * W + x0, nd);
        else { ldg_words<16>(dp + (int64_t)y0 * W + x0, nd); ldg_words<16>(dp + (int64_t)y0 * W + x0 + 4, nd + 4); }
#pragma unroll 1
        for (int yb = y0; yb < y1; yb += K) {
#pragma unroll
            for (int u = 0; u < K; u++) {
                const int y = yb + u;
                if (y < y1) {
                    float d[8];
                    expand_x(nx, ring[(u + K - 1) % K]);
                    expand8(nd, d);
                    if (y + 1 < y1) {                       // next row's vectors are in flight during this row's FMAs
                        fetch_x(y + 1 + P, nx);
                        const T* q = dp + (int64_t)(y + 1) * W + x0;
                        if (sizeof(T) == 2) ldg_words<16>(q, nd);
                        else { ldg_words<16>(q, nd); ldg_words<16>(q + 4, nd + 4); }
                    }
#pragma unroll
                    for (int i = 0; i < 8; i++) accb += d[i];
                    // column outermost: consecutive FMAs go to K*K different sums (with i innermost each sum is a dependent chain of 8)
#pragma unroll
                    for (int i = 0; i < 8; i++)
#pragma unroll
                        for (int ty = 0; ty < K; ty++)
#pragma unroll
                            for (int tx = 0; tx < K; tx++) acc[ty][tx] = fmaf(d[i], ring[(u + ty) % K][i + tx], acc[ty][tx]);
                }
            }
        }
    }
    // block reduction of the K*K + 1 sums
    __shared__ float red[4][K * K + 1];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int a = 0; a < K; a++)
#pragma unroll
        for (int b = 0; b < K; b++) {
            float v = acc[a][b];
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
            if (lane == 0) red[warp][a * K + b] = v;
        }
    {
        float v = accb;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        if (lane == 0) red[warp][K * K] = v;
    }
    __syncthreads();
    if (threadIdx.x <= K * K) {
        const float v = red[0][threadIdx.x] + red[1][threadIdx.x] + red[2][threadIdx.x] + red[3][threadIdx.x];
        if (threadIdx.x < K * K) atomicAdd(&dwt[(int64_t)c * K * K + threadIdx.x], v);
        else if (db) atomicAdd(&db[c], v);
    }
}

// ---------------------------------------------------------------------------------------------------------------------------
// Border of the data gradient of the replicate-padded fixed blur (see vfm_replicate_blur_edges in vfm_ops.h).  One thread per border
// element of a plane: 2W (first / last row) + 2(H-2) (first / last column of the rows in between).  For output j and a contributing dy
// position i = j + a of one axis, the taps that land on j after clamping form a RANGE that depends only on the class of j (first /
// interior / last) and on a: the single tap p - a in the interior, [0, p - a] on the first row / column (a >= 0), [p - a, k-1] on the
// last (a <= 0).  The CTA builds the table G[cy][a][cx][b] = sum of f over the two ranges once (3k x 3k entries) and every thread then
// does (2p+1)^2 guarded loads and FMAs.  Replaces four grouped conv_transpose2d calls (0.4 ms each in cuDNN for [64,128,2,256] slices)
// plus ~20 slice kernels of a host-side fold.
template <class T, int K>
__global__ void __launch_bounds__(256) replicate_blur_edges_kernel(const T* __restrict__ dy, T* __restrict__ dx, const float* __restrict__ f, int64_t planes, int H, int W) {
    constexpr int P = K / 2;
    __shared__ float G[3 * K][3 * K];              // [class_y * K + (a + P)][class_x * K + (b + P)]
    auto range = [](int cls, int a, int& lo, int& hi) {     // tap range of one axis; empty when lo > hi
        if (cls == 1) { lo = hi = P - a; }
        else if (cls == 0) { lo = 0; hi = (a >= 0) ? P - a : -1; }
        else { lo = (a <= 0) ? P - a : K; hi = K - 1; }
        lo = max(lo, 0); hi = min(hi, K - 1);
    };
    for (int i = threadIdx.x; i < 9 * K * K; i += blockDim.x) {
        const int r = i / (3 * K), c = i - r * (3 * K);
        int ty0, ty1, tx0, tx1;
        range(r / K, r % K - P, ty0, ty1);
        range(c / K, c % K - P, tx0, tx1);
        float sum = 0.f;
        for (int t = ty0; t <= ty1; t++) for (int u = tx0; u <= tx1; u++) sum += f[t * K + u];
        G[r][c] = sum;
    }
    __syncthreads();
    const int border = 2 * W + 2 * (H - 2);
    // grid-stride over the border elements: the table above is built once per CTA, not once per 256 elements
    for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < planes * border; idx += (int64_t)gridDim.x * blockDim.x) {
    const int64_t plane = idx / border;
    int e = (int)(idx - plane * border), jy, jx;
    const bool e_col = e >= 2 * W;
    if (e < W) { jy = 0; jx = e; }
    else if (e < 2 * W) { jy = H - 1; jx = e - W; }
    else { e -= 2 * W; jy = 1 + (e >> 1); jx = (e & 1) ? W - 1 : 0; }
    const int cy = (jy == 0) ? 0 : (jy == H - 1 ? 2 : 1), cx = (jx == 0) ? 0 : (jx == W - 1 ? 2 : 1);
    const T* dp = dy + plane * (int64_t)H * W;
    float acc = 0.f;
    if (e_col && (W & 3) == 0 && (reinterpret_cast<uintptr_t>(dy) & 15u) == 0) {
        // first / last column: adjacent lanes sit in different rows, so every scalar load of a warp touches 32 sectors (the kernel was bound
        // by exactly that: 15 loads x 32 sectors per warp).  The <= 3 columns a row contributes lie inside one aligned group of 4: one 8- /
        // 16-byte load per row instead of three scalar ones
        const int cb = (jx == 0) ? 0 : W - 4;
#pragma unroll
        for (int a = -P; a <= P; a++) {
            const int iy = jy + a;
            if (iy < 0 || iy >= H) continue;
            const float* grow = &G[cy * K + a + P][cx * K + P];
            struct alignas(4 * sizeof(T)) { T v[4]; } q;
            q = *reinterpret_cast<const decltype(q)*>(dp + (int64_t)iy * W + cb);
#pragma unroll
            for (int b = -P; b <= P; b++) {
                const int ix = jx + b;
                if (ix < 0 || ix >= W) continue;
                const int o = ix - cb;            // 0..3
                acc = fmaf(grow[b], to_acc(o == 0 ? q.v[0] : o == 1 ? q.v[1] : o == 2 ? q.v[2] : q.v[3]), acc);
            }
        }
    } else {
#pragma unroll
        for (int a = -P; a <= P; a++) {
            const int iy = jy + a;
            if (iy < 0 || iy >= H) continue;
            const float* grow = &G[cy * K + a + P][cx * K + P];
            const T* drow = dp + (int64_t)iy * W + jx;
#pragma unroll
            for (int b = -P; b <= P; b++) {
                const int ix = jx + b;
                if (ix < 0 || ix >= W) continue;
                acc = fmaf(grow[b], to_acc(drow[b]), acc);
            }
        }
    }
    dx[plane * (int64_t)H * W + (int64_t)jy * W + jx] = from_acc<T, float>(acc);
    }
}

}  // namespace
}  // namespace vfm

extern "C" int vfm_replicate_blur_edges(const vfm_replicate_blur_edges_params* p, void* stream_) {
    using namespace vfm;
    cudaStream_t stream = (cudaStream_t)stream_;
    VFM_CHECK_ARG(p != nullptr && p->dy && p->dx && p->f, "replicate_blur_edges: NULL argument");
    VFM_CHECK_ARG(p->dtype == VFM_F16 || p->dtype == VFM_F32, "replicate_blur_edges: fp16 / fp32 only");
    VFM_CHECK_ARG(p->k == 3 || p->k == 5, "replicate_blur_edges: k must be 3 or 5 (got %d)", p->k);
    VFM_CHECK_ARG(p->planes >= 1 && p->h >= p->k && p->w >= p->k, "replicate_blur_edges: planes must be at least k x k");
    const int64_t total = p->planes * (2 * (int64_t)p->w + 2 * ((int64_t)p->h - 2));
    const unsigned blocks = (unsigned)std::min<int64_t>(ceil_div64(total, 256), (int64_t)kNumSMs * 16);
    const int es = p->dtype == VFM_F16 ? 2 : 4;
    KernelTimer timer("replicate_blur_edges", stream, 0.0, (double)total * es * 2.0, "h%dw%d", p->h, p->w);
    if (p->dtype == VFM_F16) {
        if (p->k == 3) replicate_blur_edges_kernel<__half, 3><<<blocks, 256, 0, stream>>>((const __half*)p->dy, (__half*)p->dx, p->f, p->planes, p->h, p->w);
        else replicate_blur_edges_kernel<__half, 5><<<blocks, 256, 0, stream>>>((const __half*)p->dy, (__half*)p->dx, p->f, p->planes, p->h, p->w);
    } else {
        if (p->k == 3) replicate_blur_edges_kernel<float, 3><<<blocks, 256, 0, stream>>>((const float*)p->dy, (float*)p->dx, p->f, p->planes, p->h, p->w);
        else replicate_blur_edges_kernel<float, 5><<<blocks, 256, 0, stream>>>((const float*)p->dy, (float*)p->dx, p->f, p->planes, p->h, p->w);
    }
    return launch_status("replicate_blur_edges_kernel");
}

extern "C" int vfm_pixel_shuffle2(const vfm_pixel_shuffle2_params* p, void* stream_) {
    using namespace vfm;
    cudaStream_t stream = (cudaStream_t)stream_;
    VFM_CHECK_ARG(p != nullptr && p->x && p->y, "pixel_shuffle2: NULL argument");
    VFM_CHECK_ARG(p->dtype == VFM_F16 || p->dtype == VFM_F32, "pixel_shuffle2: fp16 / fp32 only");
    VFM_CHECK_ARG(p->batch >= 1 && p->out_channels >= 1 && p->in_h >= 1 && p->in_w >= 4 && p->in_w % 4 == 0, "pixel_shuffle2: in_w must be a positive multiple of 4");
    VFM_CHECK_ARG(aligned16(p->x) && aligned16(p->y), "pixel_shuffle2: x and y must be 16-byte aligned");
    const int64_t planes = (int64_t)p->batch * p->out_channels;
    const int64_t total = planes * 2 * p->in_h * (p->in_w / 4);
    const unsigned blocks = (unsigned)std::min<int64_t>(ceil_div64(total, 256), (int64_t)kNumSMs * 32);
    const int es = p->dtype == VFM_F16 ? 2 : 4;
    KernelTimer timer(p->inverse ? "pixel_unshuffle2" : "pixel_shuffle2", stream, 0.0, 2.0 * (double)planes * 4 * p->in_h * p->in_w * es, "c%dh%d", p->out_channels, p->in_h);
    if (p->dtype == VFM_F16) {
        if (p->inverse) pixel_shuffle2_kernel<__half, true><<<blocks, 256, 0, stream>>>((const __half*)p->x, (__half*)p->y, planes, p->in_h, p->in_w);
        else pixel_shuffle2_kernel<__half, false><<<blocks, 256, 0, stream>>>((const __half*)p->x, (__half*)p->y, planes, p->in_h, p->in_w);
    } else {
        if (p->inverse) pixel_shuffle2_kernel<float, true><<<blocks, 256, 0, stream>>>((const float*)p->x, (float*)p->y, planes, p->in_h, p->in_w);
        else pixel_shuffle2_kernel<float, false><<<blocks, 256, 0, stream>>>((const float*)p->x, (float*)p->y, planes, p->in_h, p->in_w);
    }
    return launch_status("pixel_shuffle2");
}

namespace vfm {
namespace {
template <class T, int K>
int launch_dw_wgrad(const vfm_depthwise_wgrad_params* p, cudaStream_t stream) {
    const int cg = p->w / 8;
    int strip_rows = 32;
    while (strip_rows > 8 && (int64_t)p->batch * p->channels * ceil_div(p->h, strip_rows) * cg < (int64_t)kNumSMs * 2048) strip_rows >>= 1;
    const int strips = ceil_div(p->h, strip_rows);
    const int64_t bx = ceil_div64((int64_t)p->batch * strips * cg, 128);
    if (bx > 0x7fffffffLL) { set_error("depthwise_wgrad: grid too large"); return VFM_ERR_INVALID; }
    dim3 grid((unsigned)bx, (unsigned)p->channels, 1);
    KernelTimer timer("depthwise_wgrad", stream, 0.0, 2.0 * (double)p->batch * p->channels * p->h * p->w * sizeof(T), "k%dw%dc%d", K, p->w, p->channels);
    dw_wgrad_kernel<T, K><<<grid, 128, 0, stream>>>((const T*)p->x, (const T*)p->dy, p->dweight, p->dbias, p->batch, p->channels, p->h, p->w, strips, strip_rows);
    return launch_status("depthwise_wgrad");
}
}  // namespace
}  // namespace vfm

extern "C" int vfm_depthwise_wgrad(const vfm_depthwise_wgrad_params* p, void* stream_) {
    using namespace vfm;
    cudaStream_t stream = (cudaStream_t)stream_;
    VFM_CHECK_ARG(p != nullptr && p->x && p->dy && p->dweight, "depthwise_wgrad: NULL argument");
    VFM_CHECK_ARG(p->dtype == VFM_F16 || p->dtype == VFM_F32, "depthwise_wgrad: fp16 / fp32 only");
    VFM_CHECK_ARG(p->k == 3 || p->k == 5 || p->k == 7, "depthwise_wgrad: k must be 3, 5 or 7");
    VFM_CHECK_ARG(p->batch >= 1 && p->channels >= 1 && p->channels <= 65535 && p->h >= 1, "depthwise_wgrad: bad shape");
    if (p->w < 8 || p->w % 8 != 0 || p->w > 1024 || !aligned16(p->x) || !aligned16(p->dy)) {
        set_error("depthwise_wgrad: needs 16-byte aligned tensors with 8 <= W <= 1024, W %% 8 == 0"); return VFM_ERR_NO_KERNEL;
    }
    // dweight / dbias are accumulated with atomics: the caller passes zero-initialised buffers (or accumulates on purpose)
    if (p->dtype == VFM_F16) {
        switch (p->k) { case 3: return launch_dw_wgrad<__half, 3>(p, stream); case 5: return launch_dw_wgrad<__half, 5>(p, stream); default: return launch_dw_wgrad<__half, 7>(p, stream); }
    }
    switch (p->k) { case 3: return launch_dw_wgrad<float, 3>(p, stream); case 5: return launch_dw_wgrad<float, 5>(p, stream); default: return launch_dw_wgrad<float, 7>(p, stream); }
}
