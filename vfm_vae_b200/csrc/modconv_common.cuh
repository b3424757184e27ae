// Shared pieces of the modulated-conv implementation: the small "coefficient" kernels (pre-normalisation, demodulation
// coefficients, gradient fix-ups) and the plan/workspace layout used by both the generic SIMT path and the tcgen05 path.
#pragma once
#include "common.cuh"

namespace vfm {
namespace modconv {

constexpr int kMaxTaps = 25;

struct TapTable {
    int ntaps;
    int off_y[kMaxTaps], off_x[kMaxTaps];   // input = (out * sn + off) / sd
    int widx[kMaxTaps];                      // index of the tap inside the [kh*kw] weight slice
};

// Carves a caller-provided workspace; all sub-buffers 256-byte aligned.
struct Carver {
    char* base; size_t off; size_t cap;
    Carver(void* p, size_t c) : base((char*)p), off(0), cap(c) {}
    template <class T> T* take(size_t n) {
        off = (off + 255) & ~(size_t)255;
        T* r = (T*)(base ? base + off : nullptr);
        off += n * sizeof(T);
        return r;
    }
    bool ok() const { return off <= cap; }
};

struct Coefs {          // fp32 scratch shared by forward and backward
    float* a;           // [O]   weight pre-normalisation (1 unless fp16 && demodulate)
    float* c;           // [N]   style pre-normalisation
    float* wsq;         // [O*I] a^2 * sum_k W^2
    float* iscale;      // [N*I] s' = c * s
    float* oscale;      // [N*O] d * a
};

inline void carve_coefs(Carver& cv, const vfm_modconv_desc& d, Coefs& k) {
    k.a = cv.take<float>(d.out_channels);
    k.c = cv.take<float>(d.batch);
    k.wsq = cv.take<float>((size_t)d.out_channels * d.in_channels);
    k.iscale = cv.take<float>((size_t)d.batch * d.in_channels);
    k.oscale = cv.take<float>((size_t)d.batch * d.out_channels);
}

// Launches the coefficient kernels: fills Coefs and dcoefs[N*O].  If `dcoefs_in` is non-NULL it is trusted (backward).
int compute_coefs(const vfm_modconv_desc& d, const float* weight, const float* styles, const Coefs& k,
                  float* dcoefs_out, const float* dcoefs_in, cudaStream_t stream);

struct ConvArgs {
    const void* in;          // [N, Cin, Hin, Win]
    void* out;               // [N, Cout, Hout, Wout]
    const float* w;          // raw weight
    int64_t w_s_co, w_s_ci;  // element strides of (this conv's out channel, this conv's in channel) inside `w`
    const float* in_scale;   // [N, Cin] or NULL
    const float* out_scale;  // [N, Cout] or NULL
    const float* add;        // noise or NULL
    int64_t add_sn, add_sh;  // strides of `add` (sn = 0 for [H,W] broadcast)
    const void* aux;         // dgrad: original x [N, Cout, Hout, Wout] (same dtype) for the dstyles reduction, or NULL
    float* aux_sum;          // [N, Cout] += sum_p aux * acc (before out_scale)
    int N, Cin, Cout, Hin, Win, Hout, Wout;
    int sn, sd;              // input coordinate = (out * sn + off) / sd, tap skipped unless divisible
    TapTable taps;
};

struct WgradArgs {
    const void* dy;          // [N, Co, Hd, Wd]   (the tensor the taps slide over is x; p indexes dy positions)
    const void* x;           // [N, Ci, Hx, Wx]
    const float* oscale;     // [N, Co] or NULL
    const float* iscale;     // [N, Ci] or NULL
    float* dw;               // fp32, element (co,ci,tap) at co*s_co + ci*s_ci + widx   (zero-initialised)
    int64_t s_co, s_ci;
    int N, Co, Ci, Hd, Wd, Hx, Wx;
    int sn, sd;              // x position = (p * sn + off) / sd
    TapTable taps;
    int chunks;              // number of pixel chunks (split-K)
    int chunk_pix;           // pixels per chunk (multiple of 32)
};

// Geometry of the transposed-conv stage for up == 2 (mirrors torch_utils/ops/conv2d_resample.py:112-126).
struct UpGeom {
    int pxt, pyt;             // conv_transpose2d padding
    int zh, zw;               // intermediate size
    int bpx0, bpx1, bpy0, bpy1;   // blur padding
};
inline UpGeom up_geometry(const vfm_modconv_desc& d) {
    UpGeom g;
    int up = d.up, kw = d.kw, kh = d.kh, fw = d.fw, fh = d.fh;
    int px0 = d.padding + (fw + up - 1) / 2, px1 = d.padding + (fw - up) / 2;
    int py0 = d.padding + (fh + up - 1) / 2, py1 = d.padding + (fh - up) / 2;
    px0 -= kw - 1; px1 -= kw - up; py0 -= kh - 1; py1 -= kh - up;
    g.pxt = max(min(-px0, -px1), 0);
    g.pyt = max(min(-py0, -py1), 0);
    g.zw = (d.in_w - 1) * up + kw - 2 * g.pxt;
    g.zh = (d.in_h - 1) * up + kh - 2 * g.pyt;
    g.bpx0 = px0 + g.pxt; g.bpx1 = px1 + g.pxt; g.bpy0 = py0 + g.pyt; g.bpy1 = py1 + g.pyt;
    return g;
}

}  // namespace modconv
}  // namespace vfm
