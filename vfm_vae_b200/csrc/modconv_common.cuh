// Shared pieces of the modulated-conv implementation: the small "coefficient" kernels (pre-normalisation, demodulation
// coefficients, gradient fix-ups), the stage-1 geometry (what "conv with up=2" means) and the workspace carver used by
// both the generic SIMT path and the tcgen05 path.
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

inline size_t esize(int dtype) { return dtype == VFM_F16 ? 2 : (dtype == VFM_F32 ? 4 : 8); }

// fused layer epilogue (inference): y = clamp(act(y + bias) * gain); optionally y = (gamma * y + residual) * res_scale
struct Epilogue {
    int enable, act;
    float alpha, gain, clamp, res_scale;
    const void* bias; const void* residual; const float* gamma;
    const float* res_a; const float* res_b;     // optional per-(n, channel) affine map of the residual (tcgen05 path only)
};
inline Epilogue no_epilogue() { Epilogue e; e.enable = 0; e.act = 1; e.alpha = 0.f; e.gain = 1.f; e.clamp = -1.f; e.res_scale = 1.f; e.bias = nullptr; e.residual = nullptr; e.gamma = nullptr; e.res_a = nullptr; e.res_b = nullptr; return e; }
template <class T> __device__ __forceinline__ float apply_epilogue(const Epilogue& e, float v, int ch, size_t idx) {
    if (e.bias) v += to_acc(((const T*)e.bias)[ch]);
    if (e.act == 3) v = (v > 0.f) ? v : v * e.alpha;
    v *= e.gain;
    if (e.clamp >= 0.f) v = fminf(fmaxf(v, -e.clamp), e.clamp);
    if (e.residual) v = (e.gamma[ch] * v + to_acc(((const T*)e.residual)[idx])) * e.res_scale;
    return v;
}

struct Coefs {          // fp32 scratch shared by forward and backward
    float* a;           // [O]   weight pre-normalisation (1 unless fp16 && demodulate)
    float* c;           // [N]   style pre-normalisation
    float* wsq;         // [O*I] a^2 * sum_k W^2
    float* iscale;      // [N*I] s' = c * s
    float* oscale;      // [N*O] d * a
    const float* d;     // [N*O] demodulation coefficients (the caller's dcoefs tensor)
};

inline void carve_coefs(Carver& cv, const vfm_modconv_desc& d, Coefs& k) {
    k.a = cv.take<float>(d.out_channels);
    k.c = cv.take<float>(d.batch);
    k.wsq = cv.take<float>((size_t)d.out_channels * d.in_channels);
    k.iscale = cv.take<float>((size_t)d.batch * d.in_channels);
    k.oscale = cv.take<float>((size_t)d.batch * d.out_channels);
    k.d = nullptr;
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

// ---- "stage 1": the dense contraction of conv2d_resample (torch_utils/ops/conv2d_resample.py:46-141) ----------------
//   up == 1          : plain conv, z == y
//   up == 2, k == 1  : 1x1 conv at input resolution, then upfirdn2d(up=2)            (reference lines 101-104)
//   up == 2, k  > 1  : conv_transpose2d(stride 2) to (2H+1)x(2W+1), then the blur    (reference lines 112-126)
struct Stage1 {
    int zh, zw;            // stage-1 output size
    int sn, sd;            // x position = (z * sn + off) / sd
    int pyt, pxt;          // transposed-conv padding
    TapTable taps;         // forward taps
    int r_up, r_px0, r_py0;      // stage-2 resampler: upfirdn2d(up = r_up, pad0 = r_p*0, gain = up^2)
    int rb_px0, rb_py0;          // its backward (torch_utils/ops/upfirdn2d.py:251-269)
};

inline Stage1 make_stage1(const vfm_modconv_desc& d) {
    Stage1 s;
    const int kh = d.kh, kw = d.kw;
    s.taps.ntaps = kh * kw;
    s.pyt = s.pxt = 0;
    if (d.up == 1) {
        s.zh = d.out_h; s.zw = d.out_w; s.sn = 1; s.sd = 1;
        for (int ky = 0; ky < kh; ky++) for (int kx = 0; kx < kw; kx++) {
            int t = ky * kw + kx;
            s.taps.off_y[t] = ky - d.padding; s.taps.off_x[t] = kx - d.padding;
            s.taps.widx[t] = d.flip_weight ? t : (kh - 1 - ky) * kw + (kw - 1 - kx);
        }
        s.r_up = 1; s.r_px0 = s.r_py0 = s.rb_px0 = s.rb_py0 = 0;
    } else if (kh == 1) {
        s.zh = d.in_h; s.zw = d.in_w; s.sn = 1; s.sd = 1;
        s.taps.off_y[0] = s.taps.off_x[0] = 0; s.taps.widx[0] = 0;
        s.r_up = 2;
        s.r_px0 = d.padding + (d.fw + 1) / 2; s.r_py0 = d.padding + (d.fh + 1) / 2;
        s.rb_px0 = d.fw - s.r_px0 - 1; s.rb_py0 = d.fh - s.r_py0 - 1;
    } else {
        const int up = d.up;
        int px0 = d.padding + (d.fw + up - 1) / 2, px1 = d.padding + (d.fw - up) / 2;
        int py0 = d.padding + (d.fh + up - 1) / 2, py1 = d.padding + (d.fh - up) / 2;
        px0 -= kw - 1; px1 -= kw - up; py0 -= kh - 1; py1 -= kh - up;
        s.pxt = max(min(-px0, -px1), 0);
        s.pyt = max(min(-py0, -py1), 0);
        s.zw = (d.in_w - 1) * up + kw - 2 * s.pxt;
        s.zh = (d.in_h - 1) * up + kh - 2 * s.pyt;
        s.sn = 1; s.sd = 2;
        for (int ky = 0; ky < kh; ky++) for (int kx = 0; kx < kw; kx++) {
            int t = ky * kw + kx;
            s.taps.off_y[t] = s.pyt - ky; s.taps.off_x[t] = s.pxt - kx;
            s.taps.widx[t] = d.flip_weight ? (kh - 1 - ky) * kw + (kw - 1 - kx) : t;
        }
        s.r_up = 1; s.r_px0 = px0 + s.pxt; s.r_py0 = py0 + s.pyt;
        s.rb_px0 = d.fw - s.r_px0 - 1; s.rb_py0 = d.fh - s.r_py0 - 1;
    }
    return s;
}

// taps of the data gradient: z position read for a given x position:  forward x = (z + off)/sd  <=>  z = x*sd - off
inline void dgrad_taps(const Stage1& s, TapTable& t, int& sn, int& sd) {
    t.ntaps = s.taps.ntaps;
    sn = s.sd; sd = 1;
    for (int i = 0; i < t.ntaps; i++) { t.off_y[i] = -s.taps.off_y[i]; t.off_x[i] = -s.taps.off_x[i]; t.widx[i] = s.taps.widx[i]; }
}

}  // namespace modconv
}  // namespace vfm
