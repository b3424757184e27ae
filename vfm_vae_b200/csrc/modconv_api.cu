// extern "C" entry points of the modulated conv: validation, workspace planning, and dispatch between the tcgen05
// implicit-GEMM path (modconv_tc.cu) and the generic SIMT path (modconv_generic.cu).
//
// forward : coefficients -> stage-1 contraction (x*s' (*) W, scaled by d[n,o])  [-> blur for up=2]  (+ noise)
// backward: coefficients -> g[n,o], dnoise -> [blur backward] -> data gradient (+ dstyles reduction) -> weight gradient
//           -> fix-ups for the gradient through d[n,o]
#include "modconv_common.cuh"

namespace vfm {
namespace modconv {

// ---- generic path (modconv_generic.cu) ----
int run_conv(int dtype, const ConvArgs& a, cudaStream_t stream);
int run_wgrad(int dtype, WgradArgs a, cudaStream_t stream);
int run_gsum(int dtype, const void* dy, const void* y, const float* noise, int64_t noise_sn, const float* dcoefs, int planes, int O, int HW, float* g, cudaStream_t stream);
int run_dnoise(int dtype, const void* dy, int N, int O, int HW, int per_sample, float* dnoise, cudaStream_t stream);
int run_gsum_dnoise(int dtype, const void* dy, const void* y, const float* noise, int64_t noise_sn, const float* dcoefs, int N, int O, int HW,
                    float* g, float* dnoise, int dnoise_per_sample, cudaStream_t stream);
bool act_grad_fused_supported(int dtype, int O, int HW, const void* dy, const void* y, const void* dz);
int run_act_grad_gsum_dnoise(int dtype, const void* dy, const void* y, const void* bias, const float* noise, int64_t noise_sn, const float* dcoefs, int N, int O, int HW,
                             float gain, float alpha, float clamp, int act, void* dz, float* g, float* gz, float* dnoise, int dnoise_per_sample, cudaStream_t stream);
int run_dw_fix(float* dw, const float* w, const float* a, const float* g, const float* dcoefs, const float* iscale, int N, int O, int I, int KK, cudaStream_t stream);
int run_ds_fix(float* ds, const float* dsum, const float* c, const float* g, const float* dcoefs, const float* iscale, const float* wsq, int N, int O, int I, int demod, cudaStream_t stream);

// ---- tensor-core path (modconv_tc.cu) ----
bool tc_supported(const vfm_modconv_desc& d);
size_t tc_workspace_bytes(const vfm_modconv_desc& d, int direction);
// stage-1 contraction: x -> z (== y for up=1; noise only added when up == 1)
int tc_stage1_forward(const vfm_modconv_desc& d, const Stage1& s, const void* x, const float* weight, const Coefs& k, void* z, int zpitch,
                      const float* noise, int64_t noise_sn, const Epilogue& ep, const float* x_scale, const float* x_shift,
                      void* ws, size_t ws_bytes, cudaStream_t stream, int keep_operand);
void tc_forward_operand(const vfm_modconv_desc& d, void* ws, size_t ws_bytes, const void** hi, const void** lo);
// gradients of the stage-1 contraction given dz: dx (+ dsum) and the main part of dweight
int tc_stage1_backward(const vfm_modconv_desc& d, const Stage1& s, const void* dz, int dz_pitch, const void* x, const float* weight, const Coefs& k,
                       void* dx, float* dsum, float* dweight, void* ws, size_t ws_bytes, cudaStream_t stream, const void* saved_xt, const void* saved_xt_lo);

static int validate(const vfm_modconv_desc& d) {
    VFM_CHECK_ARG(d.dtype == VFM_F16 || d.dtype == VFM_F32 || d.dtype == VFM_F64, "modulated_conv2d: unsupported dtype %d", d.dtype);
    VFM_CHECK_ARG(d.batch >= 1 && d.in_channels >= 1 && d.out_channels >= 1 && d.in_h >= 1 && d.in_w >= 1, "modulated_conv2d: empty tensor");
    VFM_CHECK_ARG(d.kh == d.kw && (d.kh & 1) == 1, "modulated_conv2d: kernel must be square with odd size (got %dx%d)", d.kh, d.kw);
    VFM_CHECK_ARG(d.up == 1 || d.up == 2, "modulated_conv2d: up must be 1 or 2 (got %d); down > 1 is not on the decoder path", d.up);
    VFM_CHECK_ARG(d.noise_mode >= 0 && d.noise_mode <= 2, "modulated_conv2d: bad noise_mode");
    VFM_CHECK_ARG(d.up == 1 || (d.resample_filter && d.fw >= 1 && d.fh >= 1), "modulated_conv2d: up=2 needs a 2-D resample filter");
    if (d.kh * d.kw > kMaxTaps) { set_error("modulated_conv2d: %dx%d kernels are not supported", d.kh, d.kw); return VFM_ERR_NO_KERNEL; }
    int oh, ow;
    if (d.up == 1) { oh = d.in_h + 2 * d.padding - d.kh + 1; ow = d.in_w + 2 * d.padding - d.kw + 1; }
    else {
        int px0 = d.padding + (d.fw + 1) / 2, px1 = d.padding + (d.fw - 2) / 2;
        int py0 = d.padding + (d.fh + 1) / 2, py1 = d.padding + (d.fh - 2) / 2;
        ow = d.in_w * 2 + px0 + px1 - d.fw + 1 - (d.kw - 1);
        oh = d.in_h * 2 + py0 + py1 - d.fh + 1 - (d.kh - 1);
    }
    VFM_CHECK_ARG(oh == d.out_h && ow == d.out_w, "modulated_conv2d: out size %dx%d does not match expected %dx%d", d.out_h, d.out_w, oh, ow);
    return VFM_OK;
}

static int call_upfirdn(int dtype, const void* in, void* out, const float* f, int fw, int fh, int up, int down, int px0, int py0, int flip, float gain,
                        int N, int C, int ih, int iw, int oh, int ow, const float* add, int64_t add_sn, cudaStream_t stream, int in_pitch = 0,
                        const Epilogue* ep = nullptr, int out_pitch = 0) {
    if (in_pitch == 0) in_pitch = iw;
    if (out_pitch == 0) out_pitch = ow;
    vfm_upfirdn2d_params u;
    u.x = in; u.f = f; u.y = out; u.dtype = dtype;
    u.upx = u.upy = up; u.downx = u.downy = down; u.padx0 = px0; u.pady0 = py0; u.flip = flip; u.gain = gain;
    u.in_w = iw; u.in_h = ih; u.channels = C; u.batch = N;
    u.in_stride_w = 1; u.in_stride_h = in_pitch; u.in_stride_c = (int64_t)ih * in_pitch; u.in_stride_n = (int64_t)C * ih * in_pitch;
    u.fw = fw; u.fh = fh; u.f_stride_w = 1; u.f_stride_h = fw;
    u.out_w = ow; u.out_h = oh;
    u.out_stride_w = 1; u.out_stride_h = out_pitch; u.out_stride_c = (int64_t)oh * out_pitch; u.out_stride_n = (int64_t)C * oh * out_pitch;
    u.add = add; u.add_stride_h = ow; u.add_stride_n = add_sn;
    u.ep_enable = 0; u.ep_act = 1; u.ep_alpha = 0; u.ep_gain = 1; u.ep_clamp = -1; u.ep_bias = nullptr; u.pad_mode = 0; u.f_stride_c = 0;
    if (ep && ep->enable) { u.ep_enable = 1; u.ep_act = ep->act; u.ep_alpha = ep->alpha; u.ep_gain = ep->gain; u.ep_clamp = ep->clamp; u.ep_bias = ep->bias; }
    return vfm_upfirdn2d(&u, stream);
}

static size_t generic_workspace(const vfm_modconv_desc& d, int direction) {
    Carver cv(nullptr, ~(size_t)0);
    Coefs k; carve_coefs(cv, d, k);
    Stage1 s = make_stage1(d);
    if (d.up == 2) cv.take<char>((size_t)d.batch * d.out_channels * s.zh * (size_t)((s.zw + 15) & ~15) * esize(d.dtype));
    if (direction == 1) { cv.take<float>((size_t)d.batch * d.out_channels); cv.take<float>((size_t)d.batch * d.in_channels); }
    return cv.off + 256;
}

// ---- streaming path for 1x1 convs with <= 4 output channels (modconv_pointwise.cu) ----
bool pw_supported(const vfm_modconv_desc& d);
size_t pw_workspace_bytes(const vfm_modconv_desc& d, int direction);
int pw_stage1_forward(const vfm_modconv_desc& d, const void* x, const float* weight, const Coefs& k, void* y, const float* noise, int64_t noise_sn, const Epilogue& ep, cudaStream_t stream);
int pw_stage1_backward(const vfm_modconv_desc& d, const void* dy, const void* x, const float* weight, const Coefs& k, void* dx, float* dsum, float* dweight,
                       void* ws, size_t ws_bytes, cudaStream_t stream);

static bool use_pw(const vfm_modconv_desc& d) { return !d.force_generic && pw_supported(d); }
static bool use_tc(const vfm_modconv_desc& d) { return !d.force_generic && !pw_supported(d) && tc_supported(d); }

}  // namespace modconv
}  // namespace vfm

using namespace vfm;
using namespace vfm::modconv;

extern "C" int vfm_modconv_uses_tensor_cores(const vfm_modconv_desc* d) {
    if (!d) return 0;
    return use_tc(*d) ? 1 : 0;
}

// the forward's workspace layout up to the tensor-core section (must mirror vfm_modconv_forward)
extern "C" int vfm_modconv_forward_operand(const vfm_modconv_desc* dp, void* workspace, size_t workspace_bytes, const void** hi, const void** lo) {
    VFM_CHECK_ARG(dp && hi && lo, "modconv_forward_operand: NULL argument");
    *hi = *lo = nullptr;
    const vfm_modconv_desc& d = *dp;
    if (!workspace || !use_tc(d) || (use_pw(d))) return VFM_OK;
    Carver cv(workspace, workspace_bytes);
    Coefs k; carve_coefs(cv, d, k);
    Stage1 s = make_stage1(d);
    if (d.up == 2) cv.take<char>((size_t)d.batch * d.out_channels * s.zh * (size_t)((s.zw + 15) & ~15) * esize(d.dtype));
    cv.off = (cv.off + 255) & ~(size_t)255;
    if (cv.off >= workspace_bytes) return VFM_OK;
    tc_forward_operand(d, (char*)workspace + cv.off, workspace_bytes - cv.off, hi, lo);
    return VFM_OK;
}

extern "C" size_t vfm_modconv_workspace_bytes(const vfm_modconv_desc* d, int direction) {
    if (!d) return 0;
    const int dir = direction == 2 ? 1 : direction;
    size_t g = generic_workspace(*d, dir);
    if (use_tc(*d)) g += tc_workspace_bytes(*d, dir);
    if (use_pw(*d)) g += pw_workspace_bytes(*d, dir);
    // fused-epilogue backward: the pre-activation gradient dz [N,O,Hout,Wout] lives in the workspace
    if (direction == 2) g += (size_t)d->batch * d->out_channels * d->out_h * d->out_w * esize(d->dtype) + 512;
    return g;
}

extern "C" int vfm_modconv_fused_backward_supported(const vfm_modconv_desc* d) {
    if (!d) return 0;
    const int HW = d->out_h * d->out_w;
    return ((d->dtype == VFM_F16 || d->dtype == VFM_F32) && HW % 8 == 0 && HW >= 2048 && (size_t)d->out_channels * 2 * sizeof(float) <= 48 * 1024) ? 1 : 0;
}

extern "C" int vfm_modconv_forward(const vfm_modconv_fwd_params* p, void* stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    VFM_CHECK_ARG(p != nullptr, "modulated_conv2d: params is NULL");
    const vfm_modconv_desc& d = p->d;
    int st = validate(d); if (st) return st;
    VFM_CHECK_ARG(p->x && p->weight && p->styles && p->y && p->dcoefs, "modulated_conv2d: x, weight, styles, y, dcoefs must be non-NULL");
    VFM_CHECK_ARG((d.noise_mode == VFM_NOISE_NONE) == (p->noise == nullptr), "modulated_conv2d: noise pointer does not match noise_mode");
    size_t need = vfm_modconv_workspace_bytes(&d, 0);
    if (!p->workspace || p->workspace_bytes < need) { set_error("modulated_conv2d: workspace too small (%zu < %zu)", p->workspace_bytes, need); return VFM_ERR_WORKSPACE; }

    Carver cv(p->workspace, p->workspace_bytes);
    Coefs k; carve_coefs(cv, d, k);
    k.d = p->dcoefs;
    Stage1 s = make_stage1(d);
    void* z = p->y;
    // stage-1 output rows are padded to 32 bytes on the tensor-core path so that the blur stages its tiles with vector loads
    const int zpitch = (d.up == 2 && use_tc(d)) ? ((s.zw + 15) & ~15) : s.zw;
    if (d.up == 2) z = cv.take<char>((size_t)d.batch * d.out_channels * s.zh * (size_t)((s.zw + 15) & ~15) * esize(d.dtype));
    st = compute_coefs(d, p->weight, p->styles, k, p->dcoefs, nullptr, stream); if (st) return st;
    const int64_t noise_sn = (d.noise_mode == VFM_NOISE_N1HW) ? (int64_t)d.out_h * d.out_w : 0;
    const float* s1_noise = (d.up == 1) ? p->noise : nullptr;
    // optional fused layer epilogue: goes into the stage-1 kernel for up=1 and into the blur for up=2
    Epilogue ep = no_epilogue();
    if (p->ep_enable) {
        VFM_CHECK_ARG(p->ep_act == 1 || p->ep_act == 3 || p->ep_act == VFM_EP_ACT_GELU, "modulated_conv2d: the fused epilogue supports linear, lrelu and gelu only");
        if (p->ep_act == VFM_EP_ACT_GELU && !(use_tc(d) && d.up == 1 && !(use_pw(d) && aligned16(p->x) && aligned16(p->y)))) {
            set_error("modulated_conv2d: the gelu epilogue is only implemented in the tcgen05 kernel (up = 1)"); return VFM_ERR_NO_KERNEL;
        }
        VFM_CHECK_ARG(p->ep_act != VFM_EP_ACT_GELU || p->ep_gain > 0, "modulated_conv2d: the gelu epilogue needs ep_gain > 0");
        VFM_CHECK_ARG(!p->ep_residual || p->ep_gamma, "modulated_conv2d: ep_residual needs ep_gamma");
        const bool fusable = (use_pw(d) && aligned16(p->x) && aligned16(p->y)) || (use_tc(d) && (d.up == 1 || !p->ep_residual));
        if (!fusable) { set_error("modulated_conv2d: no kernel fuses the epilogue for this descriptor"); return VFM_ERR_NO_KERNEL; }
        ep.enable = 1; ep.act = p->ep_act; ep.alpha = (float)p->ep_alpha; ep.gain = (float)p->ep_gain; ep.clamp = (float)p->ep_clamp;
        ep.res_scale = (float)p->ep_res_scale; ep.bias = p->ep_bias; ep.residual = p->ep_residual; ep.gamma = p->ep_gamma;
    }
    if (p->x_scale || p->x_shift || p->ep_res_affine) {
        VFM_CHECK_ARG(p->x_scale && p->x_shift, "modulated_conv2d: x_scale and x_shift must be given together");
        VFM_CHECK_ARG(!p->ep_res_affine || (p->ep_residual && d.in_channels == d.out_channels), "modulated_conv2d: ep_res_affine needs ep_residual and I == O");
        if (!use_tc(d) || (use_pw(d) && aligned16(p->x) && aligned16(p->y))) { set_error("modulated_conv2d: the input affine map is only implemented on the tcgen05 path"); return VFM_ERR_NO_KERNEL; }
        if (p->ep_res_affine) { ep.res_a = p->x_scale; ep.res_b = p->x_shift; }
    }
    const Epilogue s1_ep = (d.up == 1) ? ep : no_epilogue();

    if (use_pw(d) && aligned16(p->x) && aligned16(p->y)) {
        st = pw_stage1_forward(d, p->x, p->weight, k, z, s1_noise, noise_sn, s1_ep, stream);
        if (st) return st;
    } else if (use_tc(d)) {
        cv.off = (cv.off + 255) & ~(size_t)255;
        st = tc_stage1_forward(d, s, p->x, p->weight, k, z, zpitch, s1_noise, noise_sn, s1_ep, p->x_scale, p->x_shift,
                               (char*)p->workspace + cv.off, p->workspace_bytes - cv.off, stream, p->keep_operand);
        if (st) return st;
    } else {
        ConvArgs a;
        a.in = p->x; a.out = z; a.w = p->weight;
        a.w_s_co = (int64_t)d.in_channels * d.kh * d.kw; a.w_s_ci = (int64_t)d.kh * d.kw;
        a.in_scale = k.iscale; a.out_scale = k.oscale;
        a.add = s1_noise; a.add_sn = noise_sn; a.add_sh = s.zw;
        a.aux = nullptr; a.aux_sum = nullptr;
        a.N = d.batch; a.Cin = d.in_channels; a.Cout = d.out_channels; a.Hin = d.in_h; a.Win = d.in_w; a.Hout = s.zh; a.Wout = s.zw;
        a.sn = s.sn; a.sd = s.sd; a.taps = s.taps;
        st = run_conv(d.dtype, a, stream); if (st) return st;
    }
    if (d.up == 1) return VFM_OK;
    return call_upfirdn(d.dtype, z, p->y, d.resample_filter, d.fw, d.fh, s.r_up, 1, s.r_px0, s.r_py0, 0, (float)(d.up * d.up),
                        d.batch, d.out_channels, s.zh, s.zw, d.out_h, d.out_w, p->noise, noise_sn, stream, zpitch, &ep);
}

extern "C" int vfm_modconv_backward(const vfm_modconv_bwd_params* p, void* stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    VFM_CHECK_ARG(p != nullptr, "modulated_conv2d backward: params is NULL");
    const vfm_modconv_desc& d = p->d;
    int st = validate(d); if (st) return st;
    VFM_CHECK_ARG(p->dy && p->x && p->weight && p->styles && p->dcoefs, "modulated_conv2d backward: dy, x, weight, styles, dcoefs must be non-NULL");
    VFM_CHECK_ARG(!d.demodulate || p->y, "modulated_conv2d backward: y (forward output) is required when demodulate is on");
    VFM_CHECK_ARG(!p->dstyles || p->dx, "modulated_conv2d backward: dstyles needs dx to be computed as well");
    VFM_CHECK_ARG(!p->dnoise || d.noise_mode != VFM_NOISE_NONE, "modulated_conv2d backward: dnoise requested without noise");
    size_t need = vfm_modconv_workspace_bytes(&d, p->ep_enable ? 2 : 1);
    if (!p->workspace || p->workspace_bytes < need) { set_error("modulated_conv2d backward: workspace too small (%zu < %zu)", p->workspace_bytes, need); return VFM_ERR_WORKSPACE; }
    if (p->ep_enable) {
        VFM_CHECK_ARG(p->ep_act == 1 || p->ep_act == 3, "modulated_conv2d backward: the fused epilogue gradient supports linear and lrelu only");
        VFM_CHECK_ARG(p->ep_gain > 0 && p->y, "modulated_conv2d backward: the fused epilogue gradient needs ep_gain > 0 and the activated output y");
        if (!vfm_modconv_fused_backward_supported(&d)) { set_error("modulated_conv2d backward: no fused epilogue gradient for this descriptor"); return VFM_ERR_NO_KERNEL; }
    }

    const int N = d.batch, I = d.in_channels, O = d.out_channels, KK = d.kh * d.kw;
    Carver cv(p->workspace, p->workspace_bytes);
    Coefs k; carve_coefs(cv, d, k);
    k.d = p->dcoefs;
    Stage1 s = make_stage1(d);
    void* dz = nullptr;
    // tensor-core path: rows of the blur-backward output are padded to 32 bytes so that the streaming blur kernel writes it
    const int dz_pitch = (d.up == 2 && use_tc(d)) ? ((s.zw + 15) & ~15) : s.zw;
    if (d.up == 2) dz = cv.take<char>((size_t)N * O * s.zh * (size_t)((s.zw + 15) & ~15) * esize(d.dtype));
    float* g = cv.take<float>((size_t)N * O);
    float* dsum = cv.take<float>((size_t)N * I);
    st = compute_coefs(d, p->weight, p->styles, k, nullptr, p->dcoefs, stream); if (st) return st;

    const int HWo = d.out_h * d.out_w;
    const int64_t noise_sn = (d.noise_mode == VFM_NOISE_N1HW) ? (int64_t)HWo : 0;
    const void* dy_pre = p->dy;                 // gradient w.r.t. the conv output (before bias / activation)
    if (p->ep_enable) {
        // one pass: activation gradient -> dz (workspace), g, per-sample bias gradient, dnoise
        void* dzp_act = cv.take<char>((size_t)N * O * HWo * esize(d.dtype));
        const bool need_g = d.demodulate && (p->dweight || p->dstyles);
        const int per_sample = (d.noise_mode == VFM_NOISE_N1HW);
        if (p->dnoise) VFM_CUDA_OK(cudaMemsetAsync(p->dnoise, 0, sizeof(float) * (size_t)HWo * (per_sample ? N : 1), stream));
        st = run_act_grad_gsum_dnoise(d.dtype, p->dy, p->y, p->ep_bias, p->noise, noise_sn, p->dcoefs, N, O, HWo, (float)p->ep_gain, (float)p->ep_alpha,
                                      (float)p->ep_clamp, p->ep_act, dzp_act, need_g ? g : nullptr, p->dbias_no, p->dnoise, per_sample, stream);
        if (st) return st;
        dy_pre = dzp_act;
    } else {
        const bool need_g = d.demodulate && (p->dweight || p->dstyles);
        const int per_sample = (d.noise_mode == VFM_NOISE_N1HW);
        if (p->dnoise) VFM_CUDA_OK(cudaMemsetAsync(p->dnoise, 0, sizeof(float) * (size_t)HWo * (per_sample ? N : 1), stream));
        // one pass over dy for both reductions where the tensors allow it
        st = (need_g || p->dnoise) ? run_gsum_dnoise(d.dtype, p->dy, p->y, p->noise, noise_sn, p->dcoefs, N, O, HWo, need_g ? g : nullptr, p->dnoise, per_sample, stream)
                                   : VFM_OK;
        if (st == VFM_ERR_NO_KERNEL) {
            st = VFM_OK;
            if (need_g) { st = run_gsum(d.dtype, p->dy, p->y, p->noise, noise_sn, p->dcoefs, N * O, O, HWo, g, stream); if (st) return st; }
            if (p->dnoise) { st = run_dnoise(d.dtype, p->dy, N, O, HWo, per_sample, p->dnoise, stream); if (st) return st; }
        }
        if (st) return st;
    }

    // gradient w.r.t. the stage-1 output
    const void* dzp = dy_pre;
    if (d.up == 2) {
        // backward of upfirdn2d (torch_utils/ops/upfirdn2d.py:251-269): swap up/down, flip the filter, same gain
        st = call_upfirdn(d.dtype, dy_pre, dz, d.resample_filter, d.fw, d.fh, 1, s.r_up, s.rb_px0, s.rb_py0, 1, (float)(d.up * d.up),
                          N, O, d.out_h, d.out_w, s.zh, s.zw, nullptr, 0, stream, 0, nullptr, dz_pitch);
        if (st) return st;
        dzp = dz;
    }
    if (p->dstyles) VFM_CUDA_OK(cudaMemsetAsync(dsum, 0, sizeof(float) * (size_t)N * I, stream));
    if (p->dweight) VFM_CUDA_OK(cudaMemsetAsync(p->dweight, 0, sizeof(float) * (size_t)O * I * KK, stream));

    if (use_pw(d) && aligned16(p->x) && aligned16(dy_pre) && (!p->dx || aligned16(p->dx))) {
        cv.off = (cv.off + 255) & ~(size_t)255;
        st = pw_stage1_backward(d, dzp, p->x, p->weight, k, p->dx, p->dstyles ? dsum : nullptr, p->dweight,
                                (char*)p->workspace + cv.off, p->workspace_bytes - cv.off, stream);
        if (st) return st;
    } else if (use_tc(d)) {
        cv.off = (cv.off + 255) & ~(size_t)255;
        st = tc_stage1_backward(d, s, dzp, d.up == 2 ? dz_pitch : 0, p->x, p->weight, k, p->dx, p->dstyles ? dsum : nullptr, p->dweight,
                                (char*)p->workspace + cv.off, p->workspace_bytes - cv.off, stream, p->saved_operand, p->saved_operand_lo);
        if (st) return st;
    } else {
        if (p->dx) {
            ConvArgs a;
            a.in = dzp; a.out = p->dx; a.w = p->weight;
            a.w_s_co = KK; a.w_s_ci = (int64_t)I * KK;          // roles of the channel axes are swapped
            a.in_scale = k.oscale; a.out_scale = k.iscale; a.add = nullptr; a.add_sn = a.add_sh = 0;
            a.aux = p->dstyles ? p->x : nullptr; a.aux_sum = p->dstyles ? dsum : nullptr;
            a.N = N; a.Cin = O; a.Cout = I; a.Hin = s.zh; a.Win = s.zw; a.Hout = d.in_h; a.Wout = d.in_w;
            dgrad_taps(s, a.taps, a.sn, a.sd);
            st = run_conv(d.dtype, a, stream); if (st) return st;
        }
        if (p->dweight) {
            WgradArgs w;
            w.dy = dzp; w.x = p->x; w.oscale = k.oscale; w.iscale = k.iscale; w.dw = p->dweight;
            w.s_co = (int64_t)I * KK; w.s_ci = KK;
            w.N = N; w.Co = O; w.Ci = I; w.Hd = s.zh; w.Wd = s.zw; w.Hx = d.in_h; w.Wx = d.in_w;
            w.sn = s.sn; w.sd = s.sd; w.taps = s.taps; w.chunks = 0; w.chunk_pix = 0;
            st = run_wgrad(d.dtype, w, stream); if (st) return st;
        }
    }
    if (p->dweight && d.demodulate) { st = run_dw_fix(p->dweight, p->weight, k.a, g, p->dcoefs, k.iscale, N, O, I, KK, stream); if (st) return st; }
    if (p->dstyles) { st = run_ds_fix(p->dstyles, dsum, k.c, g, p->dcoefs, k.iscale, k.wsq, N, O, I, d.demodulate, stream); if (st) return st; }
    return VFM_OK;
}
