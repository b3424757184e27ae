/*
 * vfm_ops.h -- C ABI of libvfmops.so: the B200 (sm_100a) kernels behind the VFM-VAE pixel-decoder ops.
 *
 * This is the drop-in boundary.  Every entry point below is what the reference's pybind plugin function
 * for the same op would bind if the plugin were a C library; the reference-side bindings a maintainer
 * would add are shown in INTEGRATION.md.  Reference interfaces replaced (paths relative to the
 * reference repo tianciB/VFM-VAE):
 *
 *   vfm_bias_act               <- bias_act(x,b,xref,yref,dy,grad,dim,act,alpha,gain,clamp)      torch_utils/ops/bias_act.cpp:32-90
 *   vfm_upfirdn2d              <- upfirdn2d(x,f,upx,upy,downx,downy,padx0..pady1,flip,gain)     torch_utils/ops/upfirdn2d.cpp:16-98
 *   vfm_filtered_lrelu         <- filtered_lrelu(x,fu,fd,b,si,up,down,px0..py1,sx,sy,gain,...)   torch_utils/ops/filtered_lrelu.cpp:16-208
 *   vfm_filtered_lrelu_act     <- filtered_lrelu_act_(x,si,sx,sy,gain,slope,clamp,writeSigns)   torch_utils/ops/filtered_lrelu.cpp:213-290
 *   vfm_modconv_forward/_backward <- modulated_conv2d(...) + its stock-autograd backward        networks/generator.py:46-103
 *                                    (the reference has no native code for this op: it is Python over a grouped
 *                                    cuDNN conv, torch_utils/ops/conv2d_resample.py:46-141)
 *
 * Conventions
 *   - plain pointers and sizes only; no torch types.  All data pointers are DEVICE pointers.
 *   - `stream` is a cudaStream_t passed as void* (NULL = legacy default stream).
 *   - the library never allocates device memory and never synchronises; scratch space is passed in by the
 *     caller (vfm_modconv_workspace_bytes tells how much).  All launches are stream-ordered and there is no
 *     global mutable device state (unlike the reference's filtered_lrelu __constant__ filter buffer,
 *     torch_utils/ops/filtered_lrelu.cu:77-78), so concurrent use from several streams is safe.
 *   - return value: 0 = launched; negative = error, nothing launched (VFM_ERR_*).  VFM_ERR_NO_KERNEL mirrors the
 *     reference's `return_code = -1` ("no optimised kernel, use the generic composition",
 *     torch_utils/ops/filtered_lrelu.cpp:52-56).
 *   - sizes/strides are in ELEMENTS unless a field says bytes.
 */
#ifndef VFM_OPS_H
#define VFM_OPS_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define VFM_ABI_VERSION 10

#if defined(__GNUC__)
#define VFM_API __attribute__((visibility("default")))
#else
#define VFM_API
#endif

enum vfm_dtype { VFM_F16 = 0, VFM_F32 = 1, VFM_F64 = 2 };

enum vfm_status {
    VFM_OK = 0,
    VFM_ERR_NO_KERNEL = -1,   /* no specialised kernel for these parameters */
    VFM_ERR_INVALID = -2,     /* argument check failed (see vfm_last_error) */
    VFM_ERR_CUDA = -3,        /* a CUDA runtime/driver call failed (see vfm_last_error) */
    VFM_ERR_WORKSPACE = -4    /* workspace missing or too small */
};

/* Human-readable description of the last error on the calling thread ("" if none). */
VFM_API const char* vfm_last_error(void);
VFM_API int vfm_abi_version(void);
/* Number of kernels this library has launched in this process (all ops); bench.py reports the delta. */
VFM_API uint64_t vfm_launch_count(void);

/* Optional per-kernel timing (measurement aid for bench.py; off by default and free when off).
 * vfm_timing_enable(1) clears old records and makes every hot kernel launch record a CUDA event before and after it on
 * the launching stream, tagged with the launch's ALGORITHMIC flops and bytes.  vfm_timing_report() waits for the
 * recorded events and returns one aggregated entry per kernel name (returns the number of distinct names). */
typedef struct {
    char     name[64];
    int64_t  launches;
    double   total_ms;   /* sum of event-measured durations */
    double   flops;      /* sum of algorithmic FLOPs */
    double   bytes;      /* sum of algorithmic HBM bytes */
} vfm_kernel_stat;
VFM_API void vfm_timing_enable(int on);
VFM_API int vfm_timing_report(vfm_kernel_stat* out, int max_entries);

/* ------------------------------------------------------------------------------------------------------------
 * bias_act: y = clamp(act(x + b) * gain, +-clamp)            (grad = 0)
 *           dx = dy * act'(.) * gain, 0 where |yref| >= clamp (grad = 1; here `x` is the incoming gradient)
 *           second-order term                                 (grad = 2; `x` = d_dx, `dy` = first-order dy)
 * Same argument meaning as the reference kernel parameters (torch_utils/ops/bias_act.h:12-34):
 * the bias index of flat element i is (i / step_b) % size_b.
 * Extension: if `db` is non-NULL (fp32[size_b], must be zero-initialised by the caller) the kernel also
 * accumulates the bias gradient sum_{all but dim} y into it with warp-shuffle + block reductions, which
 * replaces the separate `dx.sum(...)` reduction at torch_utils/ops/bias_act.py:170.
 */
typedef struct {
    const void* x;      /* [size_x] */
    const void* b;      /* [size_b] or NULL */
    const void* xref;   /* [size_x] or NULL */
    const void* yref;   /* [size_x] or NULL */
    const void* dy;     /* [size_x] or NULL */
    void*       y;      /* [size_x] out */
    float*      db;     /* [size_b] fp32 accumulate, or NULL */
    int32_t     dtype;  /* vfm_dtype of x/b/xref/yref/dy/y */
    int32_t     grad;   /* 0, 1, 2 */
    int32_t     act;    /* 1 linear 2 relu 3 lrelu 4 tanh 5 sigmoid 6 elu 7 selu 8 softplus 9 swish */
    double      alpha;  /* scalars travel as fp64 so that the fp64 kernels are exact (the reference passes floats) */
    double      gain;
    double      clamp;  /* < 0 = off */
    int64_t     size_x;
    int64_t     size_b;
    int64_t     step_b;
} vfm_bias_act_params;

VFM_API int vfm_bias_act(const vfm_bias_act_params* p, void* stream);

/* ------------------------------------------------------------------------------------------------------------
 * upfirdn2d: zero-insert by (upx,upy) -> pad/crop -> FIR -> decimate by (downx,downy), per (n,c) plane.
 * out = (in*up + pad0 + pad1 - taps + down) / down   (the caller computes it, torch_utils/ops/upfirdn2d.cpp:35-36).
 * Any element strides (NCHW or channels_last).  Filter is fp32 [fh,fw] with its own strides.
 */
typedef struct {
    const void*  x;
    const float* f;
    void*        y;
    int32_t      dtype;
    int32_t      upx, upy, downx, downy;
    int32_t      padx0, pady0;          /* only the leading pads matter once the output size is fixed */
    int32_t      flip;                  /* 0 = true convolution, 1 = correlation */
    double       gain;
    int32_t      in_w, in_h, channels, batch;
    int64_t      in_stride_w, in_stride_h, in_stride_c, in_stride_n;
    int32_t      fw, fh;
    int64_t      f_stride_w, f_stride_h;
    int32_t      out_w, out_h;
    int64_t      out_stride_w, out_stride_h, out_stride_c, out_stride_n;
    /* Extension (NULL = off): fp32 addend broadcast over c: y += add[n*add_stride_n + oy*add_stride_h + ox]
     * (add_stride_n = 0 broadcasts over n too).  Used by the up=2 modulated conv to fold the `x.add_(noise)` of
     * networks/generator.py:101-102 into the blur. */
    const float* add;
    int64_t      add_stride_h;
    int64_t      add_stride_n;
    /* Extension (ep_enable = 0 = off): fused bias + activation on the result, y = clamp(act(y + ep_bias[c]) * ep_gain, +-ep_clamp),
     * ep_act 1 = linear, 3 = lrelu(ep_alpha).  Lets the up=2 modulated conv fold the following bias_act into its blur. */
    int32_t      ep_enable;
    int32_t      ep_act;
    double       ep_alpha, ep_gain, ep_clamp;
    const void*  ep_bias;     /* [channels], dtype of x, or NULL */
    /* 0 = zero padding (the reference op); 1 = replicate (clamp-to-edge) padding for same-size blurs with up = down = 1 and
     * <= 5x5 taps: F.pad(x, mode='replicate') + depthwise conv with a fixed kernel, the blur behind the pixel-shuffle upsampler
     * (networks/utils/convnext_utils.py:250-255).  Returns VFM_ERR_NO_KERNEL where the streaming kernel does not apply. */
    int32_t      pad_mode;
    /* != 0: f holds one fh x fw filter per channel, f_stride_c elements apart: a depthwise conv with learned taps (fp16 / fp32, 3x3 / 5x5 / 7x7,
     * "same" zero padding: padx0 = fw/2), the dwconv of the ConvNeXt synthesis layers (networks/utils/convnext_utils.py:99,128);
     * together with flip = 1 (correlation) and the ep_bias epilogue it is nn.Conv2d(C, C, k, padding=k//2, groups=C).  Inference. */
    int64_t      f_stride_c;
} vfm_upfirdn2d_params;

VFM_API int vfm_upfirdn2d(const vfm_upfirdn2d_params* p, void* stream);

/* ------------------------------------------------------------------------------------------------------------
 * filtered_lrelu: y = down_fd( clamp( lrelu( up_fu(x + b) * up^2 * gain ), +-clamp ) ), one fused kernel.
 * Sign tensor: uint8 [N,C,s_h,s_w_bytes], 2 bits per element of the upsampled intermediate, 4 per byte,
 * bit0 = negative, bit1 = clamped -- the reference format (torch_utils/ops/filtered_lrelu.cpp:87-94,
 * filtered_lrelu.cu:494-505) so tensors saved by either implementation are interchangeable.
 * fu/fd: fp32, separable when f*_h == 0 ([taps]) else full 2-D [h,w].
 * Returns VFM_ERR_NO_KERNEL when the parameters are outside what the fused kernel supports (up/down not in
 * {1,2,4}, more than 32 taps, fp64): the caller then composes upfirdn2d + vfm_filtered_lrelu_act + upfirdn2d
 * exactly as the reference does (torch_utils/ops/filtered_lrelu.py:223-229).
 */
typedef struct {
    const void*  x;
    void*        y;
    const void*  b;          /* [C] same dtype as x (never NULL; zeros if no bias) */
    uint8_t*     s;          /* signs in/out or NULL */
    const float* fu;
    const float* fd;
    int32_t      dtype;
    int32_t      up, down;
    int32_t      fu_w, fu_h; /* fu_h == 0 -> separable */
    int32_t      fd_w, fd_h;
    int64_t      fu_stride_w, fu_stride_h, fd_stride_w, fd_stride_h;
    int32_t      pad_x0, pad_y0;
    float        gain, slope, clamp;
    int32_t      flip;
    int32_t      write_signs, read_signs;
    int32_t      x_w, x_h, channels, batch;
    int64_t      x_stride_w, x_stride_h, x_stride_c, x_stride_n;
    int32_t      y_w, y_h;
    int64_t      y_stride_w, y_stride_h, y_stride_c, y_stride_n;
    int64_t      b_stride;
    int32_t      s_w_bytes, s_h;   /* sign tensor row length in bytes and height */
    int32_t      s_ofs_x, s_ofs_y; /* offset between upsampled coordinates and sign coordinates */
    int32_t      s_w_active;       /* active width in ELEMENTS (write: yw*down-(down-1)+fd_w-1; read: s_w_bytes*4) */
    /* Extension (NULL = off): fp32 [C], ACCUMULATED into (pass zeros): y_sum[c] += sum over n, h, w of the stored output.  In the backward
     * pass (the same op with up/down swapped, filtered_lrelu.py:252-263) the output is dx, so this is the bias gradient
     * db = dx.sum([0,2,3]) of filtered_lrelu.py:266 without a second pass over dx. */
    float*       y_sum;
} vfm_filtered_lrelu_params;

VFM_API int vfm_filtered_lrelu(const vfm_filtered_lrelu_params* p, void* stream);

/* In-place gain * lrelu * clamp with sign write/read, used by the generic composition. */
typedef struct {
    void*    x;
    uint8_t* s;
    int32_t  dtype;
    double   gain, slope, clamp;
    int32_t  write_signs, read_signs;
    int32_t  x_w, x_h, channels, batch;
    int64_t  x_stride_w, x_stride_h, x_stride_c, x_stride_n;
    int32_t  s_w, s_h;             /* sign tensor width in ELEMENTS (multiple of 4) and height */
    int32_t  s_ofs_x, s_ofs_y;
} vfm_filtered_lrelu_act_params;

VFM_API int vfm_filtered_lrelu_act(const vfm_filtered_lrelu_act_params* p, void* stream);

/* ------------------------------------------------------------------------------------------------------------
 * modulated_conv2d (networks/generator.py:46-103), k in {1,3} (any odd k on the generic path), up in {1,2}, down = 1.
 *
 *   w'[n,o,i,k] = weight[o,i,k] * styles[n,i];  d[n,o] = rsqrt(sum_{i,k} w'^2 + 1e-8)  (demodulate)
 *   y[n,o]      = d[n,o] * conv2d_resample(x[n] * styles[n], weight) + noise
 *
 * The modulation is applied to the activation operand and the demodulation in the GEMM epilogue, so the
 * [N,O,I,kh,kw] per-sample weight tensor of the reference is never materialised.
 * fp16 inputs with demodulate use the reference's overflow pre-normalisation (generator.py:66-68) internally.
 * x, y: NCHW contiguous, `dtype`.  weight fp32 [O,I,kh,kw] contiguous; styles fp32 [N,I]; noise fp32.
 * dcoefs: fp32 [N,O] written by forward (ones if !demodulate) and consumed by backward.
 */
#define VFM_EP_ACT_GELU 10
enum vfm_noise_mode { VFM_NOISE_NONE = 0, VFM_NOISE_HW = 1 /* [Hout,Wout] */, VFM_NOISE_N1HW = 2 /* [N,1,Hout,Wout] */ };

typedef struct {
    int32_t      dtype;
    int32_t      batch, in_channels, out_channels, in_h, in_w, kh, kw;
    int32_t      up;               /* 1 or 2 */
    int32_t      padding;          /* symmetric, w.r.t. the upsampled image (kh/2 in the decoder) */
    int32_t      demodulate;
    int32_t      flip_weight;      /* 1 = correlation (F.conv2d), 0 = true convolution */
    int32_t      noise_mode;
    const float* resample_filter;  /* fp32 [fh,fw] contiguous 2-D, required when up == 2 */
    int32_t      fw, fh;
    int32_t      out_h, out_w;
    int32_t      force_generic;    /* 1 = skip the tcgen05 path (used by the parity tests to cross-check both) */
} vfm_modconv_desc;

typedef struct {
    vfm_modconv_desc d;
    const void*  x;        /* [N,I,H,W] */
    const float* weight;   /* [O,I,kh,kw] */
    const float* styles;   /* [N,I] */
    const float* noise;    /* per noise_mode or NULL */
    void*        y;        /* [N,O,Hout,Wout] out */
    float*       dcoefs;   /* [N,O] out */
    void*        workspace;
    size_t       workspace_bytes;
    /* Optional fused layer epilogue (inference).  When ep_enable != 0 the kernel that produces y also
     * applies what networks/generator.py:268-274 does after the conv:
     *     y = clamp(act(y + bias[o]) * ep_gain, +-ep_clamp);   if (ep_residual) y = (ep_gamma[o] * y + ep_residual) * ep_res_scale
     * ep_act is 1 (linear), 3 (lrelu, slope ep_alpha) or VFM_EP_ACT_GELU (exact erf gelu: the ConvNeXt layers'
     * pwconv1 -> GELU, networks/utils/convnext_utils.py:140-142).  Returns VFM_ERR_NO_KERNEL if the path chosen for this descriptor
     * cannot fuse it; the caller then composes bias_act (+ residual arithmetic) itself, as the reference does. */
    int32_t      ep_enable;
    int32_t      ep_act;
    double       ep_alpha, ep_gain, ep_clamp;   /* ep_clamp < 0 = off */
    const void*  ep_bias;        /* [O], dtype of x, or NULL */
    const void*  ep_residual;    /* [N,O,Hout,Wout], dtype of x, or NULL */
    const float* ep_gamma;       /* [O] fp32 layer scale, required with ep_residual */
    double       ep_res_scale;
    /* Optional per-(sample, channel) affine map of the input (the GroupNorm32 prologue of the residual layers,
     * networks/generator.py:261-263, in the form vfm_group_norm_affine produces): the conv sees x * x_scale[n,i] + x_shift[n,i]
     * (zero padding applies to the mapped tensor).  With ep_res_affine != 0 the same map is applied to ep_residual, which
     * then is the raw x (I == O).  Only the tcgen05 path implements it; other paths return VFM_ERR_NO_KERNEL. */
    const float* x_scale;        /* [N,I] fp32 or NULL */
    const float* x_shift;        /* [N,I] fp32 or NULL (requires x_scale) */
    int32_t      ep_res_affine;
    /* Training: != 0 asks the tcgen05 path to leave its activation operand -- x * s' as NHWC fp16 (hi [+ lo] for fp32 tensors), scaled by ONE
     * power of two for the whole batch -- intact in the workspace, so that the backward can feed it to the weight gradient instead of
     * re-reading and re-laying x (vfm_modconv_forward_operand / vfm_modconv_bwd_params::saved_operand).  The caller keeps the workspace
     * alive until the backward has run.  Ignored by the other paths and together with x_scale. */
    int32_t      keep_operand;
} vfm_modconv_fwd_params;

typedef struct {
    vfm_modconv_desc d;
    const void*  dy;       /* [N,O,Hout,Wout] */
    const void*  x;        /* [N,I,H,W] */
    const void*  y;        /* forward output (needed when demodulate: g[n,o] = sum dy*(y-noise)/d) */
    const float* weight;
    const float* styles;
    const float* noise;
    const float* dcoefs;   /* from forward */
    void*        dx;       /* [N,I,H,W] out, or NULL */
    float*       dweight;  /* [O,I,kh,kw] fp32 out, or NULL */
    float*       dstyles;  /* [N,I] fp32 out, or NULL */
    float*       dnoise;   /* fp32, shape per noise_mode, out, or NULL */
    void*        workspace;
    size_t       workspace_bytes;
    /* Optional: the forward's activation operand, as returned by vfm_modconv_forward_operand after a forward with keep_operand != 0 on the
     * same x / styles (NULL = recompute it from x).  Skips one read of x and one write + read of its NHWC copy per layer. */
    const void*  saved_operand;
    const void*  saved_operand_lo;   /* fp32 tensors: the lo half of the split */
    /* Training with the fused layer epilogue (forward ran with ep_enable != 0, no residual): `dy` and `y` then are the gradient / value of
     * the ACTIVATED output, y = clamp(act(conv + noise + ep_bias) * ep_gain, +-ep_clamp), and the backward of that bias_act
     * (torch_utils/ops/bias_act.py:158-179) is folded into the one pass over dy that the demodulation and noise reductions make anyway.
     * dbias_no receives sum_p of the pre-activation gradient per (sample, channel); the caller sums it over the batch.
     * ep_act 1 (linear) or 3 (lrelu); needs fp16 / fp32, >= 2048 output pixels (a multiple of 8), else VFM_ERR_NO_KERNEL. */
    int32_t      ep_enable;
    int32_t      ep_act;
    double       ep_alpha, ep_gain, ep_clamp;
    const void*  ep_bias;            /* [O], dtype of x, or NULL */
    float*       dbias_no;           /* [N,O] fp32 out, or NULL */
} vfm_modconv_bwd_params;

/* ------------------------------------------------------------------------------------------------------------
 * GroupNorm statistics as an affine map (see vfm_modconv_fwd_params::x_scale): for x [N,C,H*W] contiguous,
 *   scale[n,c] = rstd[n,g(c)] * gamma[c],   shift[n,c] = beta[c] - mean[n,g(c)] * scale[n,c]
 * with the biased variance and eps of torch.nn.GroupNorm evaluated in fp32 (networks/utils/shared.py GroupNorm32).
 */
typedef struct {
    const void*  x;          /* [N,C,HW] contiguous, `dtype` (f16 / f32) */
    const float* gamma;      /* [C] or NULL (= 1) */
    const float* beta;       /* [C] or NULL (= 0) */
    float*       scale;      /* [N,C] out */
    float*       shift;      /* [N,C] out */
    int32_t      dtype;
    int32_t      batch, channels, groups;
    int64_t      hw;
    double       eps;
} vfm_group_norm_affine_params;
VFM_API int vfm_group_norm_affine(const vfm_group_norm_affine_params* p, void* stream);

/* GroupNorm32 itself (networks/utils/shared.py: nn.GroupNorm evaluated in fp32, result cast back to x.dtype), forward and
 * backward, for the training path of the residual / ConvNeXt layers and the z-convs: statistics pass + one elementwise pass each way
 * (y = x*A + B;  dx = dy*P + x*Q + R with per-(n,c) coefficients), instead of cast -> moments -> apply -> cast.
 * dgamma / dbeta are returned per (sample, channel); the caller sums them over the batch. */
typedef struct {
    const void*  x;          /* [N,C,HW] contiguous, `dtype` (f16 / f32) */
    const float* gamma;      /* [C] or NULL */
    const float* beta;       /* [C] or NULL */
    void*        y;          /* forward out, `dtype` */
    float*       mean;       /* [N,groups] forward out / backward in */
    float*       rstd;       /* [N,groups] forward out / backward in */
    float*       scratch;    /* 3*N*C floats */
    const void*  dy;         /* backward in */
    void*        dx;         /* backward out */
    float*       dgamma_nc;  /* [N,C] backward out: sum_hw dy * xhat */
    float*       dbeta_nc;   /* [N,C] backward out: sum_hw dy */
    int32_t      dtype;
    int32_t      batch, channels, groups;
    int64_t      hw;
    double       eps;
} vfm_group_norm_params;
VFM_API int vfm_group_norm_forward(const vfm_group_norm_params* p, void* stream);
VFM_API int vfm_group_norm_backward(const vfm_group_norm_params* p, void* stream);

/* Row-wise helpers for the layer-scaled residual of the residual SynthesisLayers in training,  y = (gamma*y + x)*sqrt2
 * (networks/generator.py:272-274): a tensor is `rows` = N*C planes of `hw` contiguous elements of `dtype`.
 *   vfm_rows_affine: out[r,:] = a[r,:]*P[r] (+ b[r,:]*Q[r]) + R[r]     (b, Q optional, together; out has the dtype of a)
 *   vfm_rows_dot   : ((float*)out)[r] = sum_i a[r,i]*b[r,i]             (the layer-scale gradient) */
typedef struct {
    const void*  a;
    const void*  b;
    const float* P;
    const float* Q;
    const float* R;
    void*        out;
    int32_t      dtype;
    int64_t      rows, hw;
} vfm_rows_params;
VFM_API int vfm_rows_affine(const vfm_rows_params* p, void* stream);
VFM_API int vfm_rows_dot(const vfm_rows_params* p, void* stream);

/* PixelShuffle(2): y[n, c, 2h+i, 2w+j] = x[n, 4c + 2i + j, h, w];  x [batch, 4*out_channels, in_h, in_w] -> y [batch, out_channels,
 * 2 in_h, 2 in_w], both contiguous NCHW, fp16 / fp32, in_w % 4 == 0.  Replaces nn.PixelShuffle(2) in the reference's
 * SeparableUpsampleWithFixedBlur (networks/utils/convnext_utils.py:197-257); the Python mirror uses it under no_grad. */
typedef struct {
    const void* x;
    void*       y;
    int32_t     dtype;
    int32_t     batch, out_channels, in_h, in_w;
    int32_t     inverse;      /* != 0: PixelUnshuffle(2) (the backward): x is the [batch, out_channels, 2 in_h, 2 in_w] tensor, y the 4-plane one */
} vfm_pixel_shuffle2_params;
VFM_API int vfm_pixel_shuffle2(const vfm_pixel_shuffle2_params* p, void* stream);

/* Border rows / columns of the data gradient of  y = conv2d(pad(x, k/2, mode='replicate'), f[k,k], groups=C)  (the fixed blur behind
 * the pixel-shuffle upsampler, networks/utils/convnext_utils.py:250-255):
 *   dx[jy,jx] = sum_{iy,ix} dy[iy,ix] * sum_{ty,tx} f[ty,tx] [clamp(iy+ty-k/2) == jy] [clamp(ix+tx-k/2) == jx]
 * Away from the border this is the zero-padded "same" stencil over dy (one vfm_upfirdn2d pass, flip = 0); on the outermost row / column
 * of every plane the clamped taps fold back, and this call OVERWRITES those 2W + 2(H-2) elements of dx with the complete sums
 * (range sums of f from a summed-area table).  dy, dx contiguous [planes, h, w] of `dtype` (fp16 / fp32), f fp32 [k,k], k in {3, 5},
 * h, w >= k.  Replaces the four thin conv_transpose2d calls + slice arithmetic a host-side fold needs. */
typedef struct {
    const void*  dy;
    void*        dx;
    const float* f;
    int32_t      dtype;
    int32_t      k;
    int64_t      planes;
    int32_t      h, w;
} vfm_replicate_blur_edges_params;
VFM_API int vfm_replicate_blur_edges(const vfm_replicate_blur_edges_params* p, void* stream);

/* Weight and bias gradient of the depthwise k x k conv (k = 3, 5, 7; "same" zero padding; see vfm_upfirdn2d_params::f_stride_c for
 * the forward; the data gradient is the forward with flip = 0):
 *   dweight[c,ty,tx] += sum_{n,y,x} dy[n,c,y,x] * x[n,c,y+ty-k/2,x+tx-k/2],   dbias[c] += sum_{n,y,x} dy[n,c,y,x]   (dbias optional)
 * x, dy contiguous NCHW of `dtype` (fp16 / fp32), W % 8 == 0; dweight [C,k,k] and dbias [C] are fp32 and ACCUMULATED into
 * (atomics): pass zeroed buffers.  Replaces the stock autograd of nn.Conv2d(C, C, k, groups=C) in the ConvNeXt layers
 * (networks/utils/convnext_utils.py:99,128).  VFM_ERR_NO_KERNEL for other shapes (the caller keeps the stock op). */
typedef struct {
    const void* x;
    const void* dy;
    float*      dweight;
    float*      dbias;
    int32_t     dtype;
    int32_t     batch, channels, h, w, k;
} vfm_depthwise_wgrad_params;
VFM_API int vfm_depthwise_wgrad(const vfm_depthwise_wgrad_params* p, void* stream);

/* ------------------------------------------------------------------------------------------------------------
 * Gradient exchange, the post-all-reduce pass of the reference's sync_grads (training/training_loop.py:281-289) over the
 * flat fp32 gradient buffer, in place and in ONE pass:
 *     g = g / world_size;   if (use_gain) g = g * gain;   g = nan_to_num(g, nan, posinf, neginf)     (reference: 0, 1e5, -1e5)
 * `grads` must be 16-byte aligned.  The all-reduce itself is torch.distributed / NCCL plumbing (vfm_vae_b200/sync.py).
 */
typedef struct {
    float*   grads;       /* [numel] fp32, in place */
    int64_t  numel;
    int32_t  world_size;  /* >= 1 */
    int32_t  use_gain;    /* 0 = the reference's `gain is None` */
    double   gain;
    double   nan, posinf, neginf;
} vfm_grad_finalize_params;
VFM_API int vfm_grad_finalize(const vfm_grad_finalize_params* p, void* stream);

/* ------------------------------------------------------------------------------------------------------------
 * Decode I/O path: decoder output -> the bytes a PNG encoder takes, in one pass.  Replaces, bit for bit,
 *     images = ((images + 1) / 2).clamp(0, 1)            tools/decode/decode_latents_to_images.py:92  (tools/reconstruct likewise)
 *     to_pil_image(img.clamp(0, 1))                      safe_save, :20-24  == img.mul(255).byte() transposed to HWC (truncation)
 * y[n,h,w,c] = (uint8) trunc( clamp((x[n,c,h,w] + pre_add) / pre_div, 0, 1) * scale ),  evaluated in fp32 in exactly that order
 * (reference values: pre_add 1, pre_div 2, scale 255).  x: contiguous NCHW, fp16 or fp32; y: contiguous NHWC uint8; C in {1, 3, 4}.
 */
typedef struct {
    const void* x;          /* [N,C,H,W] */
    uint8_t*    y;          /* [N,H,W,C] out */
    int32_t     dtype;      /* vfm_dtype of x */
    int32_t     batch, channels, height, width;
    double      pre_add, pre_div, scale;
} vfm_image_to_u8_params;
VFM_API int vfm_image_to_u8(const vfm_image_to_u8_params* p, void* stream);

/* direction: 0 = forward, 1 = backward, 2 = backward with the fused layer epilogue (vfm_modconv_bwd_params::ep_enable).  Returns bytes
 * (0 is a valid answer). */
VFM_API size_t vfm_modconv_workspace_bytes(const vfm_modconv_desc* d, int direction);
/* 1 if vfm_modconv_backward can take ep_enable != 0 for this descriptor (fused training layer), else 0. */
VFM_API int vfm_modconv_fused_backward_supported(const vfm_modconv_desc* d);
VFM_API int vfm_modconv_forward(const vfm_modconv_fwd_params* p, void* stream);
VFM_API int vfm_modconv_backward(const vfm_modconv_bwd_params* p, void* stream);
/* 1 if the tcgen05/TMEM implicit-GEMM path will be used for this descriptor, 0 if the generic SIMT kernel. */
VFM_API int vfm_modconv_uses_tensor_cores(const vfm_modconv_desc* d);
/* Where a forward with keep_operand != 0 left its activation operand inside `workspace` (the pointer given to that forward):
 * *hi (and *lo for fp32 tensors, else NULL), or both NULL when the path taken for this descriptor keeps none.  Pure address arithmetic. */
VFM_API int vfm_modconv_forward_operand(const vfm_modconv_desc* d, void* workspace, size_t workspace_bytes, const void** hi, const void** lo);

#ifdef __cplusplus
}
#endif
#endif /* VFM_OPS_H */
