"""Style-modulated / demodulated conv2d on sm_100a.  Drop-in for the reference helper networks/generator.py:46-103
(same signature and argument meaning), implemented as ONE custom autograd Function over the native
``modconv_plugin.forward/backward`` entry points instead of stock autograd through a grouped cuDNN conv:

* the per-sample weight tensor [N,O,I,kh,kw] is never built: styles scale the activation operand, the demodulation
  coefficients d[n,o] scale the accumulator in the GEMM epilogue, noise is added there too;
* backward = one data-gradient conv (with the dstyles reduction in its epilogue) + one batched weight-gradient GEMM
  + two tiny fix-up kernels for the gradient through d[n,o] (formulas: SURVEY.md section 8a).

``fused_modconv`` is accepted and ignored: both reference branches compute the same function and there is only one
kernel here.  ``down`` must be 1 (the decoder never uses anything else).  weight/styles/noise are consumed in fp32
(they are fp32 parameters in the decoder); second-order gradients are not implemented."""
import torch

from ... import custom_ops

_plugin = None

#: set True (tests only) to run every call through the generic SIMT kernel instead of the tcgen05 path
force_generic = False

#: training: keep the forward's NHWC activation operand (x * s', fp16) alive for the weight gradient instead of re-laying x in the backward:
#: one read of x and one write + read of its NHWC copy less per layer, for N*H*W*C*2 bytes of extra live memory per layer (~10 GB for the
#: f16d32 decoder at batch 64).  Set False to trade the time back for the memory.
keep_forward_operand = True


def _init():
    global _plugin
    if _plugin is None:
        _plugin = custom_ops.get_plugin(module_name='modconv_plugin')
    return True


def _noise_canon(noise, n, oh, ow):
    """-> fp32 contiguous [oh,ow] or [n,1,oh,ow] view of whatever broadcastable noise the caller passed."""
    if noise is None:
        return None
    t = noise.to(torch.float32)
    if t.dim() <= 2:
        return t.expand(oh, ow).contiguous()
    t = t.expand(n, 1, oh, ow) if t.dim() == 4 else t.reshape(-1, 1, oh, ow).expand(n, 1, oh, ow)
    if t.stride(0) == 0:     # broadcast over the batch -> the cheaper [H,W] mode
        return t[0, 0].contiguous()
    return t.contiguous()


class _ModulatedConv2d(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, weight, styles, noise, up, padding, resample_filter, demodulate, flip_weight):
        n = x.shape[0]
        xc = x.contiguous()
        w32 = weight.detach().to(torch.float32).contiguous()
        s32 = styles.detach().to(torch.float32).contiguous()
        f32 = resample_filter.to(device=x.device, dtype=torch.float32).contiguous() if (up > 1 and resample_filter is not None) else None
        kh = weight.shape[2]
        oh = x.shape[2] * up + (2 * padding - kh + 1 if up == 1 else 0)
        ow = x.shape[3] * up + (2 * padding - weight.shape[3] + 1 if up == 1 else 0)
        if up > 1:
            oh = x.shape[2] * up + 2 * padding - (kh - 1)
            ow = x.shape[3] * up + 2 * padding - (weight.shape[3] - 1)
        n32 = _noise_canon(noise.detach() if noise is not None else None, n, oh, ow)
        keep = bool(keep_forward_operand and ctx.needs_input_grad[1] and not force_generic)
        out = _plugin.forward(xc, w32, s32, n32, up, padding, f32, demodulate, flip_weight, force_generic, keep_operand=keep)
        y, dcoefs = out[0], out[1]
        ctx.saved_operand = out[2] if keep else None           # (workspace tensor, hi, lo): the tensor reference keeps the bytes alive
        ctx.save_for_backward(xc, w32, s32, n32 if n32 is not None else torch.empty([0]), dcoefs,
                              y if demodulate else torch.empty([0]), f32 if f32 is not None else torch.empty([0]))
        ctx.cfg = (up, padding, demodulate, flip_weight)
        ctx.in_dtypes = (x.dtype, weight.dtype, styles.dtype, noise.dtype if noise is not None else None)
        ctx.noise_shape = tuple(noise.shape) if noise is not None else None
        return y

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, dy):
        xc, w32, s32, n32, dcoefs, y, f32 = ctx.saved_tensors
        up, padding, demodulate, flip_weight = ctx.cfg
        n32 = n32 if n32.numel() else None
        f32 = f32 if f32.numel() else None
        y = y if y.numel() else None
        need = ctx.needs_input_grad
        dx, dw, ds, dn = _plugin.backward(dy.contiguous(), xc, y, w32, s32, n32, dcoefs, up, padding, f32, demodulate, flip_weight,
                                          need_dx=need[0], need_dweight=need[1], need_dstyles=need[2],
                                          need_dnoise=(need[3] and n32 is not None), force_generic=force_generic,
                                          saved_operand=ctx.saved_operand)
        ctx.saved_operand = None
        xd, wd, sd, nd = ctx.in_dtypes
        if dx is not None and not need[0]:
            dx = None
        if dw is not None:
            dw = dw.to(wd)
        if ds is not None:
            ds = ds.to(sd)
        if dn is not None:
            dn = _undo_noise_broadcast(dn, ctx.noise_shape, nd)
        return dx, dw, ds, dn, None, None, None, None, None


def _undo_noise_broadcast(dn, shape, dtype):
    """gradient of the canonical [H,W] / [N,1,H,W] noise -> the shape the caller passed"""
    full = dn if dn.dim() == 4 else dn[None, None]
    while full.dim() > len(shape):
        full = full.sum(0)
    for i, sz in enumerate(shape):
        if sz == 1 and full.shape[i] != 1:
            full = full.sum(i, keepdim=True)
    return full.reshape(shape).to(dtype)


class _FusedModconvBiasAct(torch.autograd.Function):
    """One legacy synthesis layer for TRAINING: modulated conv + noise + bias_act (networks/generator.py:264-270) as one autograd node.
    Forward = the conv (blur for up=2) kernel with the bias / lrelu / gain / clamp epilogue: only the activated output is written and kept.
    Backward = the bias_act gradient folded into the single pass over the incoming gradient that the demodulation / noise reductions make
    anyway (csrc/modconv_generic.cu act_grad_gsum_dnoise_kernel), then the usual data / weight gradient kernels."""

    @staticmethod
    def forward(ctx, x, weight, styles, bias, noise, up, padding, resample_filter, flip_weight, act, alpha, gain, clamp):
        n = x.shape[0]
        xc = x.contiguous()
        w32 = weight.detach().to(torch.float32).contiguous()
        s32 = styles.detach().to(torch.float32).contiguous()
        f32 = resample_filter.to(device=x.device, dtype=torch.float32).contiguous() if (up > 1 and resample_filter is not None) else None
        kh, kw = weight.shape[2], weight.shape[3]
        oh = x.shape[2] * up + 2 * padding - (kh - 1)
        ow = x.shape[3] * up + 2 * padding - (kw - 1)
        n32 = _noise_canon(noise.detach() if noise is not None else None, n, oh, ow)
        b = bias.detach().to(x.dtype).contiguous() if bias is not None else None
        ep = dict(act=act, alpha=alpha, gain=gain, clamp=clamp, bias=b)
        keep = bool(keep_forward_operand and ctx.needs_input_grad[1])
        out = _plugin.forward(xc, w32, s32, n32, up, padding, f32, True, flip_weight, False, epilogue=ep, keep_operand=keep)
        if out is None:
            raise RuntimeError('fused synthesis layer: no kernel fuses this call (checked by the caller)')
        y, dcoefs = out[0], out[1]
        ctx.saved_operand = out[2] if keep else None
        ctx.save_for_backward(xc, w32, s32, n32 if n32 is not None else torch.empty([0]), dcoefs, y, f32 if f32 is not None else torch.empty([0]),
                              b if b is not None else torch.empty([0]))
        ctx.cfg = (up, padding, flip_weight, act, alpha, gain, clamp)
        ctx.in_dtypes = (weight.dtype, styles.dtype, bias.dtype if bias is not None else None, noise.dtype if noise is not None else None)
        ctx.noise_shape = tuple(noise.shape) if noise is not None else None
        return y

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, dy):
        xc, w32, s32, n32, dcoefs, y, f32, b = ctx.saved_tensors
        up, padding, flip_weight, act, alpha, gain, clamp = ctx.cfg
        n32 = n32 if n32.numel() else None
        f32 = f32 if f32.numel() else None
        b = b if b.numel() else None
        need = ctx.needs_input_grad
        dx, dw, ds, dn, db_no = _plugin.backward(dy.contiguous(), xc, y, w32, s32, n32, dcoefs, up, padding, f32, True, flip_weight,
                                                 need_dx=need[0], need_dweight=need[1], need_dstyles=need[2], need_dnoise=(need[4] and n32 is not None),
                                                 saved_operand=ctx.saved_operand, epilogue=dict(act=act, alpha=alpha, gain=gain, clamp=clamp, bias=b))
        ctx.saved_operand = None
        wd, sd, bd, nd = ctx.in_dtypes
        dx = dx if need[0] else None
        dw = dw.to(wd) if dw is not None else None
        ds = ds.to(sd) if ds is not None else None
        db = db_no.sum(0).to(bd) if (need[3] and bd is not None) else None
        dn = _undo_noise_broadcast(dn, ctx.noise_shape, nd) if dn is not None else None
        return dx, dw, ds, db, dn, None, None, None, None, None, None, None, None


def fused_synthesis_layer_train(x, weight, styles, bias, noise=None, up=1, padding=0, resample_filter=None, flip_weight=True, act='lrelu', alpha=0.2,
                                gain=1.0, clamp=None):
    """Training-time fusion of ``bias_act(modulated_conv2d(x, weight, styles, noise, up, ...), bias, act, gain, clamp)`` (demodulated 3x3 layers of
    the legacy decoder) with full first-order autograd for x, weight, styles, bias and noise.  Returns None when the kernels cannot fuse the
    call (the caller then composes the two ops as the reference does): needs CUDA fp16 / fp32, the tcgen05 path, act in {linear, lrelu},
    gain > 0 and >= 2048 output pixels."""
    if x.device.type != 'cuda' or act not in ('linear', 'lrelu') or not gain > 0 or x.dtype not in (torch.float16, torch.float32) or force_generic:
        return None
    _init()
    xs = x.detach()
    w32 = weight.detach().to(torch.float32)
    f32 = resample_filter.to(device=x.device, dtype=torch.float32).contiguous() if (up > 1 and resample_filter is not None) else None
    kw = dict(up=int(up), padding=int(padding), demodulate=True, flip_weight=bool(flip_weight), resample_filter=f32)
    if not (_plugin.uses_tensor_cores(xs, w32, **kw) and _plugin.fused_backward_supported(xs, w32, **kw)):
        return None
    return _FusedModconvBiasAct.apply(x, weight, styles, bias, noise, int(up), int(padding), resample_filter, bool(flip_weight), act, float(alpha),
                                      float(gain), None if clamp is None else float(clamp))


def modulated_conv2d(x, weight, styles, noise=None, up=1, down=1, padding=0, resample_filter=None, demodulate=True,
                     flip_weight=True, fused_modconv=True):
    """Arguments as in the reference (networks/generator.py:46-58)."""
    assert isinstance(x, torch.Tensor) and x.ndim == 4
    batch_size = x.shape[0]
    out_channels, in_channels, kh, kw = weight.shape
    assert x.shape[1] == in_channels, f'x has {x.shape[1]} channels, weight expects {in_channels}'
    assert tuple(styles.shape) == (batch_size, in_channels), f'styles must be [{batch_size}, {in_channels}]'
    if down != 1:
        raise NotImplementedError('vfm_vae_b200.modulated_conv2d: down > 1 is not on the decoder path and is not implemented')
    if x.device.type != 'cuda':
        raise RuntimeError('vfm_vae_b200.modulated_conv2d has no reference/CPU implementation: CUDA tensors only '
                           '(the CPU oracle is oracle/ref_ops.py, for tests).')
    _init()
    return _ModulatedConv2d.apply(x, weight, styles, noise, int(up), int(padding), resample_filter, bool(demodulate), bool(flip_weight))


def fused_modconv_bias_act(x, weight, styles, bias, noise=None, up=1, padding=0, resample_filter=None, demodulate=True,
                           flip_weight=True, act='lrelu', alpha=0.2, gain=1.0, clamp=None, residual=None, gamma=None, res_scale=1.0,
                           group_norm=None):
    """Inference-only fusion of one legacy synthesis layer:

        x = GroupNorm32(x)        (group_norm=dict(weight, bias, num_groups, eps); residual layers)  networks/generator.py:261-263
        y = modulated_conv2d(x, weight, styles, noise, up, ...)                      networks/generator.py:264-265
        y = bias_act(y, bias, act=act, gain=gain, clamp=clamp)                       networks/generator.py:268-270
        y = (gamma * y + residual) * res_scale            (residual layers only)     networks/generator.py:272-274

    in the kernel that writes y (conv epilogue for up=1, blur epilogue for up=2, streaming kernel for ToRGB).  No autograd.
    With ``group_norm`` the normalisation costs one statistics pass over x: it becomes a per-(sample, channel) affine map that the
    conv's operand pre-pass applies to x and its epilogue applies to the residual; ``residual`` must then be ``x`` itself (raw).
    Returns None when no kernel can fuse this call; the caller then composes the unfused ops exactly as the reference does.
    """
    assert act in ('linear', 'lrelu', 'gelu')
    if x.device.type != 'cuda' or torch.is_grad_enabled() and (x.requires_grad or weight.requires_grad or styles.requires_grad):
        return None
    if residual is not None and up != 1:
        return None
    if group_norm is not None and (up != 1 or (residual is not None and residual is not x)):
        return None
    _init()
    n = x.shape[0]
    kh = weight.shape[2]
    xc = x.contiguous()
    w32 = weight.detach().to(torch.float32).contiguous()
    s32 = styles.detach().to(torch.float32).contiguous()
    f32 = resample_filter.to(device=x.device, dtype=torch.float32).contiguous() if (up > 1 and resample_filter is not None) else None
    if up == 1:
        oh, ow = x.shape[2] + 2 * padding - kh + 1, x.shape[3] + 2 * padding - weight.shape[3] + 1
    else:
        oh, ow = x.shape[2] * up + 2 * padding - (kh - 1), x.shape[3] * up + 2 * padding - (weight.shape[3] - 1)
    n32 = _noise_canon(noise.detach() if noise is not None else None, n, oh, ow)
    ep = dict(act=act, alpha=alpha, gain=gain, clamp=clamp, bias=bias.detach().to(x.dtype).contiguous() if bias is not None else None,
              residual=residual.detach().contiguous() if residual is not None else None, gamma=gamma, res_scale=res_scale)
    x_affine = None
    if group_norm is not None:
        if not _plugin.uses_tensor_cores(xc, w32, up=up, padding=padding, demodulate=demodulate, flip_weight=flip_weight, noise=n32, resample_filter=f32):
            return None
        x_affine = _plugin.group_norm_affine(xc, group_norm.get('weight'), group_norm.get('bias'), group_norm['num_groups'], group_norm.get('eps', 1e-5))
        if residual is not None:
            ep['residual'] = xc
            ep['residual_affine'] = True
    out = _plugin.forward(xc, w32, s32, n32, int(up), int(padding), f32, bool(demodulate), bool(flip_weight), force_generic, epilogue=ep,
                          x_affine=x_affine)
    return None if out is None else out[0]


def modulated_pointwise_conv2d(x, weight, style, bias=None, demodulate=True):
    """Drop-in for networks/utils/convnext_utils.py:36 (the ConvNeXt layers' 1x1 modulated conv): the same function as
    ``modulated_conv2d`` with a 1x1 kernel (its fp16 pre-normalisation ``(1/I)**0.5 / max|w|`` is generator.py:66-68 with kh = kw = 1),
    plus the broadcast bias.  Runs on the tcgen05 kernel when the channel counts are multiples of 128; autograd as modulated_conv2d.
    Like the reference under autocast, the contraction runs in the autocast dtype."""
    if torch.is_autocast_enabled() and x.is_cuda:
        x = x.to(torch.get_autocast_dtype('cuda'))
    y = modulated_conv2d(x, weight, style, noise=None, up=1, padding=0, demodulate=demodulate, flip_weight=True)
    if bias is not None:
        y = y + bias          # [1,O,1,1] fp32 parameter: promotes like the reference does
    return y


def fused_convnext_mlp(x, x_in, norm_weight, norm_bias, num_groups, eps, w1, b1, style, w2, b2, gamma, demodulate=True):
    """Inference-only fusion of the pointwise half of a ConvNeXt synthesis layer (networks/utils/convnext_utils.py:138-147):

        h = gelu(modulated_pointwise_conv2d(GroupNorm32(x), w1, style, b1));   y = gamma * (conv1x1(h, w2) + b2) + x_in

    as: one GroupNorm statistics pass (vfm_group_norm_affine) + two tcgen05 1x1 convs, the first with the normalisation folded
    into its operand pre-pass and bias + GELU in its epilogue, the second with bias, layer scale and the residual in its
    epilogue.  x: the depthwise-conv output (+noise), x_in: the layer input; both [N,C,H,W] in the same fp16/fp32 dtype.
    Returns None when the shapes do not map onto the tensor-core kernel (the caller composes the reference ops)."""
    if x.device.type != 'cuda' or torch.is_grad_enabled() and (x.requires_grad or w1.requires_grad or w2.requires_grad or style.requires_grad):
        return None
    _init()
    c, c4 = w1.shape[1], w1.shape[0]
    if c % 128 != 0 or c4 % 128 != 0 or x.dtype not in (torch.float16, torch.float32) or x_in.dtype != x.dtype:
        return None
    xc = x.contiguous()
    w1f = w1.detach().to(torch.float32).contiguous()
    s32 = style.detach().to(torch.float32).contiguous()
    if not _plugin.uses_tensor_cores(xc, w1f, up=1, padding=0, demodulate=demodulate):
        return None
    aff = _plugin.group_norm_affine(xc, norm_weight, norm_bias, num_groups, eps)
    ep1 = dict(act='gelu', gain=1.0, clamp=None, bias=b1.detach().reshape(-1).to(x.dtype).contiguous() if b1 is not None else None)
    out = _plugin.forward(xc, w1f, s32, None, 1, 0, None, bool(demodulate), True, force_generic, epilogue=ep1, x_affine=aff)
    if out is None:
        return None
    h = out[0]
    w2f = w2.detach().to(torch.float32).contiguous()
    ones = torch.ones([x.shape[0], c4], dtype=torch.float32, device=x.device)
    g = gamma if gamma is not None else torch.ones([c], dtype=torch.float32, device=x.device)
    ep2 = dict(act='linear', gain=1.0, clamp=None, bias=b2.detach().reshape(-1).to(x.dtype).contiguous() if b2 is not None else None,
               residual=x_in.detach().contiguous(), gamma=g, res_scale=1.0)
    out = _plugin.forward(h, w2f, ones, None, 1, 0, None, False, True, force_generic, epilogue=ep2)
    return None if out is None else out[0]
