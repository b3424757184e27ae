"""Host-side mirror of the reference's legacy pixel decoder (``use_convnext=False``): the caller of the hot-path ops.

This is the path BASELINE.json's north_star describes: networks/generator.py ``SynthesisNetwork`` built from
``SynthesisLayer`` / ``ToRGBLayer`` (generator.py:188-310), ``SynthesisBlock`` (generator.py:320-576) and the z-concat /
self-attention / multi-scale-output plumbing around them (generator.py:655-912, networks/utils/gigagan_utils.py,
networks/utils/convnext_utils.py:197-257, networks/utils/shared.py).  Every modulated conv, bias_act and upfirdn2d goes
through the sm_100a kernels; the rest (1x1/depthwise z-convs, GroupNorm, self-attention at <= 32x32 tokens, pixel
shuffle) is out-of-scope glue and uses stock torch modules, exactly as the reference does.

Module and parameter names equal the reference's, so a reference ``state_dict`` loads with ``strict=True`` (tested
against a golden checkpoint in tests/test_decoder.py).  The reference decoder itself also runs unchanged on these
kernels via ``vfm_vae_b200.integration.install()``; this mirror exists because the reference cannot travel to the
benchmark box.

``ops`` injection: the constructor takes an ``ops`` namespace (default: the CUDA ops).  Tests pass the CPU oracle there
to check this file's host logic against the golden vectors without a GPU; the product default never touches the oracle.
"""
import math
import os
from types import SimpleNamespace

import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F


def default_ops():
    from .torch_utils.ops import bias_act, upfirdn2d
    from .torch_utils.ops.modulated_conv2d import (modulated_conv2d, fused_modconv_bias_act, modulated_pointwise_conv2d, fused_convnext_mlp,
                                                   fused_synthesis_layer_train)
    return SimpleNamespace(fused_layer=fused_modconv_bias_act, fused_layer_train=None if os.environ.get('VFM_NO_FUSED_TRAIN') else fused_synthesis_layer_train,
                           bias_act=bias_act.bias_act, def_gain=lambda act: bias_act.activation_funcs[act].def_gain,
                           setup_filter=upfirdn2d.setup_filter, upsample2d=upfirdn2d.upsample2d, blur2d_replicate=upfirdn2d.blur2d_replicate,
                           depthwise_conv2d=upfirdn2d.depthwise_conv2d, pixel_shuffle2=upfirdn2d.pixel_shuffle2,
                           modulated_conv2d=modulated_conv2d, modulated_pointwise_conv2d=modulated_pointwise_conv2d,
                           fused_convnext_mlp=None if os.environ.get('VFM_NO_FUSED_CONVNEXT') else fused_convnext_mlp)


# ------------------------------------------------------------------------------------------- small shared layers

class Conv1x1(nn.Conv2d):
    """A plain 1x1 conv of the glue layers (attention projections / feed-forward, z-convs, the upsampler's pointwise conv) --
    same parameters and state-dict names as ``nn.Conv2d(cin, cout, 1)``.

    The reference trains with TF32 off (training/training_loop.py:504-505), which leaves cuDNN with its fp32 SIMT kernels for
    these layers: forward, data and weight gradients together were ~30% of the training step.  In exactly that strict-fp32
    setting (fp32 CUDA tensors, no autocast, ``torch.backends.cudnn.allow_tf32 == False``) the conv runs on the library's
    modulated-conv kernels with unit styles and no demodulation instead: tcgen05 with the 2-term fp16 split (error ~1e-6, i.e.
    fp32-grade like the kernels it replaces), forward and backward.  With TF32 allowed (PyTorch's default, the decode tools)
    cuDNN's own tensor-core kernels are faster and the stock path is kept."""

    def forward(self, x):
        if (x.is_cuda and x.dtype == torch.float32 and not torch.is_autocast_enabled() and not torch.backends.cudnn.allow_tf32
                and self.in_channels % 128 == 0 and self.out_channels % 128 == 0):
            from .torch_utils.ops.modulated_conv2d import modulated_conv2d      # CUDA only: never reached by the CPU (oracle) runs
            ones = torch.ones([x.shape[0], self.in_channels], dtype=torch.float32, device=x.device)
            y = modulated_conv2d(x.contiguous(), self.weight, ones, demodulate=False)
            return y if self.bias is None else y + self.bias.reshape(1, -1, 1, 1)
        return super().forward(x)


class DepthwiseConv2d(nn.Conv2d):
    """``nn.Conv2d(C, C, k, padding=k // 2, groups=C)`` of the glue layers (z-convs, generator.py:729-783; the pixel-shuffle upsampler,
    convnext_utils.py:219) -- same parameters and state-dict names.  On CUDA it runs on the library's streaming stencil kernel
    (``depthwise_conv2d``: forward, data and weight gradients) wherever that applies.

    This is not only faster: PyTorch's stock fp16 depthwise conv in NCHW layout is WRONG on this platform (B200, torch 2.11.0+cu128, cuDNN
    9.22): for images of 32x32 and larger it returns garbage / NaN that depends on the allocator's state (tools/stock_fp16_probe.py;
    channels_last is fine).  Where the library kernel does not apply, fp16 inputs therefore take the channels_last stock kernel."""

    def __init__(self, channels, kernel_size=3, bias=False, ops=None):
        super().__init__(channels, channels, kernel_size, padding=kernel_size // 2, groups=channels, bias=bias)
        self.ops = ops

    def forward(self, x):
        if x.is_cuda:
            if torch.is_autocast_enabled():
                x = x.to(torch.get_autocast_dtype('cuda'))
            dw = getattr(self.ops, 'depthwise_conv2d', None) if self.ops is not None else None
            y = dw(x.contiguous(), self.weight, self.bias) if dw is not None else None
            if y is not None:
                return y
            if x.dtype == torch.float16:
                w = self.weight.to(x.dtype)
                b = self.bias.to(x.dtype) if self.bias is not None else None
                return F.conv2d(x.contiguous(memory_format=torch.channels_last), w, b, padding=self.padding, groups=self.groups).contiguous()
        return super().forward(x)


class FullyConnectedLayer(nn.Module):
    """Equalised-lr linear layer (networks/utils/shared.py:24-106)."""

    def __init__(self, in_features, out_features, bias=True, activation='linear', lr_multiplier=1.0, weight_init=1.0, bias_init=0.0):
        super().__init__()
        self.in_features, self.out_features, self.activation = in_features, out_features, activation
        self.weight = nn.Parameter(torch.randn(out_features, in_features) * (weight_init / lr_multiplier))
        self.bias = nn.Parameter(torch.full((out_features,), bias_init / lr_multiplier)) if bias else None
        self.weight_gain = lr_multiplier / math.sqrt(in_features)
        self.bias_gain = lr_multiplier

    def forward(self, x):
        w = self.weight.to(x.dtype) * self.weight_gain
        b = self.bias.to(x.dtype) * self.bias_gain if self.bias is not None else None
        if self.activation == 'linear':
            return torch.addmm(b.unsqueeze(0), x, w.t()) if b is not None else x.matmul(w.t())
        y = x.matmul(w.t())
        if b is not None:
            y = y + b
        return {'relu': F.relu, 'lrelu': lambda t: F.leaky_relu(t, 0.2), 'gelu': F.gelu}[self.activation](y)


class StyleSplit(nn.Module):
    """w -> 3C -> m1*m2+m3 (networks/utils/shared.py:166-175)."""

    def __init__(self, in_channels, out_channels, **kwargs):
        super().__init__()
        self.proj = FullyConnectedLayer(in_channels, 3 * out_channels, **kwargs)

    def forward(self, x):
        m1, m2, m3 = self.proj(x).chunk(3, 1)
        return m1 * m2 + m3


class GroupNorm32(nn.GroupNorm):
    """networks/utils/shared.py GroupNorm32: fp32 statistics, result in x.dtype.  CUDA tensors go through the library's
    two-pass kernels (forward and backward); everything else through the stock module, exactly as the reference."""

    def forward(self, x):
        if x.is_cuda:
            from .torch_utils.ops import group_norm as _gn          # CUDA only: never reached by the CPU (oracle) runs
            if _gn.supported(x, self.num_groups):
                return _gn.group_norm32(x, self.num_groups, self.weight, self.bias, self.eps)
        return super().forward(x.float()).type(x.dtype)


class ChannelRMSNorm(nn.Module):
    def __init__(self, dim):
        super().__init__()
        self.scale = dim ** 0.5
        self.gamma = nn.Parameter(torch.ones(dim, 1, 1))

    def forward(self, x):
        return F.normalize(x, dim=1) * self.scale * self.gamma


class SelfAttention(nn.Module):
    """networks/utils/gigagan_utils.py:46-91 (null key/value, SDPA)."""

    def __init__(self, dim, dim_head=64, heads=8):
        super().__init__()
        self.heads = heads
        inner = dim_head * heads
        self.norm = ChannelRMSNorm(dim)
        self.to_q = Conv1x1(dim, inner, 1, bias=False)
        self.to_k = Conv1x1(dim, inner, 1, bias=False)
        self.to_v = Conv1x1(dim, inner, 1, bias=False)
        self.null_kv = nn.Parameter(torch.randn(2, heads, dim_head) * 0.02)
        self.to_out = Conv1x1(inner, dim, 1, bias=False)
        nn.init.zeros_(self.to_out.weight)

    def forward(self, fmap):
        B, _, H, W = fmap.shape
        h = self.heads
        fmap = self.norm(fmap)

        def split(t):   # b (h d) x y -> b h (x y) d
            return t.reshape(B, h, -1, H * W).transpose(2, 3)

        q, k, v = split(self.to_q(fmap)), split(self.to_k(fmap)), split(self.to_v(fmap))
        nk, nv = (t[None, :, None, :].expand(B, -1, -1, -1).to(k.dtype) for t in self.null_kv)
        k = torch.cat((nk, k), dim=-2)
        v = torch.cat((nv, v), dim=-2)
        # q is a transposed view: with a contiguous last dimension SDPA picks its memory-efficient kernel instead of the
        # materialise-the-scores math path (same function, 2.4x faster at 32x32 tokens in fp32)
        out = F.scaled_dot_product_attention(q.contiguous(), k, v)
        out = out.transpose(2, 3).reshape(B, -1, H, W)
        return self.to_out(out)


class SelfAttentionBlock(nn.Module):
    def __init__(self, dim, dim_head=64, heads=8, ff_mult=4):
        super().__init__()
        self.attn = SelfAttention(dim=dim, dim_head=dim_head, heads=heads)
        hidden = int(dim * ff_mult)
        proj2 = Conv1x1(hidden, dim, 1)
        nn.init.zeros_(proj2.weight)
        self.ff = nn.Sequential(ChannelRMSNorm(dim), Conv1x1(dim, hidden, 1), nn.GELU(), proj2)

    def forward(self, x):
        x = self.attn(x) + x
        return self.ff(x) + x


_BLUR_TAPS = {'3x3': [1, 2, 1], '4x4': [1, 3, 3, 1], '5x5': [1, 4, 6, 4, 1]}


class SeparableUpsampleWithFixedBlur(nn.Module):
    """GN -> dw3x3 -> 1x1 -> PixelShuffle(2) -> replicate-pad -> fixed binomial blur
    (networks/utils/convnext_utils.py:197-257).  Used by the legacy path only for ``last_upsample_conv``."""

    def __init__(self, in_channels, out_channels, upscale_factor=2, blur_kernel='3x3', pre_normalize=True, use_gaussian_blur=True, ops=None):
        super().__init__()
        self.ops = ops
        self.out_channels, self.pre_normalize, self.use_gaussian_blur = out_channels, pre_normalize, use_gaussian_blur
        nc = in_channels if pre_normalize else out_channels
        self.norm = nn.GroupNorm(min(32, nc // 4), nc)
        self.depthwise = DepthwiseConv2d(in_channels, 3, bias=False, ops=ops)
        self.pointwise = Conv1x1(in_channels, out_channels * upscale_factor ** 2, 1, bias=False)
        self.shuffle = nn.PixelShuffle(upscale_factor)
        if use_gaussian_blur:
            k = torch.tensor(_BLUR_TAPS[blur_kernel] if isinstance(blur_kernel, str) else blur_kernel, dtype=torch.float32)
            k2 = torch.outer(k, k)
            k2 = k2 / k2.sum()
            kh, kw = k2.shape
            ph, pw = (kh - 1) // 2, (kw - 1) // 2
            self.pad = (pw, pw + int(kw % 2 == 0), ph, ph + int(kh % 2 == 0))
            self.register_buffer('blur_weight', k2[None, None].repeat(out_channels, 1, 1, 1))

    # On CUDA the GroupNorm (stock: cast -> fp32 moments/normalise -> cast under autocast), the 3x3 depthwise conv and the
    # PixelShuffle run on the library's kernels, forward and backward (statistics in fp32, each result rounded to the activation dtype exactly where the
    # reference's autocast rounds it); 4.7 ms of the 56 ms f16d32-D decode step were these three stock ops at the two largest blocks.
    def _norm(self, x):
        if x.is_cuda:
            from .torch_utils.ops import group_norm as _gn          # forward and backward kernels
            if _gn.supported(x, self.norm.num_groups):
                return _gn.group_norm32(x, self.norm.num_groups, self.norm.weight, self.norm.bias, self.norm.eps)
        return self.norm(x)

    def _depthwise(self, x):
        return self.depthwise(x)

    def _shuffle(self, x):
        ps = getattr(self.ops, 'pixel_shuffle2', None) if (self.ops is not None and x.is_cuda and self.shuffle.upscale_factor == 2) else None
        y = ps(x) if ps is not None else None
        return y if y is not None else self.shuffle(x)

    def forward(self, x):
        amp_dtype = torch.get_autocast_dtype('cuda') if (x.is_cuda and torch.is_autocast_enabled()) else None
        if self.pre_normalize:
            # statistics AND normalisation on the uncast input (at the fp32 -> fp16 block boundary x_sum is fp32): autocast runs the reference's
            # nn.GroupNorm in fp32 and only the following conv rounds to fp16
            x = self._norm(x)
            if amp_dtype is not None:
                x = x.to(amp_dtype)
            x = self._shuffle(self.pointwise(self._depthwise(x)))
        else:
            if amp_dtype is not None:
                x = x.to(amp_dtype)
            x = self._norm(self._shuffle(self.pointwise(self._depthwise(x))))
        if self.use_gaussian_blur:
            fused = getattr(self.ops, 'blur2d_replicate', None) if self.ops is not None else None
            if fused is not None and x.is_cuda:
                y = fused(x, self.blur_weight[0, 0], self.pad)       # replicate pad + fixed blur in one pass (forward and backward)
                if y is not None:
                    return y
            x = F.conv2d(F.pad(x, self.pad, mode='replicate'), self.blur_weight.to(x.dtype) if not torch.is_autocast_enabled() else self.blur_weight,
                         groups=self.out_channels)
        return x


# ------------------------------------------------------------------------------------------- hot-path layers

class SynthesisLayer(nn.Module):
    """modulated 3x3 conv (+2x up) -> noise -> bias + lrelu + clamp [-> layer-scaled residual]   (generator.py:188-281)"""

    def __init__(self, in_channels, out_channels, w_dim, resolution, kernel_size=3, up=1, use_noise=True, activation='lrelu',
                 resample_filter=(1, 3, 3, 1), conv_clamp=None, channels_last=False, layer_scale_init=1e-5, residual=False,
                 gn_groups=32, ops=None):
        super().__init__()
        if residual:
            assert in_channels == out_channels and up == 1
        self.ops = ops
        self.in_channels, self.out_channels, self.w_dim, self.resolution, self.up = in_channels, out_channels, w_dim, resolution, up
        self.use_noise, self.activation, self.conv_clamp, self.residual = use_noise, activation, conv_clamp, residual
        self.register_buffer('resample_filter', ops.setup_filter(list(resample_filter)))
        self.padding = kernel_size // 2
        self.act_gain = ops.def_gain(activation)
        if use_noise:
            self.register_buffer('noise_const', torch.randn([resolution, resolution]))
            self.noise_strength = nn.Parameter(torch.zeros([]))
        self.affine = StyleSplit(w_dim, in_channels, bias_init=1)
        self.weight = nn.Parameter(torch.randn([out_channels, in_channels, kernel_size, kernel_size]))
        self.bias = nn.Parameter(torch.zeros([out_channels]))
        if residual:
            self.norm = GroupNorm32(gn_groups, out_channels)
            self.gamma = nn.Parameter(layer_scale_init * torch.ones([1, out_channels, 1, 1]))

    def forward(self, x, w, noise_mode='const', fused_modconv=True, gain=1):
        assert noise_mode in ('const', 'random', 'none')
        dtype = x.dtype
        r_in = self.resolution // self.up
        assert x.shape[1] == self.in_channels and x.shape[2] == r_in and x.shape[3] == r_in, f'bad input shape {tuple(x.shape)}'
        noise = None
        if self.use_noise and noise_mode == 'random':
            noise = torch.randn([x.shape[0], 1, self.resolution, self.resolution], device=x.device) * self.noise_strength
        if self.use_noise and noise_mode == 'const':
            noise = self.noise_const * self.noise_strength
        styles = self.affine(w)
        act_clamp = self.conv_clamp * gain if self.conv_clamp is not None else None
        fused = getattr(self.ops, 'fused_layer', None)
        can_fuse = fused is not None and not torch.is_grad_enabled() and self.activation in ('linear', 'lrelu')
        if can_fuse and self.residual:
            # inference, residual layer: GroupNorm32 -> conv + noise + bias_act -> layer-scaled residual with ONE statistics pass
            # over x; the normalisation itself is folded into the conv's operand pre-pass and epilogue
            y = fused(x, self.weight, styles, self.bias, noise=noise, up=self.up, padding=self.padding, resample_filter=self.resample_filter,
                      flip_weight=True, act=self.activation, gain=self.act_gain * gain, clamp=act_clamp, residual=x, gamma=self.gamma,
                      res_scale=float(np.sqrt(2)),
                      group_norm=dict(weight=self.norm.weight, bias=self.norm.bias, num_groups=self.norm.num_groups, eps=self.norm.eps))
            if y is not None:
                return y
        if self.residual:
            x = self.norm(x)
        if can_fuse:
            # inference: conv + noise + bias_act (+ layer-scaled residual) in one kernel epilogue
            y = fused(x, self.weight, styles, self.bias, noise=noise, up=self.up, padding=self.padding, resample_filter=self.resample_filter,
                      flip_weight=(self.up == 1), act=self.activation, gain=self.act_gain * gain, clamp=act_clamp,
                      residual=x if self.residual else None, gamma=self.gamma if self.residual else None, res_scale=float(np.sqrt(2)))
            if y is not None:
                return y
        y = None
        fused_train = getattr(self.ops, 'fused_layer_train', None)
        if fused_train is not None and torch.is_grad_enabled() and self.activation in ('linear', 'lrelu') and x.is_cuda:
            # training: conv + noise + bias_act as ONE autograd node (only the activated output is written and kept; the bias_act gradient is
            # folded into the backward's single pass over dy); None where the kernels do not fuse the shape
            y = fused_train(x, self.weight, styles, self.bias, noise=noise, up=self.up, padding=self.padding, resample_filter=self.resample_filter,
                            flip_weight=(self.up == 1), act=self.activation, gain=self.act_gain * gain, clamp=act_clamp)
        if y is None:
            y = self.ops.modulated_conv2d(x=x, weight=self.weight, styles=styles, noise=noise, up=self.up, padding=self.padding,
                                          resample_filter=self.resample_filter, flip_weight=(self.up == 1), fused_modconv=fused_modconv)
            y = y.to(dtype)
            y = self.ops.bias_act(y, self.bias.to(x.dtype), act=self.activation, gain=self.act_gain * gain, clamp=act_clamp)
        if self.residual:
            if y.is_cuda:
                from .torch_utils.ops import layer_scale as _ls         # CUDA only: never reached by the CPU (oracle) runs
                if _ls.supported(y, x, self.gamma):
                    return _ls.layer_scale_residual(y, x, self.gamma, float(np.sqrt(2)))
            y = (self.gamma * y).to(dtype).add_(x).mul(np.sqrt(2))
        return y


class ToRGBLayer(nn.Module):
    """1x1 modulated conv without demodulation -> bias + clamp   (generator.py:284-312)"""

    def __init__(self, in_channels, out_channels, w_dim, kernel_size=1, conv_clamp=None, channels_last=False, ops=None):
        super().__init__()
        self.ops = ops
        self.in_channels, self.out_channels, self.w_dim, self.conv_clamp = in_channels, out_channels, w_dim, conv_clamp
        self.affine = StyleSplit(w_dim, in_channels, bias_init=1)
        self.weight = nn.Parameter(0.1 * torch.randn([out_channels, in_channels, kernel_size, kernel_size]))
        self.bias = nn.Parameter(torch.zeros([out_channels]))
        self.weight_gain = 1 / np.sqrt(in_channels * (kernel_size ** 2))

    def forward(self, x, w):
        styles = self.affine(w) * self.weight_gain
        fused = getattr(self.ops, 'fused_layer', None)
        if fused is not None and not torch.is_grad_enabled():
            y = fused(x, self.weight, styles, self.bias, demodulate=False, act='linear', gain=1.0, clamp=self.conv_clamp)
            if y is not None:
                return y
        x = self.ops.modulated_conv2d(x=x, weight=self.weight, styles=styles, demodulate=False)
        return self.ops.bias_act(x, self.bias.to(x.dtype), clamp=self.conv_clamp)


# ------------------------------------------------------------------------------------------- ConvNeXt-variant layers
# (use_convnext=True: what the shipped YAMLs run; SURVEY.md 8f row 1.  networks/utils/convnext_utils.py:60-187)

class ModulatedPointwiseConv2DLayer(nn.Module):
    def __init__(self, in_channels, out_channels, demodulate=True, ops=None):
        super().__init__()
        self.ops, self.demodulate = ops, demodulate
        self.weight = nn.Parameter(torch.empty([out_channels, in_channels, 1, 1]))
        self.bias = nn.Parameter(torch.zeros(1, out_channels, 1, 1))
        nn.init.trunc_normal_(self.weight, std=0.02)

    def forward(self, x, style):
        return self.ops.modulated_pointwise_conv2d(x, self.weight, style, self.bias, self.demodulate)


class ConvNeXtSynthesisLayer(nn.Module):
    """dw kxk conv (+bilinearly resized const noise) -> GroupNorm32 -> modulated 1x1 conv C->4C -> GELU -> 1x1 conv 4C->C -> gamma*y + x"""

    def __init__(self, channels, w_dim, kernel_size, layer_scale_init=1e-5, demodulate=True, block_index=0, legacy=False, ops=None):
        super().__init__()
        self.ops, self.legacy, self.channels, self.kernel_size = ops, legacy, channels, kernel_size
        self.affine_pw1 = StyleSplit(w_dim, channels, bias_init=1)
        self.dwconv = DepthwiseConv2d(channels, kernel_size, bias=True, ops=ops)      # its stock fallback is fp16-safe (see the class)
        nn.init.trunc_normal_(self.dwconv.weight, std=0.02)
        nn.init.constant_(self.dwconv.bias, 0)
        if legacy:
            resolution = 8 * 2 ** block_index
            self.register_buffer('noise_const', torch.randn([resolution, resolution]))
            self.noise_strength = nn.Parameter(torch.zeros([]))
        self.pwconv1 = ModulatedPointwiseConv2DLayer(channels, 4 * channels, demodulate, ops=ops)
        self.pwconv2 = Conv1x1(4 * channels, channels, kernel_size=1)
        nn.init.trunc_normal_(self.pwconv2.weight, std=0.02)
        nn.init.zeros_(self.pwconv2.bias)
        self.norm = GroupNorm32(min(32, channels // 4), channels)
        self.act = nn.GELU()
        self.gamma = nn.Parameter(layer_scale_init * torch.ones([1, channels, 1, 1])) if layer_scale_init > 0 else None

    def forward(self, x, w):
        dtype = x.dtype
        x_in = x
        style = self.affine_pw1(w)
        noise = None
        if self.legacy:
            noise = self.noise_const[None, None] * self.noise_strength
            noise = F.interpolate(noise, size=x.shape[2:], mode='bilinear', align_corners=False)
        dw = getattr(self.ops, 'depthwise_conv2d', None)
        y = None
        if dw is not None and x.is_cuda:
            # k x k depthwise conv + bias + noise on the streaming stencil kernel, one rounding of the sum to the activation dtype (the
            # reference promotes `dwconv(x) + noise` to fp32 and feeds GroupNorm32 / the autocast cast with it: same statistics up to
            # that rounding, a third of the bytes); autograd: data gradient on the same kernel + vfm_depthwise_wgrad
            xin = x.to(torch.get_autocast_dtype('cuda')) if torch.is_autocast_enabled() else x
            y = dw(xin, self.dwconv.weight, self.dwconv.bias, noise)
        noise_done = y is not None
        x = y if y is not None else self.dwconv(x)
        fused = getattr(self.ops, 'fused_convnext_mlp', None)
        if fused is not None and not torch.is_grad_enabled():
            # inference: one GroupNorm statistics pass + two tensor-core 1x1 convs with everything else in their epilogues
            xd = torch.add(x, noise.to(x.dtype)) if (self.legacy and not noise_done) else x
            y = fused(xd.to(dtype), x_in, self.norm.weight, self.norm.bias, self.norm.num_groups, self.norm.eps, self.pwconv1.weight,
                      self.pwconv1.bias, style, self.pwconv2.weight, self.pwconv2.bias, self.gamma, self.pwconv1.demodulate)
            if y is not None:
                return y
        if self.legacy and not noise_done:
            x = x + noise
        x = self.norm(x)
        x = self.pwconv1(x, style)
        x = self.act(x)
        x = self.pwconv2(x)
        if self.gamma is not None:
            x = self.gamma * x
        return (x + x_in).to(dtype)


class ConvNeXtToRGBLayer(nn.Module):
    """1x1 modulated conv without demodulation + broadcast bias   (convnext_utils.py:145-187)"""

    def __init__(self, in_channels, out_channels, w_dim, kernel_size=1, ops=None):
        super().__init__()
        self.ops, self.in_channels, self.out_channels, self.kernel_size = ops, in_channels, out_channels, kernel_size
        self.weight = nn.Parameter(torch.randn(out_channels, in_channels, kernel_size, kernel_size) * 0.1)
        self.bias = nn.Parameter(torch.zeros(1, out_channels, 1, 1))
        self.affine = StyleSplit(w_dim, in_channels, bias_init=1)
        self.weight_gain = 1 / np.sqrt(in_channels * kernel_size ** 2)

    def forward(self, x, w):
        style = self.affine(w) * self.weight_gain
        if torch.is_autocast_enabled() and x.is_cuda:
            x = x.to(torch.get_autocast_dtype('cuda'))
        y = self.ops.modulated_conv2d(x=x, weight=self.weight, styles=style, demodulate=False)
        return y + self.bias


class SynthesisBlock(nn.Module):
    """conv0 (up 2) + 2*num_res_blocks convs (plain, residual alternating, gain sqrt(1/2)) + optional self-attention +
    multi-scale ToRGB on the running feature sum   (generator.py:320-576, legacy branch)"""

    def __init__(self, block_index, in_channels, out_channels, last_out_channels, w_dim, resolution, img_channels, is_last,
                 num_res_blocks=1, use_multiscale_output=False, architecture='skip', resample_filter=(1, 3, 3, 1), conv_clamp=None,
                 use_fp16=False, attn_depth=0, attn_heads=8, attn_ff_mult=4, use_gaussian_blur=True, use_convnext=False,
                 add_additional_convnext=False, legacy=False, is_first=False, ops=None, **layer_kwargs):
        super().__init__()
        assert architecture in ('orig', 'skip')
        assert architecture == 'skip' or not use_multiscale_output
        if in_channels == 0:
            raise NotImplementedError('SynthesisInput blocks (in_channels == 0) are not used by the f16d32 configs')
        self.ops = ops
        self.in_channels, self.out_channels, self.last_out_channels = in_channels, out_channels, last_out_channels
        self.w_dim, self.resolution, self.img_channels, self.is_last = w_dim, resolution, img_channels, is_last
        self.architecture, self.use_fp16, self.use_multiscale_output = architecture, use_fp16, use_multiscale_output
        if not use_convnext:                       # the reference registers the buffer only for the legacy layers (generator.py:378-380)
            self.register_buffer('resample_filter', ops.setup_filter(list(resample_filter)))
        blur_kernel = '3x3' if block_index <= 2 else '5x5'
        self.use_convnext = use_convnext
        kernel_size = 5 if block_index <= 1 else 7
        convs = []
        if use_convnext:
            self.seperate_upsample_conv = SeparableUpsampleWithFixedBlur(in_channels, out_channels, upscale_factor=2, pre_normalize=not is_first,
                                                                         use_gaussian_blur=use_gaussian_blur, blur_kernel=blur_kernel, ops=ops)
            self.conv0 = ConvNeXtSynthesisLayer(out_channels, w_dim=w_dim, kernel_size=kernel_size, block_index=block_index, legacy=legacy, ops=ops)
            for _ in range(num_res_blocks):
                for _ in range(3 if block_index <= 3 and add_additional_convnext else 2):
                    convs.append(ConvNeXtSynthesisLayer(out_channels, w_dim=w_dim, kernel_size=kernel_size, block_index=block_index, legacy=legacy, ops=ops))
        else:
            self.conv0 = SynthesisLayer(in_channels, out_channels, w_dim=w_dim, resolution=resolution, up=2, resample_filter=resample_filter,
                                        conv_clamp=conv_clamp, ops=ops, **layer_kwargs)
            for _ in range(num_res_blocks):
                convs.append(SynthesisLayer(out_channels, out_channels, w_dim=w_dim, resolution=resolution, conv_clamp=conv_clamp, ops=ops, **layer_kwargs))
                convs.append(SynthesisLayer(out_channels, out_channels, w_dim=w_dim, resolution=resolution, conv_clamp=conv_clamp, residual=True, ops=ops, **layer_kwargs))
        self.convs1 = nn.ModuleList(convs)
        self.num_conv = 1 + len(convs)
        self.num_torgb = 0
        if is_last or architecture == 'skip':
            self.torgb = (ConvNeXtToRGBLayer(out_channels, img_channels, w_dim=w_dim, ops=ops) if use_convnext else
                          ToRGBLayer(out_channels, img_channels, w_dim=w_dim, conv_clamp=conv_clamp, ops=ops))
            self.num_torgb = 1
        if use_multiscale_output and last_out_channels is not None:
            self.last_upsample_conv = SeparableUpsampleWithFixedBlur(last_out_channels, out_channels, upscale_factor=2,
                                                                     use_gaussian_blur=use_gaussian_blur, blur_kernel=blur_kernel, ops=ops)
        self.self_attns = nn.ModuleList([SelfAttentionBlock(out_channels, dim_head=out_channels // attn_heads, heads=attn_heads, ff_mult=attn_ff_mult)
                                         for _ in range(attn_depth)]) if attn_depth > 0 else None

    def forward(self, x, x_sum, img, ws, force_fp32=False, fused_modconv=True, **layer_kwargs):
        w_iter = iter(ws.unbind(dim=1))
        on_cuda = ws.device.type == 'cuda'
        fp16 = on_cuda and self.use_fp16 and not force_fp32
        dtype = torch.float16 if fp16 else torch.float32
        amp = dict(device_type='cuda', enabled=fp16, dtype=torch.float16)
        x = x.to(dtype=dtype)
        if self.use_convnext:
            with torch.amp.autocast(**amp):
                x = self.seperate_upsample_conv(x)
                x = self.conv0(x, next(w_iter))
                for conv in self.convs1:
                    x = conv(x, next(w_iter))
        else:
            x = self.conv0(x, next(w_iter), fused_modconv=fused_modconv, **layer_kwargs)
            for conv in self.convs1:
                x = conv(x, next(w_iter), fused_modconv=fused_modconv, gain=np.sqrt(0.5), **layer_kwargs)
        if self.self_attns is not None:
            with torch.amp.autocast(**amp):
                for attn in self.self_attns:
                    x = attn(x)
        x = x.to(dtype=dtype)
        if self.use_multiscale_output:
            with torch.amp.autocast(**amp):
                x_sum = self.last_upsample_conv(x_sum) + x if self.last_out_channels is not None else x
                img = self.torgb(x_sum, next(w_iter))
            img = img.to(dtype=torch.float32)
        else:
            if img is not None:
                img = self.ops.upsample2d(img, self.resample_filter)
            if self.is_last or self.architecture == 'skip':
                y = self.torgb(x, next(w_iter)).to(dtype=torch.float32)
                img = img.add_(y) if img is not None else y
        assert x.dtype == dtype
        return x, x_sum, img


class SynthesisNetwork(nn.Module):
    """Legacy (use_convnext=False) pixel decoder: z [N,z_dim,zr,zr] + ws [N,num_ws,w_dim] -> img [N,3,R,R] fp32 and the
    lower-resolution multi-scale images   (generator.py:655-912)."""

    def __init__(self, w_dim, img_resolution, img_channels=3, channel_base=32768, channel_max=512, num_fp16_res=3, conv_clamp=None,
                 num_blocks=6, num_res_blocks=3, z_resolution=16, z_dim=8, concat_z_block_indices=(), concat_z_mapped_dims=(),
                 how_to_process_concat_z='unshuffle', activation_for_concat_z='gelu', use_multiscale_output=False,
                 attn_block_indices=(), attn_depths=(), use_self_attn=False, use_cross_attn=False, use_convnext=False,
                 use_gaussian_blur=True, c_dim=0, add_additional_convnext=False, legacy=False, ops=None, **block_kwargs):
        super().__init__()
        if use_cross_attn:
            raise NotImplementedError('cross-attention is unused by the f16d32 configs (conditional: False)')
        assert how_to_process_concat_z == 'unshuffle', 'only the unshuffle z-processing of the shipped configs is mirrored'
        ops = ops if ops is not None else default_ops()
        self.ops = ops
        self.c_dim, self.w_dim, self.img_resolution, self.img_channels = c_dim, w_dim, img_resolution, img_channels
        self.num_blocks, self.num_fp16_res = num_blocks, num_fp16_res
        self.z_resolution, self.z_dim = z_resolution, z_dim
        self.concat_z_block_indices = list(concat_z_block_indices)
        self.use_multiscale_output = use_multiscale_output
        res0 = img_resolution // (2 ** (num_blocks - 1))
        self.block_resolutions = [res0 * 2 ** i for i in range(num_blocks)]
        scale = img_resolution / 256
        channels = {i: min(channel_base // int(r / scale), channel_max) for i, r in enumerate(self.block_resolutions)}
        fp16_idx = num_blocks - num_fp16_res

        self.z_convs = nn.ModuleDict()
        z_dims = {}
        for idx in self.concat_z_block_indices:
            res = self.block_resolutions[idx]
            mapped = concat_z_mapped_dims[idx] if len(concat_z_mapped_dims) > 0 else None
            layers = []
            if res < z_resolution * 2:
                f = int(z_resolution / res * 2)
                cin = int(z_dim * f ** 2)
                out = mapped if mapped is not None else cin
                layers += [nn.PixelUnshuffle(f), self._conv3x3(cin, out, activation_for_concat_z), self._conv1x1(out, out)]
            elif res == z_resolution * 2:
                out = mapped if mapped is not None else z_dim
                layers += [self._conv3x3(z_dim, out, activation_for_concat_z), self._conv1x1(out, out)]
            else:
                f = int(res / z_resolution / 2)
                out = mapped if mapped is not None else z_dim
                layers += [self._conv3x3(z_dim, int(out * f ** 2), activation_for_concat_z), nn.PixelShuffle(f), self._conv1x1(out, out)]
            self.z_convs[str(idx)] = nn.Sequential(*layers)
            z_dims[idx] = out

        self.blocks = nn.ModuleDict()
        self.num_ws = 0
        for idx in range(num_blocks):
            cin = (channels[idx - 1] if idx > 0 else 0) + z_dims.get(idx, 0)
            depth = attn_depths[list(attn_block_indices).index(idx)] if (use_self_attn and idx in attn_block_indices) else 0
            block = SynthesisBlock(block_index=idx, in_channels=cin, out_channels=channels[idx],
                                   last_out_channels=channels[idx - 1] if idx > 0 else None, w_dim=w_dim,
                                   resolution=self.block_resolutions[idx], img_channels=img_channels, is_last=(idx == num_blocks - 1),
                                   use_fp16=(idx >= fp16_idx), conv_clamp=conv_clamp, num_res_blocks=num_res_blocks,
                                   use_multiscale_output=use_multiscale_output, use_gaussian_blur=use_gaussian_blur,
                                   use_convnext=use_convnext, add_additional_convnext=add_additional_convnext, legacy=legacy, is_first=(idx == 0),
                                   attn_depth=depth, ops=ops, **block_kwargs)
            self.num_ws += block.num_conv + block.num_torgb
            self.blocks[str(idx)] = block

    @staticmethod
    def _act(name):
        return {'lrelu': lambda: nn.LeakyReLU(negative_slope=0.2), 'silu': nn.SiLU, 'gelu': nn.GELU}[name]()

    def _conv3x3(self, cin, cout, activation):
        return nn.Sequential(DepthwiseConv2d(cin, 3, bias=False, ops=self.ops), Conv1x1(cin, cout, 1, bias=False),
                             GroupNorm32(min(32, cout), cout), self._act(activation))

    def _conv1x1(self, cin, cout):
        return nn.Sequential(Conv1x1(cin, cout, 1, bias=False), GroupNorm32(min(32, cout), cout))

    def decode_graph(self, z, ws, **block_kwargs):
        """Inference forward through a cached ``DecodeGraph`` (captured on first use per input shape / dtype / device).  Same
        result as ``forward`` under ``no_grad`` -- the same kernels on the same inputs -- returned in the graph's static buffers."""
        key = (tuple(z.shape), z.dtype, tuple(ws.shape), ws.dtype, z.device, tuple(sorted(block_kwargs.items())))
        cache = self.__dict__.setdefault('_decode_graphs', {})
        g = cache.get(key)
        if g is None:
            g = cache[key] = DecodeGraph(self, z, ws, **block_kwargs)
        return g(z, ws)

    def forward(self, z, ws, text=None, text_mask=None, **block_kwargs):
        ws = ws.to(torch.float32)
        x = x_sum = img = None
        multiscale = []
        w_idx = 0
        for idx in range(self.num_blocks):
            block = self.blocks[str(idx)]
            n_w = block.num_conv + block.num_torgb
            cur_ws = ws.narrow(1, w_idx, n_w)
            w_idx += n_w
            if idx in self.concat_z_block_indices:
                with torch.amp.autocast('cuda', enabled=bool(block.use_fp16 and z.device.type == 'cuda'), dtype=torch.float16):
                    zc = self.z_convs[str(idx)](z)
                    x = torch.cat([x, zc], dim=1) if x is not None else zc
            x, x_sum, img = block(x, x_sum, img, cur_ws, **block_kwargs)
            if not block.is_last:
                multiscale.append(img)
        return img, multiscale[::-1]


class DecodeGraph:
    """The whole inference forward of a ``SynthesisNetwork`` as ONE captured CUDA graph (``SynthesisNetwork.decode_graph``).

    A decode step is ~900 kernel launches, and the 8x8..32x32 blocks are launch-bound (their kernels finish faster than the
    host can issue them): replaying the captured step removes that host time (+3 % images/s at batch 64, 256x256).  The graph
    owns static input / output buffers: ``__call__`` copies z and ws in (stream-ordered, no sync), replays, and returns the
    static outputs -- valid until the next call, so copy them if they must outlive it.  The kernels read the parameters from
    their own storage at replay time, so in-place parameter updates are seen; re-capture after replacing a parameter tensor."""

    def __init__(self, net, z, ws, **block_kwargs):
        assert z.is_cuda and ws.is_cuda, 'decode_graph: CUDA tensors only'
        self.z, self.ws = z.clone(), ws.clone()
        self.key = (tuple(z.shape), z.dtype, tuple(ws.shape), ws.dtype, z.device)
        with torch.no_grad():
            side = torch.cuda.Stream(device=z.device)
            side.wait_stream(torch.cuda.current_stream(z.device))
            with torch.cuda.stream(side):                     # warm-up outside the capture: cuDNN autotuning, lazy plugin init
                for _ in range(2):
                    net(self.z, self.ws, **block_kwargs)
            torch.cuda.current_stream(z.device).wait_stream(side)
            self.graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(self.graph):
                self.img, self.multi = net(self.z, self.ws, **block_kwargs)

    def __call__(self, z, ws):
        self.z.copy_(z, non_blocking=True)
        self.ws.copy_(ws, non_blocking=True)
        self.graph.replay()
        return self.img, self.multi


#: SynthesisNetwork kwargs of the shipped f16d32 configs with use_convnext=False (configs/vfm_vae_f16d32_siglip2_stage_1_*.yaml:32-99)
F16D32_LEGACY_KWARGS = dict(
    w_dim=512, img_resolution=256, img_channels=3, z_resolution=16, z_dim=512,
    concat_z_block_indices=[0, 1, 2, 3], concat_z_mapped_dims=[512, 256, 128, 128], how_to_process_concat_z='unshuffle',
    activation_for_concat_z='lrelu', attn_block_indices=[0, 1, 2], attn_depths=[2, 2, 2], use_self_attn=True, use_cross_attn=False,
    use_convnext=False, use_multiscale_output=True, num_blocks=6, num_fp16_res=3, conv_clamp=256, channel_base=32768,
    channel_max=512, num_res_blocks=2, architecture='skip')

#: the same with the ConvNeXt layers of the shipped configs (use_convnext: True, add_additional_convnext: True, legacy noise)
F16D32_CONVNEXT_KWARGS = dict(F16D32_LEGACY_KWARGS, use_convnext=True, add_additional_convnext=True, legacy=True, use_gaussian_blur=True)
