"""2-D convolution with optional 2x upsampling.  Mirror of the reference entry point
torch_utils/ops/conv2d_resample.py:46 (same signature).  The reference composes cuDNN conv / conv_transpose with
upfirdn2d; here the supported cases run on the same sm_100a implicit-GEMM kernel as ``modulated_conv2d`` (an
unmodulated conv is the modulated one with unit styles and demodulation off), so the up=2 semantics -- transposed conv
then [1,3,3,1] blur with the padding arithmetic of reference lines 82-126 -- live in exactly one place (csrc/modconv_api.cu).

Supported: groups == 1, down == 1, up in {1, 2}, symmetric integer padding.  Everything else raises."""
import torch

from .modulated_conv2d import modulated_conv2d
from .upfirdn2d import _parse_padding


def conv2d_resample(x, w, f=None, up=1, down=1, padding=0, groups=1, flip_weight=True, flip_filter=False):
    assert isinstance(x, torch.Tensor) and x.ndim == 4
    assert isinstance(w, torch.Tensor) and w.ndim == 4
    assert f is None or (isinstance(f, torch.Tensor) and f.ndim in [1, 2] and f.dtype == torch.float32)
    px0, px1, py0, py1 = _parse_padding(padding)
    if groups != 1 or down != 1 or up not in (1, 2) or not (px0 == px1 == py0 == py1) or flip_filter:
        raise NotImplementedError('vfm_vae_b200.conv2d_resample: only groups=1, down=1, up in {1,2}, symmetric padding, '
                                  'flip_filter=False are implemented (the decoder path uses nothing else)')
    if up > 1 and f is not None and f.ndim == 1:
        f = f.ger(f)
    styles = torch.ones([x.shape[0], x.shape[1]], dtype=torch.float32, device=x.device)
    return modulated_conv2d(x, w, styles, noise=None, up=up, padding=px0, resample_filter=f, demodulate=False, flip_weight=flip_weight)
