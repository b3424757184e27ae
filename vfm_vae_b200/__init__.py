"""vfm_vae_b200 -- B200 (sm_100a) kernels for the VFM-VAE pixel-decoder hot path.

Layout (only what the path needs):
  csrc/            CUDA kernels + the C ABI (include/vfm_ops.h)  -> lib/libvfmops.so
  _lib.py          ctypes binding of the C ABI (fails loudly; no CPU fallback)
  plugins.py       plugin objects with the reference's pybind call signatures
  custom_ops.py    get_plugin() mirror of torch_utils/custom_ops.py:59
  torch_utils/ops/ host-side mirror of the reference op wrappers (same names, signatures, autograd structure)
  decoder.py       host-side mirror of the reference SynthesisNetwork (legacy and ConvNeXt variants) that calls the ops
  sync.py          gradient exchange (reference sync_grads semantics) for the batch-sharded multi-GPU step
  integration.py   drop the kernels into an unmodified reference checkout
"""
from .torch_utils.ops import bias_act, upfirdn2d, filtered_lrelu, conv2d_resample, fma  # noqa: F401
from .torch_utils.ops.modulated_conv2d import modulated_conv2d, modulated_pointwise_conv2d  # noqa: F401

__all__ = ['bias_act', 'upfirdn2d', 'filtered_lrelu', 'conv2d_resample', 'fma', 'modulated_conv2d', 'modulated_pointwise_conv2d']
