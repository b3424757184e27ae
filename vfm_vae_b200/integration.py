"""Drop the sm_100a kernels into an UNMODIFIED checkout of the reference (tianciB/VFM-VAE).

    import sys; sys.path.insert(0, '/path/to/VFM-VAE')
    import vfm_vae_b200.integration as integ
    integ.install()            # before the first CUDA call of any op
    from networks.generator import Generator   # ... the reference decoder, losses and training loop run unchanged

What it patches (and nothing else):

1. ``torch_utils.custom_ops.get_plugin`` (reference torch_utils/custom_ops.py:59) -> ``vfm_vae_b200.custom_ops.get_plugin``.
   The reference op wrappers keep calling ``_plugin.bias_act(...)``, ``_plugin.upfirdn2d(...)``,
   ``_plugin.filtered_lrelu(...)`` / ``.filtered_lrelu_act_(...)`` with their original positional signatures; those calls
   now land in libvfmops.so through the C ABI instead of a JIT-compiled pybind module.  The wrappers' own autograd
   Functions, caches and ``impl='ref'`` CPU path are untouched.
2. ``networks.generator.modulated_conv2d`` (reference networks/generator.py:46) -> ``vfm_vae_b200.modulated_conv2d``
   for CUDA tensors; CPU tensors keep going to the reference function (its pure-PyTorch body *is* its reference path).
   ``SynthesisLayer``/``ToRGBLayer`` look the name up in module globals at call time, so they need no change.

3. ``networks.utils.convnext_utils.modulated_pointwise_conv2d`` (reference convnext_utils.py:36, the 1x1 modulated conv of the
   ConvNeXt layers every shipped config runs) -> ``vfm_vae_b200.modulated_pointwise_conv2d`` for CUDA tensors, same rule.

``uninstall()`` restores all three.  Nothing here imports ``oracle/``.
"""
import importlib

from . import custom_ops as _our_custom_ops

_saved = {}


def install(patch_modconv=True):
    ref_custom_ops = importlib.import_module('torch_utils.custom_ops')
    if 'get_plugin' not in _saved:
        _saved['get_plugin'] = ref_custom_ops.get_plugin
    ref_custom_ops.get_plugin = _our_custom_ops.get_plugin
    # plugins already loaded by the reference wrappers (module-level ``_plugin`` caches) are dropped so that the next
    # ``_init()`` goes through the patched loader
    for name in ('bias_act', 'upfirdn2d', 'filtered_lrelu'):
        mod = importlib.import_module(f'torch_utils.ops.{name}')
        mod._plugin = None
    if patch_modconv:
        gen = importlib.import_module('networks.generator')
        from .torch_utils.ops.modulated_conv2d import modulated_conv2d as ours
        if 'modulated_conv2d' not in _saved:
            _saved['modulated_conv2d'] = gen.modulated_conv2d
        ref_fn = _saved['modulated_conv2d']

        def modulated_conv2d(x, weight, styles, *args, **kwargs):
            if x.device.type == 'cuda':
                return ours(x, weight, styles, *args, **kwargs)
            return ref_fn(x, weight, styles, *args, **kwargs)

        modulated_conv2d.__wrapped__ = ref_fn
        gen.modulated_conv2d = modulated_conv2d

        cnx = importlib.import_module('networks.utils.convnext_utils')
        from .torch_utils.ops.modulated_conv2d import modulated_pointwise_conv2d as ours_pw
        if 'modulated_pointwise_conv2d' not in _saved:
            _saved['modulated_pointwise_conv2d'] = cnx.modulated_pointwise_conv2d
        ref_pw = _saved['modulated_pointwise_conv2d']

        def modulated_pointwise_conv2d(x, weight, style, bias=None, demodulate=True):
            if x.device.type == 'cuda':
                return ours_pw(x, weight, style, bias, demodulate)
            return ref_pw(x, weight, style, bias, demodulate)

        modulated_pointwise_conv2d.__wrapped__ = ref_pw
        cnx.modulated_pointwise_conv2d = modulated_pointwise_conv2d
    return True


def uninstall():
    if 'get_plugin' in _saved:
        importlib.import_module('torch_utils.custom_ops').get_plugin = _saved.pop('get_plugin')
        for name in ('bias_act', 'upfirdn2d', 'filtered_lrelu'):
            importlib.import_module(f'torch_utils.ops.{name}')._plugin = None
    if 'modulated_conv2d' in _saved:
        importlib.import_module('networks.generator').modulated_conv2d = _saved.pop('modulated_conv2d')
    if 'modulated_pointwise_conv2d' in _saved:
        importlib.import_module('networks.utils.convnext_utils').modulated_pointwise_conv2d = _saved.pop('modulated_pointwise_conv2d')
