"""TEST / MEASUREMENT INFRASTRUCTURE ONLY -- loader for the UNMODIFIED reference staged under oracle/_ref/.

``tools/stage_reference.py`` (run by ``__graft_entry__.build()`` in the build container) copies the reference's
``networks/``, ``torch_utils/`` and ``dnnlib/`` packages byte for byte into the git-ignored ``oracle/_ref/``; this module
puts that directory on ``sys.path`` and hands back the reference's own modules.  Only ``tests/``, ``smoke()`` and the CPU /
reference legs of ``bench.py`` may import it; the product (``vfm_vae_b200/``) never does.

What callers get from the real reference:
  * ``networks.generator.SynthesisNetwork`` (reference networks/generator.py:655-912) -- the decoder whose impl='ref'
    CPU path is the cpu_baseline / ``--impl reference`` arm, and which runs unchanged on the sm_100a kernels after
    ``vfm_vae_b200.integration.install()``;
  * ``torch_utils.ops.{bias_act,upfirdn2d,filtered_lrelu,conv2d_resample}`` -- the reference's own ``_*_ref`` functions.
"""
import importlib
import json
import os
import sys
import warnings

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.join(HERE, '_ref')

#: reference SynthesisNetwork kwargs of the shipped f16d32 configs (configs/vfm_vae_f16d32_siglip2_stage_1_*.yaml:32-99) with
#: use_convnext=False (D-legacy, the variant north_star describes); c_dim=0 because the configs are unconditional (generator.py:1034-1037)
F16D32_LEGACY_KWARGS = dict(
    c_dim=0, w_dim=512, img_resolution=256, img_channels=3, z_resolution=16, z_dim=512,
    concat_z_block_indices=[0, 1, 2, 3], concat_z_mapped_dims=[512, 256, 128, 128], how_to_process_concat_z='unshuffle',
    activation_for_concat_z='lrelu', attn_block_indices=[0, 1, 2], attn_depths=[2, 2, 2], use_self_attn=True, use_cross_attn=False,
    use_convnext=False, use_multiscale_output=True, num_blocks=6, num_fp16_res=3, conv_clamp=256, channel_base=32768,
    channel_max=512, num_res_blocks=2, architecture='skip')
F16D32_CONVNEXT_KWARGS = dict(F16D32_LEGACY_KWARGS, use_convnext=True, add_additional_convnext=True, legacy=True, use_gaussian_blur=True)


def available():
    return os.path.isfile(os.path.join(ROOT, 'MANIFEST.json')) and os.path.isdir(os.path.join(ROOT, 'networks'))


def manifest():
    return json.load(open(os.path.join(ROOT, 'MANIFEST.json')))


def load():
    """Put the staged reference on sys.path (once) and return its ``networks.generator`` module."""
    if not available():
        raise RuntimeError('the reference is not staged: run `python tools/stage_reference.py` in the build container '
                           '(oracle/_ref/ is git-ignored and travels to the GPU box with the snapshot)')
    for p in (ROOT, os.path.join(ROOT, '_shims')):
        if p not in sys.path:
            sys.path.insert(0, p)
    with warnings.catch_warnings():
        warnings.simplefilter('ignore')
        gen = importlib.import_module('networks.generator')
    assert os.path.abspath(gen.__file__).startswith(ROOT), f'networks.generator was imported from {gen.__file__}, not the staged reference'
    return gen


def ops():
    """The reference's own op modules (torch_utils.ops.*) from the staged copy."""
    load()
    from types import SimpleNamespace
    names = ('bias_act', 'upfirdn2d', 'filtered_lrelu', 'conv2d_resample', 'fma')
    return SimpleNamespace(**{n: importlib.import_module(f'torch_utils.ops.{n}') for n in names})
