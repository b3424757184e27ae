"""install()/uninstall() against the real, unmodified reference checkout (build container only: the GPU box has no
/root/reference, so this test skips there)."""
import os
import sys

import pytest
import torch

from conftest import REPO

REF = os.environ.get('VFM_REFERENCE', '/root/reference')
pytestmark = pytest.mark.skipif(not os.path.isdir(os.path.join(REF, 'torch_utils')), reason='reference checkout not present')


@pytest.fixture()
def ref_on_path():
    added = [os.path.join(REPO, 'tools', 'ref_shims'), REF]
    for p in added:
        sys.path.insert(0, p)
    yield
    for p in added:
        sys.path.remove(p)


def test_install_patches_loader_and_modconv(ref_on_path):
    import warnings
    warnings.filterwarnings('ignore')
    import torch_utils.custom_ops as ref_custom_ops
    import networks.generator as ref_gen
    from torch_utils.ops import bias_act as ref_bias_act
    import vfm_vae_b200.integration as integ
    from vfm_vae_b200.plugins import bias_act_plugin, upfirdn2d_plugin, filtered_lrelu_plugin

    orig_loader, orig_modconv = ref_custom_ops.get_plugin, ref_gen.modulated_conv2d
    integ.install()
    try:
        # the reference wrappers ask for plugins with their own keyword arguments; they now get ours
        got = ref_custom_ops.get_plugin(module_name='bias_act_plugin', sources=['bias_act.cpp', 'bias_act.cu'], headers=['bias_act.h'],
                                        source_dir='.', extra_cuda_cflags=['--use_fast_math'])
        assert got is bias_act_plugin
        assert ref_custom_ops.get_plugin(module_name='upfirdn2d_plugin', sources=[]) is upfirdn2d_plugin
        assert ref_custom_ops.get_plugin(module_name='filtered_lrelu_plugin', sources=[]) is filtered_lrelu_plugin
        ref_bias_act._init()
        assert ref_bias_act._plugin is bias_act_plugin
        # CPU tensors still take the reference's own path, bit for bit
        x = torch.randn(2, 4, 6, 6)
        w = torch.randn(3, 4, 3, 3)
        s = torch.randn(2, 4) + 1
        assert torch.equal(ref_gen.modulated_conv2d(x, w, s, padding=1), orig_modconv(x, w, s, padding=1))
        assert torch.equal(ref_bias_act.bias_act(x, torch.zeros(4), act='lrelu'), torch.nn.functional.leaky_relu(x, 0.2) * (2 ** 0.5))
        assert ref_gen.modulated_conv2d is not orig_modconv
        import networks.utils.convnext_utils as ref_cnx
        wp, bp = torch.randn(8, 4, 1, 1), torch.randn(1, 8, 1, 1)
        assert hasattr(ref_cnx.modulated_pointwise_conv2d, '__wrapped__')
        assert torch.equal(ref_cnx.modulated_pointwise_conv2d(x, wp, s, bp), ref_cnx.modulated_pointwise_conv2d.__wrapped__(x, wp, s, bp))
    finally:
        integ.uninstall()
    assert ref_custom_ops.get_plugin is orig_loader and ref_gen.modulated_conv2d is orig_modconv
    import networks.utils.convnext_utils as ref_cnx
    assert not hasattr(ref_cnx.modulated_pointwise_conv2d, '__wrapped__')


def test_plugin_signatures_match_reference_pybind():
    """Positional arity of each plugin entry point equals the reference's pybind function (plus optional extensions)."""
    import inspect
    from vfm_vae_b200.plugins import bias_act_plugin, upfirdn2d_plugin, filtered_lrelu_plugin
    def required(fn):
        return [p.name for p in inspect.signature(fn).parameters.values() if p.default is inspect.Parameter.empty]
    assert required(bias_act_plugin.bias_act) == ['x', 'b', 'xref', 'yref', 'dy', 'grad', 'dim', 'act', 'alpha', 'gain', 'clamp']
    assert required(upfirdn2d_plugin.upfirdn2d) == ['x', 'f', 'upx', 'upy', 'downx', 'downy', 'padx0', 'padx1', 'pady0', 'pady1', 'flip', 'gain']
    assert required(filtered_lrelu_plugin.filtered_lrelu) == ['x', 'fu', 'fd', 'b', 'si', 'up', 'down', 'px0', 'px1', 'py0', 'py1', 'sx', 'sy',
                                                              'gain', 'slope', 'clamp', 'flip_filters', 'writeSigns']
    assert required(filtered_lrelu_plugin.filtered_lrelu_act_) == ['x', 'si', 'sx', 'sy', 'gain', 'slope', 'clamp', 'writeSigns']
