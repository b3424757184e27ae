"""CPU-side checks of the drop-in boundary: the C-ABI library builds/loads, exports every symbol include/vfm_ops.h
declares, its structs have the layout the ctypes binding assumes, and the product refuses to run without CUDA."""
import ctypes
import os
import re
import subprocess
import sys

import pytest
import torch

from conftest import REPO

HEADER = os.path.join(REPO, 'include', 'vfm_ops.h')


def _declared_symbols():
    src = open(HEADER).read()
    return sorted(set(re.findall(r'VFM_API\s+[\w\s\*]+?\b(vfm_\w+)\s*\(', src)))


def test_library_exports_every_declared_symbol():
    from vfm_vae_b200 import _lib
    lib = _lib.load()
    declared = _declared_symbols()
    assert len(declared) >= 11
    assert sorted(_lib.SYMBOLS) == declared
    for name in declared:
        assert hasattr(lib, name), name
    assert lib.vfm_abi_version() == 10
    assert _lib.last_error() == ''


def test_struct_layout_matches_header(tmp_path):
    from vfm_vae_b200 import _lib
    structs = {
        'vfm_bias_act_params': _lib.BiasActParams, 'vfm_upfirdn2d_params': _lib.Upfirdn2dParams,
        'vfm_filtered_lrelu_params': _lib.FilteredLreluParams, 'vfm_filtered_lrelu_act_params': _lib.FilteredLreluActParams,
        'vfm_modconv_desc': _lib.ModconvDesc, 'vfm_modconv_fwd_params': _lib.ModconvFwdParams,
        'vfm_modconv_bwd_params': _lib.ModconvBwdParams, 'vfm_group_norm_affine_params': _lib.GroupNormAffineParams, 'vfm_group_norm_params': _lib.GroupNormParams, 'vfm_rows_params': _lib.RowsParams,
        'vfm_pixel_shuffle2_params': _lib.PixelShuffle2Params, 'vfm_replicate_blur_edges_params': _lib.ReplicateBlurEdgesParams, 'vfm_depthwise_wgrad_params': _lib.DepthwiseWgradParams,
        'vfm_grad_finalize_params': _lib.GradFinalizeParams, 'vfm_image_to_u8_params': _lib.ImageToU8Params,
    }
    lines = ['#include <stdio.h>', '#include <stddef.h>', '#include "vfm_ops.h"', 'int main(void){']
    for cname, st in structs.items():
        lines.append(f'printf("{cname} %zu\\n", sizeof({cname}));')
        for fname, _ in st._fields_:
            if fname == 'd':
                continue
            lines.append(f'printf("{cname}.{fname} %zu\\n", offsetof({cname}, {fname}));')
    lines.append('return 0;}')
    src = tmp_path / 'layout.c'
    src.write_text('\n'.join(lines))
    exe = tmp_path / 'layout'
    subprocess.check_call(['gcc', '-I', os.path.join(REPO, 'include'), str(src), '-o', str(exe)])
    out = dict(l.split() for l in subprocess.check_output([str(exe)]).decode().splitlines())
    for cname, st in structs.items():
        assert int(out[cname]) == ctypes.sizeof(st), cname
        for fname, _ in st._fields_:
            if fname == 'd':
                continue
            assert int(out[f'{cname}.{fname}']) == getattr(st, fname).offset, f'{cname}.{fname}'


def test_argument_errors_are_reported_without_a_gpu():
    """NULL / invalid parameter blocks are rejected before any CUDA call (no compute without a GPU)."""
    from vfm_vae_b200 import _lib
    lib = _lib.load()
    assert lib.vfm_bias_act(None, None) == _lib.VFM_ERR_INVALID
    assert 'NULL' in _lib.last_error()
    p = _lib.Upfirdn2dParams()
    assert lib.vfm_upfirdn2d(ctypes.byref(p), None) == _lib.VFM_ERR_INVALID
    q = _lib.FilteredLreluParams()
    assert lib.vfm_filtered_lrelu(ctypes.byref(q), None) == _lib.VFM_ERR_INVALID
    d = _lib.ModconvDesc()
    d.dtype, d.batch, d.in_channels, d.out_channels, d.in_h, d.in_w, d.kh, d.kw, d.up = 1, 1, 4, 4, 8, 8, 3, 3, 3
    fp = _lib.ModconvFwdParams()
    fp.d = d
    assert lib.vfm_modconv_forward(ctypes.byref(fp), None) == _lib.VFM_ERR_INVALID
    assert 'up must be 1 or 2' in _lib.last_error()


def test_product_refuses_cpu_tensors():
    import vfm_vae_b200 as V
    x = torch.randn(1, 2, 4, 4)
    with pytest.raises(RuntimeError, match='CUDA'):
        V.bias_act.bias_act(x, None, act='lrelu')
    with pytest.raises(RuntimeError, match='CUDA'):
        V.upfirdn2d.upfirdn2d(x, None)
    with pytest.raises(RuntimeError, match='CUDA'):
        V.filtered_lrelu.filtered_lrelu(x)
    with pytest.raises(RuntimeError, match='CUDA'):
        V.modulated_conv2d(x, torch.randn(2, 2, 3, 3), torch.ones(1, 2), padding=1)


def test_product_never_imports_the_oracle():
    """No module under vfm_vae_b200/ may import, call, link or execute anything under oracle/ (static scan + runtime)."""
    pkg = os.path.join(REPO, 'vfm_vae_b200')
    for root, _, files in os.walk(pkg):
        for f in files:
            if f.endswith(('.py', '.cu', '.cuh', '.h', '.cpp')):
                src = open(os.path.join(root, f)).read()
                assert not re.search(r'^\s*(from|import)\s+oracle\b', src, re.M), os.path.join(root, f)
    code = 'import sys; import vfm_vae_b200, vfm_vae_b200.decoder, vfm_vae_b200.sync; assert not any(m == "oracle" or m.startswith("oracle.") for m in sys.modules)'
    subprocess.check_call([sys.executable, '-c', code], cwd=REPO)


def test_setup_filter_matches_golden():
    from conftest import golden
    from vfm_vae_b200.torch_utils.ops import upfirdn2d as U
    G = golden('upfirdn2d')
    assert torch.equal(U.setup_filter([1, 3, 3, 1]), G.t('sf_a'))
    assert torch.allclose(U.setup_filter([1, 2, 1], gain=4), G.t('sf_b'), rtol=1e-6)
    assert torch.equal(U.setup_filter(list(range(1, 13))), G.t('sf_c'))
    assert torch.equal(U.setup_filter([1, 3, 3, 1], flip_filter=True, normalize=False), G.t('sf_d'))
