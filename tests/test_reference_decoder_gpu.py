"""SURVEY.md 8 row a16 -- the reference's OWN decoder, unchanged, on these kernels on a GPU.

``oracle/_ref`` holds the unmodified reference packages (tools/stage_reference.py).  ``vfm_vae_b200.integration.install()``
swaps the plugin loader (torch_utils/custom_ops.py:59) and the ``modulated_conv2d`` helper (networks/generator.py:46); then
the reference's ``SynthesisNetwork`` -- ``SynthesisLayer.forward`` / ``ToRGBLayer.forward`` / ``SynthesisBlock.forward``
(networks/generator.py:240-276, 306-310, 483-576) and the reference's own ``bias_act`` / ``upfirdn2d`` Python wrappers with
their autograd Functions -- runs on CUDA with every plugin call landing in libvfmops.so, and is compared with the same
network's own CPU (impl='ref') output and parameter gradients.
"""
import contextlib
import io

import pytest
import torch

from conftest import rel_err

pytestmark = pytest.mark.gpu
DEV = 'cuda'


@pytest.fixture(scope='module')
def ref_installed():
    from oracle import reference as R
    if not R.available():
        pytest.skip('the reference is not staged under oracle/_ref (run tools/stage_reference.py in the build container)')
    gen = R.load()
    import vfm_vae_b200.integration as integ
    integ.install()
    torch.backends.cudnn.allow_tf32 = False          # as the reference's training loop (training/training_loop.py:504-505)
    torch.backends.cuda.matmul.allow_tf32 = False
    yield R, gen
    integ.uninstall()


def _build(gen, kw, seed):
    torch.manual_seed(seed)
    with contextlib.redirect_stdout(io.StringIO()):
        net = gen.SynthesisNetwork(**kw)
    g = torch.Generator().manual_seed(seed + 1)
    with torch.no_grad():
        for name, p in net.named_parameters():
            if name.endswith('noise_strength'):
                p.fill_(0.1)
            elif name.endswith('.bias') and p.ndim == 1 and 'affine' not in name and 'norm' not in name:
                p.copy_(torch.randn(p.shape, generator=g) * 0.1)
            elif name.endswith('gamma') and p.ndim == 4:
                p.fill_(0.3)
            elif name.endswith('to_out.weight') or (name.endswith('.3.weight') and '.ff.' in name):
                p.copy_(torch.randn(p.shape, generator=g) * 0.02)
    return net


def _loss(img, multi):
    return (img.square().mean() + sum(m.square().mean() for m in multi)) * 4096.0     # keeps d(loss)/d(img) out of the fp16 subnormals


def _all_fp32(net):
    """The tools' num_fp16_res=0 configuration on an already-built network (SynthesisBlock.use_fp16, generator.py:499-510; `force_fp32`
    alone would leave block 3's z-convs under fp16 autocast, generator.py:897)."""
    old = [b.use_fp16 for b in net.blocks.values()]
    for b in net.blocks.values():
        b.use_fp16 = False
    return old


def _restore_fp16(net, flags):
    for b, f in zip(net.blocks.values(), flags):
        b.use_fp16 = f


def _stock_fp16_depthwise_is_broken():
    from test_benchmark_config_gpu import _stock_fp16_depthwise_is_broken as probe
    return probe()


def _block_outputs(net, z, ws):
    """-> (img, [x after each SynthesisBlock]): the hot-path activations, without the x_sum / ToRGB branch that runs through the reference's
    stock pixel-shuffle upsampler."""
    xs, hooks = [], []
    for b in net.blocks.values():
        hooks.append(b.register_forward_hook(lambda mod, inp, out: xs.append(out[0].detach().float())))
    with torch.no_grad():
        img, _ = net(z, ws, None, None)
    for h in hooks:
        h.remove()
    return img, xs


def _launches():
    from vfm_vae_b200 import _lib
    return _lib.launch_count()


SMALL_KW = dict(c_dim=0, w_dim=64, img_resolution=64, img_channels=3, z_resolution=8, z_dim=16,
                concat_z_block_indices=[0, 1], concat_z_mapped_dims=[128, 128], how_to_process_concat_z='unshuffle',
                activation_for_concat_z='lrelu', attn_block_indices=[0], attn_depths=[1], use_self_attn=True,
                use_cross_attn=False, use_convnext=False, use_multiscale_output=True, num_blocks=4, num_fp16_res=2,
                conv_clamp=256, channel_base=32768, channel_max=128, num_res_blocks=2, architecture='skip')


def test_reference_decoder_small_runs_on_the_kernels(ref_installed):
    """128-channel reference decoder (tcgen05-eligible widths), forward + parameter gradients, fp32 and fp16 blocks."""
    R, gen = ref_installed
    net = _build(gen, SMALL_KW, 11)
    g = torch.Generator().manual_seed(12)
    z = torch.randn(2, 16, 8, 8, generator=g)
    ws = torch.randn(2, net.num_ws, 64, generator=g)
    img_r, multi_r = net(z, ws, None, None)                       # CPU tensors: the reference's own impl='ref' path
    names = [n for n, _ in net.named_parameters() if n.endswith(('conv0.weight', 'convs1.1.bias', 'torgb.weight', 'noise_strength', 'affine.proj.weight'))]
    params = dict(net.named_parameters())
    gr = [t.clone() for t in torch.autograd.grad(_loss(img_r, multi_r), [params[n] for n in names])]
    img_r, multi_r = img_r.detach(), [m.detach() for m in multi_r]
    net = net.to(DEV)
    l0 = _launches()
    params = dict(net.named_parameters())
    flags = _all_fp32(net)
    img, multi = net(z.to(DEV), ws.to(DEV), None, None)
    gg = torch.autograd.grad(_loss(img, multi), [params[n] for n in names])
    _restore_fp16(net, flags)
    assert _launches() - l0 > 100, 'the reference decoder must have launched libvfmops kernels'
    assert rel_err(img, img_r) <= 2e-5
    for a, b in zip(multi, multi_r):
        assert rel_err(a, b) <= 2e-5
    for n, a, b in zip(names, gg, gr):
        if n.endswith('noise_strength'):
            continue          # a scalar: the sum of dy * noise_const over every pixel, i.e. pure cancellation (|sum| << sum |.|)
        assert rel_err(a, b) <= 2e-3, n                              # network-level fp32 gradients: see tests/test_decoder.py on lrelu sign flips
    # blocks 2-3 in fp16, as the reference does on CUDA: the hot-path activations block by block against the all-fp32 run
    flags = _all_fp32(net)
    _, xs32 = _block_outputs(net, z.to(DEV), ws.to(DEV))
    _restore_fp16(net, flags)
    img16, xs16 = _block_outputs(net, z.to(DEV), ws.to(DEV))
    assert img16.dtype == torch.float32
    for a, b in zip(xs16, xs32):
        assert rel_err(a, b) <= 3e-3
    if not _stock_fp16_depthwise_is_broken():        # the image also passes through the reference's STOCK fp16 upsampler (depthwise conv)
        assert rel_err(img16, img_r) <= 3e-3


def test_reference_decoder_f16d32_runs_on_the_kernels(ref_installed, capsys):
    """The benchmarked network: the reference's SynthesisNetwork(**f16d32 kwargs, use_convnext=False), N=2, on CUDA vs on CPU."""
    R, gen = ref_installed
    net = _build(gen, R.F16D32_LEGACY_KWARGS, 13)
    g = torch.Generator().manual_seed(14)
    z = torch.randn(2, 512, 16, 16, generator=g)
    ws = torch.randn(2, net.num_ws, 512, generator=g)
    names = ['blocks.5.convs1.3.weight', 'blocks.4.conv0.weight', 'blocks.3.convs1.0.weight', 'blocks.1.conv0.weight', 'blocks.5.convs1.2.bias',
             'blocks.5.torgb.weight', 'blocks.2.convs1.1.gamma']
    img_r, multi_r = net(z, ws, None, None)
    params = dict(net.named_parameters())
    gr = [t.clone() for t in torch.autograd.grad(_loss(img_r, multi_r), [params[n] for n in names])]
    img_r, multi_r = img_r.detach(), [m.detach() for m in multi_r]
    net = net.to(DEV)
    params = dict(net.named_parameters())
    from vfm_vae_b200 import _lib
    lib = _lib.load()
    lib.vfm_timing_enable(1)
    flags = _all_fp32(net)
    img, multi = net(z.to(DEV), ws.to(DEV), None, None)
    gg = torch.autograd.grad(_loss(img, multi), [params[n] for n in names])
    _restore_fp16(net, flags)
    torch.cuda.synchronize()
    lib.vfm_timing_enable(0)
    buf = (_lib.KernelStat * 512)()
    kernels = {buf[i].name.decode() for i in range(min(lib.vfm_timing_report(buf, 512), 512))}
    assert any(k.startswith('modconv_tc_fwd') for k in kernels) and any(k.startswith('modconv_tc_wgrad') for k in kernels), kernels
    assert any(k.startswith('bias_act') for k in kernels) and any(k.startswith('upfirdn2d') for k in kernels), kernels
    e_img = rel_err(img, img_r)
    e_multi = [rel_err(a, b) for a, b in zip(multi, multi_r)]
    e_grads = {n: rel_err(a, b) for n, a, b in zip(names, gg, gr)}
    # num_fp16_res=3: blocks 3-5 in fp16.  Hot-path activations block by block (fp16 run vs all-fp32 run on the GPU); the final image only
    # where the reference's STOCK fp16 upsampler works (its fp16 depthwise conv is broken on B200 / torch 2.11 / cuDNN 9.22: tools/stock_fp16_probe.py)
    flags = _all_fp32(net)
    _, xs32 = _block_outputs(net, z.to(DEV), ws.to(DEV))
    _restore_fp16(net, flags)
    img16, xs16 = _block_outputs(net, z.to(DEV), ws.to(DEV))
    e_blocks = [rel_err(a, b) for a, b in zip(xs16, xs32)]
    e16 = rel_err(img16, img_r)
    broken = _stock_fp16_depthwise_is_broken()
    with capsys.disabled():
        print(f'\n[a16 reference decoder f16d32 on libvfmops] fp32 img {e_img:.3g} multi {max(e_multi):.3g} grads {e_grads}  fp16: block outputs {e_blocks} '
              f'img {e16:.3g} (stock fp16 depthwise conv broken on this box: {broken})')
    assert e_img <= 1e-4 and max(e_multi) <= 1e-4
    assert max(e_grads.values()) <= 2e-3
    assert max(e_blocks) <= 5e-3, e_blocks
    if not broken:
        assert e16 <= 5e-3


def test_reference_convnext_decoder_runs_on_the_kernels(ref_installed):
    """The shipped-config variant (use_convnext=True): its modulated_pointwise_conv2d goes through the tcgen05 kernel (k = 1)."""
    R, gen = ref_installed
    kw = dict(SMALL_KW, use_convnext=True, add_additional_convnext=True, legacy=True, use_gaussian_blur=True)
    net = _build(gen, kw, 15)
    g = torch.Generator().manual_seed(16)
    z = torch.randn(2, 16, 8, 8, generator=g)
    ws = torch.randn(2, net.num_ws, 64, generator=g)
    with torch.no_grad():
        img_r, _ = net(z, ws, None, None)
        net = net.to(DEV)
        l0 = _launches()
        _all_fp32(net)
        img, _ = net(z.to(DEV), ws.to(DEV), None, None)
        assert _launches() - l0 > 10
    assert rel_err(img, img_r) <= 2e-5
