"""Decode I/O path (vfm_vae_b200/decode_io.py; reference tools/decode/decode_latents_to_images.py:20-100): the uint8 conversion kernel is
bit-identical to the reference's ``((img + 1) / 2).clamp(0, 1)`` -> ``to_pil_image`` arithmetic, and the pipelined tool writes the same
files (names, rank split, pixels) as the reference's serial loop."""
import os

import numpy as np
import pytest
import torch

from vfm_vae_b200 import decode_io as D


def _reference_bytes(images):
    """What the reference writes into the PNG: ((x + 1) / 2).clamp(0, 1), then torchvision's to_pil_image float path = mul(255).byte(), HWC."""
    t = ((images.float() + 1) / 2).clamp(0, 1)
    return t.mul(255).byte().permute(0, 2, 3, 1).contiguous()


def test_file_split_matches_reference(tmp_path):
    for i in (3, 1, 2, 0, 4):
        (tmp_path / f'part_{i:02d}.safetensors').write_bytes(b'')
    (tmp_path / 'notes.txt').write_text('x')
    a, total = D.list_latent_files(str(tmp_path), 0, 2)
    b, _ = D.list_latent_files(str(tmp_path), 1, 2)
    assert total == 5 and a == ['part_00.safetensors', 'part_02.safetensors', 'part_04.safetensors'] and b == ['part_01.safetensors', 'part_03.safetensors']
    empty = tmp_path / 'empty'
    empty.mkdir()
    with pytest.raises(AssertionError):          # the reference asserts that the directory holds latent files
        D.list_latent_files(str(empty), 0, 1)


def test_images_to_uint8_refuses_cpu_tensors():
    with pytest.raises(RuntimeError, match='CUDA'):
        D.images_to_uint8(torch.zeros(1, 3, 4, 4))


@pytest.mark.gpu
@pytest.mark.parametrize('dtype', [torch.float32, torch.float16], ids=['fp32', 'fp16'])
@pytest.mark.parametrize('shape', [(2, 3, 64, 64), (1, 3, 256, 256), (3, 1, 8, 8), (2, 4, 16, 20), (2, 3, 5, 7), (1, 3, 3, 3)])
def test_images_to_uint8_bit_exact(shape, dtype):
    g = torch.Generator().manual_seed(0)
    x = (torch.rand(shape, generator=g) * 3 - 1.5).to(dtype)
    # exact bin boundaries and the clamp edges: x = 2 k / 255 - 1 and its fp32 neighbours
    k = torch.arange(0, 256, dtype=torch.float32)
    edge = (2 * k / 255 - 1)
    edge = torch.cat([edge, torch.nextafter(edge, torch.tensor(2.0)), torch.nextafter(edge, torch.tensor(-2.0)), torch.tensor([-1.0, 1.0, 0.0, -0.0, 5.0, -5.0])])
    flat = x.flatten()
    m = min(flat.numel(), edge.numel())
    flat[:m] = edge[:m].to(dtype)
    x = flat.reshape(shape).cuda()
    got = D.images_to_uint8(x)
    want = _reference_bytes(x)
    assert got.dtype == torch.uint8 and got.shape == want.shape
    assert torch.equal(got, want)
    # an unaligned view (element offset 1) takes the scalar path
    buf = torch.empty(x.numel() + 1, dtype=dtype, device='cuda')
    xv = buf[1:].view(shape)
    xv.copy_(x)
    assert torch.equal(D.images_to_uint8(xv), want)


@pytest.mark.gpu
def test_decode_latents_to_images_matches_reference_loop(tmp_path):
    from PIL import Image
    from safetensors.torch import load_file, save_file
    g = torch.Generator().manual_seed(1)
    in_dir, out_dir = tmp_path / 'latents', tmp_path / 'png'
    in_dir.mkdir()
    sizes = [5, 3, 4]
    for i, n in enumerate(sizes):
        save_file({'latents': torch.randn(n, 4, 8, 8, generator=g), 'labels': torch.zeros(n)}, str(in_dir / f'shard_{i}.safetensors'))
    save_file({'other': torch.zeros(1)}, str(in_dir / 'shard_9.safetensors'))          # no 'latents': skipped, like the reference
    w = torch.randn(3, 4, 1, 1, generator=g).cuda()

    def decode_fn(latents, labels):
        return torch.nn.functional.interpolate(torch.nn.functional.conv2d(latents, w), scale_factor=4, mode='nearest')

    for rank, world in ((0, 1), (1, 2)):
        out = out_dir / f'r{rank}w{world}'
        saved = D.decode_latents_to_images(decode_fn, str(in_dir), str(out), batch_size_per_gpu=2, rank=rank, world_size=world, workers=4)
        files, _ = D.list_latent_files(str(in_dir), rank, world)
        # the reference's serial loop, restated
        want, idx = {}, 0
        for f in files:
            data = load_file(str(in_dir / f))
            if 'latents' not in data:
                continue
            lat = data['latents'].cuda()
            for s in range(0, lat.size(0), 2):
                img = _reference_bytes(decode_fn(lat[s:s + 2], None)).cpu().numpy()
                for i in range(img.shape[0]):
                    want[f'rank{rank:02d}_{idx + i:06d}.png'] = img[i]
                idx += img.shape[0]
        assert saved == len(want)
        assert sorted(os.listdir(out)) == sorted(want)
        for name, px in want.items():
            assert np.array_equal(np.asarray(Image.open(out / name)), px), name
    # max_images_per_gpu stops mid-batch exactly like the reference
    out = out_dir / 'capped'
    assert D.decode_latents_to_images(decode_fn, str(in_dir), str(out), batch_size_per_gpu=4, max_images_per_gpu=6, workers=2) == 6
    assert sorted(os.listdir(out)) == [f'rank00_{i:06d}.png' for i in range(6)]
