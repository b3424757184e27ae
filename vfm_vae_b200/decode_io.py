"""Decode I/O path: ``.safetensors`` latents -> decoder -> ``((img + 1) / 2).clamp(0, 1)`` -> PNG files
(reference tools/decode/decode_latents_to_images.py:27-100, the same tail as tools/reconstruct/reconstruct.py).

Once the decoder runs at > 1000 images/s, the reference's per-image tail -- ``img_tensor.cpu()`` (a blocking 786 KB fp32 copy per image),
``to_pil_image`` (multiply, cast and transpose on the host) and a PNG encode, all serial on the Python thread -- is what bounds the tool.
Here the tail is a pipeline:

    GPU      decode batch k+1                      | vfm_image_to_u8: NCHW fp32 -> NHWC uint8 in ONE pass (csrc/image_io.cu), bit-identical
    copy     D2H of batch k's uint8 pixels (a quarter of the fp32 bytes) into one of two pinned buffers, on a side stream
    host     PNG-encode batch k-1 on a thread pool (zlib releases the GIL), file names as the reference writes them

``decode_latents_to_images`` keeps the reference function's behaviour: files ``sorted(...)[rank::world_size]``, the ``latents`` /
``labels`` keys, batches of ``batch_size_per_gpu``, ``max_images_per_gpu``, output names ``rank{rank:02d}_{index:06d}.png``.
The decoder itself is passed in as a callable (latents, labels) -> images in [-1, 1]: the reference's ``Generator.decode`` wraps the LDM
adapter and the mapping network (out of the hot-path scope) around the pixel decoder.
"""
import ctypes as C
import os
from concurrent.futures import ThreadPoolExecutor

import torch

from . import _lib

_DT = {torch.float16: _lib.VFM_F16, torch.float32: _lib.VFM_F32}


def images_to_uint8(images, out=None, pre_add=1.0, pre_div=2.0, scale=255.0):
    """[N,C,H,W] fp16/fp32 CUDA images in [-1, 1] -> [N,H,W,C] uint8, ``trunc(clamp((x + 1) / 2, 0, 1) * 255)`` evaluated in fp32 exactly as
    the reference's ``((images + 1) / 2).clamp(0, 1)`` followed by torchvision's ``to_pil_image`` (``mul(255).byte()``).  CUDA only."""
    if not (isinstance(images, torch.Tensor) and images.is_cuda):
        raise RuntimeError('vfm_vae_b200.decode_io.images_to_uint8 runs the sm_100a kernel on CUDA tensors only (no CPU fallback)')
    if images.dim() != 4 or images.dtype not in _DT:
        raise RuntimeError('images_to_uint8: expected a [N,C,H,W] float16 / float32 tensor')
    x = images.contiguous()
    n, c, h, w = x.shape
    if out is None:
        out = torch.empty([n, h, w, c], dtype=torch.uint8, device=x.device)
    elif not (out.is_cuda and out.dtype == torch.uint8 and out.is_contiguous() and tuple(out.shape) == (n, h, w, c)):
        raise RuntimeError('images_to_uint8: `out` must be a contiguous CUDA uint8 tensor of shape [N,H,W,C]')
    if x.numel() == 0:
        return out
    p = _lib.ImageToU8Params()
    p.x, p.y, p.dtype = x.data_ptr(), out.data_ptr(), _DT[x.dtype]
    p.batch, p.channels, p.height, p.width = n, c, h, w
    p.pre_add, p.pre_div, p.scale = float(pre_add), float(pre_div), float(scale)
    with torch.cuda.device(x.device):
        _lib.check(_lib.load().vfm_image_to_u8(C.byref(p), C.c_void_p(torch.cuda.current_stream(x.device).cuda_stream)), 'image_to_u8')
    return out


def _save_png(array, path, compress_level):
    from PIL import Image
    img = Image.fromarray(array if array.shape[2] != 1 else array[:, :, 0])
    img.save(path, compress_level=compress_level)
    img.close()


class PngSink:
    """Asynchronous tail of the decode tools: uint8 conversion on the GPU, double-buffered pinned D2H on a side stream, PNG encoding on
    worker threads.  ``put(images, paths)`` returns as soon as the conversion and the copy are enqueued; ``close()`` drains."""

    def __init__(self, device, workers=None, compress_level=6, slots=2):
        self.device = torch.device(device)
        self.pool = ThreadPoolExecutor(max_workers=workers or min(32, (os.cpu_count() or 4)))
        self.copy_stream = torch.cuda.Stream(device=self.device)
        self.compress_level = compress_level          # PIL's default, as the reference's img.save(path)
        self.slots = [dict(host=None, dev=None, event=None, futures=[]) for _ in range(slots)]
        self.turn = 0
        self.saved = 0

    def _wait(self, slot):
        for f in slot['futures']:
            f.result()                                 # re-raises encoder / file errors
        slot['futures'] = []

    def put(self, images, paths):
        assert images.shape[0] == len(paths)
        slot = self.slots[self.turn]
        self.turn = (self.turn + 1) % len(self.slots)
        self._wait(slot)                               # the slot's previous batch has been encoded: its pinned buffer is free
        n, c, h, w = images.shape
        if slot['dev'] is None or slot['dev'].shape[0] < n or tuple(slot['dev'].shape[1:]) != (h, w, c):
            slot['dev'] = torch.empty([n, h, w, c], dtype=torch.uint8, device=self.device)
            slot['host'] = torch.empty([n, h, w, c], dtype=torch.uint8).pin_memory()
        dev, host = slot['dev'][:n], slot['host'][:n]
        images_to_uint8(images, out=dev)
        ready = torch.cuda.Event()
        ready.record(torch.cuda.current_stream(self.device))
        with torch.cuda.stream(self.copy_stream):
            self.copy_stream.wait_event(ready)
            host.copy_(dev, non_blocking=True)
            done = torch.cuda.Event()
            done.record(self.copy_stream)
        arr = host.numpy()

        def job(i, path):
            done.synchronize()                         # host-side wait in the worker, never on the decoding thread
            _save_png(arr[i], path, self.compress_level)
        slot['futures'] = [self.pool.submit(job, i, path) for i, path in enumerate(paths)]
        self.saved += n

    def close(self):
        for slot in self.slots:
            self._wait(slot)
        self.pool.shutdown(wait=True)


def list_latent_files(input_dir, rank=0, world_size=1):
    """The reference's split: every ``.safetensors`` file of the directory, sorted, strided by rank (decode_latents_to_images.py:45-49)."""
    files = sorted(f for f in os.listdir(input_dir) if f.endswith('.safetensors'))
    assert files, f'No .safetensors files found in {input_dir}'
    return files[rank::world_size], len(files)


@torch.no_grad()
def decode_latents_to_images(decode_fn, input_dir, output_dir, batch_size_per_gpu=32, rank=0, world_size=1, device='cuda',
                             max_images_per_gpu=None, workers=None, compress_level=6):
    """Mirror of the reference's ``run_latent_decoding`` (decode_latents_to_images.py:27-100) with the pipelined tail described above.
    ``decode_fn(latents, labels) -> images`` in [-1, 1], [B,C,H,W] on ``device``.  Returns the number of images written by this rank."""
    from safetensors.torch import load_file
    os.makedirs(output_dir, exist_ok=True)
    files, _ = list_latent_files(input_dir, rank, world_size)
    device = torch.device(device)
    sink = PngSink(device, workers=workers, compress_level=compress_level)
    global_index = saved = 0
    try:
        for file in files:
            if max_images_per_gpu is not None and saved >= max_images_per_gpu:
                break
            try:
                data = load_file(os.path.join(input_dir, file))
            except Exception as e:                       # noqa: BLE001  (the reference skips unreadable files with a warning)
                print(f'Failed to load {file}: {e}')
                continue
            if 'latents' not in data:
                print(f"Missing 'latents' in {file}")
                continue
            latents = data['latents'].to(device, non_blocking=True)
            labels = data['labels'].to(device) if 'labels' in data else torch.zeros(latents.size(0), device=device)
            for start in range(0, latents.size(0), batch_size_per_gpu):
                if max_images_per_gpu is not None and saved >= max_images_per_gpu:
                    break
                end = min(start + batch_size_per_gpu, latents.size(0))
                images = decode_fn(latents[start:end], labels[start:end])
                n = images.shape[0]
                if max_images_per_gpu is not None:
                    n = min(n, max_images_per_gpu - saved)
                paths = [os.path.join(output_dir, f'rank{rank:02d}_{global_index + i:06d}.png') for i in range(n)]
                sink.put(images[:n], paths)
                saved += n
                global_index += images.shape[0]
    finally:
        sink.close()
    return saved
