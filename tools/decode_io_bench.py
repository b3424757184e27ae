#!/usr/bin/env python
"""Decode tool end to end (safetensors latents -> f16d32 D-legacy decoder -> PNG files): the reference's serial tail
(tools/decode/decode_latents_to_images.py:89-99: per image .cpu(), to_pil_image, save) against vfm_vae_b200.decode_io's pipeline, same
decoder, same files.  Prints one JSON line; diagnostic (profiles/), not the bench contract."""
import argparse, json, os, shutil, sys, tempfile, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from PIL import Image
from safetensors.torch import load_file, save_file
from vfm_vae_b200 import decode_io as D
from vfm_vae_b200.decoder import SynthesisNetwork, F16D32_LEGACY_KWARGS

ap = argparse.ArgumentParser()
ap.add_argument('--images', type=int, default=512)
ap.add_argument('--batch', type=int, default=32)
ap.add_argument('--workers', type=int, default=None)
a = ap.parse_args()
dev = torch.device('cuda')
torch.backends.cudnn.benchmark = True
torch.manual_seed(0)
net = SynthesisNetwork(**F16D32_LEGACY_KWARGS).to(dev).eval().requires_grad_(False)
ws = torch.randn(1, net.num_ws, 512, device=dev)

def decode_fn(latents, labels):
    # stand-in for Generator.decode: the LDM adapter / mapping network are out of scope; latents ARE the decoder's z here
    n = latents.shape[0]
    if n == a.batch:
        return net.decode_graph(latents, ws.expand(n, -1, -1).contiguous())[0]
    return net(latents, ws.expand(n, -1, -1))[0]

root = tempfile.mkdtemp(prefix='vfm_decode_io_')
try:
    in_dir = os.path.join(root, 'latents'); os.makedirs(in_dir)
    per_file = 128
    g = torch.Generator().manual_seed(1)
    for i in range(a.images // per_file):
        save_file({'latents': torch.randn(per_file, 512, 16, 16, generator=g)}, os.path.join(in_dir, f'part_{i:03d}.safetensors'))
    with torch.no_grad():
        decode_fn(torch.randn(a.batch, 512, 16, 16, device=dev), None); torch.cuda.synchronize()

    def reference_tail(out_dir):
        os.makedirs(out_dir)
        idx = 0
        with torch.no_grad():
            for f in sorted(os.listdir(in_dir)):
                lat = load_file(os.path.join(in_dir, f))['latents'].to(dev)
                for s in range(0, lat.size(0), a.batch):
                    images = ((decode_fn(lat[s:s + a.batch], None) + 1) / 2).clamp(0, 1)
                    for i, t in enumerate(images):
                        arr = t.cpu().clamp(0, 1).mul(255).byte().permute(1, 2, 0).numpy()     # torchvision.to_pil_image's float path
                        img = Image.fromarray(arr); img.save(os.path.join(out_dir, f'rank00_{idx + i:06d}.png')); img.close()
                    idx += images.size(0)
        return idx

    t0 = time.perf_counter(); n_ref = reference_tail(os.path.join(root, 'ref')); t_ref = time.perf_counter() - t0
    t0 = time.perf_counter(); n_ours = D.decode_latents_to_images(decode_fn, in_dir, os.path.join(root, 'ours'), batch_size_per_gpu=a.batch, workers=a.workers); t_ours = time.perf_counter() - t0
    same = all(open(os.path.join(root, 'ref', f), 'rb').read() == open(os.path.join(root, 'ours', f), 'rb').read() for f in sorted(os.listdir(os.path.join(root, 'ref')))[:64])
    # decoder alone, for scale
    z = torch.randn(a.batch, 512, 16, 16, device=dev); torch.cuda.synchronize(); t0 = time.perf_counter()
    with torch.no_grad():
        for _ in range(a.images // a.batch): decode_fn(z, None)
    torch.cuda.synchronize(); t_dec = time.perf_counter() - t0
    print(json.dumps({'what': 'decode tool end to end, f16d32 D-legacy 256x256 fp16 blocks, batch %d' % a.batch, 'images': n_ours, 'host_cores': os.cpu_count(),
                      'reference_serial_tail_img_s': n_ref / t_ref, 'pipelined_img_s': n_ours / t_ours, 'decoder_only_img_s': a.images / t_dec,
                      'png_files_byte_identical_first_64': bool(same)}))
finally:
    shutil.rmtree(root, ignore_errors=True)
