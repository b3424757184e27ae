#!/usr/bin/env python
"""Stage the UNMODIFIED reference decoder path under oracle/_ref/ so that it travels to the GPU box.

    python tools/stage_reference.py            # /root/reference -> oracle/_ref/   (also run by __graft_entry__.build())

The reference (tianciB/VFM-VAE) is a plain Python package without setup.py; its pixel-decoder path needs only the three
package directories below (networks/, torch_utils/ incl. the plugin .cpp/.cu sources its own loader JIT-compiles, dnnlib/).
They are copied byte for byte -- nothing is edited -- into the git-ignored, NOT gpurun-ignored ``oracle/_ref/`` together with
the two-symbol test-only ``timm`` shim (tools/ref_shims/timm: ``trunc_normal_`` and ``get_norm_layer``, which the reference
imports at module import time but which take no part in the decoder's arithmetic).  No reference source enters the git
history; ``oracle/_ref/MANIFEST.json`` records the sha256 of every staged file for the parity tests to check.

Consumers (all test / measurement infrastructure, never the product): ``bench.py --impl reference`` and the
``cpu_baseline`` leg (the reference's own impl='ref' CPU path), tests/test_reference_decoder_gpu.py (the reference's
SynthesisNetwork on CUDA through vfm_vae_b200.integration.install()), tools/ref_gpu_bench.py (the reference's stock GPU
path as context).
"""
import hashlib
import json
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(HERE)
REF = os.environ.get('VFM_REFERENCE', '/root/reference')
DST = os.path.join(REPO, 'oracle', '_ref')
PACKAGES = ('networks', 'torch_utils', 'dnnlib')
KEEP_EXT = ('.py', '.cpp', '.cu', '.h')


def stage(verbose=True):
    """-> True if oracle/_ref/ is (now) populated, False if there is no reference checkout to stage from."""
    if not os.path.isdir(os.path.join(REF, 'torch_utils')):
        if verbose:
            print(f'stage_reference: no reference checkout at {REF}; keeping whatever is in {DST}')
        return os.path.isfile(os.path.join(DST, 'MANIFEST.json'))
    manifest = {}
    tmp = DST + '.tmp'
    shutil.rmtree(tmp, ignore_errors=True)
    os.makedirs(tmp)
    for pkg in PACKAGES:
        for root, dirs, files in os.walk(os.path.join(REF, pkg)):
            dirs[:] = [d for d in dirs if d != '__pycache__']
            for f in files:
                if not f.endswith(KEEP_EXT):
                    continue
                src = os.path.join(root, f)
                rel = os.path.relpath(src, REF)
                dst = os.path.join(tmp, rel)
                os.makedirs(os.path.dirname(dst), exist_ok=True)
                shutil.copyfile(src, dst)
                manifest[rel] = hashlib.sha256(open(src, 'rb').read()).hexdigest()
    shutil.copytree(os.path.join(HERE, 'ref_shims', 'timm'), os.path.join(tmp, '_shims', 'timm'),
                    ignore=shutil.ignore_patterns('__pycache__'))
    json.dump({'source': REF, 'files': manifest}, open(os.path.join(tmp, 'MANIFEST.json'), 'w'), indent=1, sort_keys=True)
    shutil.rmtree(DST, ignore_errors=True)
    os.rename(tmp, DST)
    if verbose:
        print(f'stage_reference: {len(manifest)} files -> {DST}')
    return True


if __name__ == '__main__':
    sys.exit(0 if stage() else 1)
