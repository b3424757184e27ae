#!/usr/bin/env python
"""Debug aid: is the eager decode step bit-deterministic, and where does a CUDA-graph replay first differ from it?"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from vfm_vae_b200.decoder import SynthesisNetwork, F16D32_LEGACY_KWARGS, DecodeGraph

torch.backends.cudnn.benchmark = '--no-bench' not in sys.argv
torch.manual_seed(0)
B = 8
net = SynthesisNetwork(**F16D32_LEGACY_KWARGS).cuda().eval().requires_grad_(False)
z = torch.randn(B, 512, 16, 16, device='cuda'); ws = torch.randn(B, net.num_ws, 512, device='cuda')
with torch.no_grad():
    outs = []
    for i in range(4):
        img, multi = net(z, ws)
        outs.append([img.clone()] + [m.clone() for m in multi])
    for i in range(1, 4):
        print('eager run', i, 'vs 0:', [bool(torch.equal(a, b)) for a, b in zip(outs[i], outs[0])])
    g = DecodeGraph(net, z, ws)
    for i in range(3):
        img, multi = g(z, ws)
        torch.cuda.synchronize()
        print('graph replay', i, 'vs eager 0:', [bool(torch.equal(a, b)) for a, b in zip([img] + list(multi), outs[0])],
              [float((a - b).abs().max()) for a, b in zip([img] + list(multi), outs[0])])
