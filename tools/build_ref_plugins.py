#!/usr/bin/env python
"""Compile the reference's OWN bias_act / upfirdn2d plugin sources, unmodified and with the flags its wrappers pass
(torch_utils/ops/bias_act.py:41-47, upfirdn2d.py:25-32), for sm_100 into the git-ignored oracle/_ref_plugins/ -- only for
tools/ref_gpu_bench.py (the reference's stock GPU path as a context number).  Needed because the reference's loader
(torch_utils/custom_ops.py:141) cannot import what it builds under torch >= 2.5; run in the build container (nvcc, no GPU needed):

    TORCH_CUDA_ARCH_LIST=10.0 python tools/build_ref_plugins.py
"""
import os
import sys

os.environ.setdefault('TORCH_CUDA_ARCH_LIST', '10.0')
import torch.utils.cpp_extension as ce  # noqa: E402

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.path.join(REPO, 'oracle', '_ref', 'torch_utils', 'ops')
if not os.path.isdir(SRC):
    sys.exit('the reference is not staged (python tools/stage_reference.py)')
FLAGS = ['--use_fast_math', '--allow-unsupported-compiler']
for name, srcs in (('bias_act_plugin', ['bias_act.cpp', 'bias_act.cu']), ('upfirdn2d_plugin', ['upfirdn2d.cpp', 'upfirdn2d.cu'])):
    bd = os.path.join(REPO, 'oracle', '_ref_plugins', name)
    os.makedirs(bd, exist_ok=True)
    ce.load(name=name, sources=[os.path.join(SRC, s) for s in srcs], extra_cuda_cflags=FLAGS, build_directory=bd, verbose=False)
    print('built', os.path.join(bd, name + '.so'))
