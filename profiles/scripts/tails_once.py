#!/usr/bin/env python
"""One warm-up + one launch of every decoder-tail kernel at its bench shape, for `ncu --set full` captures:
    ncu --set full --clock-control none --import-source on -k regex:'tail|bn_relu_split|tc_gemm|concat_cast' \
        --launch-skip <warm-up launches> -o gpurun_out/r1_tails python profiles/scripts/tails_once.py"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from segland_b200 import ops  # noqa: E402

dev = 'cuda'
g = torch.Generator(dev).manual_seed(0)
x = torch.randn(16, 192, 256, 256, device=dev, generator=g)
gamma, beta = torch.rand(192, device=dev) + 0.5, torch.randn(192, device=dev)
maps = [torch.randn(8, 96, 256, 256, device=dev, generator=g) for _ in range(4)]
xc = torch.randn(32, 512, 128, 128, device=dev, generator=g)
W, b = torch.randn(512, 512, device=dev) / 22.6, torch.randn(512, device=dev)
bn = (torch.rand(512, device=dev) + 0.5, torch.randn(512, device=dev), torch.randn(512, device=dev), torch.rand(512, device=dev) + 0.5, 1e-5)
tail = ops.ConvTail(W, b, bn=bn)
for _ in range(2):
    ops.layernorm_tail(x, gamma, beta, 1e-5)
    ops.sum_tail(maps)
    tail(xc)
    ops.bn_relu_tail(xc, bn)
    torch.cuda.synchronize()
