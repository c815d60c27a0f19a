#!/usr/bin/env python
"""Decoder tails: achieved bandwidth / throughput of the fused tails next to the stock eager
PyTorch sequence they replace (same GPU, same tensors; the eager lines are context, not a target).
    python profiles/bench_tails.py          (on the GPU box)
Algorithmic bytes per element: 4 (fp32 in) + 2 (bf16 out); sum tail 4*M + 2; the conv tail additionally does
2*Cin*Cout FLOP per pixel (3 split-bf16 passes executed)."""
import os
import sys

import torch
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'profiles'))
from segland_b200 import ops  # noqa: E402
from bench_ops import report, timeit  # noqa: E402


def main():
    dev = 'cuda'
    g = torch.Generator(dev).manual_seed(0)
    conv_only = '--conv-only' in sys.argv
    for name, C, hw, T in () if conv_only else (('ConvNeXt-T C=192 256^2', 192, 256, 16), ('Swin-T/S C=96 256^2', 96, 256, 32),
                           ('wide head C=480 256^2 (no reference model: 32-pixel tiles)', 480, 256, 8)):
        x = torch.randn(T, C, hw, hw, device=dev, generator=g)
        gamma, beta = torch.rand(C, device=dev) + 0.5, torch.randn(C, device=dev)
        out = torch.empty(T, C, hw, hw, dtype=torch.bfloat16, device=dev)
        t = timeit(lambda: ops.layernorm_tail(x, gamma, beta, 1e-5, out=out))
        report(f'layernorm_tail {name}', t, x.numel() * 6, T)
        t = timeit(lambda: F.layer_norm(x.permute(0, 2, 3, 1), (C,), gamma, beta, 1e-5).permute(0, 3, 1, 2)
                   .to(torch.bfloat16).contiguous(), iters=5)
        report(f'  eager: layer_norm(permute) -> permute -> bf16 contiguous', t, x.numel() * 6, T)
        del x, out
    for name, C, hw, T, M in () if conv_only else (('Swin-T/S C=96 256^2 M=4', 96, 256, 16, 4), ('LSK-T C=192 256^2 M=4', 192, 256, 8, 4)):
        maps = [torch.randn(T, C, hw, hw, device=dev, generator=g) for _ in range(M)]
        out = torch.empty(T, C, hw, hw, dtype=torch.bfloat16, device=dev)
        t = timeit(lambda: ops.sum_tail(maps, out=out))
        report(f'sum_tail {name}', t, maps[0].numel() * (4 * M + 2), T)
        t = timeit(lambda: torch.stack(maps, dim=-1).sum(-1).to(torch.bfloat16), iters=5)
        report(f'  eager: stack(..., -1).sum(-1) -> bf16', t, maps[0].numel() * (4 * M + 2), T)
        del maps, out
    if not conv_only:
        x = torch.randn(32, 256, 128, 128, device=dev, generator=g)
        bn = torch.nn.BatchNorm2d(256).to(dev).eval()
        bn.running_mean.normal_(0, 0.2); bn.running_var.uniform_(0.5, 1.5)
        out = torch.empty(32, 256, 128, 128, dtype=torch.bfloat16, device=dev)
        t = timeit(lambda: ops.bn_relu_tail(x, bn, out=out))
        report('bn_relu_tail DeepLab C=256 128^2', t, x.numel() * 6, 32)
        with torch.no_grad():
            t = timeit(lambda: F.relu(bn(x)).to(torch.bfloat16), iters=5)
        report('  eager: bn -> relu -> bf16', t, x.numel() * 6, 32)
        del x, out
        maps = [torch.randn(8, c, 256, 256, device=dev, generator=g) for c in (32, 64, 128, 256)]
        out = torch.empty(8, 480, 256, 256, dtype=torch.bfloat16, device=dev)
        t = timeit(lambda: ops.concat_tail(maps, out=out))
        report('concat_tail HRNet-w32 C=480 256^2', t, out.numel() * 6, 8)
        t = timeit(lambda: torch.cat(maps, 1).to(torch.bfloat16), iters=5)
        report('  eager: cat -> bf16', t, out.numel() * 6, 8)
        del maps, out
    for name, C, hw, T in (('PSPNet C=512 128^2', 512, 128, 32), ('PSPNet C=512 64^2 (512^2 tiles)', 512, 64, 64)):
        x = torch.randn(T, C, hw, hw, device=dev, generator=g)
        bn = torch.nn.BatchNorm2d(C).to(dev).eval()
        bn.running_mean.normal_(0, 0.2); bn.running_var.uniform_(0.5, 1.5)
        conv = torch.nn.Conv2d(C, C, 1).to(dev)
        tail = ops.ConvTail(conv.weight, conv.bias, bn=(bn.weight, bn.bias, bn.running_mean, bn.running_var, bn.eps))
        out = torch.empty(T, C, hw, hw, dtype=torch.bfloat16, device=dev)
        t = timeit(lambda: tail(x, out=out))
        flops = 2 * C * C * hw * hw * T
        print(f'{"conv_tail BN+ReLU+1x1 conv+bias -> bf16 " + name:58s} {t * 1e6 / T:9.2f} us/tile {T / t:10.0f} tiles/s '
              f'{flops / t / 1e12:7.1f} TFLOP/s algorithmic ({3 * flops / t / 1e12:.0f} executed), {x.numel() * 6 / t / 1e9:.0f} GB/s in+out')
        if conv_only:
            continue
        with torch.no_grad():
            for tf32 in (True, False):
                torch.backends.cudnn.allow_tf32 = tf32
                t = timeit(lambda: conv(F.relu(bn(x))).to(torch.bfloat16), iters=5)
                print(f'{"  eager: bn -> relu -> conv1x1 -> bf16 (" + ("TF32" if tf32 else "fp32") + " conv)":58s} '
                      f'{t * 1e6 / T:9.2f} us/tile {T / t:10.0f} tiles/s {flops / t / 1e12:7.1f} TFLOP/s')
            torch.backends.cudnn.allow_tf32 = True
        del x, out


if __name__ == '__main__':
    main()
