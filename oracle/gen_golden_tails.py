"""Generate tests/golden/tails.npz by running the REFERENCE's decoders (authoring container only; see
oracle/gen_golden.py for the import shim and conventions).

    python oracle/gen_golden_tails.py      # needs /root/reference

What is executed from the reference (eval mode, CPU fp32), with forward hooks capturing the tensor that enters
the tail this repo replaces:
  networks/pspnet_pop.py    PSPModule.forward; hook on bottleneck[0] (the 3x3 convolution): its output is the tail's
                            input x, the module's output is BN -> ReLU -> 1x1 conv + bias of it (:19-22)
  networks/convnext_pop.py  FPN_Seg_OCR_Decoder.forward; hook on .conv: LayerNorm over channels of it (:26-27)
  networks/swin_pop.py      UperNet_Decoder_Plus.forward; hooks on fpn_convs[i]: the output is
                            stack(interpolate(fpn_outs)).sum(-1) (:163-172)
  networks/deeplab_pop.py   _ASPP.forward; hook on fc.conv: the output is relu(bn(.)) of it (:12-29,61,66)
  networks/seghr_pop.py     HRFPN_Seg_Decoder.forward: cat of x[0] and the interpolated x[1:] (:8-24)
BatchNorm running statistics and affine parameters are randomised so the normalisation is not an identity.
"""
from __future__ import annotations

import os
import sys

import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
sys.path.insert(0, os.path.join(REPO, 'oracle'))

import gen_golden as gg  # noqa: E402


def randomise_norms(mod, gen):
    for m in mod.modules():
        if isinstance(m, nn.BatchNorm2d):
            m.running_mean.copy_(0.2 * torch.randn(m.num_features, generator=gen))
            m.running_var.copy_(0.5 + torch.rand(m.num_features, generator=gen))
            m.weight.data.copy_(1.0 + 0.3 * torch.randn(m.num_features, generator=gen))
            m.bias.data.copy_(0.2 * torch.randn(m.num_features, generator=gen))
        if isinstance(m, nn.LayerNorm):
            m.weight.data.copy_(1.0 + 0.3 * torch.randn(m.normalized_shape[0], generator=gen))
            m.bias.data.copy_(0.2 * torch.randn(m.normalized_shape[0], generator=gen))


def capture(module):
    box = []
    h = module.register_forward_hook(lambda m, i, o: box.append(o.detach().clone()))
    return box, h


def main():
    gg.import_reference()
    import networks.pspnet_pop as pspnet_pop
    import networks.convnext_pop as convnext_pop
    import networks.swin_pop as swin_pop
    arrays = {}

    def psp_case(name, feat_in, C, B, h, w, seed):
        gen = torch.Generator().manual_seed(seed)
        torch.manual_seed(seed)
        mod = pspnet_pop.PSPModule(feat_in, out_features=C).eval()
        randomise_norms(mod, gen)
        box, hk = capture(mod.bottleneck[0])
        with torch.no_grad():
            out = mod(torch.randn(B, feat_in, h, w, generator=gen))
        hk.remove()
        bn, conv = mod.bottleneck[1], mod.bottleneck[3]
        arrays.update({f'{name}_x': box[0].numpy(), f'{name}_out': out.numpy(),
                       f'{name}_bn_w': bn.weight.detach().numpy(), f'{name}_bn_b': bn.bias.detach().numpy(),
                       f'{name}_bn_m': bn.running_mean.numpy(), f'{name}_bn_v': bn.running_var.numpy(),
                       f'{name}_bn_eps': np.float64(bn.eps), f'{name}_W': conv.weight.detach().reshape(C, C).numpy(),
                       f'{name}_bias': conv.bias.detach().numpy()})
        print(name, tuple(box[0].shape), '->', tuple(out.shape), 'rms %.3f' % out.pow(2).mean().sqrt().item())

    psp_case('psp_c64', 16, 64, 2, 24, 24, seed=301)        # N = 576: ragged last 128-pixel tile
    psp_case('psp_c512', 8, 512, 1, 16, 16, seed=303)       # the PSPNet width, two n-tiles of 256
    psp_case('psp_c96', 8, 96, 1, 8, 24, seed=305)          # C % 64 != 0: partial k-block and partial n-tile

    def ln_case(name, chans, C, B, h, w, seed):
        gen = torch.Generator().manual_seed(seed)
        torch.manual_seed(seed)
        mod = convnext_pop.FPN_Seg_OCR_Decoder(sum(chans), C).eval()
        randomise_norms(mod, gen)
        box, hk = capture(mod.conv)
        xs = [torch.randn(B, c, max(1, h >> i), max(1, w >> i), generator=gen) * (1.0 + i) for i, c in enumerate(chans)]
        with torch.no_grad():
            out = mod(xs)
        hk.remove()
        arrays.update({f'{name}_x': box[0].numpy(), f'{name}_out': out.contiguous().numpy(),
                       f'{name}_gamma': mod.norm.weight.detach().numpy(), f'{name}_beta': mod.norm.bias.detach().numpy(),
                       f'{name}_eps': np.float64(mod.norm.eps)})
        print(name, tuple(box[0].shape), '->', tuple(out.shape), 'rms %.3f' % out.pow(2).mean().sqrt().item())

    ln_case('ln_c192', (8, 16, 32, 64), 192, 2, 16, 24, seed=311)      # ConvNeXt-T width; N = 384 (6 x 64-pixel tiles)
    ln_case('ln_c96', (8, 16, 32, 64), 96, 1, 8, 13 * 8, seed=313)     # Swin-T width; N = 832 = 13 x 64
    ln_case('ln_c480', (8, 16, 32, 64), 480, 1, 8, 9, seed=315)        # wide: 16-pixel tiles, ragged (N = 72)
    ln_case('ln_c40', (8, 8, 8, 8), 40, 1, 5, 8, seed=317)             # tiny, N = 40 < one tile

    def sum_case(name, filters, dim, B, h, w, seed):
        gen = torch.Generator().manual_seed(seed)
        torch.manual_seed(seed)
        mod = swin_pop.UperNet_Decoder_Plus(list(filters), dim).eval()
        randomise_norms(mod, gen)
        hooks = [capture(m) for m in mod.fpn_convs]
        xs = [torch.randn(B, c, h >> i, w >> i, generator=gen) for i, c in enumerate(filters)]
        with torch.no_grad():
            out = mod(xs)
        maps = []
        for box, hk in hooks:
            hk.remove()
            f = box[0]
            if f.shape[-2:] != xs[0].shape[-2:]:                       # swin_pop.py:165-168, stock interpolate
                f = F.interpolate(f, size=xs[0].shape[-2:], mode='bilinear', align_corners=True)
            maps.append(f)
        for i, f in enumerate(maps):
            arrays[f'{name}_map{i}'] = f.numpy()
        arrays[f'{name}_out'] = out.numpy()
        print(name, len(maps), 'maps', tuple(out.shape))

    sum_case('sum_swin', (16, 32, 64, 128), 16, 1, 48, 48, seed=321)

    # _ASPP (deeplab_pop.py:46-66): hook on fc.conv; the module output is relu(bn(.)) of it
    import networks.deeplab_pop as deeplab_pop
    import networks.seghr_pop as seghr_pop
    gen = torch.Generator().manual_seed(331)
    torch.manual_seed(331)
    aspp = deeplab_pop._ASPP(24, 64, rates=[1, 2, 3]).eval()
    randomise_norms(aspp, gen)
    box, hk = capture(aspp.fc.conv)
    with torch.no_grad():
        out = aspp(torch.randn(2, 24, 12, 10, generator=gen))
    hk.remove()
    bn = aspp.fc.bn
    arrays.update({'aspp_x': box[0].numpy(), 'aspp_out': out.numpy(), 'aspp_bn_w': bn.weight.detach().numpy(),
                   'aspp_bn_b': bn.bias.detach().numpy(), 'aspp_bn_m': bn.running_mean.numpy(),
                   'aspp_bn_v': bn.running_var.numpy(), 'aspp_bn_eps': np.float64(bn.eps)})
    print('aspp', tuple(box[0].shape), '->', tuple(out.shape))
    # HRFPN_Seg_Decoder (seghr_pop.py:8-24): inputs at four scales; the maps entering the concatenation are x[0] and
    # the stock align_corners interpolations of x[1:], recomputed here with the same call
    dec = seghr_pop.HRFPN_Seg_Decoder().eval()
    xs = [torch.randn(2, c, 16 >> i, 24 >> i, generator=gen) for i, c in enumerate((8, 16, 32, 64))]
    with torch.no_grad():
        out = dec(xs)
    maps = [xs[0]] + [F.interpolate(x, size=xs[0].shape[-2:], mode='bilinear', align_corners=True) for x in xs[1:]]
    for i, m in enumerate(maps):
        arrays[f'cat_hr_map{i}'] = m.numpy()
    arrays['cat_hr_out'] = out.numpy()
    print('cat_hr', [tuple(m.shape) for m in maps], '->', tuple(out.shape))
    path = os.path.join(REPO, 'tests', 'golden', 'tails.npz')
    np.savez_compressed(path, **arrays)
    print(path, '%.2f MB' % (os.path.getsize(path) / 1e6))


if __name__ == '__main__':
    main()
