"""Generate tests/golden/train_grads.npz by running the REFERENCE's training-mode forwards + autograd
(authoring container only; see oracle/gen_golden.py for the import shim and conventions).

    python oracle/gen_golden_train.py      # needs /root/reference

What is executed from the reference:
  networks/pspnet_pop.py   GFSS_Model.forward -> forward_novel (train mode, is_ft) and forward_base (train mode)
                           with identity backbone/decoder, criterion = loss/criterion.py OrthLoss;
                           total_loss.backward() gives the gradients of novel_emb, classifier_n, classifier,
                           base_emb (base mode) and of the input features.
Cases:
  ft_c64    C=64,  Kb=7, Kn=4, 2 novel + 2 base images, 64x64 labels, stride 8 (N=64: CUDA-core forward)
  ft_c64b   C=64,  Kb=7, Kn=4, 1 + 1 images, 128x128 labels, stride 8 (N=256: tensor-core forward)
  ft_c96    C=96,  Kb=5, Kn=3, 1 + 1 images, 96x64 labels, stride 4 (non-square, ragged GEMM tiles)
  base_c64  C=64,  Kb=7, 2 images, 64x64, stride 8 (forward_base: base_emb + classifier + features train)
"""
from __future__ import annotations

import os
import sys

import numpy as np
import torch

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
sys.path.insert(0, os.path.join(REPO, 'oracle'))

from segland_b200 import synth  # noqa: E402
import gen_golden as gg  # noqa: E402


def conv_grads(seq):
    return [seq[i].weight.grad.reshape(seq[i].weight.shape[0], -1).squeeze(0).numpy().copy() for i in (0, 2, 4)]


def main():
    pspnet_pop, _, _, criterion, _ = gg.import_reference()
    out_dir = os.path.join(REPO, 'tests', 'golden')
    crit = criterion.OrthLoss(ignore_index=255)
    arrays = {}

    def ft_case(name, C, Kb, Kn, nb, H, W, stride, seed):
        gen = torch.Generator().manual_seed(seed)
        st = synth.make_head_state(C, Kb, Kn, seed=seed)
        st.novel_emb = st.novel_emb + 0.3 * torch.randn(Kn, C, generator=gen)
        model = gg.build_ref_model(pspnet_pop, st, criterion=crit)
        model.train()
        lab_n = synth.make_labels(nb, H, W, 1 + Kb + Kn, seed=seed, coarse=8)
        lab_b = synth.make_labels(nb, H, W, 1 + Kb, seed=seed + 1, coarse=8)
        img_n = synth.make_features(lab_n, st, stride, seed=seed).float().requires_grad_(True)
        img_b = synth.make_features(lab_b, st, stride, seed=seed + 1).float().requires_grad_(True)
        mask_n, mask_b = lab_n.long(), lab_b.long()
        mask_b_before = mask_b.clone()
        loss = model(img_n, mask_n, img_b, mask_b)                  # forward_novel (pspnet_pop.py:191-245)
        loss['total_loss'].backward()
        g1, g2, g3 = conv_grads(model.classifier)
        n1, n2, n3 = conv_grads(model.classifier_n)
        arrays.update({f'{name}_img_n_bits': img_n.detach().to(torch.bfloat16).view(torch.int16).numpy(),
                       f'{name}_img_b_bits': img_b.detach().to(torch.bfloat16).view(torch.int16).numpy(),
                       f'{name}_mask_n': mask_n.numpy(), f'{name}_mask_b_before': mask_b_before.numpy(),
                       f'{name}_mask_b_after': mask_b.numpy(), f'{name}_stride': np.int64(stride),
                       f'{name}_total': loss['total_loss'].detach().numpy(), f'{name}_seg': loss['seg_loss'].detach().numpy(),
                       f'{name}_orth': loss['orth_loss'].detach().numpy(),
                       f'{name}_g_novel_emb': model.novel_emb.grad.numpy().copy(),
                       f'{name}_g_W1': g1, f'{name}_g_W2': g2, f'{name}_g_w3': g3,
                       f'{name}_g_W1n': n1, f'{name}_g_W2n': n2, f'{name}_g_w3n': n3,
                       f'{name}_g_img_n': img_n.grad.numpy().copy(), f'{name}_g_img_b': img_b.grad.numpy().copy()})
        arrays.update(gg.state_arrays(name + '_', st))
        print(name, 'total %.6f' % float(loss['total_loss']), '|g novel| %.3e' % model.novel_emb.grad.abs().max().item(),
              '|g W1n| %.3e |g W2n| %.3e |g w3n| %.3e |g W1| %.3e |g img| %.3e' % (
                  np.abs(n1).max(), np.abs(n2).max(), np.abs(n3).max(), np.abs(g1).max(), img_n.grad.abs().max().item()))

    ft_case('ft_c64', 64, 7, 4, 2, 64, 64, 8, seed=81)
    ft_case('ft_c64b', 64, 7, 4, 1, 128, 128, 8, seed=83)
    ft_case('ft_c96', 96, 5, 3, 1, 96, 64, 4, seed=85)

    # forward_base in train mode (train_base.py:259): base_emb, classifier and the features all get gradients
    gen = torch.Generator().manual_seed(91)
    st = synth.make_head_state(64, 7, 0, seed=91)
    st.base_emb = st.base_emb + 0.3 * torch.randn(7, 64, generator=gen)
    model = gg.build_ref_model(pspnet_pop, st, criterion=crit)
    model.train()
    lab = synth.make_labels(2, 64, 64, 8, seed=91, coarse=8)
    img = synth.make_features(lab, st, 8, seed=91).float().requires_grad_(True)
    loss = model(img, lab.long())                                   # forward_base (pspnet_pop.py:161-189)
    loss['total_loss'].backward()
    g1, g2, g3 = conv_grads(model.classifier)
    arrays.update({'base_c64_img_bits': img.detach().to(torch.bfloat16).view(torch.int16).numpy(),
                   'base_c64_mask': lab.numpy(), 'base_c64_stride': np.int64(8),
                   'base_c64_total': loss['total_loss'].detach().numpy(), 'base_c64_seg': loss['seg_loss'].detach().numpy(),
                   'base_c64_orth': loss['orth_loss'].detach().numpy(),
                   'base_c64_g_base_emb': model.base_emb.grad.numpy().copy(),
                   'base_c64_g_W1': g1, 'base_c64_g_W2': g2, 'base_c64_g_w3': g3,
                   'base_c64_g_img': img.grad.numpy().copy()})
    arrays.update(gg.state_arrays('base_c64_', st))
    print('base_c64 total %.6f |g base| %.3e |g W1| %.3e' % (float(loss['total_loss']),
                                                            model.base_emb.grad.abs().max().item(), np.abs(g1).max()))
    path = os.path.join(out_dir, 'train_grads.npz')
    np.savez_compressed(path, **arrays)
    print(path, '%.2f MB' % (os.path.getsize(path) / 1e6))


if __name__ == '__main__':
    main()
