#!/usr/bin/env python
"""How many tensor-core passes does the background MLP need?  For the three modes of sl_pop_bg_tc --
precise (2+3 split-bf16 passes), mid (2+2: fp16 hi/lo hidden layer x fp16 W2) and balanced (2+1: fp16 hidden layer x
fp16 W2) -- measure, against the fp32 CPU oracle (or the reference-generated golden logits):
  * rel-to-max of the background logit and of the whole logit tensor,
  * the element-wise slack  max(|err| - (1e-3 |ref| + 1e-3 rms(ref))) / rms  (<= 0 passes the element-wise bound),
  * arg-max agreement of the up-sampled prediction with the oracle's and the largest top-2 gap (relative to max|logit|)
    among disagreeing pixels (near-tie confinement),
on the five reference goldens, the eight shapes of profiles/precision_probe.py and two full 1024^2 PSPNet tiles, plus
the kernel time of each mode on the bench shape.    python profiles/scripts/pass_probe.py   (GPU box)"""
import os
import sys

import numpy as np
import torch
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tests'))
from helpers import bf16_from_bits, state_from_npz  # noqa: E402
from oracle import ref_ops  # noqa: E402
from segland_b200 import ops, synth  # noqa: E402

MODES = (('precise', '2+3'), ('mid', '2+2'), ('balanced', '2+1'))


def report(name, st, feats, ref_logits, out_size):
    ref = ref_logits.double()
    rms = ref.pow(2).mean().sqrt()
    ref_hr = F.interpolate(ref_logits.float(), size=out_size, mode='bilinear', align_corners=True)
    ref_pred = ref_hr.argmax(1)
    top2 = ref_hr.topk(2, dim=1)[0]
    for mode, passes in MODES:
        head = ops.PopHead(st.base_emb, st.cls, st.novel_emb, st.cls_n, bg_mode='tc', tc_precision=mode)
        lg = head(feats.cuda())
        out = lg.double().cpu()
        err = (out - ref).abs()
        rel_bg = (err[:, 0].max() / ref[:, 0].abs().max()).item()
        rel_all = (err.max() / ref.abs().max()).item()
        slack = ((err - (1e-3 * ref.abs() + 1e-3 * rms)).max() / rms).item()
        pred = ops.upsample_argmax(lg, out_size)['pred'].cpu().long()
        dis = pred != ref_pred
        agree = 1.0 - dis.float().mean().item()
        gap = ((top2[:, 0] - top2[:, 1])[dis].max() / ref_hr.abs().max()).item() if dis.any() else 0.0
        print(f'{name:34s} {mode:8s} ({passes}) bg rel-to-max {rel_bg:.2e}  all {rel_all:.2e}  elementwise slack/rms '
              f'{slack:+.2e} {"ok  " if slack <= 0 else "OVER"}  argmax agree {100 * agree:8.4f}%  worst gap {gap:.1e}')


def main():
    torch.manual_seed(0)
    print('--- reference goldens (logits produced by the reference itself)')
    for case in ('head_base_c64', 'head_ft_c64', 'head_base_c512', 'head_ft_c192_s4', 'head_ft_c96_rand'):
        z = np.load(os.path.join(ROOT, 'tests', 'golden', case + '.npz'))
        st = state_from_npz(z)
        if st.base_emb.shape[1] % 32:
            continue
        feats = bf16_from_bits(z['feats_bf16_bits'])
        report(case, st, feats, torch.from_numpy(z['logits']), z['labels'].shape[-2:])
    print('--- precision_probe shapes (random-init heads, oracle on the CPU)')
    for C, Kn, hw, scale in ((512, 0, 64, 1.0), (512, 4, 64, 1.0), (192, 4, 64, 1.0), (256, 4, 32, 1.0), (128, 4, 32, 1.0),
                             (64, 4, 16, 1.0), (512, 0, 32, 100.0), (512, 0, 32, 1e-3)):
        st = synth.make_head_state(C, 7, Kn, seed=7 + C)
        labels = synth.make_labels(1, hw * 8, hw * 8, st.n_classes, seed=C, coarse=8)
        feats = (synth.make_features(labels, st, 8, seed=C).float() * scale).to(torch.bfloat16)
        ref = ref_ops.ref_head(feats.float(), st.base_emb, st.novel_emb, st.cls, st.cls_n)
        report(f'C={C} Kn={Kn} N={hw * hw} scale={scale:g}', st, feats, ref, (hw * 8, hw * 8))
    print('--- full 1024^2 PSPNet tiles (C=512, 128^2 features)')
    for name, st in (('trained-like head (bench)', synth.make_trained_like_state(512, 7, 0, seed=1234)),
                     ('random-init head', synth.make_head_state(512, 7, 0, seed=99)),
                     ('random-init ft head, randn feats', synth.make_head_state(512, 7, 4, seed=98))):
        labels = synth.make_labels(1, 1024, 1024, st.n_classes, seed=5)
        feats = synth.make_random_features(1, 512, 128, 128, seed=5) if 'randn' in name else \
            synth.make_features(labels, st, 8, seed=5)
        ref = ref_ops.ref_head(feats.float(), st.base_emb, st.novel_emb, st.cls, st.cls_n)
        report(name, st, feats, ref, (1024, 1024))
    print('--- kernel time on the bench shape (32 tiles, C=512, 128^2), CUDA events')
    st = synth.make_trained_like_state(512, 7, 0, seed=1234)
    feats = torch.randn(32, 512, 128, 128, device='cuda').to(torch.bfloat16)
    lg = torch.empty(32, 8, 128, 128, device='cuda')
    for mode, passes in MODES:
        head = ops.PopHead(st.base_emb, st.cls, None, None, bg_mode='tc', tc_precision=mode)
        for _ in range(3):
            head.bg_tc(feats, lg)
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(20):
            head.bg_tc(feats, lg)
        b.record()
        torch.cuda.synchronize()
        ms = a.elapsed_time(b) / 20
        fl = (4 * 512 * 512 + 2 * 512) * 128 * 128 * 32
        print(f'bg kernel {mode:8s} ({passes}): {ms:.3f} ms / 32 tiles = {fl / ms / 1e9:.0f} TFLOP/s algorithmic')


if __name__ == '__main__':
    main()
