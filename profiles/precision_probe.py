#!/usr/bin/env python
"""Measured error of the tensor-core background-MLP precision modes against the fp32 CPU oracle.
Run on the GPU box:  python profiles/precision_probe.py  (prints one line per case/mode)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import ref_ops  # noqa: E402
from segland_b200 import ops, synth  # noqa: E402


def main():
    cases = [(512, 0, 64, 1.0), (512, 4, 64, 1.0), (192, 4, 64, 1.0), (256, 4, 32, 1.0), (128, 4, 32, 1.0),
             (64, 4, 16, 1.0), (512, 0, 32, 100.0), (512, 0, 32, 1e-3)]
    for C, Kn, hw, scale in cases:
        st = synth.make_head_state(C, 7, Kn, seed=7 + C)
        labels = synth.make_labels(1, hw * 8, hw * 8, st.n_classes, seed=C, coarse=8)
        feats = (synth.make_features(labels, st, 8, seed=C).float() * scale).to(torch.bfloat16)
        ref = ref_ops.ref_head(feats.float(), st.base_emb, st.novel_emb, st.cls, st.cls_n)[:, 0].double()
        rms = ref.pow(2).mean().sqrt()
        for mode, prec in (('simt', 'precise'), ('tc', 'precise'), ('tc', 'balanced')):
            head = ops.PopHead(st.base_emb, st.cls, st.novel_emb, st.cls_n, bg_mode=mode, tc_precision=prec)
            out = head(feats.cuda())[:, 0].double().cpu()
            err = (out - ref).abs()
            relmax = (err.max() / ref.abs().max()).item()
            slack = (err - (1e-3 * ref.abs() + 1e-3 * rms)).max().item() / rms.item()
            print(f'C={C:3d} Kn={Kn} N={hw * hw:5d} scale={scale:g} {mode:4s}/{prec:8s} rel-to-max {relmax:.2e} '
                  f'elementwise-slack/rms {slack:+.2e} ({"ok" if slack <= 0 else "OVER"})')


if __name__ == '__main__':
    main()
