"""Full-size (configs[1]) golden logits: tests/golden/head_base_c512_full.npz (authoring container only).

    python oracle/gen_golden_full.py       # needs /root/reference

The reference's own GFSS_Model.forward_base (networks/pspnet_pop.py:161-182, identity backbone/decoder) is run on whole
1024^2-tile feature maps [1,512,128,128]; its logits [1,8,128,128] fp32 (0.5 MB per case), the prediction map the
script's post-processing gives (eval_base.py:168-170: F.interpolate(align_corners=True) -> np.argmax) and the confusion
matrix (eval_base.py:172-178 -> utils/pyt_utils.py:182-200) are stored.  The 16.8 MB of bf16 features per case are NOT
stored: they come from the seeded generator in segland_b200.synth (torch CPU RNG), and the file carries a SHA-256 of
their bit pattern so a drifted generator is detected instead of producing a bogus mismatch.
"""
from __future__ import annotations

import hashlib
import os
import sys

import numpy as np
import torch
import torch.nn.functional as F

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)

from oracle import gen_golden  # noqa: E402
from segland_b200 import synth  # noqa: E402

CASES = (('trained', 'trained-like head (bench workload)', 1234, 77),
         ('random', 'random-init head (near-ties everywhere)', 99, 78))


def make_case(kind, head_seed, data_seed):
    st = synth.make_trained_like_state(512, 7, 0, seed=head_seed) if kind == 'trained' else \
        synth.make_head_state(512, 7, 0, seed=head_seed)
    labels = synth.make_labels(1, 1024, 1024, st.n_classes, seed=data_seed)
    feats = synth.make_features(labels, st, 8, seed=data_seed)
    return st, labels, feats


def feature_digest(feats):
    return hashlib.sha256(feats.view(torch.int16).numpy().tobytes()).hexdigest()


def main():
    pspnet_pop, _, _, _, pyt_utils = gen_golden.import_reference()
    out = {}
    for kind, desc, head_seed, data_seed in CASES:
        st, labels, feats = make_case(kind, head_seed, data_seed)
        model = gen_golden.build_ref_model(pspnet_pop, st).eval()
        with torch.no_grad():
            logits = model(feats.float())                                   # forward_base: [1,8,128,128]
            up = F.interpolate(logits, size=labels.shape[-2:], mode='bilinear', align_corners=True)   # eval_base.py:168
        pred = np.asarray(np.argmax(up.numpy(), axis=1), dtype=np.uint8)    # eval_base.py:170
        gt = labels.numpy().astype(int)
        keep = gt != 255
        cm = pyt_utils.get_confusion_matrix(gt[keep], pred[keep], st.n_classes)
        out.update({f'{kind}_logits': logits.numpy(), f'{kind}_pred': pred, f'{kind}_cm': cm,
                    f'{kind}_seeds': np.array([head_seed, data_seed]), f'{kind}_feat_sha256': np.array(feature_digest(feats))})
        print(kind, desc, 'logits', tuple(logits.shape), 'mIoU-ish diag frac', float(np.trace(cm) / cm.sum()))
    path = os.path.join(REPO, 'tests', 'golden', 'head_base_c512_full.npz')
    np.savez_compressed(path, **out)
    print('wrote', path, os.path.getsize(path), 'bytes')


if __name__ == '__main__':
    main()
