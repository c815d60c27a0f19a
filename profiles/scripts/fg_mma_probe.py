#!/usr/bin/env python
"""sl_pop_fg_lowres: FFMA2 kernel (SL_FG_MMA=0) vs mma.sync kernel (SL_FG_MMA=1) at the ft-mode (K = 11) and base-mode
(K = 7) shapes: us per 1024^2 tile and fraction of the measured HBM copy peak."""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from segland_b200 import _cabi, ops, synth  # noqa: E402

PEAK = json.load(open(os.path.join(ROOT, 'MEASURED_PEAKS.json')))['hbm_gbs'] if os.path.exists(
    os.path.join(ROOT, 'MEASURED_PEAKS.json')) else 6650.0


def timeit(fn, iters=30, warm=5):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(iters):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / iters * 1e-3


for name, C, hw, Kn in (('PSPNet ft C=512', 512, 128, 4), ('ConvNeXt-T ft C=192', 192, 256, 4), ('Swin ft C=96', 96, 256, 4),
                        ('PSPNet base C=512 (K=7)', 512, 128, 0), ('HRNet-w32 ft C=480', 480, 256, 4)):
    st = synth.make_head_state(C, 7, Kn, seed=2)
    T = 32 if C * hw * hw * 2 * 32 <= (1 << 30) else 16
    f = torch.randn(T, C, hw, hw, device='cuda').to(torch.bfloat16)
    head = ops.PopHead(st.base_emb, st.cls, st.novel_emb, st.cls_n, bg_mode='simt')
    lg = torch.empty(T, head.n_classes, hw, hw, device='cuda')
    byts = T * (C * hw * hw * 2 + head.K * hw * hw * 4)
    for mode in (0, 1):
        _cabi.set_env(SL_FG_MMA=mode)
        t = timeit(lambda: head(f, out=lg, fg_only=True))
        print(f'{name:28s} K={head.K:2d} {"mma.sync" if mode else "FFMA2   "}: {t * 1e6 / T:7.2f} us/tile {byts / t / 1e9:6.0f} GB/s = '
              f'{100 * byts / t / 1e9 / PEAK:5.1f} % of the HBM peak')
    _cabi.set_env(SL_FG_MMA=None)
