#!/usr/bin/env python
"""Post-processing A/B on the bench shapes: per-cell pruning kernel (default) vs the row-cached kernel
(SL_POST_PRUNE=0), confusion fused vs second launch (SL_POST_FUSED_CM), and sl_window_accumulate bandwidth."""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from segland_b200 import _cabi, ops, synth  # noqa: E402

PEAK = json.load(open(os.path.join(ROOT, 'MEASURED_PEAKS.json')))['hbm_gbs'] if os.path.exists(
    os.path.join(ROOT, 'MEASURED_PEAKS.json')) else 6650.0


def timeit(fn, iters=20, warm=3):
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


def main():
    dev = 'cuda'
    for name, C, hw, Kb, Kn, stride in (('configs[1] PSPNet base K=8 x8', 512, 128, 7, 0, 8),
                                        ('configs[3] ConvNeXt ft K=12 x4', 192, 256, 7, 4, 4)):
        T = 32
        st = synth.make_trained_like_state(C, Kb, Kn, seed=1234)
        K = st.n_classes
        head = ops.PopHead(st.base_emb, st.cls, st.novel_emb, st.cls_n)
        data = []
        for coarse in (32, 8):                       # 32-pixel label regions (the bench workload) and 128-pixel regions
            labels_h = synth.make_labels(8, 1024, 1024, K, seed=1234, coarse=coarse)
            feats = synth.make_features(labels_h, st, stride, seed=1234).to(dev).repeat(4, 1, 1, 1)
            data.append((f'trained-like, {1024 // coarse}-px regions', head(feats), labels_h.to(dev).repeat(4, 1, 1)))
            del feats
        data.append(('pure randn logits', torch.randn_like(data[0][1]), data[0][2]))
        for data_name, x, labels in data:
            for prune in (1, 0):
                for fused in (None, 0, 1):
                    _cabi.set_env(SL_POST_PRUNE=prune, SL_POST_FUSED_CM=fused)
                    cm = torch.zeros(K, K, dtype=torch.int64, device=dev)
                    t = timeit(lambda: ops.upsample_argmax(x, (1024, 1024), label=labels, cm=cm))
                    t2 = timeit(lambda: ops.upsample_argmax(x, (1024, 1024)))
                    byts = T * (K * hw * hw * 4 + 2 * 1024 * 1024)
                    print(f'{name:32s} {data_name:32s} prune={prune} fused_cm={fused}: pred+cm {t * 1e3:7.3f} ms '
                          f'({byts / t / 1e9:6.0f} GB/s = {100 * byts / t / 1e9 / PEAK:4.1f}% HBM)  pred only {t2 * 1e3:7.3f} ms')
        _cabi.set_env(SL_POST_PRUNE=None, SL_POST_FUSED_CM=None)
    # ---- sliding-window aggregation: 1024^2 tile, stride-4 model (canvas 256^2), 512-px crops at stride 256 (3x3),
    # 2 views, K = 12: 18 entries per tile
    for name, tile, crop, stride, ms, flips in (('x4 model 3x3 windows x 2 views', 1024, 512, 256, 4, (0, 1)),
                                                ('x8 model 3x3 windows x 2 views', 1024, 512, 256, 8, (0, 1)),
                                                ('x4 model ragged stride 384, 1 view', 1024, 512, 384, 4, (0,))):
        plan = ops.WindowPlan((tile, tile), (crop, crop), (stride, stride), ms)
        B, K = 32, 12
        E = plan.n_windows * len(flips)
        hc, wc = plan.crop_lr_hw
        crops = torch.randn(B, E, K, hc, wc, device=dev)
        out = torch.empty(B, K, *plan.canvas_hw, device=dev)
        t = timeit(lambda: ops.window_accumulate(crops, plan, flips, out=out))
        byts = crops.numel() * 4 + out.numel() * 4
        print(f'window_accumulate {name:40s} {t * 1e6 / B:8.2f} us/tile {byts / t / 1e9:6.0f} GB/s = '
              f'{100 * byts / t / 1e9 / PEAK:4.1f}% HBM ({byts / 1e6:.0f} MB)')


if __name__ == '__main__':
    main()
