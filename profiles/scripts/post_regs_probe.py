#!/usr/bin/env python
"""Post-processing A/B on the bench shapes: register-resident kernel (post_regs.cu, default) vs the row-cached
shared-memory kernel (SL_POST_REGS=0), with and without the fused confusion counts; 32 tiles, CUDA-event timed."""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from segland_b200 import _cabi, ops, synth  # noqa: E402

PEAK = json.load(open(os.path.join(ROOT, 'MEASURED_PEAKS.json')))['hbm_gbs']


def timeit(fn, iters=30, warm=5):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda._sleep(4_000_000)                 # ~2 ms: the host queues every launch before the first one starts
    a.record()
    for _ in range(iters):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / iters * 1e-3


def main():
    dev = 'cuda'
    T = int(os.environ.get('PROBE_TILES', '32'))
    for name, C, hw, Kb, Kn, stride in (('configs[1] PSPNet base K=8 x8', 512, 128, 7, 0, 8),
                                        ('PSPNet ft K=12 x8', 512, 128, 7, 4, 8),
                                        ('configs[3] ConvNeXt ft K=12 x4', 192, 256, 7, 4, 4)):
        st = synth.make_trained_like_state(C, Kb, Kn, seed=1234)
        K = st.n_classes
        head = ops.PopHead(st.base_emb, st.cls, st.novel_emb, st.cls_n)
        data = []
        for coarse in (32, 8):
            labels_h = synth.make_labels(8, 1024, 1024, K, seed=1234, coarse=coarse)
            feats = synth.make_features(labels_h, st, stride, seed=1234).to(dev).repeat(T // 8, 1, 1, 1)
            data.append((f'trained-like, {1024 // coarse}-px regions', head(feats), labels_h.to(dev).repeat(T // 8, 1, 1)))
            del feats
        data.append(('pure randn logits', torch.randn_like(data[0][1]), data[0][2]))
        for data_name, x, labels in data:
            ref = None
            for regs in (1, 0):
                _cabi.set_env(SL_POST_REGS=regs)
                cm = torch.zeros(K, K, dtype=torch.int64, device=dev)
                out = ops.upsample_argmax(x, (1024, 1024), label=labels, cm=cm)
                if ref is None:
                    ref = (out['pred'].clone(), cm.clone())
                else:
                    assert torch.equal(ref[0], out['pred']) and torch.equal(ref[1], cm), 'kernels disagree'
                # straight through the C ABI with preallocated outputs: the Python wrapper (allocation, checks) costs
                # ~0.1 ms per call, as much as the kernel
                pred = out['pred']
                st_ = torch.cuda.current_stream().cuda_stream
                B_, _, h_, w_ = x.shape
                t = timeit(lambda: _cabi.call('sl_upsample_argmax', _cabi.ptr(x), B_, K, h_, w_, 1024, 1024, _cabi.ptr(labels),
                                              255, _cabi.ptr(pred), None, None, None, _cabi.ptr(cm), st_))
                t2 = timeit(lambda: _cabi.call('sl_upsample_argmax', _cabi.ptr(x), B_, K, h_, w_, 1024, 1024, None,
                                               255, _cabi.ptr(pred), None, None, None, None, st_))
                byts = T * (K * hw * hw * 4 + 2 * 1024 * 1024)
                print(f'{name:32s} {data_name:32s} regs={regs}: pred+cm {t * 1e3:7.3f} ms '
                      f'({byts / t / 1e9:6.0f} GB/s = {100 * byts / t / 1e9 / PEAK:4.1f}% HBM)  pred only {t2 * 1e3:7.3f} ms',
                      flush=True)
        _cabi.set_env(SL_POST_REGS=None)


if __name__ == '__main__':
    main()
