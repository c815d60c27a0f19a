#!/usr/bin/env python
"""Soft-output (probability map) path of sl_upsample_argmax: register-resident kernel (default at K = 12) vs the row-cached
shared-memory kernel (SL_POST_REGS=0); K = 12 x4 (configs[3]) and K = 8 x8, 16 and 32 tiles, CUDA-event timed."""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from segland_b200 import _cabi  # noqa: E402

PEAK = json.load(open(os.path.join(ROOT, 'MEASURED_PEAKS.json')))['hbm_gbs']


def timeit(fn, iters=20, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda._sleep(4_000_000)
    a.record()
    for _ in range(iters):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / iters * 1e-3


def main():
    st = torch.cuda.current_stream().cuda_stream
    for K, hw in ((12, 256), (8, 128)):
        for T in (16, 32):
            lg = torch.randn(T, K, hw, hw, device='cuda')
            pred = torch.empty(T, 1024, 1024, dtype=torch.uint8, device='cuda')
            probs = torch.empty(T, K, 1024, 1024, device='cuda')
            byts = T * (K * hw * hw * 4 + K * 1024 * 1024 * 4 + 1024 * 1024)
            ref = None
            for regs in (1, 0):
                _cabi.set_env(SL_POST_REGS=regs)
                fn = lambda: _cabi.call('sl_upsample_argmax', _cabi.ptr(lg), T, K, hw, hw, 1024, 1024, None, 255,
                                        _cabi.ptr(pred), None, _cabi.ptr(probs), None, None, st)
                fn()
                if ref is None:
                    ref = probs.clone()
                else:
                    assert torch.equal(ref, probs), 'kernels disagree'
                t = timeit(fn)
                print(f'K={K} x{1024 // hw} T={T} regs={regs}: {t * 1e6 / T:7.2f} us/tile  {byts / t / 1e9:6.0f} GB/s = '
                      f'{100 * byts / t / 1e9 / PEAK:4.1f}% of the HBM copy peak', flush=True)
            del lg, pred, probs, ref
    _cabi.set_env(SL_POST_REGS=None)


if __name__ == '__main__':
    main()
