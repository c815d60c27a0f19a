#!/usr/bin/env python
"""A/B of the two operand schedules of bg_pair_kernel (SL_TC_PAIR=1: one (A, B) pair per pass and stage; default: one
copy of every operand tile of a k-block per stage): max relative difference of the background logits and CUDA-event time."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from segland_b200 import ops, synth  # noqa: E402


def timeit(fn, iters=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(iters):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / iters


for C, T, hw, prec in ((512, 32, 128, 'precise'), (512, 32, 128, 'balanced'), (192, 16, 256, 'precise'), (480, 4, 256, 'precise'),
                       (256, 32, 128, 'precise'), (512, 3, 120, 'precise')):
    st = synth.make_head_state(C, 7, 4, seed=C)
    f = synth.make_random_features(T, C, hw, hw, seed=C).cuda()
    head = ops.PopHead(st.base_emb, st.cls, st.novel_emb, st.cls_n, bg_mode='tc', tc_precision=prec)
    outs, times = {}, {}
    for mode in ('1', '2'):
        os.environ['SL_TC_PAIR'] = mode; __import__('segland_b200._cabi', fromlist=['x']).lib().sl_env_reload()
        lg = torch.empty(T, 12, hw, hw, device='cuda')
        head(f, out=lg)
        torch.cuda.synchronize()
        outs[mode] = lg[:, 0].clone()
        times[mode] = timeit(lambda: head.bg_tc(f, lg))
    d = (outs['1'] - outs['2']).abs().max().item() / outs['1'].abs().max().item()
    print(f'C={C:4d} T={T:3d} hw={hw} {prec:9s} rel diff {d:.2e}   pair {times["1"]:.4f} ms   dedup {times["2"]:.4f} ms   '
          f'({times["1"] / times["2"]:.3f}x)')
os.environ.pop('SL_TC_PAIR', None); __import__('segland_b200._cabi', fromlist=['x']).lib().sl_env_reload()
