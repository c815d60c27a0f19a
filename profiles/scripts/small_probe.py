"""Weights-resident small-C background kernel (SL_TC_SMALL=1) vs the streaming kernels: parity + time."""
import os, sys, torch
sys.path.insert(0, '.')
from segland_b200 import ops, synth
for C, B, h, w in [(96, 16, 256, 256), (128, 16, 256, 256), (64, 16, 256, 256), (32, 4, 64, 64), (96, 3, 24, 16), (128, 1, 16, 8)]:
    st = synth.make_head_state(C, 7, 4, seed=2)
    head = ops.PopHead(st.base_emb, st.cls, st.novel_emb, st.cls_n, bg_mode='tc')
    feats = synth.make_random_features(B, C, h, w, seed=1).cuda()
    outs, times = {}, {}
    for small in ('0', '1'):
        os.environ['SL_TC_SMALL'] = small; __import__('segland_b200._cabi', fromlist=['x']).lib().sl_env_reload()
        lg = torch.zeros(B, 12, h, w, device='cuda')
        for _ in range(3): head.bg_tc(feats, lg)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
        e0.record()
        for _ in range(10): head.bg_tc(feats, lg)
        e1.record(); torch.cuda.synchronize()
        outs[small], times[small] = lg[:, 0].clone(), e0.elapsed_time(e1) / 10 / B * 1e3
    d = (outs['0'] - outs['1']).abs().max().item()
    gb = C * h * w * 2 / (times['1'] * 1e-6) / 1e9
    print(f"C={C} B={B} {h}x{w}: streaming {times['0']:.2f} us/tile  resident {times['1']:.2f} us/tile ({gb:.0f} GB/s)  "
          f"max|diff| {d:.2e} of {outs['0'].abs().max().item():.3f}", flush=True)
