"""Compare the cta_group::2 background kernel (SL_TC_PAIR=1) with the single-CTA one: parity + time."""
import os, sys, torch
sys.path.insert(0, '.')
from segland_b200 import ops, synth

def run(C, B, h, w, prec):
    st = synth.make_trained_like_state(C, Kn=4) if C >= 128 else synth.make_head_state(C, Kn=4)
    head = ops.PopHead(st.base_emb, st.cls, st.novel_emb, st.cls_n, device='cuda', bg_mode='tc', tc_precision=prec)
    feats = synth.make_random_features(B, C, h, w).cuda()
    outs, times = {}, {}
    for pair in ('0', '1'):
        os.environ['SL_TC_PAIR'] = pair; __import__('segland_b200._cabi', fromlist=['x']).lib().sl_env_reload()
        lg = torch.zeros(B, 12, h, w, device='cuda')
        head.bg_tc(feats, lg)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
        e0.record()
        for _ in range(20):
            head.bg_tc(feats, lg)
        e1.record(); torch.cuda.synchronize()
        outs[pair], times[pair] = lg[:, 0].clone(), e0.elapsed_time(e1) / 20
    d = (outs['0'] - outs['1']).abs().max().item()
    print(f"C={C} B={B} {h}x{w} {prec}: single {times['0']:.4f} ms  pair {times['1']:.4f} ms  max|diff| {d:.3e} "
          f"ref max {outs['0'].abs().max().item():.3f}", flush=True)

for cfg in [(512, 32, 128, 128, 'precise'), (512, 32, 128, 128, 'balanced'), (512, 1, 128, 128, 'precise'),
            (512, 3, 24, 16, 'precise'), (96, 8, 256, 256, 'precise'), (480, 4, 64, 64, 'precise'),
            (192, 5, 40, 48, 'precise'), (32, 2, 16, 24, 'precise')]:
    run(*cfg)
