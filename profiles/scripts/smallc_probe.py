"""Where does the small-C background kernel spend its time?  SL_TC_DEBUG knock-outs (results invalid)."""
import os, sys, torch
sys.path.insert(0, '.')
from segland_b200 import ops, synth
for C in (96, 192):
    st = synth.make_head_state(C, 7, 4, seed=2)
    head = ops.PopHead(st.base_emb, st.cls, st.novel_emb, st.cls_n, bg_mode='tc')
    feats = synth.make_random_features(16, C, 256, 256, seed=1).cuda()
    lg = torch.zeros(16, 12, 256, 256, device='cuda')
    for pair in ('1', '0'):
        for dbg in ('0', '1', '2', '3', '4', '7'):
            if pair == '1' and dbg != '0':
                continue
            os.environ['SL_TC_PAIR'] = pair; os.environ['SL_TC_DEBUG'] = dbg; __import__('segland_b200._cabi', fromlist=['x']).lib().sl_env_reload()
            for _ in range(3): head.bg_tc(feats, lg)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
            e0.record()
            for _ in range(10): head.bg_tc(feats, lg)
            e1.record(); torch.cuda.synchronize()
            print(f'C={C} pair={pair} debug={dbg}: {e0.elapsed_time(e1) / 10 / 16 * 1e3:.2f} us/tile', flush=True)
