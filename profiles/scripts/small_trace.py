import os, sys, torch
sys.path.insert(0, '.')
from segland_b200 import ops, synth
C = int(sys.argv[1]) if len(sys.argv) > 1 else 96
st = synth.make_head_state(C, 7, 4, seed=2)
head = ops.PopHead(st.base_emb, st.cls, st.novel_emb, st.cls_n, bg_mode='tc')
feats = synth.make_random_features(16, C, 256, 256, seed=1).cuda()
lg = torch.zeros(16, 12, 256, 256, device='cuda')
for _ in range(2): head.bg_tc(feats, lg)
dbg = torch.zeros(64, 16, dtype=torch.int64, device='cuda')
os.environ['SL_SMALL_DBG'] = str(dbg.data_ptr()); __import__('segland_b200._cabi', fromlist=['x']).lib().sl_env_reload()
head.bg_tc(feats, lg); torch.cuda.synchronize()
d = dbg.cpu()
names = ['g1:start', 'g1:tempty1', 'g1:xfull', 'g2:tempty2', 'g2:h1', 'g2:issued', 'e1:tfull1', 'e1:converted', 'e1:tfull2', 'e1:stored', 'e2:tfull2', 'e2:done']
t0 = d[20, 0].item()
for s in range(20, 25):
    print(f'tile {s}: ' + '  '.join(f'{n}={(d[s, i].item() - t0)}' for i, n in enumerate(names)))
