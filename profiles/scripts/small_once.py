import sys, torch
sys.path.insert(0, '.')
from segland_b200 import ops, synth
C = int(sys.argv[1]) if len(sys.argv) > 1 else 96
st = synth.make_head_state(C, 7, 4, seed=2)
head = ops.PopHead(st.base_emb, st.cls, st.novel_emb, st.cls_n, bg_mode='tc')
feats = synth.make_random_features(16, C, 256, 256, seed=1).cuda()
lg = torch.zeros(16, 12, 256, 256, device='cuda')
for _ in range(3): head.bg_tc(feats, lg)
torch.cuda.synchronize()
