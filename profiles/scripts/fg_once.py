import sys, torch
sys.path.insert(0, '.')
from segland_b200 import ops, synth
Kn = int(sys.argv[1]) if len(sys.argv) > 1 else 0
st = synth.make_head_state(512, 7, Kn, seed=2)
f = torch.randn(32, 512, 128, 128, device='cuda').to(torch.bfloat16)
head = ops.PopHead(st.base_emb, st.cls, st.novel_emb, st.cls_n, bg_mode='simt')
lg = torch.empty(32, head.n_classes, 128, 128, device='cuda')
for _ in range(4):
    head(f, out=lg, fg_only=True)
torch.cuda.synchronize()
