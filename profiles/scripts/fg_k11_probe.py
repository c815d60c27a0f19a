import sys, torch
sys.path.insert(0, '/root/repo')
from segland_b200 import ops, synth
st = synth.make_head_state(512, 7, 4, seed=2)
f = torch.randn(32, 512, 128, 128, device='cuda').to(torch.bfloat16)
head = ops.PopHead(st.base_emb, st.cls, st.novel_emb, st.cls_n)
lg = torch.empty(32, 12, 128, 128, device='cuda')
for _ in range(5):
    head(f, out=lg, fg_only=True)
torch.cuda.synchronize()
