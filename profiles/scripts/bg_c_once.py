"""One background-MLP launch at a given width (for ncu): python bg_c_once.py C hw tiles"""
import sys, torch
sys.path.insert(0, '.')
from segland_b200 import ops, synth
C, hw, T = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
st = synth.make_head_state(C, 7, 4, seed=2)
head = ops.PopHead(st.base_emb, st.cls, st.novel_emb, st.cls_n)
f = torch.randn(T, C, hw, hw, device='cuda').to(torch.bfloat16)
lg = torch.empty(T, head.n_classes, hw, hw, device='cuda')
for _ in range(3):
    head.bg_tc(f, lg)
torch.cuda.synchronize()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
for _ in range(20):
    head.bg_tc(f, lg)
b.record(); torch.cuda.synchronize()
ms = a.elapsed_time(b) / 20
fl = (4.0 * C * C + 2 * C) * hw * hw * T
print(f'C={C} hw={hw} T={T}: {ms:.4f} ms, {fl / ms / 1e9:.0f} TFLOP/s algorithmic, {fl * 2.5 / ms / 1e9:.0f} executed (5 passes)')
