"""One training-mode head forward+backward per shape (for ncu launch lists)."""
import sys, torch
sys.path.insert(0, '.')
from segland_b200 import ops, synth
for C, B, h, w in [(512, 2, 128, 128), (96, 2, 256, 256)]:
    st = synth.make_head_state(C, 7, 4, seed=5).to('cuda')
    feats = synth.make_random_features(B, C, h, w, seed=5).cuda()
    g = torch.randn(B, 12, h, w, device='cuda')
    for need in (False, True):
        for it in range(3):
            novel = st.novel_emb.clone().requires_grad_(True)
            cls_n = tuple(t.clone().requires_grad_(True) for t in st.cls_n)
            f = feats.float().requires_grad_(need)
            ops.pop_head_train(f, st.base_emb, st.cls, novel, cls_n).backward(g)
    torch.cuda.synchronize()
