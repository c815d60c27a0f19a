"""Tensor-core (split-bf16) vs exact-fp32 backward of the POP head: per-gradient relative error."""
import sys, torch
sys.path.insert(0, '.')
from segland_b200 import ops, synth

def grads(C, Kb, Kn, B, h, w, mode, seed=1):
    st = synth.make_head_state(C, Kb, Kn, seed=seed).to('cuda')
    feats = synth.make_random_features(B, C, h, w, seed=seed).cuda().float().requires_grad_(True)
    g = torch.randn(B, 1 + Kb + Kn, h, w, device='cuda', generator=torch.Generator('cuda').manual_seed(seed))
    novel = st.novel_emb.clone().requires_grad_(True)
    cls_n = tuple(t.clone().requires_grad_(True) for t in st.cls_n)
    cls = tuple(t.clone().requires_grad_(True) for t in st.cls)
    out = ops.pop_head_train(feats, st.base_emb, cls, novel, cls_n, bg_mode=mode)
    out.backward(g)
    return dict(novel=novel.grad, W1n=cls_n[0].grad, W2n=cls_n[1].grad, w3n=cls_n[2].grad, W1=cls[0].grad, feat=feats.grad)

for C, B, h, w in [(64, 2, 16, 16), (192, 2, 32, 32), (512, 2, 32, 32), (512, 2, 128, 128), (96, 2, 256, 256), (480, 1, 24, 40)]:
    a, b = grads(C, 7, 4, B, h, w, 'auto'), grads(C, 7, 4, B, h, w, 'simt')
    msg = []
    for k in a:
        err = (a[k] - b[k]).abs().double()
        ref = b[k].double()
        rms = ref.pow(2).mean().sqrt()
        relmax = (err.max() / ref.abs().max()).item()
        l2 = (err.pow(2).sum().sqrt() / ref.pow(2).sum().sqrt()).item()
        frac = (err > 1e-3 * ref.abs() + 1e-3 * rms).double().mean().item()
        msg.append(f'{k}: max {relmax:.1e} L2 {l2:.1e} out-of-bound {frac:.1e}')
    print(f'C={C} B={B} {h}x{w}: ' + ' | '.join(msg), flush=True)
