"""Time the training-mode head (forward + sl_pop_head_bwd) at the ft_pop shapes, per kernel group, and an
eager-PyTorch materialising formulation of the same maths on the same GPU for context."""
import sys, torch
import torch.nn.functional as F
sys.path.insert(0, '.')
from segland_b200 import ops, synth


def eager_head(f, base, novel, cls, cls_n):
    """[B,C,h,w] -> [B,1+Kb+Kn,h,w] the way networks/pspnet_pop.py:199-219 computes it (rank-1 tensors + 1x1 convs)."""
    B, C, h, w = f.shape
    q = f.flatten(2).float()
    s1, s2 = F.normalize(base, dim=-1), F.normalize(novel, dim=-1)
    fg1 = (s1 @ q).unsqueeze(2) * s1.unsqueeze(-1)
    fg2 = (s2 @ q).unsqueeze(2) * s2.unsqueeze(-1)
    bg = (q - fg1.sum(1) - fg2.sum(1)).unsqueeze(1)
    mlp = lambda x, ws: F.conv2d(F.relu(F.conv2d(F.relu(F.conv2d(x, ws[0].view(C, C, 1, 1))), ws[1].view(C, C, 1, 1))), ws[2].view(1, C, 1, 1))
    p1 = mlp(fg1.reshape(-1, C, h, w), cls).view(B, -1, h, w)
    p2 = mlp(torch.cat([bg, fg2], 1).reshape(-1, C, h, w), cls_n).view(B, -1, h, w)
    return torch.cat([p2[:, :1], p1, p2[:, 1:]], 1)


def timeit(fn, n=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


for name, C, B, h, w in [('pspnet_pop 1024^2 bs1', 512, 2, 128, 128), ('swin-s 1024^2 bs1', 96, 2, 256, 256),
                         ('convnext-t 1024^2 bs1', 192, 2, 256, 256)]:
    st = synth.make_head_state(C, 7, 4, seed=5).to('cuda')
    feats = synth.make_random_features(B, C, h, w, seed=5).cuda()
    g = torch.randn(B, 12, h, w, device='cuda')
    novel = st.novel_emb.clone().requires_grad_(True)
    cls_n = tuple(t.clone().requires_grad_(True) for t in st.cls_n)

    def ours(need_dfeat=False):
        f = feats.float().requires_grad_(need_dfeat)
        out = ops.pop_head_train(f, st.base_emb, st.cls, novel, cls_n)
        out.backward(g)

    def ours_fwd():
        with torch.no_grad():
            ops.pop_head_train(feats, st.base_emb, st.cls, novel, cls_n)

    def eager():
        out = eager_head(feats, st.base_emb, novel, st.cls, cls_n)
        out.backward(g)

    t_fwd, t_all, t_all_df = timeit(ours_fwd), timeit(ours), timeit(lambda: ours(True))
    try:
        t_eager = timeit(eager, 3)
    except torch.OutOfMemoryError:
        t_eager = float('nan')
    flops = 10 * C * C * B * h * w            # fwd 4C^2N + bwd 6C^2N
    print(f'{name}: ours fwd {t_fwd:.3f} ms, fwd+bwd {t_all:.3f} ms ({flops / t_all / 1e9:.1f} TFLOP/s fp32-equivalent), '
          f'+d_feat {t_all_df:.3f} ms; eager materialising fwd+bwd {t_eager:.2f} ms ({t_eager / t_all:.1f}x)', flush=True)
