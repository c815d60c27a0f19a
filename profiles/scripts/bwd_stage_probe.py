"""Compare the intermediate activations (h1, dz2, dz1) of the tensor-core backward with the exact path."""
import sys, torch
import torch.nn.functional as F
sys.path.insert(0, '.')
from segland_b200 import ops, synth, _cabi
from segland_b200._cabi import call, ptr, int_array

def run(C, B, h, w, mode, K=11, seed=1):
    st = synth.make_head_state(C, 7, 4, seed=seed).to('cuda')
    feats = synth.make_random_features(B, C, h, w, seed=seed).cuda().contiguous()
    N = h * w
    g = torch.randn(B, 12, h, w, device='cuda', generator=torch.Generator('cuda').manual_seed(seed))
    s_hat = F.normalize(torch.cat([st.base_emb, st.novel_emb]), dim=-1).contiguous()
    alpha = torch.linspace(0.1, 1, K, device='cuda'); beta = torch.linspace(1, 0.2, K, device='cuda')
    W1, W2, w3 = st.cls_n
    W1p = (W1 - (W1 @ s_hat.t()) @ s_hat).contiguous()
    new = lambda *s: torch.zeros(*s, device='cuda')
    outs = [new(K, C), new(K), new(K), new(C, C), new(C, C), new(C)]
    d_feat = new(B, C, h, w)
    ws = torch.zeros(_cabi.lib().sl_pop_head_bwd_ws_bytes(B, C, N, K) // 4, device='cuda')
    call('sl_pop_head_bwd', ptr(feats), B, C, N, ptr(s_hat), ptr(alpha), ptr(beta), K, int_array(list(range(1, 12))),
         ptr(W1p), ptr(W2.contiguous()), ptr(w3.contiguous()), ptr(g), 12, 0, *[ptr(o) for o in outs], ptr(d_feat),
         mode, ptr(ws), torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    px = B * N
    act = ws[2 * px * K: 2 * px * K + (7 * px * C + 1) // 2]
    if mode == 1:
        h1, dz2, dz1 = (act[i * px * C:(i + 1) * px * C].view(px, C) for i in range(3))
    else:
        a16 = act.view(torch.bfloat16)
        parts = [a16[i * px * C:(i + 1) * px * C].view(px, C).double() for i in range(7)]
        h1, dz2, dz1 = parts[0] + parts[1] + parts[6], parts[2] + parts[3], parts[4] + parts[5]
    q = feats.double().flatten(2).permute(0, 2, 1).reshape(px, C)
    z1 = q @ W1p.double().t()
    h1_ref = z1.clamp_min(0)
    e = (h1.double() - h1_ref).abs().max().item() / h1_ref.abs().max().item()
    flips1 = ((h1 > 0) != (z1 > 0)).sum().item()
    z2 = h1_ref @ W2.double().t()
    g0 = g[:, 0].reshape(px, 1).double()
    dz2_ref = torch.where(z2 > 0, g0 * w3.double().view(1, C), torch.zeros((), device='cuda', dtype=torch.float64))
    flips2 = ((dz2.double() - dz2_ref).abs() > 1e-3 * dz2_ref.abs().max()).sum().item()
    print(f'   mode {mode}: h1 vs float64 {e:.2e}, z1 sign flips {flips1}, dz2 mismatches vs float64 {flips2} of {px * C}')
    return dict(h1=h1.float().clone(), dz2=dz2.float().clone(), dz1=dz1.float().clone(), dW1p=outs[3], dW2=outs[4], dw3=outs[5], d_feat=d_feat)

for C, B, h, w in [(64, 2, 16, 16), (192, 2, 32, 32), (512, 2, 32, 32), (512, 2, 128, 128)]:
    a, b = run(C, B, h, w, 0), run(C, B, h, w, 1)
    msg = []
    for k in a:
        err = (a[k] - b[k]).abs()
        msg.append(f'{k} {err.max().item() / b[k].abs().max().item():.2e}')
        if k in ('dz2', 'dz1') and err.max() > 1e-3 * b[k].abs().max():
            bad = (err > 1e-3 * b[k].abs().max()).nonzero()
            msg.append(f'[bad {len(bad)} rows {len(bad[:, 0].unique())}]')
        if k == 'h1':
            # exact float64 reference of z1 for both
            pass
    print(f'C={C} B={B} {h}x{w}: ' + ' '.join(msg), flush=True)
