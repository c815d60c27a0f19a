"""GPU: the register-resident up-sample/arg-max kernel (post_regs.cu, the default prediction-only path at K = 8 / 12
when up-sampling by >= 2x) must return EXACTLY what the row-cached kernel (SL_POST_REGS=0) returns -- predictions and
fused confusion counts -- on ordinary, adversarial and non-finite inputs and on ragged geometries, and its counts must
equal get_confusion_matrix (utils/pyt_utils.py:182-200) on the same maps.  eval_base.py:168-178, eval_ft.py:168-183."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

from oracle import ref_ops
from segland_b200 import _cabi, synth

pytestmark = pytest.mark.gpu


@pytest.fixture(scope='module')
def ops():
    from segland_b200 import ops as _ops
    _ops.check_device()
    return _ops


def both(ops, lg, size, label=None, want_pred=True):
    """(register kernel, row-cached kernel) outputs for the same input."""
    K = lg.shape[1]
    res = []
    try:
        for regs in (1, 0):
            _cabi.set_env(SL_POST_REGS=regs)
            cm = torch.zeros(K, K, dtype=torch.int64, device='cuda') if label is not None else None
            out = ops.upsample_argmax(lg, size, label=label, cm=cm, want_pred=want_pred)
            res.append((out['pred'].clone() if want_pred else None, None if cm is None else cm.clone()))
    finally:
        _cabi.set_env(SL_POST_REGS=None)
    return res


def smooth_logits(B, K, h, w, seed, coarse=8, noise=0.3):
    g = torch.Generator().manual_seed(seed)
    lab = torch.randint(0, K, (B, coarse, coarse), generator=g)
    lab = F.interpolate(lab[:, None].float(), size=(h, w), mode='nearest')[:, 0].long()
    lg = noise * torch.randn(B, K, h, w, generator=g)
    lg.scatter_add_(1, lab[:, None], torch.full((B, 1, h, w), 4.0))
    return lg


@pytest.mark.parametrize('K,h,w,H,W', [
    (8, 128, 128, 1024, 1024),      # configs[1]: x8
    (12, 256, 256, 1024, 1024),     # configs[3]: x4
    (12, 128, 128, 1024, 1024),     # PSPNet ft
    (8, 125, 120, 1000, 960),       # bands and CTAs that end inside the image
    (8, 64, 300, 130, 2052),        # more than one CTA per row, ragged last CTA, x2.03 vertically
    (12, 60, 33, 300, 132),
    (8, 9, 7, 18, 16),              # exactly x2 / just above
    (8, 1, 1, 8, 8),                # a single source pixel (scale 0 -> falls back to the row-cached kernel)
    (12, 2, 2, 64, 64),
])
def test_regs_equals_row_cached(ops, K, h, w, H, W):
    for seed, maker in ((1, lambda: smooth_logits(2, K, h, w, 1)),
                        (2, lambda: torch.randn(2, K, h, w, generator=torch.Generator().manual_seed(2)))):
        lg = maker().cuda()
        g = torch.Generator().manual_seed(seed)
        label = torch.randint(0, K, (2, H, W), generator=g).to(torch.uint8)
        label[torch.rand(2, H, W, generator=g) < 0.02] = 255
        (p1, c1), (p0, c0) = both(ops, lg, (H, W), label.cuda())
        assert torch.equal(p1, p0), f'{(p1 != p0).sum().item()} pixels differ'
        assert torch.equal(c1, c0)
        cm_ref = sum(ref_ops.ref_confusion(label[i].numpy(), p1[i].cpu().numpy(), K) for i in range(2))
        assert np.array_equal(c1.cpu().numpy().astype(np.float64), cm_ref)
        # counting without a prediction buffer
        (_, c1n), (_, c0n) = both(ops, lg, (H, W), label.cuda(), want_pred=False)
        assert torch.equal(c1n, c0n) and torch.equal(c1n, c1)
        # against the oracle's F.interpolate -> np.argmax (disagreements only at near-ties)
        ref = ref_ops.ref_upsample_argmax(lg.cpu(), (H, W)) if hasattr(ref_ops, 'ref_upsample_argmax') else None
        if ref is not None:
            assert (p1.cpu().numpy() == np.asarray(ref)).mean() >= 0.9995


def test_regs_adversarial_ties_and_ulps(ops):
    K, h, w, H, W = 8, 32, 32, 256, 256
    g = torch.Generator().manual_seed(5)
    base = torch.randn(1, 1, h, w, generator=g)
    cases = {}
    cases['all equal'] = base.expand(1, K, h, w).contiguous()
    cases['constant'] = torch.full((1, K, h, w), 0.25)
    up = torch.nextafter(base, torch.full_like(base, 10.0))
    dn = torch.nextafter(base, torch.full_like(base, -10.0))
    cases['higher index one ulp above'] = torch.cat([base] * 4 + [up] * 4, 1)
    cases['lower index one ulp above'] = torch.cat([up] * 4 + [base] * 4, 1)
    cases['alternating ulps'] = torch.cat([base, up, dn, up, base, dn, up, base], 1)
    big = base * 1e30
    cases['huge magnitudes'] = torch.cat([big, big * 1.000001, big, -big, big * 0.5, big, big * 1.000001, big], 1)
    tiny = base * 1e-38
    cases['denormal range'] = torch.cat([tiny, tiny * 1.5, tiny, tiny * 0.5, -tiny, tiny, tiny * 1.5, tiny], 1)
    zeros = torch.zeros(1, K, h, w)
    zeros[:, 3] = -0.0
    cases['signed zeros'] = zeros
    near = smooth_logits(1, K, h, w, 9, noise=1e-7)
    near[:, 5] = near[:, 2]
    cases['duplicate classes'] = near
    for name, lg in cases.items():
        (p1, _), (p0, _) = both(ops, lg.cuda(), (H, W))
        assert torch.equal(p1, p0), f'{name}: {(p1 != p0).sum().item()} pixels differ'
    (p1, _), _ = both(ops, cases['all equal'].cuda(), (H, W))
    assert int(p1.max()) == 0                                    # first maximum on exact ties
    (p1, _), _ = both(ops, cases['duplicate classes'].cuda(), (H, W))
    assert int((p1 == 5).sum()) == 0                             # the duplicate with the higher index never wins


def test_regs_non_finite(ops):
    for K in (8, 12):
        h, w, H, W = 32, 32, 128, 128
        g = torch.Generator().manual_seed(6)
        lg = smooth_logits(2, K, h, w, 6)
        idx = torch.randint(0, lg.numel(), (200,), generator=g)
        flat = lg.view(-1)
        flat[idx[:80]] = float('nan')
        flat[idx[80:140]] = float('inf')
        flat[idx[140:]] = float('-inf')
        (p1, _), (p0, _) = both(ops, lg.cuda(), (H, W))
        assert torch.equal(p1, p0)
        lg2 = lg[:, :, :16, :16].contiguous()
        up = F.interpolate(lg2, size=(64, 64), mode='bilinear', align_corners=True)
        (q1, _), (q0, _) = both(ops, lg2.cuda(), (64, 64))
        assert torch.equal(q1, q0)
        ref = np.argmax(up.numpy(), axis=1)
        finite = np.isfinite(up.numpy()).all(axis=1)
        assert (q1.cpu().numpy()[finite] == ref[finite]).mean() >= 0.999


def test_regs_is_the_default_at_the_bench_shape_and_matches_oracle(ops):
    """configs[1] geometry, 8 tiles (the size from which the library picks this kernel on its own)."""
    st = synth.make_trained_like_state(512, 7, 0, seed=1234)
    labels = synth.make_labels(8, 1024, 1024, st.n_classes, seed=78)
    feats = synth.make_features(labels, st, 8, seed=78)
    head = ops.PopHead(st.base_emb, st.cls, None, None)
    lg = head(feats.cuda())
    K = st.n_classes
    outs = []
    try:
        for regs in (None, 0):                                   # library default, then the row-cached kernel
            _cabi.set_env(SL_POST_REGS=regs)
            cm = torch.zeros(K, K, dtype=torch.int64, device='cuda')
            outs.append((ops.upsample_argmax(lg, (1024, 1024), label=labels.cuda(), cm=cm)['pred'].clone(), cm))
    finally:
        _cabi.set_env(SL_POST_REGS=None)
    (p1, c1), (p0, c0) = outs
    assert torch.equal(p1, p0) and torch.equal(c1, c0)
    cm_ref = sum(ref_ops.ref_confusion(labels[i].numpy(), p1[i].cpu().numpy(), K) for i in range(8))
    assert np.array_equal(c1.cpu().numpy().astype(np.float64), cm_ref)
    up = F.interpolate(lg.cpu(), size=(1024, 1024), mode='bilinear', align_corners=True)
    ref = np.argmax(up.numpy(), axis=1)
    assert (p1.cpu().numpy() == ref).mean() >= 0.9999


def soft_both(ops, lg, size, label=None):
    K = lg.shape[1]
    res = []
    try:
        for regs in (1, 0):                                      # register kernel (K = 12) / row-cached kernel
            _cabi.set_env(SL_POST_REGS=regs)
            cm = torch.zeros(K, K, dtype=torch.int64, device='cuda') if label is not None else None
            out = ops.upsample_argmax(lg, size, label=label, cm=cm, want_conf=True, want_probs=True)
            res.append({k: v.clone() for k, v in out.items()} | ({'cm': cm.clone()} if cm is not None else {}))
    finally:
        _cabi.set_env(SL_POST_REGS=None)
    return res


@pytest.mark.parametrize('K,h,w,H,W', [
    (12, 256, 256, 1024, 1024),     # configs[3]: ConvNeXt / Swin ft, x4, probability maps
    (8, 128, 128, 1024, 1024),
    (12, 60, 33, 300, 132),
    (8, 125, 120, 1000, 960),
    (12, 9, 7, 18, 16),
])
def test_regs_soft_outputs_equal_row_cached(ops, K, h, w, H, W):
    """Probability maps and confidences of the register kernel == the row-cached soft path bit for bit (same ex2 / rcp
    formulas, same order), and == torch softmax of F.interpolate within the approximation error (spec: this repo)."""
    for seed, lg in ((1, smooth_logits(2, K, h, w, 1)), (2, torch.randn(2, K, h, w, generator=torch.Generator().manual_seed(2)) * 3)):
        g = torch.Generator().manual_seed(seed)
        label = torch.randint(0, K, (2, H, W), generator=g).to(torch.uint8)
        a, b = soft_both(ops, lg.cuda(), (H, W), label.cuda())
        for k in ('pred', 'conf', 'probs', 'cm'):
            assert torch.equal(a[k], b[k]), k
        up = F.interpolate(lg, size=(H, W), mode='bilinear', align_corners=True)
        ref = torch.softmax(up, dim=1)
        assert (a['probs'].cpu() - ref).abs().max().item() <= 2e-6
        assert (a['conf'].cpu() - ref.max(dim=1).values).abs().max().item() <= 2e-6


def test_regs_soft_non_finite(ops):
    K, h, w, H, W = 12, 32, 32, 128, 128
    g = torch.Generator().manual_seed(7)
    lg = smooth_logits(2, K, h, w, 7)
    idx = torch.randint(0, lg.numel(), (100,), generator=g)
    lg.view(-1)[idx[:50]] = float('nan')
    lg.view(-1)[idx[50:]] = float('inf')
    a, b = soft_both(ops, lg.cuda(), (H, W))
    assert torch.equal(a['pred'], b['pred'])
    fin = torch.isfinite(b['probs']).all(dim=1) & torch.isfinite(a['probs']).all(dim=1)
    assert fin.float().mean().item() > 0.5
    pa, pb = a['probs'].permute(0, 2, 3, 1)[fin], b['probs'].permute(0, 2, 3, 1)[fin]
    assert (pa - pb).abs().max().item() <= 1e-6
    assert torch.equal(torch.isfinite(a['probs']), torch.isfinite(b['probs']))
