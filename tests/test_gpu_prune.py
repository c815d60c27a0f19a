"""GPU: the per-cell class-pruning kernel behind sl_upsample_argmax's prediction-only path (post_prune.cu) must
return EXACTLY what evaluating all K classes at every pixel returns (the row-cached kernel, SL_POST_PRUNE=0), on
ordinary, adversarial (exact ties, one-ulp differences, huge / tiny magnitudes) and non-finite inputs, and its fused
confusion counts must equal get_confusion_matrix (utils/pyt_utils.py:182-200) on the same maps."""
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


def both(ops, lg, size, label=None):
    """(pruned, row-cached) outputs for the same input."""
    K = lg.shape[1]
    res = []
    try:
        for prune in (1, 0):
            _cabi.set_env(SL_POST_PRUNE=prune)
            cm = torch.zeros(K, K, dtype=torch.int64, device='cuda') if label is not None else None
            out = ops.upsample_argmax(lg, size, label=label, cm=cm)
            res.append((out['pred'].clone(), None if cm is None else cm.clone()))
    finally:
        _cabi.set_env(SL_POST_PRUNE=None)
    return res


def smooth_logits(B, K, h, w, seed, coarse=8, noise=0.3):
    g = torch.Generator().manual_seed(seed)
    lab = torch.randint(0, K, (B, coarse, coarse), generator=g)
    lab = F.interpolate(lab[:, None].float(), size=(h, w), mode='nearest')[:, 0].long()
    lg = noise * torch.randn(B, K, h, w, generator=g)
    lg.scatter_add_(1, lab[:, None], torch.full((B, 1, h, w), 4.0))
    return lg


@pytest.mark.parametrize('K,h,w,H,W', [
    (8, 128, 128, 1024, 1024),      # configs[1]: x8, the exact-K instantiation
    (12, 256, 256, 1024, 1024),     # configs[3]: x4
    (12, 128, 128, 1024, 1024),
    (5, 64, 64, 512, 512),          # run-time K
    (8, 125, 120, 1000, 960),       # band / strip boundaries that are not powers of two
    (12, 60, 33, 300, 132),         # x5 / x4, W % 4 == 0 only just, ragged CTA
    (3, 17, 9, 90, 36),
    (31, 32, 32, 128, 128),         # the largest class count the kernel takes
])
def test_pruned_equals_full_evaluation(ops, K, h, w, H, W):
    for seed, maker in ((1, lambda: smooth_logits(2, K, h, w, 1)),                    # homogeneous regions + edges
                        (2, lambda: torch.randn(2, K, h, w, generator=torch.Generator().manual_seed(2)))):   # all hard
        lg = maker().cuda()
        g = torch.Generator().manual_seed(seed)
        label = torch.randint(0, K, (2, H, W), generator=g).to(torch.uint8)
        label[torch.rand(2, H, W, generator=g) < 0.02] = 255
        (p1, c1), (p0, c0) = both(ops, lg, (H, W), label.cuda())
        assert torch.equal(p1, p0), f'{(p1 != p0).sum().item()} pixels differ'
        assert torch.equal(c1, c0)
        cm_ref = sum(ref_ops.ref_confusion(label[i].numpy(), p1[i].cpu().numpy(), K) for i in range(2))
        assert np.array_equal(c1.cpu().numpy().astype(np.float64), cm_ref)


def test_pruned_adversarial_ties_and_ulps(ops):
    K, h, w, H, W = 8, 32, 32, 256, 256
    g = torch.Generator().manual_seed(5)
    base = torch.randn(1, 1, h, w, generator=g)
    cases = {}
    cases['all equal'] = base.expand(1, K, h, w).contiguous()
    cases['constant'] = torch.full((1, K, h, w), 0.25)
    up = torch.nextafter(base, torch.full_like(base, 10.0))
    dn = torch.nextafter(base, torch.full_like(base, -10.0))
    cases['higher index one ulp above'] = torch.cat([base] * 4 + [up] * 4, 1)      # strict-margin rule must not prune
    cases['lower index one ulp above'] = torch.cat([up] * 4 + [base] * 4, 1)
    cases['alternating ulps'] = torch.cat([base, up, dn, up, base, dn, up, base], 1)
    big = base * 1e30
    cases['huge magnitudes'] = torch.cat([big, big * 1.000001, big, -big, big * 0.5, big, big * 1.000001, big], 1)
    cases['beyond the overflow guard'] = torch.cat([base * 3e37] * 8, 1) + torch.randn(1, K, h, w, generator=g) * 1e36
    tiny = base * 1e-38
    cases['denormal range'] = torch.cat([tiny, tiny * 1.5, tiny, tiny * 0.5, -tiny, tiny, tiny * 1.5, tiny], 1)
    zeros = torch.zeros(1, K, h, w)
    zeros[:, 3] = -0.0
    cases['signed zeros'] = zeros
    near = smooth_logits(1, K, h, w, 9, noise=1e-7)
    near[:, 5] = near[:, 2]                                                        # exact duplicates of a winning class
    cases['duplicate classes'] = near
    for name, lg in cases.items():
        (p1, _), (p0, _) = both(ops, lg.cuda(), (H, W))
        assert torch.equal(p1, p0), f'{name}: {(p1 != p0).sum().item()} pixels differ'
    # first maximum on exact ties
    (p1, _), _ = both(ops, cases['all equal'].cuda(), (H, W))
    assert int(p1.max()) == 0


def test_pruned_non_finite(ops):
    K, h, w, H, W = 12, 32, 32, 128, 128
    g = torch.Generator().manual_seed(6)
    lg = smooth_logits(2, K, h, w, 6)
    idx = torch.randint(0, lg.numel(), (200,), generator=g)
    flat = lg.view(-1)
    flat[idx[:80]] = float('nan')
    flat[idx[80:140]] = float('inf')
    flat[idx[140:]] = float('-inf')
    (p1, _), (p0, _) = both(ops, lg.cuda(), (H, W))
    assert torch.equal(p1, p0)
    # identity-sized check against np.argmax itself (NaN wins, first maximum)
    lg2 = lg[:, :, :16, :16].contiguous()
    up = F.interpolate(lg2, size=(64, 64), mode='bilinear', align_corners=True)
    (q1, _), (q0, _) = both(ops, lg2.cuda(), (64, 64))
    assert torch.equal(q1, q0)
    ref = np.argmax(up.numpy(), axis=1)
    finite = np.isfinite(up.numpy()).all(axis=1)
    assert (q1.cpu().numpy()[finite] == ref[finite]).mean() >= 0.999


def test_pruned_bench_shape_matches_oracle(ops):
    """configs[1] geometry on a trained-like head: pruned prediction == full evaluation, fused confusion == oracle."""
    st = synth.make_trained_like_state(512, 7, 0, seed=1234)
    labels = synth.make_labels(2, 1024, 1024, st.n_classes, seed=77)
    feats = synth.make_features(labels, st, 8, seed=77)
    head = ops.PopHead(st.base_emb, st.cls, None, None)
    lg = head(feats.cuda())
    (p1, c1), (p0, c0) = both(ops, lg, (1024, 1024), labels.cuda())
    assert torch.equal(p1, p0) and torch.equal(c1, c0)
    cm_ref = sum(ref_ops.ref_confusion(labels[i].numpy(), p1[i].cpu().numpy(), st.n_classes) for i in range(2))
    assert np.array_equal(c1.cpu().numpy().astype(np.float64), cm_ref)
