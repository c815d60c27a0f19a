"""Shared test helpers: golden-fixture decoding and tolerance definitions."""
import numpy as np
import torch

from segland_b200.synth import HeadState

# north_star: "logits, probabilities and prototypes match within 1e-3 relative (bf16 features,
# fp32 accumulation)".  We test BOTH readings: max error relative to the tensor's max magnitude,
# and element-wise |a-b| <= RTOL*|b| + RTOL*rms(b).
RTOL = 1e-3


def bf16_from_bits(arr):
    return torch.from_numpy(np.ascontiguousarray(arr)).view(torch.bfloat16)


def state_from_npz(z, prefix=''):
    t = lambda k: torch.from_numpy(z[prefix + k].copy())
    cls = (t('W1'), t('W2'), t('w3'))
    if (prefix + 'novel_emb') in z.files:
        return HeadState(t('base_emb'), t('novel_emb'), cls, (t('W1n'), t('W2n'), t('w3n')))
    return HeadState(t('base_emb'), None, cls, None)


def rel_err(a, b):
    a = torch.as_tensor(a, dtype=torch.float64)
    b = torch.as_tensor(b, dtype=torch.float64)
    return ((a - b).abs().max() / b.abs().max().clamp_min(1e-30)).item()


def assert_close_rel(a, b, rtol=RTOL, what=''):
    a = torch.as_tensor(a, dtype=torch.float64).cpu()
    b = torch.as_tensor(b, dtype=torch.float64).cpu()
    assert a.shape == b.shape, f'{what}: shape {tuple(a.shape)} vs {tuple(b.shape)}'
    rms = b.pow(2).mean().sqrt()
    err = (a - b).abs()
    bound = rtol * b.abs() + rtol * rms
    worst = (err - bound).max().item()
    assert worst <= 0, f'{what}: elementwise bound exceeded by {worst:.3e}; rel-to-max {rel_err(a, b):.3e}'
    assert rel_err(a, b) <= rtol, f'{what}: rel-to-max {rel_err(a, b):.3e} > {rtol}'


def argmax_agreement(pred, ref_pred, logits_hr=None, tie_tol=1e-4):
    """Fraction of agreeing pixels; when upsampled logits are given, also checks every
    disagreement is a near-tie (top-2 gap <= tie_tol * max|logit|)."""
    pred = np.asarray(pred)
    ref_pred = np.asarray(ref_pred)
    agree = float((pred == ref_pred).mean())
    if logits_hr is not None and agree < 1.0:
        lg = np.asarray(logits_hr)                     # [B,K,H,W]
        bad = np.nonzero(pred != ref_pred)
        a = lg[bad[0], pred[bad].astype(np.int64), bad[1], bad[2]]
        b = lg[bad[0], ref_pred[bad].astype(np.int64), bad[1], bad[2]]
        gap = np.abs(a - b).max()
        assert gap <= tie_tol * np.abs(lg).max(), f'argmax disagreement is not a near-tie: gap {gap}'
    return agree
