"""Seeded synthetic OEM-shaped inputs.  Torch-CPU only; used by the
tests, bench.py and the golden-vector generator so every consumer sees the same bytes.

OEM constants: 7 base + 4 novel classes (dataset/oem.py:32-34), ignore label 255
(dataset/oem.py:15), 1024x1024 tiles (scripts/evaluate_oem.sh:17).
"""
from __future__ import annotations

import math
from dataclasses import dataclass

import torch
import torch.nn.functional as F

KB_OEM, KN_OEM, IGNORE_LABEL = 7, 4, 255


@dataclass
class HeadState:
    """The state-dict entries the POP head reads, as plain fp32 tensors."""
    base_emb: torch.Tensor                 # [Kb,C]
    novel_emb: torch.Tensor | None         # [Kn,C] or None (base mode)
    cls: tuple                             # classifier   (W1 [C,C], W2 [C,C], w3 [C])
    cls_n: tuple | None                    # classifier_n (same shapes) or None

    @property
    def n_classes(self):
        return 1 + self.base_emb.shape[0] + (0 if self.novel_emb is None else self.novel_emb.shape[0])

    def to(self, device):
        mv = lambda t: None if t is None else t.to(device)
        return HeadState(mv(self.base_emb), mv(self.novel_emb), tuple(mv(t) for t in self.cls),
                         None if self.cls_n is None else tuple(mv(t) for t in self.cls_n))


def _conv1x1_weight(gen, c_out, c_in):
    # nn.Conv2d default init = kaiming_uniform_(a=sqrt(5)) -> U(-1/sqrt(fan_in), 1/sqrt(fan_in))
    bound = 1.0 / math.sqrt(c_in)
    return (torch.rand(c_out, c_in, generator=gen) * 2 - 1) * bound


def _orthogonal(gen, rows, cols):
    # nn.init.orthogonal_ (pspnet_pop.py:64,68) restated with an explicit generator
    a = torch.randn(cols, rows, generator=gen)
    q, r = torch.linalg.qr(a)
    q = q * torch.sign(torch.diagonal(r)).unsqueeze(0)
    return q.t().contiguous()


def make_head_state(C, Kb=KB_OEM, Kn=0, seed=1234, proto_scale=1.0):
    gen = torch.Generator().manual_seed(seed)
    base = _orthogonal(gen, Kb, C) * proto_scale
    cls = (_conv1x1_weight(gen, C, C), _conv1x1_weight(gen, C, C), _conv1x1_weight(gen, 1, C).reshape(C))
    if Kn == 0:
        return HeadState(base, None, cls, None)
    novel = _orthogonal(gen, Kn, C) * proto_scale
    # classifier_n starts as a copy of classifier (init_cls_n, pspnet_pop.py:76-78); perturb so
    # the two MLPs are distinguishable in tests.
    cls_n = tuple(t + 0.01 * torch.randn(t.shape, generator=gen) * t.abs().mean() * 10 for t in cls)
    return HeadState(base, novel, cls, cls_n)


def make_trained_like_state(C, Kb=KB_OEM, Kn=0, seed=1234, n_bg_units=64, bg_gain=5.0, noise=0.02, base=None, novel=None):
    """A head whose argmax is meaningful on make_features() tiles, so a sweep has a non-trivial mIoU
    (random-init MLPs give alpha_k of random sign and mIoU ~ chance).  Construction: hidden units
    0..K-1 of layer 1 are the normalised prototypes, so a foreground vector p*s_k lights exactly unit
    k (alpha_k ~ 1, beta_k ~ 0); n_bg_units further units are random directions orthogonal to every
    prototype and see only the background residual, giving class 0 a logit ~ 0.4*bg_gain on pure-noise
    pixels; layer 2 passes both groups through and w3 sums them.  Everything gets `noise` x default-init
    jitter so no weight is exactly zero.  In ft mode classifier_n is built the same way (classes Kb..K-1
    and the background) and classifier handles the base classes."""
    gen = torch.Generator().manual_seed(seed)
    base = _orthogonal(gen, Kb, C) if base is None else base.clone()       # given prototypes: any [Kb,C] / [Kn,C] rows
    novel = (_orthogonal(gen, Kn, C) if novel is None else novel.clone()) if Kn else None
    protos = base if novel is None else torch.cat([base, novel], 0)
    K = protos.shape[0]
    s_hat = F.normalize(protos, p=2, dim=-1)
    r = torch.randn(n_bg_units, C, generator=gen)
    r = r - (r @ s_hat.t()) @ s_hat                                  # orthogonal to every prototype
    r = F.normalize(r, p=2, dim=-1)

    def build():
        W1 = noise * _conv1x1_weight(gen, C, C)
        W2 = noise * _conv1x1_weight(gen, C, C)
        w3 = noise * _conv1x1_weight(gen, 1, C).reshape(C)
        W1[:K] += s_hat
        W1[K:K + n_bg_units] += r
        idx = torch.arange(K + n_bg_units)
        W2[idx, idx] += 1.0
        w3[:K] += 1.0
        w3[K:K + n_bg_units] += bg_gain / n_bg_units
        return (W1, W2, w3)

    assert K + n_bg_units <= C
    if Kn == 0:
        return HeadState(base, None, build(), None)
    return HeadState(base, novel, build(), build())


def make_labels(T, H, W, n_classes, seed=1234, coarse=32, ignore_frac=0.01):
    """uint8 [T,H,W]: nearest-upsampled coarse random class field, ~1% pixels = 255."""
    gen = torch.Generator().manual_seed(seed + 1)
    field = torch.randint(0, n_classes, (T, 1, coarse, coarse), generator=gen).float()
    lab = F.interpolate(field, size=(H, W), mode='nearest').squeeze(1).to(torch.uint8)
    if ignore_frac > 0:
        lab[torch.rand(T, H, W, generator=gen) < ignore_frac] = IGNORE_LABEL
    return lab


def make_features(labels, state: HeadState, stride, seed=1234, signal=4.0, dtype=torch.bfloat16):
    """bf16 [T,C,h,w] = signal * s_hat[label at stride] + randn (class 0 / ignore: noise only)."""
    gen = torch.Generator().manual_seed(seed + 2)
    T, H, W = labels.shape
    h, w = H // stride, W // stride
    protos = state.base_emb if state.novel_emb is None else torch.cat([state.base_emb, state.novel_emb], 0)
    s_hat = F.normalize(protos, p=2, dim=-1)                      # [K,C]
    C = s_hat.shape[1]
    lab_lr = labels[:, stride // 2::stride, stride // 2::stride][:, :h, :w].long()
    table = torch.cat([torch.zeros(1, C), s_hat], 0)              # class 0 -> zeros
    idx = torch.where((lab_lr >= table.shape[0]), torch.zeros_like(lab_lr), lab_lr)
    feats = torch.empty(T, C, h, w, dtype=dtype)
    for t in range(T):                                            # per tile: bounds peak memory
        f = torch.randn(C, h, w, generator=gen)
        f += signal * table[idx[t]].permute(2, 0, 1)
        feats[t] = f.to(dtype)
    return feats


def make_random_features(T, C, h, w, seed=1234, dtype=torch.bfloat16):
    """Pure randn features: the worst case for argmax near-ties."""
    gen = torch.Generator().manual_seed(seed + 3)
    return torch.randn(T, C, h, w, generator=gen).to(dtype)


def make_support_masks(T, H, W, seed=1234, coarse=16):
    """fp32 {0,1} masks [T,1,H,W] for masked-average pooling (5-shot support sets)."""
    gen = torch.Generator().manual_seed(seed + 4)
    field = (torch.rand(T, 1, coarse, coarse, generator=gen) < 0.3).float()
    return F.interpolate(field, size=(H, W), mode='nearest')


def make_logit_stacks(M, K, H, W, seed=1234):
    """M fp32 logit stacks [K,H,W] for fusemat-style fusion (one tile)."""
    gen = torch.Generator().manual_seed(seed + 5)
    return [torch.randn(K, H, W, generator=gen) for _ in range(M)]
