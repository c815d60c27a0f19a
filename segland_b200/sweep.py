"""Data-parallel tile sweeps: tiles are independent, so the path shards with
no data-path collective.  Each rank owns a contiguous range of tile indices (no
DistributedSampler padding, which would double-count tiles), accumulates its confusion matrix
on the device for the whole sweep, and takes part in ONE integer all-reduce at the end
(1,152 bytes for OEM's 12 classes).  The reference never reduces its eval confusion matrix
across ranks (eval_base.py:132-133,201), so distributed eval is new functionality here.
"""
from __future__ import annotations

import torch
import torch.distributed as dist

from . import ops


def shard_range(n_items, rank, world_size):
    """Contiguous, non-overlapping, exhaustive partition of range(n_items); sizes differ by <= 1."""
    if not (0 <= rank < world_size):
        raise ValueError('rank out of range')
    base, rem = divmod(n_items, world_size)
    start = rank * base + min(rank, rem)
    return range(start, start + base + (1 if rank < rem else 0))


def all_reduce_sum_(t, group=None):
    """In-place SUM all-reduce when a process group is up; identity otherwise.  Integer tensors
    stay integer, so confusion matrices reduce bit-exactly in any rank order."""
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
    return t


class TileEvaluator:
    """eval_base.py:162-199 / eval_ft.py:162-202 after the decoder, for a stream of tile batches:
    head -> align-corners upsample -> argmax -> confusion accumulation, all on the device."""

    def __init__(self, head: ops.PopHead, out_size, ignore_label=ops.IGNORE_LABEL):
        self.head = head
        self.out_size = (int(out_size[0]), int(out_size[1]))
        self.ignore_label = ignore_label
        K = head.n_classes
        self.cm = torch.zeros(K, K, dtype=torch.int64, device=head.device)
        self._logits = None

    def reset(self):
        self.cm.zero_()

    def step(self, features, labels=None, want_pred=True, **kw):
        """features [B,C,h,w] bf16 (device, or pinned host -> async H2D), labels [B,H,W] uint8.
        Returns the upsample_argmax dict (device tensors); no host synchronisation."""
        feats = ops._cuda(features, torch.bfloat16)
        B, _, h, w = feats.shape
        if self._logits is None or self._logits.shape[0] != B or self._logits.shape[-2:] != (h, w):
            self._logits = torch.empty(B, self.head.n_classes, h, w, dtype=torch.float32, device=feats.device)
        logits = self.head(feats, out=self._logits)
        return ops.upsample_argmax(logits, self.out_size, label=labels, cm=self.cm if labels is not None else None,
                                   ignore_label=self.ignore_label, want_pred=want_pred, **kw)

    def finalize(self, base_classes, group=None):
        """One all-reduce for the whole sweep, then the mIoU split of eval_ft.py:196-202.
        Returns (cm int64 [K,K] on the device, (base, novel, total, per-class) float64)."""
        all_reduce_sum_(self.cm, group)
        return self.cm, ops.miou_from_confusion(self.cm, base_classes)


def prototype_mean_all_reduce_(per_image_sum, count, group=None):
    """Multi-GPU masked-average-pool prototypes: MAP is a mean of PER-IMAGE ratios
    (networks/pspnet.py:14-15), so every support image lives wholly on one rank; ranks sum
    their per-image prototypes [C] and image counts, all-reduce both, then divide."""
    all_reduce_sum_(per_image_sum, group)
    all_reduce_sum_(count, group)
    return per_image_sum / count.to(per_image_sum.dtype)
