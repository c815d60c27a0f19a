"""CPU oracle for the POP-head + post-processing hot path (TEST INFRASTRUCTURE ONLY).

A plain torch-CPU / numpy fp32 restatement of the reference's algorithm, function by
function, each citing the SegLand file:line it follows.  Only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline / ``--impl reference`` legs
may import this package; nothing under ``segland_b200/`` does.

Pinning: the reference ships no tests or golden vectors, so the oracle
is pinned against outputs of the reference's own functions, imported from
``/root/reference`` by ``oracle/gen_golden.py`` in the authoring container and committed
as ``tests/golden/*.npz``; ``tests/test_oracle_golden.py`` replays them.

The arithmetic deliberately keeps the reference's *materialising* formulation
(rank-1 [B,K,C,N] tensors through real 1x1 convolutions) instead of the algebraic
collapse the CUDA path uses, so the two are independent derivations.
"""
from __future__ import annotations

import numpy as np
import torch
import torch.nn.functional as F

IGNORE_LABEL = 255  # dataset/oem.py:15


# --------------------------------------------------------------------------- head
def ref_orthogonal_decompose(feats, bases_b, bases_n=None):
    """networks/pspnet_pop.py:95-121.  feats [B,C,N]; bases_* [1,K,C]."""
    q = feats.to(torch.float)
    s1 = F.normalize(bases_b.to(torch.float), p=2, dim=-1)
    proj1 = torch.matmul(s1, q)                                  # [B,Kb,N]
    out_fg_b = proj1.unsqueeze(2) * s1.unsqueeze(-1)             # [B,Kb,C,N]
    out_bg = q - out_fg_b.sum(1)
    if bases_n is None:
        return out_fg_b, out_bg.unsqueeze(1)
    s2 = F.normalize(bases_n, p=2, dim=-1)
    proj2 = torch.matmul(s2, q)
    out_fg_n = proj2.unsqueeze(2) * s2.unsqueeze(-1)
    out_bg = out_bg - out_fg_n.sum(1)
    return out_fg_b, out_fg_n, out_bg.unsqueeze(1)


def ref_classifier(x, W1, W2, w3):
    """The bias-free Conv1x1-ReLU-Conv1x1-ReLU-Conv1x1 stack, pspnet_pop.py:46-52.
    x [M,C,h,w]; W1,W2 [C,C] (or [C,C,1,1]); w3 [C] (or [1,C,1,1])."""
    C = x.shape[1]
    x = F.relu(F.conv2d(x, W1.reshape(C, C, 1, 1)))
    x = F.relu(F.conv2d(x, W2.reshape(C, C, 1, 1)))
    return F.conv2d(x, w3.reshape(1, C, 1, 1))


def ref_head_base(features, base_emb, cls):
    """forward_base after the decoder, pspnet_pop.py:171-182.
    features [B,C,h,w]; base_emb [Kb,C]; cls = (W1,W2,w3).  -> [B,1+Kb,h,w]."""
    outs = []
    for f in features.split(1, dim=0):                            # per image: bounds memory
        B, C, h, w = f.shape
        cls_emb = base_emb.unsqueeze(0)
        n_class = 1 + cls_emb.shape[1]
        feats_fg, feats_bg = ref_orthogonal_decompose(f.flatten(2), cls_emb)
        feats_all = torch.cat([feats_bg, feats_fg], dim=1).contiguous().view(B * n_class, C, h, w)
        outs.append(ref_classifier(feats_all, *cls).view(B, n_class, h, w))
    return torch.cat(outs, dim=0)


def ref_head_all(features, base_emb, novel_emb, cls, cls_n):
    """forward_all after the decoder, pspnet_pop.py:143-159.  -> [B,1+Kb+Kn,h,w]
    (channel order [bg(classifier_n), base.. (classifier), novel.. (classifier_n)])."""
    outs = []
    n_base, n_novel = base_emb.shape[0], novel_emb.shape[0]
    for f in features.split(1, dim=0):
        B, C, h, w = f.shape
        out_fg_b, out_fg_n, feats_bg = ref_orthogonal_decompose(
            f.flatten(2), base_emb.unsqueeze(0), novel_emb.unsqueeze(0))
        preds1 = ref_classifier(out_fg_b.contiguous().view(B * n_base, C, h, w), *cls).view(B, n_base, h, w)
        feats_n = torch.cat([feats_bg, out_fg_n], dim=1).contiguous().view(B * (n_novel + 1), C, h, w)
        preds2 = ref_classifier(feats_n, *cls_n).view(B, n_novel + 1, h, w)
        outs.append(torch.cat([preds2[:, 0].unsqueeze(1), preds1, preds2[:, 1:]], dim=1))
    return torch.cat(outs, dim=0)


def ref_head(features, base_emb, novel_emb, cls, cls_n=None):
    if novel_emb is None or novel_emb.shape[0] == 0:
        return ref_head_base(features, base_emb, cls)
    return ref_head_all(features, base_emb, novel_emb, cls, cls_n)


# ------------------------------------------------------------------ post-processing
def ref_upsample(logits, size):
    """eval_base.py:168, eval_ft.py:170, ft_pop.py:329, loss/criterion.py:51."""
    return F.interpolate(input=logits, size=tuple(size), mode='bilinear', align_corners=True)


def ref_upsample_argmax(logits, size):
    """eval_base.py:168-170 / eval_ft.py:168-172: upsample, np.argmax(axis=1) -> uint8."""
    out = ref_upsample(logits, size)
    return np.asarray(np.argmax(out.cpu().numpy(), axis=1), dtype=np.uint8)


def ref_get_confusion_matrix(gt_label, pred_label, class_num):
    """utils/pyt_utils.py:182-200 (inputs already filtered by gt != ignore)."""
    index = (gt_label * class_num + pred_label).astype('int32')
    label_count = np.bincount(index)
    confusion_matrix = np.zeros((class_num, class_num))
    for i_label in range(class_num):
        for i_pred_label in range(class_num):
            cur_index = i_label * class_num + i_pred_label
            if cur_index < len(label_count):
                confusion_matrix[i_label, i_pred_label] = label_count[cur_index]
    return confusion_matrix


def ref_confusion(seg_gt, seg_pred, class_num, ignore_label=IGNORE_LABEL):
    """eval_base.py:172-178: filter ignore, then get_confusion_matrix.  np.int -> int."""
    seg_gt = np.asarray(seg_gt, dtype=int)
    keep = seg_gt != ignore_label
    return ref_get_confusion_matrix(seg_gt[keep], np.asarray(seg_pred)[keep], class_num)


def ref_miou(confusion_matrix, base_classes):
    """eval_base.py:193-199 / eval_ft.py:196-202 -> (base, novel, total, per-class)."""
    pos = confusion_matrix.sum(1)
    res = confusion_matrix.sum(0)
    tp = np.diag(confusion_matrix)
    with np.errstate(divide='ignore', invalid='ignore'):
        miou_array = tp / (pos + res - tp)
        base = np.nanmean(miou_array[:base_classes + 1])
        novel = np.nanmean(miou_array[base_classes + 1:]) if base_classes + 1 < len(miou_array) else float('nan')
        total = np.nanmean(miou_array)
    return base, novel, total, miou_array


def ref_inter_union(output, target, K, ignore_index=IGNORE_LABEL):
    """intersectionAndUnionGPU, utils/pyt_utils.py:293-305.  torch.histc rejects int64 on
    CPU, so the histograms run on float copies (values are small exact integers); the
    in-place side effect on `output` is kept."""
    assert output.dim() in [1, 2, 3]
    assert output.shape == target.shape
    output = output.reshape(-1)
    target = target.reshape(-1)
    output[target == ignore_index] = ignore_index
    intersection = output[output == target]
    area_intersection = torch.histc(intersection.float(), bins=K, min=0, max=K - 1)
    area_output = torch.histc(output.float(), bins=K, min=0, max=K - 1)
    area_target = torch.histc(target.float(), bins=K, min=0, max=K - 1)
    area_union = area_output + area_target - area_intersection
    return area_intersection, area_union, area_target


def ref_pseudo_label(preds2_base, mask_b, n_base):
    """pspnet_pop.py:221-231.  preds2_base [Bb,1+Kn,h,w] are the base images' classifier_n
    outputs; mask_b [Bb,H,W] int64 is modified in place and also returned stacked."""
    mask_new = []
    for b in range(mask_b.shape[0]):
        bg_mask = mask_b[b] == 0
        bg_out = F.interpolate(input=preds2_base[b].unsqueeze(0), size=mask_b[b].shape,
                               mode='bilinear', align_corners=True)
        bg_idx = torch.argmax(bg_out.squeeze(0), dim=0)
        bg_idx[bg_idx > 0] += n_base
        mask_b[b][bg_mask] = bg_idx[bg_mask]
        mask_new.append(mask_b[b])
    return torch.stack(mask_new, dim=0)


# ---------------------------------------------------------------- prototypes / loss
def ref_masked_average_pooling(feature, mask):
    """networks/pspnet.py:7-15.  feature [B,C,h,w]; mask [B,1,H,W] -> [1,1,C]."""
    mask = F.interpolate(mask, size=feature.shape[-2:], mode="bilinear", align_corners=True)
    masked_feature = torch.sum(feature * mask, dim=(2, 3)) / (mask.sum(dim=(2, 3)) + 1e-5)
    return masked_feature.mean(0, keepdim=True).unsqueeze(1)


def ref_proto_sim_base(base_emb):
    """pspnet_pop.py:185-186."""
    cls_emb = F.normalize(base_emb.unsqueeze(0), p=2, dim=-1).squeeze(0)
    return torch.matmul(cls_emb, cls_emb.t())


def ref_proto_sim_ft(novel_emb, base_emb):
    """pspnet_pop.py:234-239."""
    C = novel_emb.shape[-1]
    n = F.normalize(novel_emb.unsqueeze(0).to(torch.float), p=2, dim=-1).reshape(-1, C)
    all_emb = torch.cat([n, F.normalize(base_emb.to(torch.float), p=2, dim=-1)], dim=0)
    return torch.matmul(n, all_emb.t())


def ref_orth_loss(proto_sim):
    """OrthLoss.get_orth_loss, loss/criterion.py:37-43."""
    eye_sim = torch.triu(torch.ones_like(proto_sim), diagonal=1)
    return torch.abs(proto_sim[eye_sim == 1]).mean()


def ref_seg_ce(preds, target, ignore_index=IGNORE_LABEL):
    """The segmentation term of OrthLoss.forward / CELoss.forward, loss/criterion.py:17-19,51-52:
    bilinear align_corners upsample to the label size, then CrossEntropyLoss(ignore_index, mean)."""
    scale_pre = F.interpolate(input=preds, size=target.shape[1:], mode='bilinear', align_corners=True)
    return torch.nn.CrossEntropyLoss(ignore_index=ignore_index, reduction='mean')(scale_pre, target)


def ref_orth_loss_forward(preds, target, proto_sim, w=10.0, ignore_index=IGNORE_LABEL):
    """OrthLoss.forward without aux head, loss/criterion.py:45-65 (self.w = 10.0, :35)."""
    seg_loss = ref_seg_ce(preds, target, ignore_index)
    orth_loss = ref_orth_loss(proto_sim)
    return {'total_loss': seg_loss + orth_loss * w, 'seg_loss': seg_loss, 'orth_loss': orth_loss}


def ref_forward_novel(features_full, mask, mask_b, base_emb, novel_emb, cls, cls_n):
    """forward_novel after the decoder with its criterion, pspnet_pop.py:199-243, differentiable by
    autograd in the reference's materialising formulation.  features_full [B,C,h,w] = [novel images;
    base images]; mask [B/2,H,W] int64; mask_b [B/2,H,W] int64 is pseudo-labelled IN PLACE (:221-231)."""
    preds = ref_head_all(features_full, base_emb, novel_emb, cls, cls_n)
    B, n_base = preds.shape[0], base_emb.shape[0]
    preds2 = torch.cat([preds[:, :1], preds[:, 1 + n_base:]], dim=1)
    mask_new = ref_pseudo_label(preds2[B // 2:].detach(), mask_b, n_base)
    mask_all = torch.cat([mask, mask_new], dim=0)
    return ref_orth_loss_forward(preds, mask_all, ref_proto_sim_ft(novel_emb, base_emb)), preds


def ref_forward_base_loss(features, mask, base_emb, cls):
    """forward_base with a criterion, pspnet_pop.py:169-187."""
    preds = ref_head_base(features, base_emb, cls)
    return ref_orth_loss_forward(preds, mask, ref_proto_sim_base(base_emb)), preds


# --------------------------------------------------------------------------- fusion
def ref_fuse(mats, n_lists=None):
    """fusemat.py:42-48 for ONE tile: sequential += in directory order, / len(fusion_list),
    argmax(axis=0) -> uint8.  mats: list of [K,H,W] float32 numpy arrays."""
    acc = np.array(mats[0], dtype=np.float32, copy=True)
    for m in mats[1:]:
        acc += m
    n = len(mats) if n_lists is None else n_lists
    fused = acc / n
    return np.argmax(fused, axis=0).astype(np.uint8), fused


# ------------------------------------------------------------------- whole hot path
def ref_window_accumulate(crop_logits, origins_y, origins_x, flips, canvas_hw):
    """spec: this repo (the reference has no sliding-window inference: engine.py:23-143, eval_ft.py:162-172).
    The usual sliding-window composition written with PyTorch primitives: crop_logits [B,E,K,hc,wc] with entry
    e = (gy*nx + gx)*V + v; every view is un-flipped (torch.flip; bit0 = width, bit1 = height), added into the canvas
    at its window origin in entry order, and the canvas is divided by the overlap count."""
    B, E, K, hc, wc = crop_logits.shape
    h, w = canvas_hw
    canvas = torch.zeros(B, K, h, w, dtype=torch.float32)
    count = torch.zeros(h, w, dtype=torch.float32)
    e = 0
    for oy in origins_y:
        for ox in origins_x:
            for f in flips:
                dims = [d for d, bit in ((-1, 1), (-2, 2)) if f & bit]
                c = crop_logits[:, e].to(torch.float32)
                canvas[:, :, oy:oy + hc, ox:ox + wc] += torch.flip(c, dims) if dims else c
                count[oy:oy + hc, ox:ox + wc] += 1
                e += 1
    assert e == E
    return canvas / count, count


def ref_eval_tile(features, label, base_emb, novel_emb, cls, cls_n, out_size, n_classes):
    """One eval step of eval_base.py:166-178 / eval_ft.py:166-183 after the decoder:
    head -> upsample -> argmax -> confusion.  features [1,C,h,w] fp32 (bf16-rounded values),
    label [1,H,W] uint8 numpy.  Returns (pred u8 [1,H,W], cm float64 [K,K], logits)."""
    logits = ref_head(features, base_emb, novel_emb, cls, cls_n)
    pred = ref_upsample_argmax(logits, out_size)
    cm = ref_confusion(label, pred, n_classes)
    return pred, cm, logits


# --------------------------------------------------------------------------- decoder tails
def ref_tail_layernorm(x, gamma, beta, eps=1e-5):
    """FPN_Seg_OCR_Decoder.norm applied channels-last, networks/convnext_pop.py:13,27:
    self.norm(feats.permute(0, 2, 3, 1)).permute(0, 3, 1, 2).  x [B,C,h,w] fp32 -> fp32."""
    C = x.shape[1]
    return F.layer_norm(x.permute(0, 2, 3, 1), (C,), gamma, beta, eps).permute(0, 3, 1, 2)


def ref_tail_bn_relu_conv(x, bn_weight, bn_bias, bn_mean, bn_var, bn_eps, W, bias, relu=True):
    """PSPModule.bottleneck[1:4] in eval mode, networks/pspnet_pop.py:19-22: BatchNorm2d (running statistics)
    -> ReLU -> Conv2d(C, C, kernel_size=1) with bias.  x [B,Cin,h,w]; W [Cout,Cin]; -> fp32 [B,Cout,h,w]."""
    y = x
    if bn_weight is not None:
        y = F.batch_norm(y, bn_mean, bn_var, bn_weight, bn_bias, False, 0.0, bn_eps)
    if relu:
        y = F.relu(y)
    return F.conv2d(y, W.reshape(W.shape[0], W.shape[1], 1, 1), bias)


def ref_tail_sum(maps):
    """torch.stack(fpn_outs, dim=-1).sum(-1), networks/swin_pop.py:169-172 / lsk_pop.py:163-165."""
    return torch.stack(list(maps), dim=-1).sum(-1)


def ref_tail_bn_relu(x, bn_weight, bn_bias, bn_mean, bn_var, bn_eps, relu=True):
    """The `bn` + `relu` of _ConvBnReLU after _ASPP.fc's convolution, networks/deeplab_pop.py:12-29,61,66 (eval mode),
    and the last BatchNorm2d + ReLU of DoubleConv in VGGUNet.up4, networks/vggunet_pop.py:19-20."""
    y = F.batch_norm(x, bn_mean, bn_var, bn_weight, bn_bias, False, 0.0, bn_eps)
    return F.relu(y) if relu else y


def ref_tail_concat(maps):
    """torch.cat([x[0], x1, x2, x3], 1), HRFPN_Seg_Decoder.forward, networks/seghr_pop.py:23-24."""
    return torch.cat(list(maps), 1)


def bf16_ulp_report(got_bf16, ref_fp32):
    """How a bf16 feature tensor compares with round-to-nearest(ref_fp32): fraction of identical bit patterns and
    the largest |got - ref_fp32| in units of the reference element's bf16 spacing (0.5 = ideal rounding)."""
    got = got_bf16.float().double()
    ref = ref_fp32.double()
    exact = (got_bf16 == ref_fp32.to(torch.bfloat16)).float().mean().item()
    spacing = torch.pow(2.0, torch.floor(torch.log2(ref.abs().clamp_min(1e-30))) - 7)
    return exact, ((got - ref).abs() / spacing)
