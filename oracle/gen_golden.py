"""Generate tests/golden/*.npz by running the REFERENCE's own code (authoring container only).

    python oracle/gen_golden.py            # needs /root/reference; writes tests/golden/

The reference has no tests or golden vectors of its own, so the oracle in
oracle/ref_ops.py is pinned against what the reference's functions compute here, on seeded
synthetic inputs from segland_b200.synth.  /root/reference does not exist on the GPU box;
the committed .npz files travel instead.  Nothing is copied from the reference: it is
imported (with a 4-symbol in-memory `timm` shim, which only its out-of-scope backbones need)
and called.

What is executed from the reference:
  networks/pspnet_pop.py   GFSS_Model.forward_base / forward_all / forward_novel with the
                           backbone and decoder replaced by identities (input = features)
  networks/convnext_pop.py GFSS_Model.forward_all (head body identical, d_model differs)
  loss/criterion.py        OrthLoss (get_orth_loss + forward)
  utils/pyt_utils.py       get_confusion_matrix, intersectionAndUnion, intersectionAndUnionGPU
  networks/pspnet.py       masked_average_pooling
  fusemat.py               the script itself, exec'd with its placeholder paths substituted
"""
from __future__ import annotations

import os
import sys
import tempfile
import types

import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.environ.get('SEGLAND_REFERENCE', '/root/reference')
sys.path.insert(0, REPO)

from segland_b200 import synth  # noqa: E402


def import_reference():
    if not os.path.isdir(REF):
        raise SystemExit(f'{REF} not found: golden vectors can only be regenerated where the reference is mounted')
    timm = types.ModuleType('timm')
    models = types.ModuleType('timm.models')
    layers = types.ModuleType('timm.models.layers')
    registry = types.ModuleType('timm.models.registry')

    class DropPath(nn.Identity):
        def __init__(self, *a, **k):
            super().__init__()

    layers.DropPath = DropPath
    layers.to_2tuple = lambda x: x if isinstance(x, tuple) else (x, x)
    layers.trunc_normal_ = nn.init.trunc_normal_
    registry.register_model = lambda f: f
    sys.modules.update({'timm': timm, 'timm.models': models, 'timm.models.layers': layers,
                        'timm.models.registry': registry})
    sys.path.insert(0, REF)
    import networks.pspnet_pop as pspnet_pop
    import networks.convnext_pop as convnext_pop
    import networks.pspnet as pspnet
    import loss.criterion as criterion
    import utils.pyt_utils as pyt_utils
    return pspnet_pop, convnext_pop, pspnet, criterion, pyt_utils


class _IdentityBackbone(nn.Module):
    def base_forward(self, x, **kw):
        return x

    def forward(self, x, **kw):           # convnext_pop/swin_pop call self.backbone(img)
        return x


def build_ref_model(module, state: synth.HeadState, criterion=None):
    """A real reference GFSS_Model whose backbone/decoder are identities, so that
    forward_*(features) runs exactly the reference's head code on `features`."""
    cls_t = module.GFSS_Model
    m = cls_t.__new__(cls_t)
    nn.Module.__init__(m)
    C = state.base_emb.shape[1]

    def mlp(ws):
        seq = nn.Sequential(nn.Conv2d(C, C, 1, bias=False), nn.ReLU(inplace=True),
                            nn.Conv2d(C, C, 1, bias=False), nn.ReLU(inplace=True),
                            nn.Conv2d(C, 1, 1, bias=False))
        with torch.no_grad():
            seq[0].weight.copy_(ws[0].view(C, C, 1, 1))
            seq[2].weight.copy_(ws[1].view(C, C, 1, 1))
            seq[4].weight.copy_(ws[2].view(1, C, 1, 1))
        return seq

    m.backbone = _IdentityBackbone()
    m.decoder = nn.Identity()
    m.classifier = mlp(state.cls)
    is_ft = state.novel_emb is not None
    m.base_emb = nn.Parameter(state.base_emb.clone(), requires_grad=not is_ft)
    if is_ft:
        m.novel_emb = nn.Parameter(state.novel_emb.clone(), requires_grad=True)
        m.classifier_n = mlp(state.cls_n)
        m.n_novel = state.novel_emb.shape[0]
    else:
        m.novel_emb = None
        m.n_novel = 0
    m.use_base, m.is_ft, m.criterion, m.n_base = True, is_ft, criterion, state.base_emb.shape[0]
    return m


def state_arrays(prefix, st: synth.HeadState):
    d = {f'{prefix}base_emb': st.base_emb.numpy(), f'{prefix}W1': st.cls[0].numpy(),
         f'{prefix}W2': st.cls[1].numpy(), f'{prefix}w3': st.cls[2].numpy()}
    if st.novel_emb is not None:
        d.update({f'{prefix}novel_emb': st.novel_emb.numpy(), f'{prefix}W1n': st.cls_n[0].numpy(),
                  f'{prefix}W2n': st.cls_n[1].numpy(), f'{prefix}w3n': st.cls_n[2].numpy()})
    return d


def main():
    pspnet_pop, convnext_pop, pspnet, criterion, pyt_utils = import_reference()
    out_dir = os.path.join(REPO, 'tests', 'golden')
    os.makedirs(out_dir, exist_ok=True)
    torch.manual_seed(1234)
    torch.set_grad_enabled(False)

    # ---- head, base mode (forward_base) and ft mode (forward_all), pspnet_pop + convnext_pop
    def head_case(name, module, C, Kb, Kn, T, H, stride, seed, random_feats=False):
        st = synth.make_head_state(C, Kb, Kn, seed=seed)
        labels = synth.make_labels(T, H, H, st.n_classes, seed=seed, coarse=8)
        if random_feats:
            feats = synth.make_random_features(T, C, H // stride, H // stride, seed=seed)
        else:
            feats = synth.make_features(labels, st, stride, seed=seed)
        model = build_ref_model(module, st).eval()
        logits = model(feats.float())                     # bf16-rounded values, fp32 arithmetic
        up = F.interpolate(input=logits, size=labels.shape[-2:], mode='bilinear', align_corners=True)
        pred = np.asarray(np.argmax(up.cpu().numpy(), axis=1), dtype=np.uint8)
        cm = np.zeros((st.n_classes, st.n_classes))
        for t in range(T):                                # eval_base.py:172-178 (np.int -> int)
            seg_gt = np.asarray(labels[t].numpy(), dtype=int)
            keep = seg_gt != 255
            cm += pyt_utils.get_confusion_matrix(seg_gt[keep], pred[t][keep], st.n_classes)
        arrays = dict(feats_bf16_bits=feats.view(torch.int16).numpy(), labels=labels.numpy(),
                      logits=logits.numpy(), pred=pred, cm=cm,
                      Kb=np.int64(Kb), Kn=np.int64(Kn), stride=np.int64(stride))
        if C <= 64:                                        # keep the big cases small on disk
            arrays['upsampled'] = up.numpy().astype(np.float32)
        arrays.update(state_arrays('', st))
        np.savez_compressed(os.path.join(out_dir, name + '.npz'), **arrays)
        print(name, 'logits', tuple(logits.shape), 'absmax %.4f' % logits.abs().max().item(),
              'pred classes', np.unique(pred).tolist())

    head_case('head_base_c64', pspnet_pop, 64, 7, 0, 2, 64, 8, seed=11)
    head_case('head_ft_c64', pspnet_pop, 64, 7, 4, 2, 64, 8, seed=12)
    head_case('head_base_c512', pspnet_pop, 512, 7, 0, 1, 128, 8, seed=13)
    head_case('head_ft_c192_s4', convnext_pop, 192, 7, 4, 1, 64, 4, seed=14)
    head_case('head_ft_c96_rand', pspnet_pop, 96, 7, 4, 1, 64, 4, seed=15, random_feats=True)

    # ---- upsample geometry cases (non-square, odd sizes, x8 and x4, tiny inputs)
    ups = {}
    gen = torch.Generator().manual_seed(21)
    for i, (K, h, w, H, W) in enumerate([(12, 16, 16, 128, 128), (8, 9, 13, 70, 101), (12, 32, 32, 128, 128),
                                         (3, 2, 2, 17, 5), (12, 16, 16, 16, 16), (5, 7, 1, 30, 1)]):
        lg = torch.randn(1, K, h, w, generator=gen)
        up = F.interpolate(input=lg, size=(H, W), mode='bilinear', align_corners=True)
        ups[f'in{i}'] = lg.numpy()
        ups[f'out{i}'] = up.numpy()
        ups[f'pred{i}'] = np.asarray(np.argmax(up.numpy(), axis=1), dtype=np.uint8)
    np.savez_compressed(os.path.join(out_dir, 'upsample_cases.npz'), **ups)

    # ---- metrics: get_confusion_matrix, intersectionAndUnion (numpy) and ...GPU (torch, float)
    gen = torch.Generator().manual_seed(31)
    K = 12
    gt = torch.randint(0, K, (3, 96, 96), generator=gen)
    gt[torch.rand(3, 96, 96, generator=gen) < 0.05] = 255
    pr = torch.randint(0, K, (3, 96, 96), generator=gen)
    keep = (gt != 255).numpy()
    cm = pyt_utils.get_confusion_matrix(gt.numpy()[keep], pr.numpy()[keep], K)
    i_np, u_np, t_np = pyt_utils.intersectionAndUnion(pr.numpy().copy(), gt.numpy(), K, 255)
    out_f = pr.clone().float()
    i_t, u_t, t_t = pyt_utils.intersectionAndUnionGPU(out_f, gt.clone().float(), K, 255)
    np.savez_compressed(os.path.join(out_dir, 'metrics.npz'), gt=gt.numpy().astype(np.uint8),
                        pred=pr.numpy().astype(np.uint8), cm=cm, inter_np=i_np, union_np=u_np, target_np=t_np,
                        inter_t=i_t.numpy(), union_t=u_t.numpy(), target_t=t_t.numpy(),
                        output_after=out_f.numpy().astype(np.int64))

    # ---- masked_average_pooling (networks/pspnet.py:7-15)
    st = synth.make_head_state(64, 7, 4, seed=41)
    labels = synth.make_labels(5, 64, 64, 12, seed=41, coarse=8)
    feats = synth.make_features(labels, st, 8, seed=41)
    masks = synth.make_support_masks(5, 64, 64, seed=41, coarse=8)
    proto = pspnet.masked_average_pooling(feats.float(), masks)
    soft = F.interpolate(torch.rand(2, 1, 11, 7, generator=gen), size=(50, 36), mode='bilinear', align_corners=False)
    feats2 = synth.make_random_features(2, 32, 10, 12, seed=42)
    proto2 = pspnet.masked_average_pooling(feats2.float(), soft)
    np.savez_compressed(os.path.join(out_dir, 'map.npz'), feats_bf16_bits=feats.view(torch.int16).numpy(),
                        masks=masks.numpy(), proto=proto.numpy(), feats2_bf16_bits=feats2.view(torch.int16).numpy(),
                        masks2=soft.numpy(), proto2=proto2.numpy())

    # ---- OrthLoss: get_orth_loss on base [7,7] and ft [4,11] proto_sim, + forward_base/forward_novel loss dicts
    torch.set_grad_enabled(True)
    crit = criterion.OrthLoss(ignore_index=255)
    st_b = synth.make_head_state(64, 7, 0, seed=51)
    st_b.base_emb = st_b.base_emb + 0.3 * torch.randn(7, 64, generator=gen)     # not orthogonal any more
    labels = synth.make_labels(2, 64, 64, 8, seed=51, coarse=8)
    feats = synth.make_features(labels, st_b, 8, seed=51)
    mb = build_ref_model(pspnet_pop, st_b, criterion=crit).eval()
    loss_b = mb(feats.float(), labels.long())
    loss_b['total_loss'].backward()
    grad_base = mb.base_emb.grad.clone()
    sim_b = torch.matmul(F.normalize(st_b.base_emb, dim=-1), F.normalize(st_b.base_emb, dim=-1).t())
    # orth-only gradient (the term the CUDA kernel differentiates)
    be = st_b.base_emb.clone().requires_grad_(True)
    e = F.normalize(be.unsqueeze(0), p=2, dim=-1).squeeze(0)
    crit.get_orth_loss(torch.matmul(e, e.t())).backward()
    orth_grad_b = be.grad.clone()

    st_f = synth.make_head_state(64, 7, 4, seed=52)
    st_f.novel_emb = st_f.novel_emb + 0.3 * torch.randn(4, 64, generator=gen)
    mf = build_ref_model(pspnet_pop, st_f, criterion=crit)
    mf.train()
    lab_n = synth.make_labels(2, 64, 64, 12, seed=52, coarse=8)
    lab_b = synth.make_labels(2, 64, 64, 8, seed=53, coarse=8)
    img_n = synth.make_features(lab_n, st_f, 8, seed=52)
    img_b = synth.make_features(lab_b, st_f, 8, seed=53)
    mask_n = lab_n.long()
    mask_b = lab_b.long()
    mask_b_before = mask_b.clone()
    loss_f = mf(img_n.float(), mask_n, img_b.float(), mask_b)       # forward_novel; mutates mask_b
    ne = st_f.novel_emb.clone().requires_grad_(True)
    n_hat = F.normalize(ne.unsqueeze(0), p=2, dim=-1).reshape(-1, 64)
    all_emb = torch.cat([n_hat, F.normalize(st_f.base_emb, p=2, dim=-1)], dim=0)
    sim_f = torch.matmul(n_hat, all_emb.t())
    lo = crit.get_orth_loss(sim_f, is_ft=True)
    lo.backward()
    torch.set_grad_enabled(False)
    with torch.no_grad():
        mf.eval()
        preds_all = mf(torch.cat([img_n, img_b], 0).float())
    arrays = dict(sim_b=sim_b.detach().numpy(), orth_b=loss_b['orth_loss'].detach().numpy(),
                  seg_b=loss_b['seg_loss'].detach().numpy(), total_b=loss_b['total_loss'].detach().numpy(),
                  orth_grad_b=orth_grad_b.numpy(), total_grad_base=grad_base.numpy(),
                  feats_b_bits=feats.view(torch.int16).numpy(), labels_b=labels.numpy(),
                  sim_f=sim_f.detach().numpy(), orth_f=loss_f['orth_loss'].detach().numpy(),
                  seg_f=loss_f['seg_loss'].detach().numpy(), total_f=loss_f['total_loss'].detach().numpy(),
                  orth_grad_f=ne.grad.numpy(), img_n_bits=img_n.view(torch.int16).numpy(),
                  img_b_bits=img_b.view(torch.int16).numpy(), mask_n=mask_n.numpy(),
                  mask_b_before=mask_b_before.numpy(), mask_b_after=mask_b.numpy(),
                  preds_all=preds_all.numpy())
    arrays.update(state_arrays('b_', st_b))
    arrays.update(state_arrays('f_', st_f))
    np.savez_compressed(os.path.join(out_dir, 'orth_pseudo.npz'), **arrays)
    print('orth base %.6f ft %.6f ; pseudo-labelled px %d' % (
        float(loss_b['orth_loss']), float(loss_f['orth_loss']), int((mask_b != mask_b_before).sum())))

    # ---- segmentation CE tail: the reference's OrthLoss.forward / CELoss.forward on low-res logits
    torch.set_grad_enabled(True)
    gen = torch.Generator().manual_seed(71)
    ce = {}
    for i, (B, K, h, w, H, W, scale) in enumerate([(2, 12, 16, 16, 128, 128, 1.0), (1, 8, 9, 13, 70, 101, 3.0),
                                                   (2, 5, 8, 8, 8, 8, 0.5)]):
        preds = (torch.randn(B, K, h, w, generator=gen) * scale).requires_grad_(True)
        target = torch.randint(0, K, (B, H, W), generator=gen)
        target[torch.rand(B, H, W, generator=gen) < 0.07] = 255
        sim = torch.eye(4, 11)
        out = criterion.OrthLoss(ignore_index=255)(preds, target, is_ft=True, proto_sim=sim)
        out['seg_loss'].backward()
        ce.update({f'preds{i}': preds.detach().numpy(), f'target{i}': target.numpy(),
                   f'loss{i}': out['seg_loss'].detach().numpy(), f'grad{i}': preds.grad.numpy(),
                   f'total{i}': out['total_loss'].detach().numpy()})
        out2 = criterion.CELoss(ignore_index=255)(preds.detach(), target)
        assert torch.equal(out2['total_loss'], out['seg_loss'].detach())
    np.savez_compressed(os.path.join(out_dir, 'ce.npz'), **ce)
    torch.set_grad_enabled(False)
    print('ce losses', [float(ce[f'loss{i}']) for i in range(3)])

    # ---- fusemat.py: run the script itself on temporary .mat files
    import scipy.io
    from PIL import Image
    M, K, H = 3, 8, 64
    with tempfile.TemporaryDirectory() as tmp:
        dirs, stacks = [], {}
        for m in range(M):
            d = os.path.join(tmp, f'model{m}')
            os.makedirs(d)
            dirs.append(d)
            for tile in ('tile_a', 'tile_b'):
                mats = synth.make_logit_stacks(1, K, H, H, seed=60 + 7 * m + len(tile) + ord(tile[-1]))
                arr = mats[0].numpy()[None]                       # eval_base dumps [1,K,H,W]
                if tile == 'tile_b':                              # near-ties: quantise so sums collide
                    arr = np.round(arr * 2) / 2
                stacks[(m, tile)] = arr
                scipy.io.savemat(os.path.join(d, tile + '.mat'), {'outputs': arr})
        outp = os.path.join(tmp, 'fused')
        src = open(os.path.join(REF, 'fusemat.py')).read()
        src = src.replace("""        'PATH_OF_PROBABILITY_MAPS_FOR_FUSION_1',
        'PATH_OF_PROBABILITY_MAPS_FOR_FUSION_2',
        'PATH_OF_PROBABILITY_MAPS_FOR_FUSION_3',
        '...'""", ',\n'.join(repr(d) for d in dirs))
        src = src.replace("'PATH_OF_OUTPUT_PROBABILITY_MAPS'", repr(outp))
        src = src.replace('import scipy\n', 'import scipy\nimport scipy.io\n')   # scipy>=1.x lazy submodule
        src = src.replace('resize((1024, 1024)', f'resize(({H}, {H})')            # keep the fixture small
        assert repr(dirs[0]) in src and repr(outp) in src
        exec(compile(src, 'fusemat.py', 'exec'), {'__name__': '__main__'})
        arrays = {}
        for tile in ('tile_a', 'tile_b'):
            arrays[f'{tile}_pred'] = np.array(Image.open(os.path.join(outp, tile + '.png')))
            for m in range(M):
                arrays[f'{tile}_m{m}'] = stacks[(m, tile)][0]
        np.savez_compressed(os.path.join(out_dir, 'fuse.npz'), **arrays)
        print('fuse preds', {t: np.unique(arrays[f'{t}_pred']).tolist() for t in ('tile_a', 'tile_b')})

    sizes = {f: os.path.getsize(os.path.join(out_dir, f)) for f in sorted(os.listdir(out_dir))}
    print(sizes, 'total %.2f MB' % (sum(sizes.values()) / 1e6))


if __name__ == '__main__':
    main()
