#!/usr/bin/env python
"""patch() on the REAL reference on a B200: the reference's own GFSS_Model classes with their real
backbones and decoders -- pspnet_pop/resnet50 (base and ft), convnext_pop/convnext-t, swin_pop/swin-t, lsk_pop/lsk-t,
deeplab_pop/resnet50, seghr_pop/hr-w32 and hr-w18 (C = 270: outside the kernels' range, must stay on the reference) --
seeded random init, BatchNorm statistics calibrated on random images, a trained-like head on prototypes taken from the
model's own decoder features, driven through the body of eval_base.py:162-178 -- model(image) -> F.interpolate -> .cpu() -> np.argmax -> get_confusion_matrix --
unpatched vs patched on the same GPU, plus the fused path (sweep.TileEvaluator, the three-line edit of INTEGRATION.md).

    SEGLAND_REFERENCE=/path/to/SegLand python profiles/scripts/real_reference_dropin.py [n_tiles] [model,model,...]

The reference tree is not part of this repository; for this run a scratch copy travels under baseline/_ref/ (git-ignored).
"""
import os
import sys
import time

import numpy as np
import torch
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
os.environ.setdefault('SEGLAND_REFERENCE', os.path.join(ROOT, 'baseline', '_ref', 'SegLand'))
from oracle import gen_golden  # noqa: E402  (timm shim + sys.path for the reference tree)

gen_golden.REF = os.environ['SEGLAND_REFERENCE']


def eval_loop(model, images, labels, K, pyt_utils, record=None):
    """eval_base.py:162-178 with the dead `len(label.shape) < 3` guard taken (np.int -> int)."""
    cm = np.zeros((K, K))
    lat = []
    for image, label in zip(images, labels):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        image = image.cuda()
        with torch.no_grad():
            output = model(image)
            if record is not None:
                record.append(output.float().cpu())
            output = F.interpolate(input=output, size=label.shape[-2:], mode='bilinear', align_corners=True)
        seg_pred = np.asarray(np.argmax(output.cpu().numpy(), axis=1), dtype=np.uint8)
        seg_gt = np.asarray(label.numpy(), dtype=int)
        keep = seg_gt != 255
        cm += pyt_utils.get_confusion_matrix(seg_gt[keep], seg_pred[keep], K)
        torch.cuda.synchronize()
        lat.append(time.perf_counter() - t0)
    return cm, lat


def main():
    n_tiles = int(sys.argv[1]) if len(sys.argv) > 1 else 6
    pspnet_pop, convnext_pop, _, _, pyt_utils = gen_golden.import_reference()
    from segland_b200 import ops, patch as slp, sweep
    torch.backends.cudnn.benchmark = False
    import importlib
    specs = [('pspnet_pop', 'resnet50', False), ('pspnet_pop', 'resnet50', True), ('convnext_pop', 'convnext-t', True),
             ('swin_pop', 'swin-t', True), ('lsk_pop', 'lsk-t', True), ('deeplab_pop', 'resnet50', True),
             ('seghr_pop', 'hr-w32', True), ('seghr_pop', 'hr-w18', True)]
    if len(sys.argv) > 2:
        specs = [sp for sp in specs if sp[0] in sys.argv[2].split(',')]
    for mod_name, backbone, is_ft in specs:
        torch.manual_seed(0)
        K = 8 + (4 if is_ft else 0)
        net_mod = importlib.import_module('networks.' + mod_name)
        model = net_mod.GFSS_Model(n_base=7, backbone=backbone, dilated=True, os=8, n_novel=4, is_ft=is_ft)
        Cm = model.base_emb.shape[1]
        if is_ft:
            with torch.no_grad():
                torch.nn.init.orthogonal_(model.base_emb)
                for a, b in zip(model.classifier_n.parameters(), model.classifier.parameters()):
                    a.copy_(b + 0.01 * torch.randn_like(b))
        model = model.cuda()
        model.train()                                            # calibrate BatchNorm running statistics
        with torch.no_grad():
            for _ in range(30):
                slp._features(model, torch.randn(2, 3, 256, 256, device='cuda'))
        model.eval()
        g = torch.Generator().manual_seed(1)
        # smooth random images (random init + white noise gives spatially constant features: nothing to segment)
        images = [F.interpolate(torch.randn(1, 3, 24, 24, generator=g), size=(1024, 1024), mode='bicubic', align_corners=False)
                  + 0.1 * torch.randn(1, 3, 1024, 1024, generator=g) for _ in range(n_tiles)]
        # a head whose arg-max varies over the tile: prototypes = centred decoder features of random pixels of tile 0,
        # classifier(s) = the trained-like construction of segland_b200.synth on those prototypes
        from segland_b200 import synth
        with torch.no_grad():
            f0 = slp._features(model, images[0].cuda())[0].flatten(1).float().cpu()                   # [C, N]
        f0 = f0 - f0.mean(1, keepdim=True)
        pick = torch.randperm(f0.shape[1], generator=g)[:11]
        protos = torch.nn.functional.normalize(f0[:, pick].t().contiguous(), dim=-1)
        st = synth.make_trained_like_state(Cm, 7, 4 if is_ft else 0, seed=3, n_bg_units=min(64, Cm // 4), base=protos[:7],
                                           novel=protos[7:] if is_ft else None)
        with torch.no_grad():
            model.base_emb.copy_(st.base_emb)
            for seq, ws in ((model.classifier, st.cls),) + (((model.classifier_n, st.cls_n),) if is_ft else ()):
                seq[0].weight.copy_(ws[0].view(Cm, Cm, 1, 1)); seq[2].weight.copy_(ws[1].view(Cm, Cm, 1, 1))
                seq[4].weight.copy_(ws[2].view(1, Cm, 1, 1))
            if is_ft:
                model.novel_emb.copy_(st.novel_emb)
        labels = [torch.randint(0, K, (1, 1024, 1024), generator=g).to(torch.uint8) for _ in range(n_tiles)]
        name = f'{mod_name}/{backbone} C={Cm} ' + ('is_ft=True (forward_all, 12 classes)' if is_ft else 'is_ft=False (forward_base, 8 classes)')
        print(f'=== {name}, {n_tiles} tiles of 1024^2' + ('' if ops.PopHead.supports(Cm) else
              '  [C outside the kernels\' range: patch() leaves this model on the reference forward]'))
        res = {}
        for tf32 in (False, True):
            torch.backends.cudnn.allow_tf32 = tf32
            torch.backends.cuda.matmul.allow_tf32 = tf32
            tag = 'tf32 convs (PyTorch default)' if tf32 else 'fp32 convs'
            # unpatched: the reference as shipped
            eval_loop(model, images[:1], labels[:1], K, pyt_utils)                  # warm-up (cudnn autotune, allocator)
            ref_logits = []
            cm_ref, lat_ref = eval_loop(model, images, labels, K, pyt_utils, ref_logits)
            # unpatched, but the head sees the bf16-rounded decoder output (north_star's input format)
            dec_fwd = model.decoder.forward
            model.decoder.forward = lambda *a, **k: dec_fwd(*a, **k).to(torch.bfloat16).float()
            bf_logits = []
            cm_bf, _ = eval_loop(model, images, labels, K, pyt_utils, bf_logits)
            model.decoder.forward = dec_fwd
            # patched: same loop, same script body
            done = slp.patch()
            eval_loop(model, images[:1], labels[:1], K, pyt_utils)
            our_logits = []
            cm_our, lat_our = eval_loop(model, images, labels, K, pyt_utils, our_logits)
            # fused path: decoder features -> TileEvaluator (head + up-sample + argmax + confusion on the device)
            if not ops.PopHead.supports(Cm):
                slp.unpatch()
                R, O = torch.cat(ref_logits), torch.cat(our_logits)
                print(f'--- {tag}: patched == unpatched (reference forward kept): {torch.equal(R, O)}; per tile '
                      f'{1e3 * float(np.median(lat_ref)):.1f} ms')
                continue
            ev = sweep.TileEvaluator(slp.head_for(model), (1024, 1024))
            lat_fused, preds = [], []
            for it, (image, label) in enumerate(zip(images[:1] + images, labels[:1] + labels)):
                torch.cuda.synchronize()
                t0 = time.perf_counter()
                with torch.no_grad():
                    feats = slp._features(model, image.cuda(), fused_tail=True)
                    out = ev.step(feats, label.cuda())
                pred = out['pred'].cpu()
                torch.cuda.synchronize()
                if it == 0:
                    ev.reset()
                    continue
                lat_fused.append(time.perf_counter() - t0)
                preds.append(pred)
            cm_fused = ev.cm.cpu().numpy().astype(np.float64)
            slp.unpatch()
            R, Bf, O = torch.cat(ref_logits), torch.cat(bf_logits), torch.cat(our_logits)
            rel = lambda a, b: ((a - b).abs().max() / b.abs().max()).item()
            up = lambda x: F.interpolate(x, size=(1024, 1024), mode='bilinear', align_corners=True).argmax(1)
            agree_ref = (up(O) == up(R)).float().mean().item()
            agree_bf = (up(O) == up(Bf)).float().mean().item()
            miou = lambda cm: ops.miou_from_confusion(cm, 7)[2]
            print(f'--- {tag}')
            print(f'patched: {len(done)} names; head logits rel-to-max vs reference (fp32 features) {rel(O, R):.2e}, '
                  f'vs reference on bf16-rounded features {rel(O, Bf):.2e}; bf16 rounding alone moves the reference by {rel(Bf, R):.2e}')
            hist = torch.bincount(up(R).flatten(), minlength=K).float()
            print(f'arg-max agreement at 1024^2: vs reference {100 * agree_ref:.4f} %, vs reference on bf16 features {100 * agree_bf:.4f} % '
                  f'(reference class shares: {[round(v, 3) for v in (hist / hist.sum()).tolist()]})')
            print(f'mIoU (random labels, so ~chance): reference {miou(cm_ref):.6f}, reference/bf16 {miou(cm_bf):.6f}, patched {miou(cm_our):.6f}, '
                  f'fused path {miou(cm_fused):.6f}; |cm_patched - cm_bf16ref|_1 = {np.abs(cm_our - cm_bf).sum():.0f} of {cm_bf.sum():.0f} px')
            med = lambda v: 1e3 * float(np.median(v))
            print(f'per-tile latency incl. backbone (median of {n_tiles}): unpatched {med(lat_ref):.1f} ms = {1e3 / med(lat_ref):.2f} tiles/s; '
                  f'patched, script unchanged {med(lat_our):.1f} ms = {1e3 / med(lat_our):.2f} tiles/s; '
                  f'fused TileEvaluator {med(lat_fused):.1f} ms = {1e3 / med(lat_fused):.2f} tiles/s')
            res[tf32] = (med(lat_ref), med(lat_our), med(lat_fused))
        # backbone + decoder alone, for context
        with torch.no_grad():
            x = images[0].cuda()
            for _ in range(2):
                slp._features(model, x)
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            for _ in range(5):
                slp._features(model, x)
            torch.cuda.synchronize()
            print(f'backbone + decoder alone (tf32): {1e3 * (time.perf_counter() - t0) / 5:.1f} ms per tile')


if __name__ == '__main__':
    main()
