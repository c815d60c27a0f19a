#!/usr/bin/env python
"""Per-operator achieved bandwidth for the other BASELINE configs (C3 MAP prototypes, C4 ft-mode head with
probability-map output, C5 logit fusion) and the small metric kernels: one line per op,
CUDA-event timed, inputs larger than L2 or rotated so every launch streams from HBM.
    python profiles/bench_ops.py            (on the GPU box)
Algorithmic bytes are the SURVEY section 8d figures; peak = MEASURED_PEAKS.json hbm_gbs."""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from segland_b200 import ops, synth  # noqa: E402

PEAK = json.load(open(os.path.join(ROOT, 'MEASURED_PEAKS.json')))['hbm_gbs'] if os.path.exists(
    os.path.join(ROOT, 'MEASURED_PEAKS.json')) else 6650.0


def timeit(fn, iters=20, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(iters):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / iters * 1e-3


def report(name, seconds, nbytes, tiles):
    gbs = nbytes / seconds / 1e9
    print(f'{name:58s} {seconds * 1e6 / tiles:9.2f} us/tile {tiles / seconds:10.0f} tiles/s {gbs:8.0f} GB/s '
          f'{100 * gbs / PEAK:5.1f}% of {PEAK:.0f}')


def main():
    dev = 'cuda'
    g = torch.Generator(dev).manual_seed(0)
    # ---- C5: fusion, M=3 models x K=12 x 1024^2 fp32, 8 tiles (1.2 GB > L2)
    T, M, K, H = 8, 3, 12, 1024
    mats = [torch.randn(K, T * H * H, device=dev, generator=g) for _ in range(M)]
    label = torch.randint(0, K, (T * H * H,), device=dev, dtype=torch.uint8, generator=g)
    cm = torch.zeros(K, K, dtype=torch.int64, device=dev)
    t = timeit(lambda: ops.fuse_logits(mats, label=label, cm=cm))
    report('C5 fuse_logits M=3 K=12 (+label, +pred, +confusion)', t, (M * K * 4 + 2) * T * H * H, T)
    t = timeit(lambda: ops.fuse_logits(mats))
    report('C5 fuse_logits M=3 K=12 (pred only)', t, (M * K * 4 + 1) * T * H * H, T)
    del mats
    # ---- C3: masked-average-pool prototypes, 20 support tiles [20,512,128,128] + masks [20,1,1024,1024]
    st = synth.make_head_state(512, 7, 4, seed=1)
    feats = torch.randn(20, 512, 128, 128, device=dev, generator=g).to(torch.bfloat16)
    masks = (torch.rand(20, 1, 1024, 1024, device=dev, generator=g) < 0.3).float()
    t = timeit(lambda: ops.masked_average_pooling(feats, masks))
    report('C3 masked_average_pooling 20 support tiles (fp32 masks)', t, 20 * (512 * 128 * 128 * 2 + 1024 * 1024 * 4), 20)
    # ---- C4: ft-mode heads with probability-map output
    for name, C, hw in (('ConvNeXt-T C=192', 192, 256), ('Swin-T/S C=96', 96, 256), ('PSPNet ft C=512', 512, 128)):
        stc = synth.make_head_state(C, 7, 4, seed=2)
        Tc = 32 if C * hw * hw * 2 * 32 <= (1 << 30) else 16      # >= 2 x L2 of features, enough tiles to fill 148 SMs
        f = torch.randn(Tc, C, hw, hw, device=dev, generator=g).to(torch.bfloat16)
        head = ops.PopHead(stc.base_emb, stc.cls, stc.novel_emb, stc.cls_n)
        lg = torch.empty(Tc, 12, hw, hw, device=dev)
        t = timeit(lambda: head(f, out=lg, fg_only=True))
        report(f'C4 fg logits K=11 {name}', t, Tc * (C * hw * hw * 2 + 11 * hw * hw * 4), Tc)
        t = timeit(lambda: head(f, out=lg))
        flops = (4 * C * C + 2 * C) * hw * hw * Tc
        print(f'{"C4 full head (fg + bg " + ("tc" if head._use_tc(hw * hw) else "simt") + ") " + name:58s} '
              f'{t * 1e6 / Tc:9.2f} us/tile {Tc / t:10.0f} tiles/s   bg {flops / t / 1e12:7.1f} TFLOP/s algorithmic')
        t = timeit(lambda: ops.upsample_argmax(lg, (1024, 1024), want_probs=True), iters=10)
        report(f'C4 upsample+softmax prob-map out K=12 {name}', t, Tc * (12 * hw * hw * 4 + 12 * 1024 * 1024 * 4 + 1024 * 1024), Tc)
        t = timeit(lambda: ops.upsample_argmax(lg, (1024, 1024)), iters=10)
        report(f'C4 upsample+argmax (pred only) K=12 {name}', t, Tc * (12 * hw * hw * 4 + 1024 * 1024), Tc)
    # ---- training mode: head forward + backward at the ft_pop shapes (batch 1 novel + 1 base image,
    # 1024^2 crops, scripts/ft_oem.sh), device time of every launch of one autograd step
    for name, C, hw in (('PSPNet C=512', 512, 128), ('Swin-T/S C=96', 96, 256), ('ConvNeXt-T C=192', 192, 256)):
        stc = synth.make_head_state(C, 7, 4, seed=3).to(dev)
        f = synth.make_random_features(2, C, hw, hw, seed=3).to(dev)
        gout = torch.randn(2, 12, hw, hw, device=dev, generator=g)
        novel = stc.novel_emb.clone().requires_grad_(True)
        cls_n = tuple(w.clone().requires_grad_(True) for w in stc.cls_n)

        def step(mode='auto', need_dfeat=False):
            for w in (novel, *cls_n):
                w.grad = None
            x = f.float().requires_grad_(need_dfeat)
            ops.pop_head_train(x, stc.base_emb, stc.cls, novel, cls_n, bg_mode=mode).backward(gout)

        flops = 10 * C * C * 2 * hw * hw
        for label_, kw in (('tcgen05', {}), ('tcgen05 + d_feat', {'need_dfeat': True}), ('exact fp32', {'mode': 'simt'})):
            t = timeit(lambda: step(**kw), iters=10)
            print(f'{"train head fwd+bwd " + name + " (" + label_ + ")":58s} {t * 1e3:9.3f} ms/step '
                  f'{flops / t / 1e12:7.1f} TFLOP/s fp32-equivalent (10 C^2 N)')
    # ---- C4: flip test-time augmentation at feature resolution (orig + h-flip views -> mean logits)
    views = torch.randn(2, 32, 12, 256, 256, device=dev, generator=g)
    t = timeit(lambda: ops.aggregate_views(views, [0, 1]))
    report('C4 aggregate_views V=2 (orig + h-flip) K=12 256^2', t, 3 * 32 * 12 * 256 * 256 * 4, 32)
    del views
    # ---- C3: pseudo-labelling of base images (pspnet_pop.py:221-231), 5 classifier_n channels -> int64 masks
    p2 = torch.randn(8, 5, 128, 128, device=dev, generator=g)
    mk = torch.zeros(8, 1024, 1024, dtype=torch.int64, device=dev)
    t = timeit(lambda: ops.pseudo_label(p2, mk.zero_(), 7))
    report('C3 pseudo_label 1+Kn=5, 128^2 -> 1024^2 int64 masks (+memset)', t, 8 * (3 * 8 * 1024 * 1024), 8)
    # ---- f-1: cross-entropy of the up-sampled logits without materialising them (loss/criterion.py:51-52)
    for name, hw in (('128^2 -> 1024^2', 128), ('256^2 -> 1024^2', 256)):
        lgt = torch.randn(8, 12, hw, hw, device=dev, generator=g).requires_grad_(True)
        tgt = torch.randint(0, 12, (8, 1024, 1024), device=dev, generator=g)

        def ce_step():
            lgt.grad = None
            ops.seg_cross_entropy(lgt, tgt).backward()
        t = timeit(ce_step, iters=10)
        # algorithmic bytes: int64 target once per direction + low-res logits / gradients
        report(f'seg_cross_entropy fwd+bwd K=12 {name} (int64 labels)', t, 8 * (2 * 8 * 1024 * 1024 + 3 * 12 * hw * hw * 4), 8)
    # ---- metric kernels on 1024^2 label maps (32 tiles)
    gt = torch.randint(0, 12, (32, 1024, 1024), device=dev, dtype=torch.uint8, generator=g)
    pr = torch.randint(0, 12, (32, 1024, 1024), device=dev, dtype=torch.uint8, generator=g)
    t = timeit(lambda: ops.confusion_update(cm, gt, pr))
    report('confusion_update 32 x 1024^2 (random labels: worst case)', t, 2 * gt.numel(), 32)
    gt_c = (torch.arange(32 * 1024 * 1024, device=dev) // 4096 % 12).to(torch.uint8).view(32, 1024, 1024)
    t = timeit(lambda: ops.confusion_update(cm, gt_c, gt_c))
    report('confusion_update 32 x 1024^2 (coherent labels)', t, 2 * gt.numel(), 32)
    o64, t64 = pr.long(), gt.long()
    t = timeit(lambda: ops.intersectionAndUnionGPU(o64, t64, 12))
    report('intersectionAndUnionGPU 32 x 1024^2 int64', t, 2 * 8 * gt.numel(), 32)


if __name__ == '__main__':
    main()
