"""GPU parity tests for the decoder tails: sl_tail_layernorm, sl_tail_bn_relu_conv,
sl_tail_sum through segland_b200.ops, against the reference-generated golden file (tests/golden/tails.npz, written
by oracle/gen_golden_tails.py from the reference's own decoders) and the CPU oracle.

The tails emit bf16 features; the reference keeps fp32.  Parity statement (written here, as north_star asks):
  * |got - ref_fp32| <= 0.5 bf16 spacings of the reference element (ideal rounding) + ABS_SLACK * rms(ref), where
    the slack covers the fp32-level differences of the arithmetic before the rounding (1e-6 for the fp32 kernels,
    2e-5 for the split-bf16 tensor-core convolution, whose products carry ~1e-5);
  * at least 99 % (fp32 kernels: 99.9 %) of the bit patterns equal round-to-nearest-even of the reference;
  * the POP head's logits on these features match the head's logits on bf16(reference features) within 1e-3.
"""
import numpy as np
import pytest
import torch

from helpers import RTOL, assert_close_rel, rel_err
from oracle import ref_ops
from segland_b200 import synth

pytestmark = pytest.mark.gpu


@pytest.fixture(scope='module')
def ops():
    from segland_b200 import ops as _ops
    _ops.check_device()
    return _ops


def check_features(got, ref32, slack, min_exact, what):
    assert got.dtype == torch.bfloat16 and got.shape == ref32.shape and got.is_contiguous()
    got = got.cpu()
    exact, d = ref_ops.bf16_ulp_report(got, ref32)
    rms = ref32.double().pow(2).mean().sqrt()
    spacing = torch.pow(2.0, torch.floor(torch.log2(ref32.double().abs().clamp_min(1e-30))) - 7)
    err = (got.double() - ref32.double()).abs()
    worst = (err - (0.5 * spacing * (1 + 1e-9) + slack * rms)).max().item()
    assert worst <= 0, f'{what}: rounding bound exceeded by {worst:.3e} (max {d.max().item():.3f} spacings)'
    assert exact >= min_exact, f'{what}: only {exact:.5f} of the bf16 patterns are the rounded reference'
    return exact


def assert_logits_close_rerounded(a, b, what, rtol=RTOL):
    """Logits of the head on two feature tensors that differ by bf16 re-rounding (<= 1 % of the elements one spacing
    apart): within north_star's 1e-3 of the tensor maximum, rms difference below 1e-3 of the rms logit.  The
    per-element bound of helpers.assert_close_rel (1e-3 |b| + 1e-3 rms) is a statement about identical inputs: one
    flipped feature moves a logit by up to 2^-8 |q_i w_i|, a few per pixel add in quadrature, and the worst of
    ~10^5 logits sits at a few 1e-3 rms."""
    a, b = a.double().cpu(), b.double().cpu()
    assert rel_err(a, b) <= rtol, f'{what}: rel-to-max {rel_err(a, b):.3e}'
    rms_d, rms_b = (a - b).pow(2).mean().sqrt().item(), b.pow(2).mean().sqrt().item()
    assert rms_d <= rtol * rms_b, f'{what}: rms difference {rms_d:.3e} vs rms logit {rms_b:.3e}'


@pytest.mark.parametrize('name', ['ln_c192', 'ln_c96', 'ln_c480', 'ln_c40'])
def test_layernorm_tail_vs_golden(ops, golden, name):
    z = golden('tails')
    x = torch.from_numpy(z[name + '_x'])
    got = ops.layernorm_tail(x.cuda(), torch.from_numpy(z[name + '_gamma']), torch.from_numpy(z[name + '_beta']),
                             float(z[name + '_eps']))
    torch.cuda.synchronize()
    check_features(got, torch.from_numpy(z[name + '_out']), 1e-6, 0.999, name)


@pytest.mark.parametrize('C,B,h,w', [(192, 2, 64, 64), (96, 1, 40, 56), (128, 3, 8, 9 * 8), (480, 1, 24, 24), (960, 1, 8, 8),
                                     (1536, 1, 4, 6), (8, 2, 16, 16)])
def test_layernorm_tail_vs_oracle(ops, C, B, h, w):
    g = torch.Generator().manual_seed(C + h)
    x = torch.randn(B, C, h, w, generator=g) * (0.5 + torch.rand(B, 1, h, w, generator=g) * 3) + torch.randn(B, 1, h, w, generator=g)
    gamma = 1 + 0.3 * torch.randn(C, generator=g)
    beta = 0.2 * torch.randn(C, generator=g)
    ref = ref_ops.ref_tail_layernorm(x, gamma, beta, 1e-5).contiguous()
    out = torch.full((B, C, h, w), 7.0, dtype=torch.bfloat16, device='cuda')
    got = ops.layernorm_tail(x.cuda(), gamma, beta, 1e-5, out=out)
    torch.cuda.synchronize()
    assert got.data_ptr() == out.data_ptr()
    check_features(got, ref, 1e-6, 0.999, f'ln C={C}')


def test_layernorm_tail_constant_pixels_and_errors(ops):
    # a pixel whose channels are all equal has zero variance: y = beta exactly (eps keeps rstd finite)
    x = torch.full((1, 64, 4, 8), 3.25)
    gamma, beta = torch.ones(64) * 2, torch.arange(64).float() / 8
    got = ops.layernorm_tail(x.cuda(), gamma, beta, 1e-5).cpu().float()
    assert torch.equal(got, beta.view(1, 64, 1, 1).expand_as(got).to(torch.bfloat16).float())
    with pytest.raises(ValueError):
        ops.layernorm_tail(torch.zeros(1, 64, 3, 3).cuda(), gamma, beta)


@pytest.mark.parametrize('name', ['psp_c64', 'psp_c512', 'psp_c96'])
def test_conv_tail_vs_golden(ops, golden, name):
    z = golden('tails')
    t = lambda k: torch.from_numpy(z[name + k])
    tail = ops.ConvTail(t('_W'), t('_bias'), bn=(t('_bn_w'), t('_bn_b'), t('_bn_m'), t('_bn_v'), float(z[name + '_bn_eps'])))
    got = tail(t('_x').cuda())
    torch.cuda.synchronize()
    check_features(got, t('_out'), 2e-5, 0.99, name)


@pytest.mark.parametrize('Cin,Cout,B,h,w', [(512, 512, 2, 32, 32), (512, 512, 1, 128, 128), (96, 512, 1, 16, 24),
                                             (512, 96, 2, 8, 24), (40, 24, 1, 8, 8), (256, 320, 1, 20, 20)])
def test_conv_tail_vs_oracle(ops, Cin, Cout, B, h, w):
    g = torch.Generator().manual_seed(Cin + Cout + h)
    x = torch.randn(B, Cin, h, w, generator=g)
    bn = (1 + 0.3 * torch.randn(Cin, generator=g), 0.2 * torch.randn(Cin, generator=g), 0.2 * torch.randn(Cin, generator=g),
          0.5 + torch.rand(Cin, generator=g), 1e-5)
    W = torch.randn(Cout, Cin, generator=g) / Cin ** 0.5
    bias = 0.1 * torch.randn(Cout, generator=g)
    ref = ref_ops.ref_tail_bn_relu_conv(x, bn[0], bn[1], bn[2], bn[3], bn[4], W, bias)
    got = ops.ConvTail(W, bias, bn=bn)(x.cuda())
    torch.cuda.synchronize()
    check_features(got, ref, 2e-5, 0.99, f'conv {Cin}->{Cout}')


def test_conv_tail_options_and_integer_exactness(ops):
    """No BN / no ReLU / no bias variants; small-integer operands make every product and sum exact, so the bf16
    output must be bit-identical to the rounded reference."""
    g = torch.Generator().manual_seed(5)
    x = torch.randint(-3, 4, (2, 64, 8, 16), generator=g).float()
    W = torch.randint(-2, 3, (128, 64), generator=g).float()
    ref = ref_ops.ref_tail_bn_relu_conv(x, None, None, None, None, 0.0, W, None, relu=False)
    got = ops.ConvTail(W, None, bn=None, relu=False)(x.cuda()).cpu()
    assert torch.equal(got, ref.to(torch.bfloat16))
    ref = ref_ops.ref_tail_bn_relu_conv(x, None, None, None, None, 0.0, W, None, relu=True)
    got = ops.ConvTail(W, None, bn=None, relu=True)(x.cuda()).cpu()
    assert torch.equal(got, ref.to(torch.bfloat16))


def test_conv_tail_from_sequential_mirrors_reference_module(ops):
    """from_sequential on a module with the reference's bottleneck structure (pspnet_pop.py:18-23), refresh() after an
    in-place weight update."""
    torch.manual_seed(3)
    seq = torch.nn.Sequential(torch.nn.Conv2d(48, 64, 3, padding=1, bias=False), torch.nn.BatchNorm2d(64),
                              torch.nn.ReLU(inplace=True), torch.nn.Conv2d(64, 64, 1)).eval()
    seq[1].running_mean.normal_(0, 0.2); seq[1].running_var.uniform_(0.5, 1.5)
    inp = torch.randn(1, 48, 16, 16)
    with torch.no_grad():
        x = seq[0](inp)
        ref = seq[1:](x.clone())
    tail = ops.ConvTail.from_sequential(seq)
    check_features(tail(x.cuda()), ref, 2e-5, 0.99, 'from_sequential')
    with torch.no_grad():
        seq[3].weight.mul_(-0.5)
        ref2 = seq[1:](x.clone())
    tail.refresh()
    check_features(tail(x.cuda()), ref2, 2e-5, 0.99, 'after refresh')
    with pytest.raises(ValueError):
        ops.ConvTail.from_sequential(seq.train())


def test_sum_tail_vs_golden_and_oracle(ops, golden):
    z = golden('tails')
    maps = [torch.from_numpy(z[f'sum_swin_map{i}']) for i in range(4)]
    got = ops.sum_tail([m.cuda() for m in maps]).cpu()
    assert torch.equal(got, torch.from_numpy(z['sum_swin_out']).to(torch.bfloat16))      # same adds, same order
    g = torch.Generator().manual_seed(9)
    for M in (1, 2, 8):
        maps = [torch.randn(2, 96, 24, 40, generator=g) * 10 ** (m % 3) for m in range(M)]
        got = ops.sum_tail([m.cuda() for m in maps]).cpu()
        assert torch.equal(got, ref_ops.ref_tail_sum(maps).to(torch.bfloat16)), M
    with pytest.raises(Exception):
        ops.sum_tail([torch.zeros(1, 8, 8, 8).cuda()] * 9)


def test_bn_relu_tail_vs_golden_and_oracle(ops, golden):
    z = golden('tails')
    t = lambda k: torch.from_numpy(z['aspp' + k])
    bn = (t('_bn_w'), t('_bn_b'), t('_bn_m'), t('_bn_v'), float(z['aspp_bn_eps']))
    got = ops.bn_relu_tail(t('_x').cuda(), bn)
    torch.cuda.synchronize()
    check_features(got, t('_out'), 1e-6, 0.999, 'aspp')
    g = torch.Generator().manual_seed(41)
    for C, B, h, w in ((256, 2, 64, 64), (64, 1, 40, 56), (8, 3, 2, 4)):
        x = torch.randn(B, C, h, w, generator=g) * 2
        bn = (1 + 0.3 * torch.randn(C, generator=g), 0.2 * torch.randn(C, generator=g), 0.2 * torch.randn(C, generator=g),
              0.5 + torch.rand(C, generator=g), 1e-5)
        for relu in (True, False):
            ref = ref_ops.ref_tail_bn_relu(x, bn[0], bn[1], bn[2], bn[3], bn[4], relu=relu)
            check_features(ops.bn_relu_tail(x.cuda(), bn, relu=relu), ref, 1e-6, 0.999, f'bn_relu C={C}')
    mod = torch.nn.BatchNorm2d(8)
    with pytest.raises(ValueError):
        ops.bn_relu_tail(torch.zeros(1, 8, 2, 4).cuda(), mod)          # train mode: no running-statistics fold


def test_concat_tail_vs_golden_and_oracle(ops, golden):
    z = golden('tails')
    maps = [torch.from_numpy(z[f'cat_hr_map{i}']) for i in range(4)]
    got = ops.concat_tail([m.cuda() for m in maps]).cpu()
    assert torch.equal(got, torch.from_numpy(z['cat_hr_out']).to(torch.bfloat16))       # a pure conversion: bit-exact
    g = torch.Generator().manual_seed(43)
    maps = [torch.randn(3, c, 24, 40, generator=g) for c in (32, 64, 128, 256)]        # HRNet-w32's 480 channels
    out = torch.zeros(3, 480, 24, 40, dtype=torch.bfloat16, device='cuda')
    got = ops.concat_tail([m.cuda() for m in maps], out=out)
    assert torch.equal(got.cpu(), ref_ops.ref_tail_concat(maps).to(torch.bfloat16))
    assert torch.equal(ops.concat_tail([maps[1].cuda()]).cpu(), maps[1].to(torch.bfloat16))
    with pytest.raises(ValueError):
        ops.concat_tail([maps[0].cuda(), maps[1][:2].cuda()])


@pytest.mark.parametrize('kind', ['psp', 'ln'])
def test_head_on_tail_features_matches_head_on_reference_features(ops, kind):
    """End of the chain: logits of the POP head fed by the fused tail vs fed by bf16(oracle tail output)."""
    g = torch.Generator().manual_seed(17)
    C, B, h, w = (512, 1, 32, 32) if kind == 'psp' else (192, 1, 32, 48)
    st = synth.make_head_state(C, 7, 4, seed=17)
    x = torch.randn(B, C, h, w, generator=g)
    if kind == 'psp':
        bn = (1 + 0.3 * torch.randn(C, generator=g), 0.2 * torch.randn(C, generator=g), 0.2 * torch.randn(C, generator=g),
              0.5 + torch.rand(C, generator=g), 1e-5)
        W, bias = torch.randn(C, C, generator=g) / C ** 0.5, 0.1 * torch.randn(C, generator=g)
        ref = ref_ops.ref_tail_bn_relu_conv(x, bn[0], bn[1], bn[2], bn[3], bn[4], W, bias)
        feats = ops.ConvTail(W, bias, bn=bn)(x.cuda())
    else:
        gamma, beta = 1 + 0.3 * torch.randn(C, generator=g), 0.2 * torch.randn(C, generator=g)
        ref = ref_ops.ref_tail_layernorm(x, gamma, beta, 1e-5).contiguous()
        feats = ops.layernorm_tail(x.cuda(), gamma, beta, 1e-5)
    head = ops.PopHead(st.base_emb, st.cls, st.novel_emb, st.cls_n)
    mine = head(feats)
    theirs = head(ref.to(torch.bfloat16).cuda())
    oracle = ref_ops.ref_head(ref.to(torch.bfloat16).float(), st.base_emb, st.novel_emb, st.cls, st.cls_n)
    torch.cuda.synchronize()
    assert_logits_close_rerounded(mine, theirs, f'{kind}: head(tail) vs head(bf16(ref))')
    assert_logits_close_rerounded(mine, oracle, f'{kind}: head(tail) vs oracle head')


def test_tails_full_size_properties(ops):
    """BASELINE sizes (PSPNet 1024^2 tile: C=512, 128x128; ConvNeXt-T: C=192, 256x256): size-independent properties.
    LayerNorm: per-pixel mean/variance of (y - beta)/gamma are 0 / 1; batch independence.  Conv tail: linearity in
    the bias (exact shift by a power of two of a zero-weight tail) and batch independence."""
    g = torch.Generator().manual_seed(23)
    x = torch.randn(2, 192, 256, 256, generator=g).cuda()
    gamma, beta = torch.ones(192), torch.zeros(192)
    y = ops.layernorm_tail(x, gamma, beta, 1e-5)
    yf = y.float()
    assert yf.mean(1).abs().max().item() < 2e-3 and (yf.var(1, unbiased=False) - 1).abs().max().item() < 5e-3
    assert torch.equal(ops.layernorm_tail(x[1:], gamma, beta, 1e-5), y[1:])
    x = torch.randn(2, 512, 128, 128, generator=g).cuda()
    W = torch.randn(512, 512, generator=g) / 512 ** 0.5
    tail = ops.ConvTail(W, torch.zeros(512), bn=None)
    y = tail(x)
    assert torch.equal(tail(x[1:].contiguous()), y[1:])
    zero = ops.ConvTail(torch.zeros(512, 512), torch.full((512,), 0.5), bn=None)(x)
    assert torch.equal(zero.float(), torch.full_like(zero, 0.5).float())
    # against a float64 evaluation on the device (no CPU minutes): rounding bound as in check_features
    ref = (W.cuda().double() @ torch.relu(x[0]).double().flatten(1)).view(512, 128, 128)
    err = (y[0].double() - ref).abs()
    spacing = torch.pow(2.0, torch.floor(torch.log2(ref.abs().clamp_min(1e-30))) - 7)
    assert (err - 0.5 * spacing - 2e-5 * ref.pow(2).mean().sqrt()).max().item() <= 0


def test_patch_runs_decoder_tails_fused(ops):
    """segland_b200.patch with reference-shaped decoders (the real tree is not on the GPU box; the class names,
    attribute names and forward bodies follow networks/pspnet_pop.py:8-35 and networks/convnext_pop.py:8-28): in
    eval mode the patched forward swaps the tail in for the call, the head receives bf16 features, and the logits
    match the un-fused path; the module tree is untouched afterwards."""
    import sys
    import types
    import torch.nn as nn
    import torch.nn.functional as F
    from segland_b200 import patch as slp

    class PSPModule(nn.Module):
        def __init__(self, features, out_features, sizes=(1, 2)):
            super().__init__()
            self.stages = nn.ModuleList([nn.Sequential(nn.AdaptiveAvgPool2d(s), nn.Conv2d(features, out_features, 1, bias=False),
                                                       nn.BatchNorm2d(out_features), nn.ReLU(inplace=True)) for s in sizes])
            self.bottleneck = nn.Sequential(nn.Conv2d(features + len(sizes) * out_features, out_features, 3, padding=1, bias=False),
                                            nn.BatchNorm2d(out_features), nn.ReLU(inplace=True),
                                            nn.Conv2d(out_features, out_features, 1))

        def forward(self, feats):
            h, w = feats.shape[2:]
            priors = [F.interpolate(st(feats), size=(h, w), mode='bilinear', align_corners=False) for st in self.stages] + [feats]
            return self.bottleneck(torch.cat(priors, 1))

    class FPN_Seg_OCR_Decoder(nn.Module):
        def __init__(self, in_ch, out_ch):
            super().__init__()
            self.conv = nn.Conv2d(in_ch, out_ch, 1)
            self.norm = nn.LayerNorm(out_ch)

        def forward(self, x):
            feats = self.conv(x)
            return self.norm(feats.permute(0, 2, 3, 1)).permute(0, 3, 1, 2)

    class Backbone(nn.Module):
        def base_forward(self, x):
            return x

        forward = base_forward

    def make_model(modname, decoder, C):
        st = synth.make_head_state(C, 7, 4, seed=C)

        def mlp(ws):
            seq = nn.Sequential(nn.Conv2d(C, C, 1, bias=False), nn.ReLU(inplace=True), nn.Conv2d(C, C, 1, bias=False),
                                nn.ReLU(inplace=True), nn.Conv2d(C, 1, 1, bias=False))
            with torch.no_grad():
                seq[0].weight.copy_(ws[0].view(C, C, 1, 1)); seq[2].weight.copy_(ws[1].view(C, C, 1, 1))
                seq[4].weight.copy_(ws[2].view(1, C, 1, 1))
            return seq

        class GFSS_Model(nn.Module):
            def __init__(self):
                super().__init__()
                self.backbone, self.decoder = Backbone(), decoder
                self.classifier, self.classifier_n = mlp(st.cls), mlp(st.cls_n)
                self.base_emb = nn.Parameter(st.base_emb.clone(), requires_grad=False)
                self.novel_emb = nn.Parameter(st.novel_emb.clone())
                self.is_ft, self.criterion = True, None

            def forward(self, img, mask=None, img_b=None, mask_b=None):
                return 'reference path'

        GFSS_Model.__module__ = 'networks.' + modname
        return GFSS_Model

    torch.manual_seed(11)
    psp = PSPModule(32, 64)
    for m in psp.modules():
        if isinstance(m, nn.BatchNorm2d):
            m.running_mean.normal_(0, 0.2); m.running_var.uniform_(0.5, 1.5)
    fpn = FPN_Seg_OCR_Decoder(48, 192)
    classes = {'pspnet_pop': make_model('pspnet_pop', psp, 64), 'convnext_pop': make_model('convnext_pop', fpn, 192)}
    names = ['networks'] + ['networks.' + n for n in classes]
    saved = {k: sys.modules.get(k) for k in names}
    sys.modules['networks'] = types.ModuleType('networks')
    for n, cls in classes.items():
        mod = types.ModuleType('networks.' + n)
        mod.GFSS_Model = cls
        sys.modules['networks.' + n] = mod
    tf32 = torch.backends.cudnn.allow_tf32
    torch.backends.cudnn.allow_tf32 = False          # the stock 3x3 / 1x1 convolutions in fp32, as on the CPU
    try:
        for n, cls in classes.items():
            model = cls().cuda().eval()
            img = torch.randn(2, 32 if n == 'pspnet_pop' else 48, 16, 24, device='cuda')
            keys = list(model.state_dict())
            slp.patch(tails=False)
            plain = model(img)
            slp.unpatch()
            slp.patch()
            seen = []
            head = slp.head_for(model)
            orig_call = type(head).__call__
            type(head).__call__ = lambda self, f, *a, **k: (seen.append(f.dtype), orig_call(self, f, *a, **k))[1]
            try:
                fused = model(img)
            finally:
                type(head).__call__ = orig_call
            slp.unpatch()
            assert seen == [torch.bfloat16], (n, seen)                   # the tail handed bf16 features to the head
            assert_logits_close_rerounded(fused, plain, f'{n}: fused-tail logits vs stock decoder + head')
            assert list(model.state_dict()) == keys
            assert isinstance(model.decoder.bottleneck if n == 'pspnet_pop' else model.decoder.norm,
                              nn.Sequential if n == 'pspnet_pop' else nn.LayerNorm)
            # the reference computation end to end on the CPU: decoder in fp32 -> bf16 features -> oracle head
            with torch.no_grad():
                f32 = model.decoder.cpu()(img.cpu())
            model.cuda()
            ref = ref_ops.ref_head(f32.to(torch.bfloat16).float(), model.base_emb.detach().cpu(), model.novel_emb.detach().cpu(),
                                   tuple(w.detach().cpu().reshape(w.shape[0], -1).squeeze(0) for w in
                                         (model.classifier[0].weight, model.classifier[2].weight, model.classifier[4].weight)),
                                   tuple(w.detach().cpu().reshape(w.shape[0], -1).squeeze(0) for w in
                                         (model.classifier_n[0].weight, model.classifier_n[2].weight, model.classifier_n[4].weight)))
            assert_logits_close_rerounded(fused, ref, f'{n}: patched model vs CPU decoder + oracle head', rtol=2e-3)
            # a decoder left in train mode is not touched
            model.decoder.train()
            slp.patch()
            seen.clear()
            type(head).__call__ = lambda self, f, *a, **k: (seen.append(f.dtype), orig_call(self, f, *a, **k))[1]
            try:
                model(img)
            finally:
                type(head).__call__ = orig_call
                slp.unpatch()
            assert seen == [torch.float32]
    finally:
        torch.backends.cudnn.allow_tf32 = tf32
        slp.unpatch()
        for k, v in saved.items():
            if v is None:
                sys.modules.pop(k, None)
            else:
                sys.modules[k] = v
