"""Drop-in installation into an unmodified SegLand checkout.

    import segland_b200.patch as slp
    slp.patch()          # after `sys.path` contains the SegLand tree, before building models

replaces, in place,
  * `GFSS_Model.forward` of every importable `networks/*_pop.py` model: in eval mode
    (`eval_base.py:167`, `eval_ft.py:167`, `ft_pop.py:327`, `train_base.py:331`) the head after the
    decoder -- `orthogonal_decompose` + `classifier` / `classifier_n` + channel assembly
    (`pspnet_pop.py:143-159`, `:171-182`) -- runs in libsegland_b200.so; backbones and decoders stay
    stock PyTorch, except that in inference the decoder's last operators (PSPModule.bottleneck[1:4],
    `pspnet_pop.py:19-22`; FPN_Seg_OCR_Decoder.norm, `convnext_pop.py:27`; _ASPP.fc's BN + ReLU,
    `deeplab_pop.py:61`; VGGUNet.up4's last BN + ReLU, `vggunet_pop.py:79`) run fused and emit the head's bf16
    features directly (`patch(tails=False)` disables it).  Training-mode calls on CUDA tensors -- `forward_novel` (`ft_pop.py:252`,
    `pspnet_pop.py:191-245`) and `forward_base` with a criterion (`train_base.py:259`, `:161-189`) -- run
    the same head with autograd (`ops.forward_novel_train` / `ops.forward_base_train`: forward kernels +
    `sl_pop_head_bwd`) when `patch(train=True)` is given, so gradients reach `novel_emb`, `classifier_n`,
    `classifier`, `base_emb` and the decoder as in the reference.  It is OPT-IN: the kernels read bf16 features
    (north_star's format), whereas the reference's `orthogonal_decompose` up-casts to fp32 under
    `autocast(enabled=False)` (`pspnet_pop.py:95,105`), so training numerics differ at the 2^-9 level of the feature
    rounding; the default `patch()` leaves training-mode forwards on the reference.
  * `utils.pyt_utils.get_confusion_matrix` and `utils.pyt_utils.intersectionAndUnionGPU` -- in the module AND in
    every already-imported module that bound them with `from utils.pyt_utils import ...` (`eval_base.py:16`,
    `eval_ft.py:16`, `ft_pop.py`, `train_base.py` do), so the order of `patch()` and the script's imports does not
    matter;
  * `loss.criterion.OrthLoss.forward` and `OrthLoss.get_orth_loss`: the seg / aux cross-entropy terms run fused
    with the up-sampling (`loss/criterion.py:51-52,57-58`) and the orthogonality term in one launch
    (`loss/criterion.py:37-43`), both differentiable, so training loops keep working.
Models whose channel count the head kernels do not take (`ops.PopHead.supports`: C % 8 == 0, C <= 512 -- of the
reference's backbones only seghr_pop with hr-w18, d_model 270, and hr-w48, d_model 720) stay on the reference's own
forward.  What `patch()` does NOT replace: the scripts' inline post-processing (`eval_base.py:168-178`,
`eval_ft.py:168-183`: `F.interpolate` -> `.cpu()` -> `np.argmax`) is code inside the scripts' loops, not a function,
so it keeps running as written; the fused up-sample/argmax/confusion path needs the three-line edit shown in
INTEGRATION.md (`sweep.TileEvaluator`).  `unpatch()` restores the originals.
"""
from __future__ import annotations

import importlib
import sys

import torch

from . import ops

MODEL_MODULES = ('pspnet_pop', 'pspplus_pop', 'deeplab_pop', 'vggunet_pop', 'seghr_pop', 'swin_pop',
                 'convnext_pop', 'lsk_pop')
_BASE_FORWARD = ('pspnet_pop', 'pspplus_pop', 'deeplab_pop')     # ResNet backbones: backbone.base_forward(img)
_originals = []


def _features(model, img, fused_tail=False):
    """The part of forward_all/forward_base before the head (`pspnet_pop.py:141-142` and siblings).
    fused_tail (inference only): the decoder's last operators run in libsegland_b200.so and hand the head bf16
    features directly."""
    if hasattr(model, 'net') and not hasattr(model, 'decoder'):           # vggunet_pop.py
        return _decode_fused(model.net, img) if fused_tail and _tails_enabled[0] else model.net(img)
    name = type(model).__module__.rsplit('.', 1)[-1]
    feats = model.backbone.base_forward(img) if name in _BASE_FORWARD else model.backbone(img)
    if fused_tail and _tails_enabled[0]:
        return _decode_fused(model.decoder, feats)
    return model.decoder(feats)


class _ConvTailSeq(torch.nn.Module):
    """Stands in for `[conv3x3, norm, ReLU, conv1x1]` (PSPModule.bottleneck, `pspnet_pop.py:18-23`;
    PSP_Plus_Decoder.fc, `pspplus_pop.py:44-47`) during one inference call: the reference's own 3x3 convolution, then
    BN -> ReLU -> 1x1 conv + bias -> bf16 in one C-ABI call."""

    def __init__(self, seq):
        super().__init__()
        self.seq = seq
        self.sig = None
        self.tail = None

    def forward(self, x):
        x = self.seq[0](x)
        if (x.shape[2] * x.shape[3]) % 8 or not x.is_cuda:
            return self.seq[3](self.seq[2](self.seq[1](x)))
        params = [self.seq[1].weight, self.seq[1].bias, self.seq[1].running_mean, self.seq[1].running_var,
                  self.seq[3].weight, self.seq[3].bias]
        sig = tuple((p.data_ptr(), p._version) for p in params)
        if self.sig != sig:
            self.tail, self.sig = ops.ConvTail.from_sequential(self.seq), sig
        return self.tail(x)


class _LayerNormTail(torch.nn.Module):
    """Stands in for FPN_Seg_OCR_Decoder.norm (`convnext_pop.py:13`), which the decoder applies to the channels-last
    VIEW of the convolution output and permutes back (`:27`): undo the view, run the fused LayerNorm on the NCHW
    tensor, return the matching view so the decoder's own permute yields contiguous bf16 NCHW features."""

    def __init__(self, norm):
        super().__init__()
        self.norm = norm

    def forward(self, x_nhwc):
        x = x_nhwc.permute(0, 3, 1, 2)
        if not x.is_cuda or not x.is_contiguous() or (x.shape[2] * x.shape[3]) % 8:
            return self.norm(x_nhwc)
        return ops.layernorm_tail(x, self.norm.weight, self.norm.bias, self.norm.eps).permute(0, 2, 3, 1)


class _BnReluTailSeq(torch.nn.Module):
    """Stands in for a Sequential that ends `..., conv, BatchNorm2d, ReLU` (_ASPP.fc = `_ConvBnReLU`,
    `deeplab_pop.py:12-29,61`; VGGUNet.up4.conv.double_conv, `vggunet_pop.py:13-20`): everything up to and including
    the last convolution runs as it is, then BN -> ReLU -> bf16 in one C-ABI call."""

    def __init__(self, seq):
        super().__init__()
        self.seq = seq

    def forward(self, x):
        mods = list(self.seq)
        for m in mods[:-2]:
            x = m(x)
        if (x.shape[2] * x.shape[3]) % 8 or not x.is_cuda:
            return mods[-1](mods[-2](x))
        return ops.bn_relu_tail(x, mods[-2], relu=True)


def _ends_conv_bn_relu(seq):
    mods = list(seq) if isinstance(seq, torch.nn.Sequential) else []
    return (len(mods) >= 3 and isinstance(mods[-3], torch.nn.Conv2d) and isinstance(mods[-2], torch.nn.BatchNorm2d)
            and isinstance(mods[-1], torch.nn.ReLU))


def _swap_spec(dec):
    """(owner module, attribute name, stand-in module) for decoders whose tail the library fuses, else None."""
    spec = _swap_spec_local(dec)
    if spec is not None:
        return (dec,) + spec
    kind = type(dec).__name__
    if kind == 'VGGUNet':                                               # vggunet_pop.py:79 -> Up.conv -> DoubleConv
        dc = getattr(getattr(getattr(dec, 'up4', None), 'conv', None), 'double_conv', None)
        if _ends_conv_bn_relu(dc):
            return dec.up4.conv, 'double_conv', _BnReluTailSeq(dc)
    return None


def _swap_spec_local(dec):
    kind = type(dec).__name__
    if kind == '_ASPP' and _ends_conv_bn_relu(getattr(dec, 'fc', None)):
        return 'fc', _BnReluTailSeq(dec.fc)
    attr = {'PSPModule': 'bottleneck', 'PSP_Plus_Decoder': 'fc'}.get(kind)
    if attr is not None:
        seq = getattr(dec, attr, None)
        if (isinstance(seq, torch.nn.Sequential) and len(seq) == 4 and isinstance(seq[1], torch.nn.BatchNorm2d)
                and isinstance(seq[2], torch.nn.ReLU) and isinstance(seq[3], torch.nn.Conv2d)
                and seq[3].kernel_size == (1, 1) and seq[3].in_channels % 8 == 0):
            return attr, _ConvTailSeq(seq)
    if kind == 'FPN_Seg_OCR_Decoder' and isinstance(getattr(dec, 'norm', None), torch.nn.LayerNorm):
        return 'norm', _LayerNormTail(dec.norm)
    return None


def _decode_fused(dec, feats):
    """`self.decoder(features)` with the decoder's tail swapped for its fused stand-in for the duration of the call
    (the module tree, state_dict and parameters are untouched afterwards).  Decoders in train mode, and decoders
    without a fused tail, run as they are."""
    if dec.training:
        return dec(feats)
    cached = dec.__dict__.get('_sl_tail')
    if cached is None:
        cached = _swap_spec(dec) or ()
        dec.__dict__['_sl_tail'] = cached
    if not cached:
        return dec(feats)
    owner, attr, stand_in = cached
    original = owner._modules[attr]
    owner._modules[attr] = stand_in
    try:
        return dec(feats)
    finally:
        owner._modules[attr] = original


def _head_params(model):
    ps = [model.base_emb] + list(model.classifier.parameters())
    if getattr(model, 'is_ft', False):
        ps += [model.novel_emb] + list(model.classifier_n.parameters())
    return ps


def head_for(model, **kw):
    """The cached PopHead of a GFSS_Model, rebuilt when any head parameter was modified in place
    (optimizer steps bump tensor._version) or re-assigned (load_state_dict copies in place too)."""
    sig = tuple((p.data_ptr(), p._version, str(p.device)) for p in _head_params(model))
    cached = getattr(model, '_sl_head', None)
    if cached is None or cached[0] != sig:
        head = ops.PopHead.from_model(model, device=model.base_emb.device, **kw)
        object.__setattr__(model, '_sl_head', (sig, head))
        return head
    return cached[1]


def _mlp_weights(seq):
    return (seq[0].weight, seq[2].weight, seq[4].weight)


def _make_forward(orig_forward):
    def forward(self, img, mask=None, img_b=None, mask_b=None):
        if not img.is_cuda or not ops.PopHead.supports(self.base_emb.shape[-1]):
            return orig_forward(self, img, mask, img_b, mask_b)      # CPU tensors, or a width the kernels do not take
        is_ft = bool(getattr(self, 'is_ft', False))
        if is_ft and self.training:                                   # forward_novel (pspnet_pop.py:191-245)
            if not _train_enabled[0] or img_b is None or mask_b is None:
                return orig_forward(self, img, mask, img_b, mask_b)
            feats = _features(self, torch.cat([img, img_b], dim=0))
            return ops.forward_novel_train(feats, mask, mask_b, self.base_emb, self.novel_emb,
                                           _mlp_weights(self.classifier), _mlp_weights(self.classifier_n),
                                           criterion=self.criterion)
        wants_loss = self.criterion is not None and mask is not None
        if not is_ft and (wants_loss or torch.is_grad_enabled() and self.training):   # forward_base, train_base.py:259
            if not _train_enabled[0]:
                return orig_forward(self, img, mask, img_b, mask_b)
            return ops.forward_base_train(_features(self, img), mask, self.base_emb, _mlp_weights(self.classifier),
                                          criterion=self.criterion)
        with torch.no_grad():                                         # forward_all / forward_base, inference
            feats = _features(self, img, fused_tail=True)
            return head_for(self)(feats)
    forward._sl_patched = True
    return forward


_train_enabled = [False]
_tails_enabled = [True]


def _rebind_everywhere(name, original, replacement, done):
    """`from utils.pyt_utils import name` copies the binding into the importing module: rebind every such copy."""
    for mod_name, mod in list(sys.modules.items()):
        if mod is None or mod_name.startswith('segland_b200'):
            continue
        try:
            if mod.__dict__.get(name) is original:
                _originals.append((mod, name, original))
                setattr(mod, name, replacement)
                done.append(f'{mod_name}.{name}')
        except Exception:                                              # noqa: BLE001  (lazy / odd modules)
            continue


def patch(verbose=False, train=False, tails=True):
    """Install the B200 path into every importable reference module.  Returns the list of patched names.
    train=True also routes training-mode forwards (forward_novel, forward_base with a criterion) through the head
    kernels + sl_pop_head_bwd (bf16 features: opt-in, see the module docstring).
    tails=False keeps the decoders entirely stock (fp32 features, cast by the head)."""
    ops.check_device()
    _train_enabled[0] = bool(train)
    _tails_enabled[0] = bool(tails)
    done = []
    for name in MODEL_MODULES:
        try:
            mod = importlib.import_module('networks.' + name)
        except Exception:                                              # noqa: BLE001  (missing timm etc.)
            continue
        cls = getattr(mod, 'GFSS_Model', None)
        if cls is None or getattr(cls.forward, '_sl_patched', False):
            continue
        _originals.append((cls, 'forward', cls.forward))
        cls.forward = _make_forward(cls.forward)
        done.append(f'networks.{name}.GFSS_Model.forward')
    try:
        pu = importlib.import_module('utils.pyt_utils')
        for fn in ('get_confusion_matrix', 'intersectionAndUnionGPU'):
            if getattr(pu, fn, None) is not getattr(ops, fn):
                _rebind_everywhere(fn, getattr(pu, fn), getattr(ops, fn), done)
    except Exception:                                                  # noqa: BLE001
        pass
    try:
        crit = importlib.import_module('loss.criterion')
        if not getattr(crit.OrthLoss.forward, '_sl_patched', False):
            def orth_forward(self, preds, target, is_ft=False, proto_sim=None, aux_preds=None):
                if not preds.is_cuda:
                    return _orig_orth(self, preds, target, is_ft, proto_sim, aux_preds)
                return ops.orth_loss_forward(preds, target, is_ft, proto_sim, aux_preds, self.ignore_index, self.w)
            _orig_orth = crit.OrthLoss.forward
            orth_forward._sl_patched = True
            _originals.append((crit.OrthLoss, 'forward', _orig_orth))
            crit.OrthLoss.forward = orth_forward
            done.append('loss.criterion.OrthLoss.forward')

            def get_orth_loss(self, proto_sim, is_ft=False):
                if not proto_sim.is_cuda:
                    return _orig_get(self, proto_sim, is_ft)
                return ops.get_orth_loss(proto_sim, is_ft)
            _orig_get = crit.OrthLoss.get_orth_loss
            _originals.append((crit.OrthLoss, 'get_orth_loss', _orig_get))
            crit.OrthLoss.get_orth_loss = get_orth_loss
            done.append('loss.criterion.OrthLoss.get_orth_loss')
    except Exception:                                                  # noqa: BLE001
        pass
    if verbose:
        print('segland_b200.patch:', ', '.join(done) or 'nothing to patch')
    return done


def unpatch():
    while _originals:
        obj, name, val = _originals.pop()
        setattr(obj, name, val)
