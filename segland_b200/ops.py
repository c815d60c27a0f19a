"""Host-side mirror of the reference's operator surface for the POP head + post-processing path.

Every function keeps the name, argument meaning and side effects of the SegLand function it
replaces (cited per function; paths are relative to the SegLand tree) and does its arithmetic
in libsegland_b200.so through the C ABI.  PyTorch is used for device memory and streams only.
There is no CPU path: CPU tensors are moved to the current CUDA device (counted as H2D by the
callers that time it) and a missing library / non-B200 device raises.
"""
from __future__ import annotations

from dataclasses import dataclass

import numpy as np
import torch

from . import _cabi
from ._cabi import call, int_array, ptr, ptr_array

IGNORE_LABEL = 255           # dataset/oem.py:15
MAX_CLASSES = 32
MAX_CHANNELS = 512           # sl_pop_fg_lowres / sl_pop_head_bwd: C % 8 == 0, C <= 512


def _stream():
    return torch.cuda.current_stream().cuda_stream


def _cuda(t, dtype=None):
    if isinstance(t, np.ndarray):
        t = torch.from_numpy(t)
    if not t.is_cuda:
        t = t.cuda(non_blocking=True)
    if dtype is not None and t.dtype != dtype:
        t = t.to(dtype)
    return t.contiguous()


def _first_cuda_device(args, kwargs):
    for a in list(args) + list(kwargs.values()):
        if torch.is_tensor(a):
            if a.is_cuda:
                return a.device
        elif isinstance(a, (list, tuple)):
            for b in a:
                if torch.is_tensor(b) and b.is_cuda:
                    return b.device
    return None


def _on_input_device(fn):
    """Run an operator with the CUDA device of its first CUDA tensor argument current: the library launches on the
    caller's current device and stream, so a tensor on another GPU must not be launched from the wrong context.
    CPU-only arguments keep the caller's current device (they are uploaded there)."""
    import functools

    @functools.wraps(fn)
    def wrapper(*args, **kwargs):
        dev = _first_cuda_device(args, kwargs)
        if dev is None or dev.index == torch.cuda.current_device():
            return fn(*args, **kwargs)
        with torch.cuda.device(dev):
            return fn(*args, **kwargs)
    return wrapper


def check_device():
    """Raise unless the current CUDA device is a B200-class (sm_100) part."""
    if not torch.cuda.is_available():
        raise RuntimeError('segland_b200 needs a CUDA device (B200, sm_100a); there is no CPU fallback')
    call('sl_check_device')


# =============================================================================== POP head
@dataclass
class _Plan:
    s_hat: torch.Tensor
    alpha: torch.Tensor
    beta: torch.Tensor
    W1p_t: torch.Tensor
    W2_t: torch.Tensor
    w3_bg: torch.Tensor
    split: tuple | None      # (W1p_hi, W1p_lo, W2_hi, W2_lo) bf16 bit patterns
    f16: tuple | None        # (W1p_f16, W2_f16) fp16 bit patterns


class PopHead:
    """The POP head of GFSS_Model after the decoder (networks/*_pop.py; pspnet_pop.py:136-182).

    Holds the state-dict entries the head reads -- base_emb [Kb,C], novel_emb [Kn,C] (ft mode),
    classifier.{0,2,4}.weight, classifier_n.{0,2,4}.weight -- and maps features [B,C,h,w]
    (bf16) to logits [B,1+Kb(+Kn),h,w] fp32 in the reference's channel order
    [bg, base_1..Kb, novel_1..Kn] (pspnet_pop.py:159).  In ft mode the background and novel
    channels use classifier_n, base channels use classifier (pspnet_pop.py:150-157).

    Shapes: C % 8 == 0 and 8 <= C <= 512 (the foreground kernel's limit; `PopHead.supports(C)`), any h*w.
    bg_mode: 'tc'   tcgen05 tensor-core MLP (C % 32 == 0, 32 <= C <= 512; any N % 8 == 0)
             'simt' exact fp32 CUDA-core MLP (any supported C)
             'auto' tc when the shape allows, else simt
    tc_precision: 'precise' (split-bf16, 5 MMA passes, ~5e-6 of fp32; the default and the mode the
             parity claims are made for) or 'balanced' (layer 2 in single-pass fp16, 3 passes, ~4e-4
             relative to the tensor maximum: inside the 1e-3 logit bound, but near-tie argmax flips
             become ~100x more frequent) or 'mid' (hidden layer as fp16 hi + lo against fp16 W2, 4 passes);
             see include/segland_b200.h and profiles/r2_pass_probe.txt.
    """

    TC_PRECISIONS = {'precise': 0, 'balanced': 1, 'mid': 2}

    def __init__(self, base_emb, classifier, novel_emb=None, classifier_n=None, device=None, bg_mode='auto',
                 tc_precision='precise', fuse=False):
        check_device()
        dev = torch.device('cuda', torch.cuda.current_device()) if device is None else torch.device(device)
        f32 = lambda t: t.detach().to(dev, torch.float32).contiguous()
        self.device = dev
        self.base_emb = f32(base_emb)
        self.novel_emb = None if novel_emb is None or novel_emb.shape[0] == 0 else f32(novel_emb)
        self.Kb, self.C = self.base_emb.shape
        self.Kn = 0 if self.novel_emb is None else self.novel_emb.shape[0]
        self.K = self.Kb + self.Kn
        if not (1 <= self.K < MAX_CLASSES):
            raise ValueError(f'1 + Kb + Kn must be <= {MAX_CLASSES}')
        if not self.supports(self.C):
            raise ValueError(f'PopHead needs C % 8 == 0 and 8 <= C <= {MAX_CHANNELS} (got C={self.C}): the foreground '
                             'and backward kernels stage 8-channel TMA boxes and keep <= 512 prototypes columns on chip')
        self.cls = self._mlp(classifier, f32)
        if self.Kn:
            if classifier_n is None:
                raise ValueError('ft mode (novel_emb given) needs classifier_n')
            self.cls_n = self._mlp(classifier_n, f32)
        else:
            self.cls_n = None
        if bg_mode not in ('auto', 'tc', 'simt'):
            raise ValueError(bg_mode)
        self.bg_mode = bg_mode
        if tc_precision not in self.TC_PRECISIONS:
            raise ValueError(tc_precision)
        self.tc_precision = tc_precision
        # fuse=True: one-launch head (sl_pop_head_tc: fg logits computed inside the tensor-core kernel, K <= 12).
        # Measured on B200 at configs[1]: 1.22 ms vs 0.10 + 1.03 ms for the two-launch path -- the projection
        # warps' shared-memory reads compete with the MMA operand reads -- so it is off by default.
        self.fuse = bool(fuse)
        self._plan = None
        self._h1_ws = {}                 # (device index, stream handle) -> hidden-layer scratch of the tensor-core kernel
        self.refresh()

    @staticmethod
    def supports(C):
        """Channel counts the head kernels take (sl_pop_fg_lowres / sl_pop_head_bwd: C % 8 == 0, C <= 512).  Of the
        reference's backbones only seghr_pop with hr-w18 (d_model 270) and hr-w48 (720) fall outside; patch() leaves
        those models on the reference's own forward."""
        return C % 8 == 0 and 8 <= C <= MAX_CHANNELS

    def _scratch(self, feats, need_bytes):
        """Per-(device, stream) scratch: two streams driving one PopHead must not share hidden-layer tiles."""
        key = (feats.device.index, _stream())
        ws = self._h1_ws.get(key)
        if ws is None or ws.numel() * 2 < need_bytes:
            ws = self._h1_ws[key] = torch.empty(need_bytes // 2, dtype=torch.int16, device=feats.device)
        return ws

    # -- construction helpers ------------------------------------------------------------
    def _mlp(self, ws, f32):
        W1, W2, w3 = ws
        C = self.C
        W1, W2, w3 = f32(W1).reshape(C, C), f32(W2).reshape(C, C), f32(w3).reshape(C)
        return (W1, W2, w3)

    @classmethod
    def from_state_dict(cls, sd, **kw):
        """Build from a GFSS_Model state dict; checkpoints saved from DDP/DataParallel carry a
        'module.' prefix (utils/pyt_utils.py:101-106) which is stripped."""
        sd = {(k[7:] if k.startswith('module.') else k): v for k, v in sd.items()}
        get = lambda p: (sd[p + '.0.weight'], sd[p + '.2.weight'], sd[p + '.4.weight'])
        novel = sd.get('novel_emb')
        return cls(sd['base_emb'], get('classifier'), novel,
                   get('classifier_n') if 'classifier_n.0.weight' in sd else None, **kw)

    @classmethod
    def from_model(cls, model, **kw):
        """Build from a (possibly DDP-wrapped) reference GFSS_Model instance."""
        m = model.module if hasattr(model, 'module') else model
        get = lambda seq: (seq[0].weight, seq[2].weight, seq[4].weight)
        is_ft = getattr(m, 'is_ft', False) and getattr(m, 'novel_emb', None) is not None
        return cls(m.base_emb, get(m.classifier), m.novel_emb if is_ft else None,
                   get(m.classifier_n) if is_ft else None, **kw)

    @property
    def n_classes(self):
        return 1 + self.K

    # -- weight-dependent precompute (sl_pop_prepare) -------------------------------------
    def refresh(self):
        """Recompute s_hat / alpha / beta / folded weights; call after any weight update."""
        dev, C, K = self.device, self.C, self.K
        protos = (self.base_emb if self.novel_emb is None else torch.cat([self.base_emb, self.novel_emb], 0)).contiguous()
        fg = self.cls
        bg = self.cls_n if self.cls_n is not None else self.cls
        new = lambda *s, dtype=torch.float32: torch.empty(*s, dtype=dtype, device=dev)
        want_split = self.bg_mode != 'simt' and C % 32 == 0 and 32 <= C <= 512
        split = tuple(new(C, C, dtype=torch.int16) for _ in range(4)) if want_split else None
        f16 = tuple(new(C, C, dtype=torch.int16) for _ in range(2)) if want_split else None
        plan = _Plan(new(K, C), new(K), new(K), new(C, C), new(C, C), bg[2], split, f16)
        sp = split if split else (None,) * 4
        hf = f16 if f16 else (None,) * 2
        ws = new(_cabi.lib().sl_pop_prepare_ws_bytes(K, C) // 4)
        with torch.cuda.device(dev):
            call('sl_pop_prepare', ptr(protos), K, self.Kb, C, ptr(fg[0]), ptr(fg[1]), ptr(fg[2]),
                 ptr(bg[0]), ptr(bg[1]), ptr(bg[2]), ptr(plan.s_hat), ptr(plan.alpha), ptr(plan.beta),
                 ptr(plan.W1p_t), ptr(plan.W2_t), ptr(sp[0]), ptr(sp[1]), ptr(sp[2]), ptr(sp[3]), ptr(hf[0]), ptr(hf[1]),
                 ptr(ws), _stream())
        self._plan = plan
        self._ch_map = int_array([1 + k for k in range(K)])
        return self

    def _use_tc(self, N):
        if self.bg_mode == 'simt':
            return False
        ok = self._plan.split is not None and N % 8 == 0
        if self.bg_mode == 'tc' and not ok:
            raise ValueError(f'bg_mode="tc" needs C % 32 == 0, 32 <= C <= 512, N % 8 == 0 (C={self.C}, N={N})')
        return ok

    def bg_tc(self, feats, out):
        """Launch the tensor-core background MLP on bf16 features [B,C,h,w] into out[:,0]."""
        B, C, h, w = feats.shape
        N = h * w
        p = self._plan
        with torch.cuda.device(feats.device):
            ws = self._scratch(feats, _cabi.lib().sl_pop_bg_tc_ws_bytes(B, C, N))
            call('sl_pop_bg_tc', ptr(feats), B, C, N, ptr(p.split[0]), ptr(p.split[1]), ptr(p.split[2]),
                 ptr(p.split[3]), ptr(p.f16[1]), ptr(p.w3_bg), self.TC_PRECISIONS[self.tc_precision],
                 ptr(ws), ptr(out), out.shape[1], 0, _stream())

    def head_tc(self, feats, out):
        """One launch for the whole head (K <= 12): background MLP on tcgen05 + foreground logits from the
        same shared-memory feature tiles."""
        B, C, h, w = feats.shape
        N = h * w
        p = self._plan
        with torch.cuda.device(feats.device):
            ws = self._scratch(feats, _cabi.lib().sl_pop_bg_tc_ws_bytes(B, C, N))
            call('sl_pop_head_tc', ptr(feats), B, C, N, ptr(p.split[0]), ptr(p.split[1]), ptr(p.split[2]),
                 ptr(p.split[3]), ptr(p.f16[1]), ptr(p.w3_bg), self.TC_PRECISIONS[self.tc_precision], ptr(p.s_hat),
                 ptr(p.alpha), ptr(p.beta), self.K, self._ch_map, ptr(ws), ptr(out), out.shape[1], 0, _stream())

    def fused_ok(self, N):
        return self.K <= 12 and self._use_tc(N)

    def bg(self, feats, out):
        """Background logit into out[:,0] by whichever kernel the head's mode and the shape select."""
        N = feats.shape[2] * feats.shape[3]
        if self._use_tc(N):
            self.bg_tc(feats, out)
        else:
            self.bg_simt(feats, out)

    def bg_simt(self, feats, out):
        B, C, h, w = feats.shape
        p = self._plan
        with torch.cuda.device(feats.device):
            call('sl_pop_bg_simt', ptr(feats), B, C, h * w, ptr(p.W1p_t), ptr(p.W2_t), ptr(p.w3_bg),
                 ptr(out), out.shape[1], 0, _stream())

    # -- forward ---------------------------------------------------------------------------
    def __call__(self, features, out=None, fg_only=False):
        """features [B,C,h,w] bf16 CUDA (other float dtypes are cast to bf16, as north_star
        specifies bf16 features) -> logits [B,1+K,h,w] fp32.
        fg_only skips the background MLP and leaves channel 0 untouched (stage-S timing)."""
        if features.dim() != 4 or features.shape[1] != self.C:
            raise ValueError(f'features must be [B,{self.C},h,w], got {tuple(features.shape)}')
        if not features.is_cuda:
            features = features.to(self.device, non_blocking=True)
        feats = _cuda(features, torch.bfloat16)
        B, C, h, w = feats.shape
        N = h * w
        Ktot = 1 + self.K
        if N % 8:
            # TMA wants 16-byte row strides: pad the pixel axis to a multiple of 8 (one extra copy of the features;
            # feature maps of the reference's models at its tile sizes never take this path) and drop the padding
            pad = -N % 8
            padded = torch.nn.functional.pad(feats.reshape(B, C, N), (0, pad)).view(B, C, 1, N + pad)
            res = self.__call__(padded, fg_only=fg_only)[..., :N].reshape(B, Ktot, h, w)
            if out is None:
                return res
            out.copy_(res)
            return out
        if out is None:
            out = torch.empty(B, Ktot, h, w, dtype=torch.float32, device=feats.device)
        p = self._plan
        if feats.device != self.device:
            raise ValueError(f'features live on {feats.device}, this head on {self.device}')
        if not fg_only and self.fuse and self.fused_ok(N):
            self.head_tc(feats, out)
            return out
        with torch.cuda.device(feats.device):
            call('sl_pop_fg_lowres', ptr(feats), B, C, N, ptr(p.s_hat), ptr(p.alpha), ptr(p.beta), self.K,
                 ptr(out), Ktot, self._ch_map, _stream())
        if not fg_only:
            if self._use_tc(N):
                self.bg_tc(feats, out)
            else:
                self.bg_simt(feats, out)
        return out

    forward = __call__


# ==================================================================== POP head, training mode
class _PopHeadTrainFn(torch.autograd.Function):
    """preds = head(features; prototypes, classifier, classifier_n) with every step in libsegland_b200.so:
    forward = sl_pop_prepare + sl_pop_fg_lowres + sl_pop_bg_tc / sl_pop_bg_simt (the eval kernels),
    backward = sl_pop_head_bwd (per-pixel part) + sl_pop_prepare_bwd (parameter-side chain)."""

    @staticmethod
    def forward(ctx, features, protos, Kb, W1, W2, w3, W1n, W2n, w3n, bg_mode):
        feats = features.detach().to(torch.bfloat16).contiguous()
        B, C, h, w = feats.shape
        N, K = h * w, protos.shape[0]
        Ktot = 1 + K
        dev = feats.device
        f32 = lambda t: t.detach().to(torch.float32).contiguous()
        protos = f32(protos)
        fg = (f32(W1).view(C, C), f32(W2).view(C, C), f32(w3).view(C))
        bg = fg if W1n is None else (f32(W1n).view(C, C), f32(W2n).view(C, C), f32(w3n).view(C))
        new = lambda *s, dtype=torch.float32: torch.empty(*s, dtype=dtype, device=dev)
        use_tc = bg_mode != 'simt' and C % 32 == 0 and 32 <= C <= 512 and N % 8 == 0
        if bg_mode == 'tc' and not use_tc:
            raise ValueError(f'bg_mode="tc" needs C % 32 == 0, 32 <= C <= 512, N % 8 == 0 (C={C}, N={N})')
        s_hat, alpha, beta, W1p_t, W2_t = new(K, C), new(K), new(K), new(C, C), new(C, C)
        split = tuple(new(C, C, dtype=torch.int16) for _ in range(4)) if use_tc else (None,) * 4
        ws = new(_cabi.lib().sl_pop_prepare_ws_bytes(K, C) // 4)
        st = _stream()
        call('sl_pop_prepare', ptr(protos), K, Kb, C, ptr(fg[0]), ptr(fg[1]), ptr(fg[2]), ptr(bg[0]), ptr(bg[1]),
             ptr(bg[2]), ptr(s_hat), ptr(alpha), ptr(beta), ptr(W1p_t), ptr(W2_t), ptr(split[0]), ptr(split[1]),
             ptr(split[2]), ptr(split[3]), None, None, ptr(ws), st)
        out = new(B, Ktot, h, w)
        call('sl_pop_fg_lowres', ptr(feats), B, C, N, ptr(s_hat), ptr(alpha), ptr(beta), K, ptr(out), Ktot,
             int_array([1 + k for k in range(K)]), st)
        if use_tc:
            h1_ws = torch.empty(_cabi.lib().sl_pop_bg_tc_ws_bytes(B, C, N) // 2, dtype=torch.int16, device=dev)
            call('sl_pop_bg_tc', ptr(feats), B, C, N, ptr(split[0]), ptr(split[1]), ptr(split[2]), ptr(split[3]),
                 ptr(split[2]), ptr(bg[2]), 0, ptr(h1_ws), ptr(out), Ktot, 0, st)
        else:
            call('sl_pop_bg_simt', ptr(feats), B, C, N, ptr(W1p_t), ptr(W2_t), ptr(bg[2]), ptr(out), Ktot, 0, st)
        ctx.save_for_backward(feats, protos, s_hat, alpha, beta, W1p_t, *fg, *(() if W1n is None else bg))
        ctx.meta = (Kb, W1n is None, bg_mode, features.dtype, tuple(W1.shape), tuple(W2.shape), tuple(w3.shape))
        return out

    @staticmethod
    def backward(ctx, g):
        feats, protos, s_hat, alpha, beta, W1p_t, *ws_ = ctx.saved_tensors
        Kb, shared, bg_mode, feat_dtype, s1, s2, s3 = ctx.meta
        fg = tuple(ws_[:3])
        bg = fg if shared else tuple(ws_[3:])
        B, C, h, w = feats.shape
        N, K = h * w, s_hat.shape[0]
        dev = feats.device
        g = g.to(torch.float32).contiguous()
        new = lambda *s: torch.empty(*s, dtype=torch.float32, device=dev)
        need = ctx.needs_input_grad            # features, protos, Kb, W1, W2, w3, W1n, W2n, w3n, bg_mode
        d_s, d_a, d_b, dW1p, dW2d, dw3d = new(K, C), new(K), new(K), new(C, C), new(C, C), new(C)
        d_feat = new(B, C, h, w) if need[0] else None
        W1p = W1p_t.t().contiguous()
        ws = torch.empty(_cabi.lib().sl_pop_head_bwd_ws_bytes(B, C, N, K) // 4, dtype=torch.float32, device=dev)
        st = _stream()
        call('sl_pop_head_bwd', ptr(feats), B, C, N, ptr(s_hat), ptr(alpha), ptr(beta), K,
             int_array([1 + k for k in range(K)]), ptr(W1p), ptr(bg[1]), ptr(bg[2]), ptr(g), 1 + K, 0,
             ptr(d_s), ptr(d_a), ptr(d_b), ptr(dW1p), ptr(dW2d), ptr(dw3d), ptr(d_feat),
             1 if bg_mode == 'simt' else 0, ptr(ws), st)
        want_fg = (not shared) and (need[3] or need[4] or need[5])
        d_protos = new(K, C)
        gfg = (new(C, C), new(C, C), new(C)) if want_fg else (None, None, None)
        gbg = (new(C, C), new(C, C), new(C))
        ws2 = new(_cabi.lib().sl_pop_prepare_bwd_ws_bytes(K, C) // 4)
        call('sl_pop_prepare_bwd', ptr(protos), K, Kb, C, ptr(fg[0]), ptr(fg[1]), ptr(fg[2]), ptr(bg[0]), ptr(bg[1]),
             ptr(bg[2]), ptr(d_s), ptr(d_a), ptr(d_b), ptr(dW1p), ptr(dW2d), ptr(dw3d), ptr(d_protos),
             ptr(gfg[0]), ptr(gfg[1]), ptr(gfg[2]), ptr(gbg[0]), ptr(gbg[1]), ptr(gbg[2]), ptr(ws2), st)
        if d_feat is not None and feat_dtype != torch.float32:
            d_feat = d_feat.to(feat_dtype)
        if shared:
            gfg, gbg = gbg, (None, None, None)
        shp = lambda t, s: None if t is None else t.view(s)
        return (d_feat, d_protos if need[1] else None, None, shp(gfg[0], s1), shp(gfg[1], s2), shp(gfg[2], s3),
                shp(gbg[0], s1), shp(gbg[1], s2), shp(gbg[2], s3), None)


@_on_input_device
def pop_head_train(features, base_emb, classifier, novel_emb=None, classifier_n=None, bg_mode='auto'):
    """The head of forward_novel / forward_base with autograd (networks/pspnet_pop.py:199-219, :169-182):
    features [B,C,h,w] (any float dtype; cast to bf16 for the kernels, gradients flow back if it requires
    grad), base_emb [Kb,C], classifier = (W1, W2, w3) conv weights, and for ft mode novel_emb [Kn,C] +
    classifier_n.  Returns preds [B,1+Kb(+Kn),h,w] fp32 in the reference's channel order.  Forward and
    backward run in libsegland_b200.so (see _PopHeadTrainFn); PyTorch only concatenates the prototypes and
    routes the gradients.  bg_mode='simt' selects the exact-fp32 CUDA-core kernels for both directions."""
    check_device()
    ft = novel_emb is not None and novel_emb.shape[0] > 0
    Kb = base_emb.shape[0]
    protos = torch.cat([base_emb, novel_emb], 0) if ft else base_emb
    cn = tuple(classifier_n) if ft else (None, None, None)
    with torch.autocast('cuda', enabled=False):
        return _PopHeadTrainFn.apply(features, protos, Kb, classifier[0], classifier[1], classifier[2],
                                     cn[0], cn[1], cn[2], bg_mode)


def forward_novel_train(features_full, mask, mask_b, base_emb, novel_emb, classifier, classifier_n, criterion=None,
                        bg_mode='auto'):
    """forward_novel after the decoder (networks/pspnet_pop.py:199-245): head on [novel; base] images,
    pseudo-labelling of the base images' background with classifier_n's argmax (IN PLACE on mask_b, :221-231),
    then criterion(preds, cat(mask, mask_b), is_ft=True, proto_sim) or preds when there is no criterion/mask."""
    Kb = base_emb.shape[0]
    preds = pop_head_train(features_full, base_emb, classifier, novel_emb, classifier_n, bg_mode)
    B = preds.shape[0]
    with torch.no_grad():
        preds2_base = torch.cat([preds[B // 2:, :1], preds[B // 2:, 1 + Kb:]], dim=1).contiguous()
        if mask_b.shape[0] != B // 2:
            raise ValueError('mask_b must hold one mask per base image (the second half of the batch)')
        mask_new = pseudo_label(preds2_base, mask_b, Kb)
    if criterion is None or mask is None:
        return preds
    with torch.autocast('cuda', enabled=False):
        C = base_emb.shape[1]
        mask_all = torch.cat([mask, mask_new], dim=0)
        n_hat = torch.nn.functional.normalize(novel_emb.to(torch.float32), p=2, dim=-1).reshape(-1, C)
        all_emb = torch.cat([n_hat, torch.nn.functional.normalize(base_emb.to(torch.float32), p=2, dim=-1)], dim=0)
        proto_sim = torch.matmul(n_hat, all_emb.t())
        return criterion(preds.to(torch.float32), mask_all, is_ft=True, proto_sim=proto_sim)


def forward_base_train(features, mask, base_emb, classifier, criterion=None, bg_mode='auto'):
    """forward_base after the decoder with autograd (networks/pspnet_pop.py:169-189)."""
    preds = pop_head_train(features, base_emb, classifier, None, None, bg_mode)
    if criterion is None or mask is None:
        return preds
    cls_emb = torch.nn.functional.normalize(base_emb, p=2, dim=-1)
    return criterion(preds, mask, proto_sim=torch.matmul(cls_emb, cls_emb.t()))


@_on_input_device
def aggregate_views(views, flips, scale=None):
    """Test-time view aggregation (spec: this repo; the reference has none).
    views [V,B,K,h,w] fp32 logits of V views of the same tiles; flips[v] in {0,1,2,3}
    (bit0 = horizontally flipped input, bit1 = vertically flipped).  Returns the un-flipped
    mean [B,K,h,w] (scale defaults to 1/V)."""
    views = _cuda(views, torch.float32)
    V, B, K, h, w = views.shape
    out = torch.empty(B, K, h, w, dtype=torch.float32, device=views.device)
    call('sl_views_reduce', ptr(views), V, B, K, h, w, int_array(flips), float(1.0 / V if scale is None else scale),
         ptr(out), _stream())
    return out


class WindowPlan:
    """Sliding-window crop grid over one tile (spec: this repo; north_star names "engine.py's sliding-window/flip
    aggregation" but the reference evaluates whole tiles, engine.py:23-143, eval_ft.py:162-172).

    tile_hw, crop_hw and stride_hw are in IMAGE pixels, model_stride is the network's output stride (8 for the
    ResNet models, 4 for Swin/ConvNeXt/LSK/HRNet).  Window origins follow the usual rule: a regular
    grid of step `stride_hw`, the last window pulled back so that it ends at the tile border (so the last step can
    be ragged).  Everything must be a multiple of model_stride so that crops stitch on the feature grid."""

    def __init__(self, tile_hw, crop_hw, stride_hw=None, model_stride=8):
        th, tw = int(tile_hw[0]), int(tile_hw[1])
        ch, cw = int(crop_hw[0]), int(crop_hw[1])
        sh, sw = (ch, cw) if stride_hw is None else (int(stride_hw[0]), int(stride_hw[1]))
        s = int(model_stride)
        if min(th, tw, ch, cw, sh, sw, s) < 1 or ch > th or cw > tw:
            raise ValueError('WindowPlan: sizes must be positive and the crop no larger than the tile')
        if sh > ch or sw > cw:
            raise ValueError('WindowPlan: a stride larger than the crop leaves pixels uncovered')
        for v in (th, tw, ch, cw, sh, sw):
            if v % s:
                raise ValueError(f'WindowPlan: tile, crop and stride must be multiples of the model stride {s}')
        self.tile_hw, self.crop_hw, self.stride_hw, self.model_stride = (th, tw), (ch, cw), (sh, sw), s
        self.origins_y = self._origins(th, ch, sh)
        self.origins_x = self._origins(tw, cw, sw)
        if max(len(self.origins_y), len(self.origins_x)) > 32:
            raise ValueError('WindowPlan: more than 32 windows per axis')

    @staticmethod
    def _origins(size, crop, stride):
        n = max(size - crop + stride - 1, 0) // stride + 1
        return [min(i * stride, size - crop) for i in range(n)]

    @property
    def n_windows(self):
        return len(self.origins_y) * len(self.origins_x)

    def windows(self):
        """(y, x) image-pixel origin of every window, row-major: the entry order window_accumulate expects."""
        return [(y, x) for y in self.origins_y for x in self.origins_x]

    @property
    def canvas_hw(self):
        return (self.tile_hw[0] // self.model_stride, self.tile_hw[1] // self.model_stride)

    @property
    def crop_lr_hw(self):
        return (self.crop_hw[0] // self.model_stride, self.crop_hw[1] // self.model_stride)

    def crop(self, images, flips=(0,)):
        """images [B,...,H,W] -> [B, n_windows*len(flips), ..., ch, cw]: the crops (and their flipped views) in entry
        order, i.e. what a backbone is fed (plain slicing/flip; host-side plumbing, any device)."""
        out = []
        for (y, x) in self.windows():
            c = images[..., y:y + self.crop_hw[0], x:x + self.crop_hw[1]]
            for f in flips:
                dims = [d for d, bit in ((-1, 1), (-2, 2)) if f & bit]
                out.append(torch.flip(c, dims) if dims else c)
        return torch.stack(out, dim=1)


@_on_input_device
def window_accumulate(crop_logits, plan, flips=(0,), layout='bekhw', want_count=False, out=None):
    """Stitch per-crop low-res logits into one canvas per tile (spec: this repo, see WindowPlan): the un-flipped
    crops are summed at feature resolution in entry order and divided by the overlap count -- the result is bit-equal
    to `canvas[..., y:y+hc, x:x+wc] += unflip(crop)` per entry followed by `canvas / count`.
    crop_logits fp32 CUDA, layout 'bekhw' = [B,E,K,hc,wc] (a batch of tiles, each with its E = n_windows*len(flips)
    crops, entry e = window*len(flips) + view) or 'ebkhw' = [E,B,K,hc,wc].  Returns canvas [B,K,h,w] fp32 (and the
    count plane [h,w] when want_count); feed it to upsample_argmax."""
    crop_logits = _cuda(crop_logits, torch.float32)
    if crop_logits.dim() != 5 or layout not in ('bekhw', 'ebkhw'):
        raise ValueError("crop_logits must be 5-D with layout 'bekhw' or 'ebkhw'")
    if layout == 'bekhw':
        B, E, K, hc, wc = crop_logits.shape
        stride_b, stride_e = E * K * hc * wc, K * hc * wc
    else:
        E, B, K, hc, wc = crop_logits.shape
        stride_e, stride_b = B * K * hc * wc, K * hc * wc
    V = len(flips)
    if E != plan.n_windows * V or (hc, wc) != plan.crop_lr_hw:
        raise ValueError(f'expected {plan.n_windows * V} entries of {plan.crop_lr_hw} feature pixels, got {E} of {(hc, wc)}')
    h, w = plan.canvas_hw
    s = plan.model_stride
    dev = crop_logits.device
    if out is None:
        out = torch.empty(B, K, h, w, dtype=torch.float32, device=dev)
    elif tuple(out.shape) != (B, K, h, w) or out.dtype != torch.float32 or not out.is_contiguous():
        raise ValueError(f'out must be a contiguous fp32 [{B},{K},{h},{w}] tensor')
    count = torch.empty(h, w, dtype=torch.float32, device=dev) if want_count else None
    oy, ox = [y // s for y in plan.origins_y], [x // s for x in plan.origins_x]
    with torch.cuda.device(dev):
        call('sl_window_accumulate', ptr(crop_logits), stride_e, stride_b, B, K, hc, wc, int_array(oy), len(oy),
             int_array(ox), len(ox), int_array(flips), V, h, w, ptr(out), ptr(count), _stream())
    return (out, count) if want_count else out


# ================================================================== dense post-processing
@_on_input_device
def upsample_argmax(logits, size, label=None, cm=None, ignore_label=IGNORE_LABEL, want_pred=True,
                    want_conf=False, want_probs=False, want_logits=False, pred_out=None, out_bufs=None):
    """F.interpolate(logits, size, mode='bilinear', align_corners=True) -> argmax(dim=1) -> uint8
    (eval_base.py:168-170, eval_ft.py:168-172), optionally fused with the confusion-matrix update
    (eval_base.py:172-178).  logits [B,K,h,w] fp32 CUDA.

    cm: int64 [K,K] CUDA accumulator (row = gt, col = pred), updated in place; needs label
        [B,H,W] uint8.  Returns a dict with the requested outputs: 'pred' uint8 [B,H,W],
        'conf' fp32 [B,H,W] and 'probs' fp32 [B,K,H,W] (softmax; spec: this repo), 'logits'
        fp32 [B,K,H,W] (the up-sampled logits eval_base.py:190-191 dumps for fusemat).
    pred_out: optional caller-owned uint8 [B,H,W] buffer for 'pred'; out_bufs: optional dict of caller-owned buffers
    for 'pred' / 'conf' / 'probs' (no allocation for those outputs)."""
    logits = _cuda(logits, torch.float32)
    B, K, h, w = logits.shape
    H, W = int(size[0]), int(size[1])
    dev = logits.device
    out = {}
    def buf(name, want, shape, dtype):                       # caller-owned output buffer (no allocation) or a fresh one
        if not want:
            return None
        given = (out_bufs or {}).get(name)
        if name == 'pred' and given is None:
            given = pred_out
        if given is None:
            return torch.empty(shape, dtype=dtype, device=dev)
        if given.dtype != dtype or tuple(given.shape) != tuple(shape) or given.device != dev or not given.is_contiguous():
            raise ValueError(f'{name} buffer must be a contiguous {dtype} {tuple(shape)} tensor on {dev}')
        return given
    pred = buf('pred', want_pred, (B, H, W), torch.uint8)
    conf = buf('conf', want_conf, (B, H, W), torch.float32)
    probs = buf('probs', want_probs, (B, K, H, W), torch.float32)
    hr = torch.empty(B, K, H, W, dtype=torch.float32, device=dev) if want_logits else None
    if cm is not None:
        if label is None:
            raise ValueError('cm needs label')
        if cm.dtype != torch.int64 or tuple(cm.shape) != (K, K) or not cm.is_cuda or not cm.is_contiguous():
            raise ValueError('cm must be a contiguous CUDA int64 [K,K] tensor')
        label = _cuda(label, torch.uint8)
        if tuple(label.shape) != (B, H, W):
            raise ValueError(f'label must be [B,H,W]={B, H, W}, got {tuple(label.shape)}')
    call('sl_upsample_argmax', ptr(logits), B, K, h, w, H, W, ptr(label) if cm is not None else None,
         int(ignore_label), ptr(pred), ptr(conf), ptr(probs), ptr(hr), ptr(cm), _stream())
    for k, v in (('pred', pred), ('conf', conf), ('probs', probs), ('logits', hr)):
        if v is not None:
            out[k] = v
    return out


@_on_input_device
def confusion_update(cm, gt, pred, ignore_label=IGNORE_LABEL):
    """cm[gt, pred] += 1 over all pixels with gt != ignore_label (device-resident accumulation).
    gt/pred: uint8 tensors of equal shape.  Returns the number of out-of-range labels skipped
    as a 1-element CUDA int64 tensor (no host sync)."""
    gt, pred = _cuda(gt, torch.uint8), _cuda(pred, torch.uint8)
    if gt.shape != pred.shape:
        raise ValueError('gt and pred shapes differ')
    K = cm.shape[0]
    n_bad = torch.zeros(1, dtype=torch.int64, device=cm.device)
    call('sl_confusion', ptr(gt), ptr(pred), gt.numel(), K, int(ignore_label), ptr(cm), ptr(n_bad), _stream())
    return n_bad


def get_confusion_matrix(gt_label, pred_label, class_num):
    """utils/pyt_utils.py:182-200 with the reference's signature and return type: 1-D (or any
    shape) label arrays already filtered by gt != ignore (eval_base.py:175-178) -> float64
    numpy [class_num, class_num], row = gt, col = pred.  Labels equal to 255 are still
    skipped, so unfiltered maps give the same matrix as the reference's filtered call."""
    dev = torch.device('cuda', torch.cuda.current_device())
    cm = torch.zeros(class_num, class_num, dtype=torch.int64, device=dev)
    to_u8 = lambda a: _cuda(torch.as_tensor(np.asarray(a).astype(np.uint8, copy=False)) if not torch.is_tensor(a) else a,
                            torch.uint8)
    confusion_update(cm, to_u8(gt_label), to_u8(pred_label))
    return cm.cpu().numpy().astype(np.float64)


def miou_from_confusion(confusion_matrix, base_classes):
    """eval_base.py:193-199 / eval_ft.py:196-202 on a host copy of the (tiny) matrix:
    returns (base_miou, novel_miou, total_miou, miou_array) in float64."""
    cm = confusion_matrix.detach().cpu().numpy() if torch.is_tensor(confusion_matrix) else np.asarray(confusion_matrix)
    cm = cm.astype(np.float64)
    pos, res, tp = cm.sum(1), cm.sum(0), np.diag(cm)
    with np.errstate(divide='ignore', invalid='ignore'):
        miou_array = tp / (pos + res - tp)
        base = np.nanmean(miou_array[:base_classes + 1])
        novel = np.nanmean(miou_array[base_classes + 1:]) if base_classes + 1 < len(miou_array) else float('nan')
        total = np.nanmean(miou_array)
    return base, novel, total, miou_array


@_on_input_device
def intersectionAndUnionGPU(output, target, K, ignore_index=IGNORE_LABEL):
    """utils/pyt_utils.py:293-305, same signature, return values and side effect:
    output/target int64 CUDA tensors of equal shape; output[target == ignore] = ignore IN PLACE;
    returns (area_intersection, area_union, area_target) as fp32 [K] CUDA tensors."""
    assert output.dim() in [1, 2, 3]
    assert output.shape == target.shape
    if not (output.is_cuda and target.is_cuda and output.dtype == torch.int64 and target.dtype == torch.int64):
        raise ValueError('intersectionAndUnionGPU takes int64 CUDA tensors (as .max(1)[1] and the label loader give)')
    if not output.is_contiguous():
        raise ValueError('output must be contiguous (it is modified in place)')
    target = target.contiguous()
    dev = output.device
    inter, uni, tgt = (torch.empty(K, dtype=torch.float32, device=dev) for _ in range(3))
    ws = torch.empty(3 * K, dtype=torch.int64, device=dev)
    call('sl_inter_union', ptr(output), ptr(target), output.numel(), K, int(ignore_index), ptr(inter), ptr(uni),
         ptr(tgt), ptr(ws), _stream())
    return inter, uni, tgt


@_on_input_device
def pseudo_label(preds2_base, mask_b, n_base):
    """pspnet_pop.py:221-231: label the background (== 0) pixels of the base images' masks with
    argmax(upsample(classifier_n outputs)), shifting novel indices by n_base.  preds2_base
    [Bb,1+Kn,h,w] fp32; mask_b [Bb,H,W] int64 CUDA, modified IN PLACE and returned."""
    preds2_base = _cuda(preds2_base, torch.float32)
    if not (mask_b.is_cuda and mask_b.dtype == torch.int64 and mask_b.is_contiguous()):
        raise ValueError('mask_b must be a contiguous int64 CUDA tensor')
    B, K2, h, w = preds2_base.shape
    H, W = mask_b.shape[-2:]
    call('sl_pseudo_label', ptr(preds2_base), B, K2, h, w, H, W, int(n_base), ptr(mask_b), _stream())
    return mask_b


# ======================================================================== prototypes / loss
@_on_input_device
def masked_average_pooling(feature, mask, return_per_image=False):
    """networks/pspnet.py:7-15: feature [B,C,h,w] (cast to bf16), mask [B,1,H,W] float ->
    [1,1,C] fp32: mean over images of sum(f * m_lr) / (sum(m_lr) + 1e-5), m_lr = bilinear
    align_corners=True resample of mask to [h,w]."""
    feature = _cuda(feature, torch.bfloat16)
    mask = _cuda(mask, torch.float32)
    B, C, h, w = feature.shape
    H, W = mask.shape[-2:]
    if mask.shape[0] != B or mask.numel() != B * H * W:
        raise ValueError('mask must be [B,1,H,W]')
    dev = feature.device
    ws = torch.empty(B * h * w + B * ((h * w + 1023) // 1024) + 8 * B * C, dtype=torch.float32, device=dev)
    per_image = torch.empty(B, C, dtype=torch.float32, device=dev)
    proto = torch.empty(C, dtype=torch.float32, device=dev)
    call('sl_map_proto', ptr(feature), ptr(mask), B, C, h, w, H, W, ptr(ws), ptr(per_image), ptr(proto), _stream())
    out = proto.view(1, 1, C)
    return (out, per_image) if return_per_image else out


class _OrthLossFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, rows, others):
        rows_c = rows.detach().to(torch.float32).contiguous()
        Kr, C = rows_c.shape
        Ko = 0 if others is None else others.shape[0]
        oth = None if Ko == 0 else others.detach().to(torch.float32).contiguous()
        dev = rows_c.device
        sim = torch.empty(Kr, Kr + Ko, dtype=torch.float32, device=dev)
        loss = torch.empty(1, dtype=torch.float32, device=dev)
        grad = torch.empty_like(rows_c)
        call('sl_orth_loss', ptr(rows_c), Kr, ptr(oth), Ko, C, ptr(sim), ptr(loss), ptr(grad), _stream())
        ctx.save_for_backward(grad)
        ctx.mark_non_differentiable(sim)
        return loss.reshape(()), sim

    @staticmethod
    def backward(ctx, g_loss, _g_sim):
        (grad,) = ctx.saved_tensors
        return grad * g_loss, None


@_on_input_device
def orth_loss(rows, others=None):
    """proto_sim + OrthLoss.get_orth_loss (loss/criterion.py:37-43) in one kernel.
    Base training (pspnet_pop.py:185-186): orth_loss(base_emb) -> sim [Kb,Kb].
    Fine-tuning (pspnet_pop.py:234-239): orth_loss(novel_emb, base_emb) -> sim [Kn,Kn+Kb]
    (base_emb is frozen in ft mode, so only `rows` receives a gradient).
    Returns (loss scalar, proto_sim); differentiable w.r.t. rows."""
    if not rows.is_cuda:
        raise ValueError('orth_loss needs CUDA tensors')
    return _OrthLossFn.apply(rows, others)


class _OrthFromSimFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, proto_sim):
        sim = proto_sim.detach().to(torch.float32).contiguous()
        Kr, Kc = sim.shape
        loss = torch.empty(1, dtype=torch.float32, device=sim.device)
        grad = torch.empty_like(sim)
        call('sl_orth_from_sim', ptr(sim), Kr, Kc, ptr(loss), ptr(grad), _stream())
        ctx.save_for_backward(grad)
        ctx.in_dtype = proto_sim.dtype
        return loss.reshape(())

    @staticmethod
    def backward(ctx, g):
        (grad,) = ctx.saved_tensors
        return (grad * g).to(ctx.in_dtype)


@_on_input_device
def get_orth_loss(proto_sim, is_ft=False):
    """OrthLoss.get_orth_loss (loss/criterion.py:37-43), same arguments (is_ft is unused there too): the mean absolute
    value of proto_sim's entries above the diagonal, one launch, differentiable w.r.t. proto_sim."""
    if not proto_sim.is_cuda or proto_sim.dim() != 2:
        raise ValueError('get_orth_loss takes a 2-D CUDA proto_sim matrix')
    return _OrthFromSimFn.apply(proto_sim)


class _SegCEFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, preds, target, ignore_index):
        logits = preds.detach().to(torch.float32).contiguous()
        B, K, h, w = logits.shape
        tgt = target.detach().to(torch.int64).contiguous()
        H, W = tgt.shape[-2:]
        dev = logits.device
        ws = torch.empty(_cabi.lib().sl_upsample_ce_ws_bytes(B, K, w, H, W), dtype=torch.uint8, device=dev)
        loss = torch.empty(1, dtype=torch.float32, device=dev)
        n_valid = torch.empty(1, dtype=torch.int64, device=dev)
        call('sl_upsample_ce_fwd', ptr(logits), B, K, h, w, H, W, ptr(tgt), int(ignore_index), ptr(ws), ptr(loss),
             ptr(n_valid), _stream())
        ctx.save_for_backward(logits, tgt, ws, n_valid)
        ctx.ignore_index = int(ignore_index)
        ctx.in_dtype = preds.dtype
        return loss.reshape(())

    @staticmethod
    def backward(ctx, g):
        logits, tgt, ws, n_valid = ctx.saved_tensors
        B, K, h, w = logits.shape
        H, W = tgt.shape[-2:]
        grad = torch.empty_like(logits)
        g = g.detach().to(torch.float32).reshape(1).contiguous()
        call('sl_upsample_ce_bwd', ptr(logits), B, K, h, w, H, W, ptr(tgt), ctx.ignore_index, ptr(ws), ptr(n_valid),
             ptr(g), ptr(grad), _stream())
        return grad.to(ctx.in_dtype), None, None


@_on_input_device
def seg_cross_entropy(preds, target, ignore_index=IGNORE_LABEL):
    """The segmentation term of OrthLoss.forward / CELoss.forward (loss/criterion.py:17-19, :51-52):
    CrossEntropyLoss(ignore_index, 'mean') of the align-corners bilinear up-sampling of `preds`
    [B,K,h,w] to the size of `target` [B,H,W] (int64), fused: the [B,K,H,W] tensor is never written.
    Differentiable w.r.t. preds; deterministic."""
    if not preds.is_cuda or not target.is_cuda:
        raise ValueError('seg_cross_entropy needs CUDA tensors')
    if preds.dim() != 4 or target.dim() != 3 or preds.shape[0] != target.shape[0]:
        raise ValueError('preds must be [B,K,h,w] and target [B,H,W]')
    return _SegCEFn.apply(preds, target, ignore_index)


def orth_loss_forward(preds, target, is_ft=False, proto_sim=None, aux_preds=None, ignore_index=IGNORE_LABEL, w=10.0):
    """OrthLoss.forward (loss/criterion.py:45-65) with the same arguments and loss-dict keys: the seg /
    aux cross-entropy terms run fused; the orthogonality term is get_orth_loss on the (tiny) proto_sim matrix the
    model built (sl_orth_from_sim; loss/criterion.py:37-43)."""
    seg_loss = seg_cross_entropy(preds, target, ignore_index)
    orth = get_orth_loss(proto_sim, is_ft)
    if aux_preds is not None:
        aux_loss = seg_cross_entropy(aux_preds, target, ignore_index)
        total = seg_loss + orth * w + 0.4 * aux_loss
        return {'total_loss': total, 'seg_loss': seg_loss, 'aux_loss': aux_loss, 'orth_loss': orth}
    total = seg_loss + orth * w
    return {'total_loss': total, 'seg_loss': seg_loss, 'orth_loss': orth}


# =================================================================================== fusion
@_on_input_device
def fuse_logits(mats, n_lists=None, label=None, cm=None, ignore_label=IGNORE_LABEL, want_fused=False):
    """fusemat.py:42-48 for one tile (or a batch laid out as one long pixel axis): mats is the
    list of per-model logit stacks [K,H,W] (or [K,...]) fp32, summed in list order, divided by
    n_lists (= len(fusion_list); defaults to len(mats)), argmax over K -> uint8 [H,W]."""
    mats = [_cuda(m, torch.float32) for m in mats]
    K = mats[0].shape[0]
    shape = tuple(mats[0].shape[1:])
    HW = int(np.prod(shape))
    for m in mats:
        if tuple(m.shape) != (K,) + shape:
            raise ValueError('all stacks must share one shape')
    dev = mats[0].device
    pred = torch.empty(shape, dtype=torch.uint8, device=dev)
    fused = torch.empty((K,) + shape, dtype=torch.float32, device=dev) if want_fused else None
    if cm is not None:
        label = _cuda(label, torch.uint8)
    call('sl_fuse_argmax', ptr_array(mats), len(mats), K, HW, int(len(mats) if n_lists is None else n_lists),
         ptr(pred), ptr(fused), ptr(label) if cm is not None else None, int(ignore_label), ptr(cm), _stream())
    return (pred, fused) if want_fused else pred


@_on_input_device
def fuse_logits_sweep(stacks, n_lists=None, labels=None, cm=None, ignore_label=IGNORE_LABEL, want_fused=False):
    """fusemat.py:37-48 for a whole sweep: stacks is the list of per-model tensors [T,K,H,W] fp32 (what each model's
    eval sweep produced, in fusion-list order); every tile is fused exactly as `fuse_logits` fuses one -- sum in list
    order, / n_lists, first-maximum argmax -- in one C-ABI call.  labels [T,H,W] uint8 + cm int64 [K,K] accumulate the
    confusion matrix of the fused maps.  Returns pred uint8 [T,H,W] (and the fused stacks when want_fused)."""
    stacks = [_cuda(m, torch.float32) for m in stacks]
    if stacks[0].dim() != 4:
        raise ValueError('stacks must be [T,K,H,W] tensors')
    T, K, H, W = stacks[0].shape
    for m in stacks:
        if tuple(m.shape) != (T, K, H, W):
            raise ValueError('all stacks must share one shape')
    dev = stacks[0].device
    pred = torch.empty(T, H, W, dtype=torch.uint8, device=dev)
    fused = torch.empty(T, K, H, W, dtype=torch.float32, device=dev) if want_fused else None
    if cm is not None:
        labels = _cuda(labels, torch.uint8)
        if tuple(labels.shape) != (T, H, W):
            raise ValueError(f'labels must be [T,H,W]={T, H, W}')
    call('sl_fuse_argmax_tiles', ptr_array(stacks), len(stacks), T, K, H * W, int(len(stacks) if n_lists is None else n_lists),
         ptr(pred), ptr(fused), ptr(labels) if cm is not None else None, int(ignore_label), ptr(cm), _stream())
    return (pred, fused) if want_fused else pred


# ============================================================== decoder tails
def _tail_out(x, C_out, out):
    B, _, h, w = x.shape
    if out is None:
        out = torch.empty(B, C_out, h, w, dtype=torch.bfloat16, device=x.device)
    elif out.dtype != torch.bfloat16 or tuple(out.shape) != (B, C_out, h, w) or not out.is_contiguous():
        raise ValueError(f'out must be a contiguous bf16 [{B},{C_out},{h},{w}] tensor')
    return out


def _tail_in(x):
    x = _cuda(x, torch.float32)
    if x.dim() != 4:
        raise ValueError(f'expected [B,C,h,w], got {tuple(x.shape)}')
    if (x.shape[2] * x.shape[3]) % 8:
        raise ValueError('h*w must be a multiple of 8 (the head pads other sizes itself: hand it fp32 features)')
    return x


@_on_input_device
def layernorm_tail(x, weight, bias, eps=1e-5, out=None):
    """The last line of FPN_Seg_OCR_Decoder.forward (networks/convnext_pop.py:27):
    `self.norm(feats.permute(0, 2, 3, 1)).permute(0, 3, 1, 2)` with nn.LayerNorm(C), returned as the contiguous bf16
    [B,C,h,w] feature tensor PopHead consumes (the reference keeps fp32; bf16 features are north_star's format).
    x is the 1x1 convolution's fp32 output [B,C,h,w]."""
    x = _tail_in(x)
    B, C, h, w = x.shape
    out = _tail_out(x, C, out)
    gamma, beta = _cuda(weight.detach(), torch.float32), _cuda(bias.detach(), torch.float32)   # keep both alive
    if gamma.numel() != C or beta.numel() != C:
        raise ValueError(f'LayerNorm weight/bias must have {C} elements')
    call('sl_tail_layernorm', ptr(x), B, C, h * w, ptr(gamma), ptr(beta), float(eps), ptr(out), _stream())
    return out


class ConvTail:
    """BN (inference) -> ReLU -> 1x1 convolution + bias -> bf16 features: PSPModule.bottleneck[1:4]
    (networks/pspnet_pop.py:19-22) and PSP_Plus_Decoder.fc[1:4] (networks/pspplus_pop.py:44-47).
    `from_sequential(seq)` takes the reference's nn.Sequential [conv3x3, norm, ReLU, conv1x1]; `__call__(x)` takes
    the 3x3 convolution's fp32 output.  The split weights are rebuilt by `refresh()` after a weight update."""

    def __init__(self, W, bias=None, bn=None, relu=True, device=None):
        dev = torch.device('cuda', torch.cuda.current_device()) if device is None else torch.device(device)
        self.device = dev
        self._W = W
        self._bias = bias
        self._bn = bn                 # (weight, bias, running_mean, running_var, eps) or None
        self.relu = bool(relu)
        self._ws = None
        self.refresh()

    @classmethod
    def from_sequential(cls, seq, **kw):
        norm, conv = seq[1], seq[3]
        if conv.kernel_size != (1, 1) or not isinstance(seq[2], torch.nn.ReLU):
            raise ValueError('expected [conv, norm, ReLU, 1x1 conv]')
        if norm.training:
            raise ValueError('ConvTail folds running statistics: put the decoder in eval() mode')
        bn = (norm.weight, norm.bias, norm.running_mean, norm.running_var, norm.eps)
        return cls(conv.weight, conv.bias, bn=bn, relu=True, device=conv.weight.device if conv.weight.is_cuda else None, **kw)

    def refresh(self):
        f32 = lambda t: None if t is None else t.detach().to(self.device, torch.float32).contiguous()
        W = f32(self._W)
        W = W.reshape(W.shape[0], -1)
        self.C_out, self.C_in = W.shape
        self.bias = f32(self._bias)
        self.bn = None if self._bn is None else tuple(f32(t) for t in self._bn[:4]) + (float(self._bn[4]),)
        self.W_hi = torch.empty(self.C_out, self.C_in, dtype=torch.int16, device=self.device)
        self.W_lo = torch.empty_like(self.W_hi)
        with torch.cuda.device(self.device):
            call('sl_tail_conv_prepare', ptr(W), self.C_out, self.C_in, ptr(self.W_hi), ptr(self.W_lo), _stream())

    @_on_input_device
    def __call__(self, x, out=None):
        x = _tail_in(x)
        B, C, h, w = x.shape
        if C != self.C_in:
            raise ValueError(f'expected {self.C_in} input channels, got {C}')
        out = _tail_out(x, self.C_out, out)
        need = _cabi.lib().sl_tail_bn_relu_conv_ws_bytes(B, C, h * w)
        if self._ws is None or self._ws.numel() < need or self._ws.device != x.device:
            self._ws = torch.empty(need, dtype=torch.uint8, device=x.device)
        bn = self.bn or (None, None, None, None, 0.0)
        call('sl_tail_bn_relu_conv', ptr(x), B, C, h * w, ptr(bn[0]), ptr(bn[1]), ptr(bn[2]), ptr(bn[3]), float(bn[4]),
             int(self.relu), ptr(self.W_hi), ptr(self.W_lo), ptr(self.bias), self.C_out, ptr(self._ws), ptr(out), _stream())
        return out


@_on_input_device
def sum_tail(maps, out=None):
    """`torch.stack(fpn_outs, dim=-1).sum(-1)` (networks/swin_pop.py:169-172, lsk_pop.py:163-165) as bf16 features:
    maps is the list of same-shape fp32 [B,C,h,w] FPN outputs (already interpolated), added in list order."""
    maps = [_tail_in(m) for m in maps]
    for m in maps:
        if m.shape != maps[0].shape:
            raise ValueError('all maps must share one shape')
    out = _tail_out(maps[0], maps[0].shape[1], out)
    call('sl_tail_sum', ptr_array(maps), len(maps), maps[0].numel(), ptr(out), _stream())
    return out


@_on_input_device
def bn_relu_tail(x, bn, relu=True, out=None):
    """Inference BatchNorm2d -> ReLU as bf16 features: the tail of _ASPP.fc (`_ConvBnReLU`, networks/deeplab_pop.py:12-29,
    61,66) and of VGGUNet.up4's DoubleConv (networks/vggunet_pop.py:19-20).  x is the convolution's fp32 output;
    bn = the nn.BatchNorm2d module (eval mode) or a (weight, bias, running_mean, running_var, eps) tuple."""
    x = _tail_in(x)
    B, C, h, w = x.shape
    if isinstance(bn, torch.nn.Module):
        if bn.training:
            raise ValueError('bn_relu_tail uses running statistics: put the module in eval() mode')
        bn = (bn.weight, bn.bias, bn.running_mean, bn.running_var, bn.eps)
    params = [_cuda(t.detach(), torch.float32) for t in bn[:4]]          # kept alive until the launch is queued
    if any(t.numel() != C for t in params):
        raise ValueError(f'BatchNorm parameters must have {C} elements')
    out = _tail_out(x, C, out)
    call('sl_tail_bn_relu', ptr(x), B, C, h * w, ptr(params[0]), ptr(params[1]), ptr(params[2]), ptr(params[3]),
         float(bn[4]), int(bool(relu)), ptr(out), _stream())
    return out


@_on_input_device
def concat_tail(maps, out=None):
    """`torch.cat([x[0], x1, x2, x3], 1)` of HRFPN_Seg_Decoder (networks/seghr_pop.py:23-24) as bf16 features:
    maps are fp32 [B,C_m,h,w] tensors of one spatial size (already interpolated)."""
    maps = [_tail_in(m) for m in maps]
    B, _, h, w = maps[0].shape
    for m in maps:
        if m.shape[0] != B or tuple(m.shape[2:]) != (h, w):
            raise ValueError('all maps must share batch and spatial size')
    chans = [int(m.shape[1]) for m in maps]
    if out is None:
        out = torch.empty(B, sum(chans), h, w, dtype=torch.bfloat16, device=maps[0].device)
    elif out.dtype != torch.bfloat16 or tuple(out.shape) != (B, sum(chans), h, w) or not out.is_contiguous():
        raise ValueError('out must be a contiguous bf16 [B,sum(C_m),h,w] tensor')
    call('sl_tail_concat', ptr_array(maps), int_array(chans), len(maps), B, h * w, ptr(out), _stream())
    return out
