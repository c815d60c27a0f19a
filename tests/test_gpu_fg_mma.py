"""GPU: the tensor-path foreground projections (pop_fg_mma.cu, K > 8 classes; networks/pspnet_pop.py:108-109,114-115 +
the alpha/beta collapse of the classifier) against the FFMA2 kernel and the fp32 oracle, over channel counts that are
not multiples of 16, ragged pixel counts, more than 16 classes (two passes) and both forced modes."""
import pytest
import torch

from helpers import RTOL, assert_close_rel
from oracle import ref_ops
from segland_b200 import _cabi, synth

pytestmark = pytest.mark.gpu


@pytest.fixture(scope='module')
def ops():
    from segland_b200 import ops as _ops
    _ops.check_device()
    return _ops


def fg_both(ops, st, feats):
    head = ops.PopHead(st.base_emb, st.cls, st.novel_emb, st.cls_n, bg_mode='simt')
    out = {}
    try:
        for mode in (1, 0):
            _cabi.set_env(SL_FG_MMA=mode)
            lg = torch.full((feats.shape[0], head.n_classes, *feats.shape[-2:]), 7.5, device='cuda')
            head(feats, out=lg, fg_only=True)
            out[mode] = lg.clone()
    finally:
        _cabi.set_env(SL_FG_MMA=None)
    return out[1], out[0]


@pytest.mark.parametrize('C,Kb,Kn,hw,B', [
    (512, 7, 4, (32, 32), 2),      # PSPNet ft
    (192, 7, 4, (64, 64), 1),      # ConvNeXt-T ft
    (96, 7, 4, (64, 64), 2),       # Swin ft
    (24, 7, 4, (16, 8), 1),        # C % 16 != 0: the last k-step is half empty (TMA zero-fill)
    (8, 5, 4, (8, 8), 1),
    (480, 7, 4, (15, 8), 3),       # N = 120: a partial 128-pixel item
    (64, 7, 0, (16, 16), 1),       # K = 7 forced onto the tensor path (one pass, 9 idle rows)
    (64, 9, 7, (24, 8), 1),        # K = 16: a full pass
    (64, 20, 11, (16, 16), 2),     # K = 31: two passes
    (128, 7, 4, (120, 120), 1),    # 14,400 pixels: 112.5 items
])
def test_fg_mma_matches_ffma_and_oracle(ops, C, Kb, Kn, hw, B):
    st = synth.make_head_state(C, Kb, Kn, seed=C + Kb + Kn)
    feats = synth.make_random_features(B, C, hw[0], hw[1], seed=C).cuda()
    mma, ffma = fg_both(ops, st, feats)
    assert torch.equal(mma[:, 0], ffma[:, 0]) and float(mma[:, 0].max()) == 7.5          # channel 0 (bg) untouched
    assert_close_rel(mma[:, 1:].cpu(), ffma[:, 1:].cpu(), 1e-5, f'mma vs ffma C={C} K={Kb + Kn}')
    if hw[0] * hw[1] <= 4096:
        ref = ref_ops.ref_head(feats.cpu().float(), st.base_emb, st.novel_emb, st.cls, st.cls_n)
        assert_close_rel(mma[:, 1:].cpu(), ref[:, 1:], RTOL, f'mma vs oracle C={C} K={Kb + Kn}')
    # bit-reproducible
    again, _ = fg_both(ops, st, feats)
    assert torch.equal(again, mma)


def test_fg_default_is_the_tensor_path(ops):
    """Default dispatch: the mma.sync kernel for every class count (SL_FG_MMA=0 is the A/B switch)."""
    for Kn, forced in ((4, 1), (0, 1)):
        st = synth.make_head_state(64, 7, Kn, seed=3)
        feats = synth.make_random_features(1, 64, 16, 16, seed=3).cuda()
        head = ops.PopHead(st.base_emb, st.cls, st.novel_emb, st.cls_n, bg_mode='simt')
        auto = head(feats, fg_only=True)[:, 1:].clone()
        pair = fg_both(ops, st, feats)
        assert torch.equal(auto, pair[0 if forced == 1 else 1][:, 1:])
