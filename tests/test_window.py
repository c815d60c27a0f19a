"""Sliding-window / flip aggregation (spec: this repo -- the reference has no sliding-window inference,
engine.py:23-143, eval_ft.py:162-172).  CPU: the window plan and the oracle composition;
GPU: sl_window_accumulate against that composition (bit-exact) and end to end through the fused
up-sampling/argmax against F.interpolate + argmax."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

from oracle import ref_ops
from segland_b200 import ops

GEOMETRIES = [
    # tile, crop, stride, model stride, flips
    ((1024, 1024), (512, 512), (256, 256), 8, (0,)),          # regular 3 x 3 grid, 50 % overlap
    ((1024, 1024), (512, 512), (384, 384), 8, (0, 1)),        # ragged last window (0, 384, 512), h-flip TTA
    ((1024, 768), (512, 384), (344, 344), 8, (0, 1, 2, 3)),   # rectangular, origins not multiples of 4 feature px
    ((1024, 1024), (512, 512), (512, 512), 4, (0, 1)),        # no overlap, stride-4 model (Swin / ConvNeXt)
    ((640, 640), (640, 640), None, 8, (0, 2)),                # one window = plain flip TTA
]


def test_window_plan_origins():
    p = ops.WindowPlan((1024, 1024), (512, 512), (384, 384), 8)
    assert p.origins_y == [0, 384, 512] and p.origins_x == [0, 384, 512]
    assert p.n_windows == 9 and p.canvas_hw == (128, 128) and p.crop_lr_hw == (64, 64)
    assert p.windows()[:4] == [(0, 0), (0, 384), (0, 512), (384, 0)]
    assert ops.WindowPlan((512, 512), (512, 512)).windows() == [(0, 0)]
    for bad in (((1024, 1024), (500, 512), None, 8), ((1024, 1024), (512, 512), (640, 640), 8),
                ((512, 512), (1024, 1024), None, 8), ((1024, 1024), (512, 512), (100, 100), 8)):
        with pytest.raises(ValueError):
            ops.WindowPlan(*bad)


def test_window_plan_covers_every_pixel_and_crop_roundtrip():
    for tile, crop, stride, ms, flips in GEOMETRIES:
        p = ops.WindowPlan(tile, crop, stride, ms)
        cover = np.zeros(tile, dtype=np.int32)
        for (y, x) in p.windows():
            assert y % ms == 0 and x % ms == 0
            cover[y:y + crop[0], x:x + crop[1]] += 1
        assert cover.min() >= 1
        # crops cut from one image and stitched back give the image (every overlap averages equal values)
        h, w = p.canvas_hw
        img = torch.randn(2, 3, h, w)
        plan_lr = ops.WindowPlan((h, w), p.crop_lr_hw, tuple(s // ms for s in p.stride_hw), 1)
        assert [o // ms for o in p.origins_y] == plan_lr.origins_y
        crops = plan_lr.crop(img, flips)
        out, cnt = ref_ops.ref_window_accumulate(crops, plan_lr.origins_y, plan_lr.origins_x, flips, (h, w))
        assert np.array_equal(cnt.numpy(), cover[::ms, ::ms] * len(flips))
        assert torch.allclose(out, img, rtol=1e-6, atol=1e-6)


@pytest.mark.gpu
@pytest.mark.parametrize('geom', GEOMETRIES)
@pytest.mark.parametrize('layout', ['bekhw', 'ebkhw'])
def test_window_accumulate_matches_composition(geom, layout):
    tile, crop, stride, ms, flips = geom
    plan = ops.WindowPlan(tile, crop, stride, ms)
    B, K = 2, 12
    hc, wc = plan.crop_lr_hw
    E = plan.n_windows * len(flips)
    g = torch.Generator().manual_seed(7)
    crops = torch.randn(B, E, K, hc, wc, generator=g)
    ref, ref_cnt = ref_ops.ref_window_accumulate(crops, [y // ms for y in plan.origins_y],
                                                 [x // ms for x in plan.origins_x], flips, plan.canvas_hw)
    dev_in = crops.cuda() if layout == 'bekhw' else crops.transpose(0, 1).contiguous().cuda()
    got, cnt = ops.window_accumulate(dev_in, plan, flips, layout=layout, want_count=True)
    assert torch.equal(cnt.cpu(), ref_cnt)
    assert torch.equal(got.cpu(), ref), f'max diff {(got.cpu() - ref).abs().max().item():.3e}'


@pytest.mark.gpu
def test_window_accumulate_feeds_fused_upsample():
    """crops -> canvas -> fused x8 up-sampling / argmax / softmax equals the PyTorch composition on the canvas."""
    plan = ops.WindowPlan((1024, 1024), (512, 512), (384, 384), 8)
    flips = (0, 1)
    B, K = 2, 8
    g = torch.Generator().manual_seed(11)
    # crops of one smooth-ish field plus per-crop noise, so the argmax is not trivially the same in every window
    field = F.interpolate(torch.randn(B, K, 16, 16, generator=g), size=plan.canvas_hw, mode='bilinear', align_corners=True)
    lr_plan = ops.WindowPlan(plan.canvas_hw, plan.crop_lr_hw, (48, 48), 1)
    crops = lr_plan.crop(field, flips) + 0.05 * torch.randn(B, plan.n_windows * 2, K, *plan.crop_lr_hw, generator=g)
    ref_canvas, _ = ref_ops.ref_window_accumulate(crops, lr_plan.origins_y, lr_plan.origins_x, flips, plan.canvas_hw)
    canvas = ops.window_accumulate(crops.cuda(), plan, flips)
    assert torch.equal(canvas.cpu(), ref_canvas)
    out = ops.upsample_argmax(canvas, plan.tile_hw, want_probs=True, want_conf=True)
    hr = F.interpolate(ref_canvas, size=plan.tile_hw, mode='bilinear', align_corners=True)
    ref_pred = hr.argmax(1).numpy().astype(np.uint8)
    pred = out['pred'].cpu().numpy()
    agree = (pred == ref_pred).mean()
    assert agree >= 0.9999, agree
    bad = np.nonzero(pred != ref_pred)
    if len(bad[0]):
        a = hr.numpy()[bad[0], pred[bad].astype(np.int64), bad[1], bad[2]]
        b = hr.numpy()[bad[0], ref_pred[bad].astype(np.int64), bad[1], bad[2]]
        assert np.abs(a - b).max() <= 1e-5 * np.abs(hr.numpy()).max()
    probs = torch.softmax(hr, 1)
    assert (out['probs'].cpu() - probs).abs().max().item() <= 1e-3 * probs.max().item()
    assert (out['conf'].cpu() - probs.max(1)[0]).abs().max().item() <= 1e-3


@pytest.mark.gpu
def test_window_accumulate_rejects_bad_plans():
    from segland_b200._cabi import SeglandError
    plan = ops.WindowPlan((512, 512), (256, 256), (256, 256), 8)
    crops = torch.zeros(1, 4, 3, 32, 32, device='cuda')
    with pytest.raises(ValueError):
        ops.window_accumulate(crops, plan, flips=(0, 1))                  # 8 entries expected
    with pytest.raises(ValueError):
        ops.window_accumulate(crops[:, :, :, :16], plan)
    with pytest.raises(SeglandError):
        ops.window_accumulate(crops, plan, flips=(7,) * 1)                # flip code out of range
