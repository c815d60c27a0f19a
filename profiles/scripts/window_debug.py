import sys, torch
sys.path.insert(0, '.')
from oracle import ref_ops
from segland_b200 import ops
for tile, crop, stride, ms, flips in (((1024, 1024), (512, 512), (384, 384), 8, (0, 1)), ((1024, 1024), (512, 512), (384, 384), 8, (0,)),
                                      ((1024, 1024), (512, 512), (256, 256), 8, (0, 1))):
    plan = ops.WindowPlan(tile, crop, stride, ms)
    B, K = 2, 12
    hc, wc = plan.crop_lr_hw
    E = plan.n_windows * len(flips)
    crops = torch.randn(B, E, K, hc, wc, generator=torch.Generator().manual_seed(7))
    ref, ref_cnt = ref_ops.ref_window_accumulate(crops, [y // ms for y in plan.origins_y], [x // ms for x in plan.origins_x], flips, plan.canvas_hw)
    got, cnt = ops.window_accumulate(crops.cuda(), plan, flips, want_count=True)
    d = (got.cpu() - ref).abs()
    bad = d > 0
    print(stride, flips, 'max diff', d.max().item(), 'mismatches', int(bad.sum()), 'of', bad.numel(), 'count equal', torch.equal(cnt.cpu(), ref_cnt))
    if bad.any():
        ys, xs = torch.nonzero(bad.any(0).any(0), as_tuple=True)
        print('  rows', sorted(set(ys.tolist()))[:20], 'cols', sorted(set(xs.tolist()))[:40])
        i = torch.nonzero(bad)[0].tolist()
        print('  first', i, got.cpu()[tuple(i)].item(), ref[tuple(i)].item(), 'count there', ref_cnt[i[2], i[3]].item())
