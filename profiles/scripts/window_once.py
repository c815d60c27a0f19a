import sys, torch
sys.path.insert(0, '.')
from segland_b200 import ops
B = int(sys.argv[1]) if len(sys.argv) > 1 else 32
plan = ops.WindowPlan((1024, 1024), (512, 512), (256, 256), 4)
flips = (0, 1)
crops = torch.randn(B, plan.n_windows * 2, 12, *plan.crop_lr_hw, device='cuda')
out = torch.empty(B, 12, *plan.canvas_hw, device='cuda')
for _ in range(3): ops.window_accumulate(crops, plan, flips, out=out)
torch.cuda.synchronize()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
for _ in range(10): ops.window_accumulate(crops, plan, flips, out=out)
b.record(); torch.cuda.synchronize()
t = a.elapsed_time(b) / 10 * 1e-3
byts = (crops.numel() + out.numel()) * 4
print(f'B={B} {t*1e6:.1f} us {byts/t/1e9:.0f} GB/s')
