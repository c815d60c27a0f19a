import sys, torch
sys.path.insert(0, '.')
from segland_b200 import ops
g = torch.Generator('cuda').manual_seed(0)
for hw in (128, 256):
    lgt = torch.randn(8, 12, hw, hw, device='cuda', generator=g).requires_grad_(True)
    tgt = torch.randint(0, 12, (8, 1024, 1024), device='cuda', generator=g)
    for _ in range(2):
        lgt.grad = None
        ops.seg_cross_entropy(lgt, tgt).backward()
torch.cuda.synchronize()
