import sys, torch
sys.path.insert(0, '.')
from segland_b200 import ops
g = torch.Generator('cuda').manual_seed(0)
feats = torch.randn(20, 512, 128, 128, device='cuda', generator=g).to(torch.bfloat16)
masks = (torch.rand(20, 1, 1024, 1024, device='cuda', generator=g) < 0.3).float()
for _ in range(3):
    ops.masked_average_pooling(feats, masks)
torch.cuda.synchronize()
