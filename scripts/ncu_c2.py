"""Driver for ncu: two forward+backward steps of the C2 workload (fp32[64,14,256,256], feat_d=1)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import dilabhelmholtzoct_b200 as tlb
from dilabhelmholtzoct_b200.synthetic import make_batch
pred, truth = make_batch(64, 256, 256, seed=1234 + 2000, device="cuda")
p = pred.clone().requires_grad_(True)
for _ in range(2):
    p.grad = None
    tlb.topo_loss(p, truth, 0.1, feat_d=1).backward()
torch.cuda.synchronize()
print("ok")
