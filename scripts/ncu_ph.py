"""Small driver for ncu: persistence kernel on 148 pred maps (dim 1)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import dilabhelmholtzoct_b200 as tlb
from dilabhelmholtzoct_b200.synthetic import make_batch
pred, truth = make_batch(11, 256, 256, seed=1234, device="cuda")
P = pred.reshape(-1, 256, 256)[:148].contiguous()
out = tlb.persistence_pairs(P, 1)
torch.cuda.synchronize()
print("ok", len(out))
