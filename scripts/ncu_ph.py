"""Small driver for ncu: persistence kernel (dim 1) on 148 synthetic maps, one per SM.
usage: python scripts/ncu_ph.py [pred|truth]"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import dilabhelmholtzoct_b200 as tlb
from dilabhelmholtzoct_b200.synthetic import make_batch
which = sys.argv[1] if len(sys.argv) > 1 else "pred"
pred, truth = make_batch(16, 256, 256, seed=1234, device="cuda")
src = pred if which == "pred" else truth
maps = src.reshape(-1, 256, 256)
if which == "truth":  # skip the constant maps of absent classes (they take the short-cut)
    keep = (maps.amax(dim=(1, 2)) != maps.amin(dim=(1, 2))).nonzero().flatten()
    maps = maps[keep]
P = maps[:148].contiguous()
out = tlb.persistence_pairs(P, 1)
torch.cuda.synchronize()
print("ok", which, len(out))
