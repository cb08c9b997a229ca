import sys, os, ctypes
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ["TL_PROFILE"] = "1"
import torch
from scripts.stats_probe import load, run
from dilabhelmholtzoct_b200.synthetic import make_batch
pred, truth = make_batch(2, 1024, 1024, seed=1, device="cuda")
P = pred.reshape(-1, 1024, 1024).contiguous(); T = truth.reshape(-1, 1024, 1024).contiguous()
for name in ("libtopoloss_stats.so",):
    L = load(name)
    for tag, m in (("pred", P), ("truth", T)):
        for dim in (1, 0):
            run(L, m, dim, True)
            print(tag, "dim", dim, run(L, m, dim, True), flush=True)
