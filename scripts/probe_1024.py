import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import dilabhelmholtzoct_b200 as tlb
from dilabhelmholtzoct_b200.synthetic import make_batch
def timeit(fn, n=2, warm=1):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ts = []
    for _ in range(n):
        e0.record(); fn(); e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
    return min(ts)
pred, truth = make_batch(2, 1024, 1024, seed=1, device="cuda")
P = pred.reshape(-1, 1024, 1024); T = truth.reshape(-1, 1024, 1024)
X = torch.rand((28, 1024, 1024), device="cuda")
for tag, m in (("pred", P), ("truth", T), ("iid", X), ("one pred map", P[:1]), ("one truth map", T[3:4])):
    for dim in (1, 0):
        print(tag, "dim", dim, "maps", m.shape[0], "ms", round(timeit(lambda: tlb.persistence_pairs(m, dim)), 2), flush=True)
