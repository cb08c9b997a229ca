"""Quick stage timing of the forward/backward on the headline shape (not the bench)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import dilabhelmholtzoct_b200 as tlb
from dilabhelmholtzoct_b200.synthetic import make_batch

def timeit(fn, n=5, warm=2):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
    ts = []
    for _ in range(n):
        ev[0].record(); fn(); ev[1].record(); torch.cuda.synchronize(); ts.append(ev[0].elapsed_time(ev[1]))
    return min(ts), sum(ts) / len(ts)

for B in (4, 16, 64):
    pred, truth = make_batch(B, 256, 256, seed=1234, device="cuda")
    p = pred.clone().requires_grad_(True)
    def fwd():
        return tlb.topo_loss(p, truth, 0.1, feat_d=1)
    def fwdbwd():
        p.grad = None
        tlb.topo_loss(p, truth, 0.1, feat_d=1).backward()
    print("B", B, "fwd ms (min, mean)", timeit(fwd), "fwd+bwd", timeit(fwdbwd), flush=True)
    for dim in (1, 0):
        f = lambda: tlb.topological_loss._TopoLossFn.apply(pred, truth, 0.1, dim, 2, False, 0)
        print("  dim", dim, "fwd", timeit(f), flush=True)
# stage probe via the inner boundary
maps = torch.cat([pred.reshape(-1, 256, 256), truth.reshape(-1, 256, 256)])
print("pairs-only(1792 maps, H1) ms", timeit(lambda: tlb.persistence_pairs(maps, 1), n=3, warm=1))
print("pairs-only(pred 896) ms", timeit(lambda: tlb.persistence_pairs(pred.reshape(-1,256,256), 1), n=3, warm=1))
print("pairs-only(truth 896) ms", timeit(lambda: tlb.persistence_pairs(truth.reshape(-1,256,256), 1), n=3, warm=1))
x = torch.rand((896, 256, 256), device="cuda")
print("pairs-only(iid 896) ms", timeit(lambda: tlb.persistence_pairs(x, 1), n=3, warm=1))
