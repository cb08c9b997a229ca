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

import ctypes
from dilabhelmholtzoct_b200 import _lib
def profile(pred, truth, dim):
    _lib.lib().tl_set_option(_lib.OPT_PROFILE, 1)
    B, C, H, W = pred.shape
    ws, sc = tlb.topological_loss._buffers(B, C, H, W, dim, pred.device)
    loss = torch.empty((), device="cuda")
    _lib.lib().tl_forward(pred.data_ptr(), truth.data_ptr(), B, C, H, W, dim, 2.0, 0.1, 0, 0, ws.data_ptr(), ws.numel(), sc.data_ptr(), sc.numel(), loss.data_ptr(), torch.cuda.current_stream().cuda_stream)
    out = (ctypes.c_ulonglong * 8)()
    _lib.lib().tl_debug_profile.argtypes = [ctypes.c_void_p, ctypes.c_void_p]
    _lib.lib().tl_debug_profile(ws.data_ptr(), out)
    _lib.lib().tl_set_option(_lib.OPT_PROFILE, 0)
    tot = sum(out) or 1
    return [round(100.0 * v / tot, 1) for v in out], tot

for B in (16, 64):
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
print("phase share % [init, L0, flatten, census, merge, emit] pred+truth dim1:", profile(pred, truth, 1))
# stage probe via the inner boundary
maps = torch.cat([pred.reshape(-1, 256, 256), truth.reshape(-1, 256, 256)])
print("pairs-only(1792 maps, H1) ms", timeit(lambda: tlb.persistence_pairs(maps, 1), n=3, warm=1))
print("pairs-only(pred 896) ms", timeit(lambda: tlb.persistence_pairs(pred.reshape(-1,256,256), 1), n=3, warm=1))
print("pairs-only(truth 896) ms", timeit(lambda: tlb.persistence_pairs(truth.reshape(-1,256,256), 1), n=3, warm=1))
x = torch.rand((896, 256, 256), device="cuda")
print("pairs-only(iid 896) ms", timeit(lambda: tlb.persistence_pairs(x, 1), n=3, warm=1))
