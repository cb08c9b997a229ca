"""C5-resolution and H0 timing probes (not the bench): fwd+bwd on 1024x1024 maps, H0 at 256x256."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import dilabhelmholtzoct_b200 as tlb
from dilabhelmholtzoct_b200.synthetic import make_batch

def timeit(fn, n=3, warm=1):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ts = []
    for _ in range(n):
        e0.record(); fn(); e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
    return min(ts)

for B, HW, dim in ((4, 1024, 1), (16, 1024, 1), (64, 256, 0), (64, 50, 1), (64, 50, 0), (64, 128, 1)):
    pred, truth = make_batch(B, HW, HW, seed=1, device="cuda")
    p = pred.clone().requires_grad_(True)
    def f():
        p.grad = None
        tlb.topo_loss(p, truth, 0.1, feat_d=dim).backward()
    ms = timeit(f)
    print(f"B={B} {HW}x{HW} feat_d={dim}: fwd+bwd {ms:.2f} ms -> {B*14/ms*1e3:.0f} maps/s, ws {sum(t.numel() for t in tlb.topological_loss._buffers(B,14,HW,HW,dim,torch.device('cuda',0)))/1e9:.2f} GB", flush=True)
    del pred, truth, p
    torch.cuda.empty_cache()
