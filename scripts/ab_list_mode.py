"""A/B of TL_OPT_LIST_MODE (L2 treatment of the crossing-edge list) on the C2 workload, in ONE process: the option is read
per call.  `python scripts/ab_list_mode.py` times 20 steps per mode (CUDA events) and checks loss / gradient against mode 0
bit for bit; `ncu ... python scripts/ab_list_mode.py ncu` runs 2 steps per mode for the DRAM byte counters."""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import dilabhelmholtzoct_b200 as tlb
from dilabhelmholtzoct_b200 import _lib
from dilabhelmholtzoct_b200.synthetic import make_batch

MODES = [int(x) for x in os.environ.get("AB_MODES", "0,1,3,5,7").split(",")]
under_ncu = len(sys.argv) > 1 and sys.argv[1] == "ncu"
L = _lib.lib()
pred, truth = make_batch(64, 256, 256, seed=1234 + 2000, device="cuda")
p = pred.clone().requires_grad_(True)


def step():
    p.grad = None
    loss = tlb.topo_loss(p, truth, 0.1, feat_d=1)
    loss.backward()
    return loss


ref = None
for m in MODES:
    assert L.tl_set_option(_lib.OPT_LIST_MODE, m) == 0
    if under_ncu:
        for _ in range(2):
            step()
        torch.cuda.synchronize()
        print("mode", m, "ok", flush=True)
        continue
    for _ in range(5):
        step()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    best = 1e9
    for rep in range(3):
        e0.record()
        for _ in range(20):
            loss = step()
        e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1) / 20)
    g = p.grad.clone()
    if ref is None:
        ref = (loss.clone(), g)
    same_l, same_g = bool(torch.equal(loss, ref[0])), bool(torch.equal(g, ref[1]))
    dg = float((g - ref[1]).abs().max()) / float(ref[1].abs().max())
    print(f"mode {m}: {best:.4f} ms/step (best of 3 x 20), loss {float(loss.detach()):.7f}, vs the first mode: loss identical {same_l}, "
          f"gradient identical {same_g} (max rel diff {dg:.2e}, support equal {bool(torch.equal(g != 0, ref[1] != 0))})", flush=True)
    # fingerprint for comparisons ACROSS library variants (separate processes): exact loss bits, gradient support size, |g| sum
    print(f"   fingerprint: loss {float(loss.detach()).hex()} nnz {int((g != 0).sum())} sum|g| {float(g.double().abs().sum()):.12e}", flush=True)
tlb.check_status()
