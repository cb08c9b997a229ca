"""F1 measurement: the reference call site  topo_loss(sigmoid(masks), gt, 0.1, feat_d=1, interp=50)
on [64, 14, 496, 512] logits -- two-step form (torch.sigmoid + F.interpolate + topo_loss) vs the fused
form (topo_loss_from_logits), and the resample kernels alone with their HBM figures."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import dilabhelmholtzoct_b200 as tlb


def timeit(fn, n=10, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(n):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    ts.sort()
    return ts[len(ts) // 2]


B, C, H, W, S = 64, 14, 496, 512, 50
from dilabhelmholtzoct_b200.synthetic import make_batch
gen = torch.Generator(device="cuda").manual_seed(99)
pred, gt = make_batch(B, H, W, seed=1234 + 6000, device="cuda")   # OCT-like layers + blobs, call-site size
logits = torch.logit(pred.clamp(1e-6, 1 - 1e-6)).contiguous()
del pred
x = logits.clone().requires_grad_(True)


def two_step():
    x.grad = None
    tlb.topo_loss(torch.sigmoid(x), gt, 0.1, feat_d=1, interp=S).backward()


def fused():
    x.grad = None
    tlb.topo_loss_from_logits(x, gt, 0.1, feat_d=1, interp=S).backward()


g = torch.randn((B, C, S, S), device="cuda", generator=gen)
xs = logits.clone().requires_grad_(True)
res = {
    "shape": [B, C, H, W], "interp": S,
    "two_step_ms": timeit(two_step), "fused_ms": timeit(fused),
    "resample_fwd_pred_ms": timeit(lambda: tlb.resample(logits, S, sigmoid=True)),
    "resample_fwd_truth_ms": timeit(lambda: tlb.resample(gt, S)),
    "torch_sigmoid_interp_fwd_ms": timeit(lambda: torch.nn.functional.interpolate(torch.sigmoid(logits), size=(S, S), mode="bilinear", align_corners=True)),
}
y = tlb.resample(xs, S, sigmoid=True)
res["resample_bwd_ms"] = timeit(lambda: torch.autograd.grad(y, xs, g, retain_graph=True))
yt = torch.nn.functional.interpolate(torch.sigmoid(xs), size=(S, S), mode="bilinear", align_corners=True)
res["torch_sigmoid_interp_bwd_ms"] = timeit(lambda: torch.autograd.grad(yt, xs, g, retain_graph=True))
# ---- F3: SAM post-processing chain 256x256 -> 1024x1024 -> crop 992x1024 -> 496x512 (training_utils.py:57-59)
import torch.nn.functional as F
pm = torch.randn((B, C, 256, 256), device="cuda", generator=gen)
gm = torch.randn((B, C, H, W), device="cuda", generator=gen)


def torch_chain(t):
    m = F.interpolate(t, (1024, 1024), mode="bilinear", align_corners=False)
    m = m[..., :992, :1024]
    return F.interpolate(m, (H, W), mode="bilinear", align_corners=False)


pa = pm.clone().requires_grad_(True)
pb = pm.clone().requires_grad_(True)
res["postprocess_fwd_ms"] = timeit(lambda: tlb.postprocess_masks(pm, (992, 1024), (H, W)))
res["torch_postprocess_fwd_ms"] = timeit(lambda: torch_chain(pm))
ya = tlb.postprocess_masks(pa, (992, 1024), (H, W))
res["postprocess_bwd_ms"] = timeit(lambda: torch.autograd.grad(ya, pa, gm, retain_graph=True))
yb = torch_chain(pb)
res["torch_postprocess_bwd_ms"] = timeit(lambda: torch.autograd.grad(yb, pb, gm, retain_graph=True))
del yb
res["postprocess_fwd_GBps"] = (B * C * (256 * 256 + H * W) * 4) / (res["postprocess_fwd_ms"] * 1e-3) / 1e9   # read source + write output
res["postprocess_bwd_GBps"] = (B * C * (256 * 256 + H * W) * 4) / (res["postprocess_bwd_ms"] * 1e-3) / 1e9   # read grad_out + write grad_in
# ---- F2: DiceCELoss(sigmoid=True) on the post-processed masks (training_utils.py:62), fused kernel vs the PyTorch ops
from oracle.dice_ce_oracle import dice_ce   # (a measurement script may use the oracle as the PyTorch-ops arm)
lx = torch.randn((B, C, H, W), device="cuda", generator=gen).requires_grad_(True)
tg = (torch.rand((B, C, H, W), device="cuda", generator=gen) < 0.3).float()
def dc_fused():
    lx.grad = None
    tlb.dice_ce_loss(lx, tg).backward()
def dc_torch():
    lx.grad = None
    dice_ce(lx, tg).backward()
res["dice_ce_fused_fwd_bwd_ms"] = timeit(dc_fused)
res["dice_ce_torch_fwd_bwd_ms"] = timeit(dc_torch)
res["dice_ce_fused_GBps"] = (B * C * H * W * 4 * 5) / (res["dice_ce_fused_fwd_bwd_ms"] * 1e-3) / 1e9   # fwd reads 2, bwd reads 2 + writes 1
bytes_grad = B * C * H * W * 4
res["resample_bwd_write_GBps"] = bytes_grad / (res["resample_bwd_ms"] * 1e-3) / 1e9
res["dense_grad_bytes"] = bytes_grad
print(json.dumps(res))
