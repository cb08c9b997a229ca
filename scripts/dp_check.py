"""Data-parallel check on real GPUs (torchrun --nproc-per-node N): the sharded training step
(parallel.training_step: scalar-loss all-reduce + decoder-gradient all-reduce over NCCL) must reproduce
the single-process step on the full batch -- same loss, same updated mask-decoder parameters."""
import os, sys, copy, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.distributed as dist
import torch.nn.functional as F
from tests.test_training_step import _tiny_sam, _batch, _seg_loss, _model_inputs, _SamWithSizes
from dilabhelmholtzoct_b200.parallel import training_step, shard_batch

rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr)
dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False
B = int(os.environ.get("DP_IMAGES_PER_RANK", "2")) * world
inputs, gt = _batch(B=B, N=3, size=64)
base = _tiny_sam().cuda()
# sharded step
m = copy.deepcopy(base)
opt = torch.optim.SGD(m.mask_decoder.parameters(), lr=0.1)
sl = shard_batch(B, rank, world)
loc = {k: v[sl].cuda() for k, v in _model_inputs(inputs).items()}
stats = {}
loss_dp = training_step(_SamWithSizes(m), loc, gt[sl].cuda(), opt, None, global_batch=B,
                        decoder_params=m.mask_decoder.parameters(), stats=stats)
# single-process step on the full batch (no process group used: world forced to 1 by a fresh wrapper call)
m1 = copy.deepcopy(base)
opt1 = torch.optim.SGD(m1.mask_decoder.parameters(), lr=0.1)
full = {k: v.cuda() for k, v in _model_inputs(inputs).items()}
import dilabhelmholtzoct_b200.parallel as par
_w = par._world
par._world = lambda group: 1
loss_1 = training_step(_SamWithSizes(m1), full, gt.cuda(), opt1, None, global_batch=B,
                       decoder_params=m1.mask_decoder.parameters())
par._world = _w
num = den = 0.0
for a, b, c in zip(m.mask_decoder.parameters(), m1.mask_decoder.parameters(), base.mask_decoder.parameters()):
    num += float(((a - c) - (b - c)).pow(2).sum()); den += float((b - c).pow(2).sum())
# timing of the data-parallel step at the C3 per-GPU batch (32 images x 14 prompts, 496x512 originals) with the
# real-size mask decoder (4 058 340 parameters = 16.2 MB of fp32 gradients per all-reduce)
def _time(fn, n=5, warm=2):
    for _ in range(warm):
        fn()
    dist.barrier(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record(); dist.barrier(); torch.cuda.synchronize()
    t = torch.tensor([e0.elapsed_time(e1) / n], device="cuda")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t)

# Model of the timed step (DP_MODEL): "tiny" = stand-in vision encoder + the real mask decoder (default, cheap);
# "vit_b" = transformers' SamModel(SamConfig()) = SAM ViT-B (93.7 M parameters, BASELINE configs[2], random init: no weights
# offline); "vit_l" = SAM ViT-L (hidden 1024, 24 layers, 16 heads, global attention at 5/11/17/23; configs[3]).
# DP_PROMPT = boxes | points (training.py --prompt).  DP_LOCAL_BATCH: images per GPU (C3: 32, C4 at 8 GPUs: 64).
model_kind, prompt = os.environ.get("DP_MODEL", "tiny"), os.environ.get("DP_PROMPT", "boxes")
Bl, Nmax = int(os.environ.get("DP_LOCAL_BATCH", "32")), 14
n_time, n_warm = (5, 2) if model_kind == "tiny" else (3, 1)


def _big_model():
    if model_kind == "tiny":
        return copy.deepcopy(base)
    from transformers import SamConfig, SamModel
    from transformers.models.sam.configuration_sam import SamVisionConfig
    torch.manual_seed(0)
    cfg = SamConfig() if model_kind == "vit_b" else SamConfig(vision_config=SamVisionConfig(
        hidden_size=1024, num_hidden_layers=24, num_attention_heads=16, global_attn_indexes=[5, 11, 17, 23], mlp_dim=4096))
    m_ = SamModel(cfg)
    for name, p_ in m_.named_parameters():  # prepare_model, training_utils.py:277-279
        if name.startswith("vision_encoder") or name.startswith("prompt_encoder"):
            p_.requires_grad_(False)
    return m_.cuda()


class _Sam(torch.nn.Module):  # training_step calls model(**inputs): drop the two size entries the HF model does not take
    def __init__(self, sam):
        super().__init__()
        self.sam = sam

    def forward(self, pixel_values, input_boxes=None, input_points=None, reshaped_input_sizes=None, original_sizes=None, multimask_output=False):
        if input_points is not None:
            return self.sam(pixel_values=pixel_values, input_points=input_points, multimask_output=multimask_output)
        return self.sam(pixel_values=pixel_values, input_boxes=input_boxes, multimask_output=multimask_output)


g = torch.Generator().manual_seed(100 + rank)
big = {"pixel_values": torch.randn((Bl, 3, 1024, 1024), generator=g).cuda(),
       "reshaped_input_sizes": torch.tensor([[992, 1024]] * Bl).cuda(), "original_sizes": torch.tensor([[496, 512]] * Bl).cuda()}
if prompt == "points":
    big["input_points"] = (torch.rand((Bl, Nmax, 1, 2), generator=g) * 900 + 50).cuda()
else:
    big["input_boxes"] = (torch.rand((Bl, Nmax, 4), generator=g) * 500 + torch.tensor([0.0, 0, 400, 400])).cuda()
# ground truth: OCT-like layers + blobs (dilabhelmholtzoct_b200.synthetic), one component mask per prompt -- NOT iid noise,
# whose diagrams (hundreds of points on both sides) would make the exact assignment, not the step, the thing measured
from dilabhelmholtzoct_b200.synthetic import make_labels
_gen = torch.Generator().manual_seed(300 + rank)
_lab = make_labels(Bl, 496, 512, _gen, n_classes=Nmax)
gt_big = torch.nn.functional.one_hot(_lab, Nmax).permute(0, 3, 1, 2).float().contiguous().cuda()
mt = _big_model()
optt = torch.optim.Adam(mt.mask_decoder.parameters(), lr=1e-3)
dec = list(mt.mask_decoder.parameters())
st2 = {}
step_fn = lambda: training_step(_Sam(mt), big, gt_big, optt, None, global_batch=Bl * world, decoder_params=dec, stats=st2)
ms_step = _time(step_fn, n=n_time, warm=n_warm)
ms_step_no_topo = _time(lambda: training_step(_Sam(mt), big, gt_big, optt, None, topological=False, global_batch=Bl * world, decoder_params=dec),
                        n=n_time, warm=1)
# the part of the step this repo owns: post-processing + DiceCE + topological loss, forward and backward to the decoder output
import dilabhelmholtzoct_b200 as tlb
_t256 = torch.nn.functional.interpolate(gt_big, (256, 256), mode="bilinear", align_corners=False)
low = (4.0 * (2.0 * _t256 - 1.0) + 0.5 * torch.randn((Bl, Nmax, 256, 256), device="cuda", generator=torch.Generator(device="cuda").manual_seed(5 + rank))).requires_grad_(True)
def loss_side():
    low.grad = None
    masks = tlb.postprocess_masks(low, (992, 1024), (496, 512))
    (tlb.dice_ce_loss(masks, gt_big) + tlb.topo_loss_from_logits(masks, gt_big, 0.1, feat_d=1, interp=50)).backward()
def loss_side_torch():
    low.grad = None
    m_ = F.interpolate(low, (1024, 1024), mode="bilinear", align_corners=False)[..., :992, :1024]
    m_ = F.interpolate(m_, (496, 512), mode="bilinear", align_corners=False)
    s_ = torch.sigmoid(m_)
    ax = (2, 3)
    dice = (1 - (2 * (s_ * gt_big).sum(ax) + 1e-5) / (gt_big.sum(ax) + s_.sum(ax) + 1e-5)).mean()
    (dice + F.cross_entropy(m_, gt_big) + tlb.topo_loss(s_, gt_big, 0.1, feat_d=1, interp=50)).backward()
ms_loss_side = _time(loss_side, n=10, warm=3)
ms_loss_side_torch = _time(loss_side_torch, n=10, warm=3)
from dilabhelmholtzoct_b200.parallel import allreduce_gradients
ms_ar = _time(lambda: allreduce_gradients(dec), n=20, warm=3)
out = {"rank": rank, "world": world, "loss_dp": float(loss_dp), "loss_single": float(loss_1),
       "rel_update_err": (num / max(den, 1e-30)) ** 0.5,
       "step": {"model": model_kind, "prompt": prompt, "model_params": sum(p.numel() for p in mt.parameters()),
                "local_batch": Bl, "global_batch": Bl * world, "prompts": Nmax, "original_size": [496, 512],
                "ms_per_step_max_over_ranks": ms_step, "ms_per_step_without_topo": ms_step_no_topo,
                "images_per_s": Bl * world / (ms_step * 1e-3),
                "decoder_params": sum(p.numel() for p in dec), "grad_allreduce_bytes": st2.get("grad_allreduce_bytes"),
                "grad_allreduce_ms": ms_ar,
                "loss_side_ms": ms_loss_side, "loss_side_pytorch_ops_ms": ms_loss_side_torch,
                "grad_allreduce_GBps": (st2.get("grad_allreduce_bytes") or 0) / max(ms_ar, 1e-9) / 1e6,
                "note": "DP_MODEL=tiny: tiny random vision encoder + the real-size SAM mask decoder; vit_b / vit_l: transformers SamModel, random init (no weights offline); "
                        "postprocess_masks + dice_ce_loss + fused topo loss (interp=50) + SUM all-reduce of decoder gradients; "
                        "loss_side = everything between the decoder output and its gradient (this repo's kernels vs PyTorch ops "
                        "around the CUDA topological loss); the step time itself is dominated by the (frozen) vision encoder's forward pass"}}
gathered = [None] * world
dist.all_gather_object(gathered, out)
if rank == 0:
    print(json.dumps(gathered))
    assert all(abs(g["loss_dp"] - g["loss_single"]) <= 1e-4 * abs(g["loss_single"]) for g in gathered)
    assert all(g["rel_update_err"] <= 1e-3 for g in gathered)
    print("dp check ok")
dist.destroy_process_group()
