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
B = 2 * world
inputs, gt = _batch(B=B, N=3, size=64)
base = _tiny_sam().cuda()
# sharded step
m = copy.deepcopy(base)
opt = torch.optim.SGD(m.mask_decoder.parameters(), lr=0.1)
sl = shard_batch(B, rank, world)
loc = {k: v[sl].cuda() for k, v in _model_inputs(inputs).items()}
loss_dp = training_step(_SamWithSizes(m), loc, gt[sl].cuda(), opt, _seg_loss, global_batch=B,
                        decoder_params=m.mask_decoder.parameters())
# single-process step on the full batch (no process group used: world forced to 1 by a fresh wrapper call)
m1 = copy.deepcopy(base)
opt1 = torch.optim.SGD(m1.mask_decoder.parameters(), lr=0.1)
full = {k: v.cuda() for k, v in _model_inputs(inputs).items()}
import dilabhelmholtzoct_b200.parallel as par
_w = par._world
par._world = lambda group: 1
loss_1 = training_step(_SamWithSizes(m1), full, gt.cuda(), opt1, _seg_loss, global_batch=B,
                       decoder_params=m1.mask_decoder.parameters())
par._world = _w
num = den = 0.0
for a, b, c in zip(m.mask_decoder.parameters(), m1.mask_decoder.parameters(), base.mask_decoder.parameters()):
    num += float(((a - c) - (b - c)).pow(2).sum()); den += float((b - c).pow(2).sum())
out = {"rank": rank, "world": world, "loss_dp": float(loss_dp), "loss_single": float(loss_1),
       "rel_update_err": (num / max(den, 1e-30)) ** 0.5}
gathered = [None] * world
dist.all_gather_object(gathered, out)
if rank == 0:
    print(json.dumps(gathered))
    assert all(abs(g["loss_dp"] - g["loss_single"]) <= 1e-4 * abs(g["loss_single"]) for g in gathered)
    assert all(g["rel_update_err"] <= 1e-3 for g in gathered)
    print("dp check ok")
dist.destroy_process_group()
