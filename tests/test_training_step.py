"""The training-step wrapper (parallel.training_step) against a literal copy of the reference's loop
body (/root/reference/octsam/models/training_utils.py:55-68) on a SAM whose mask decoder is the real one
(4 058 340 parameters, as ViT-B's) behind a tiny random vision encoder -- BASELINE configs[0] in spirit:
bs = 2, one synthetic 256x256 batch, boxes prompt.  On CPU the CPU oracle stands in for the CUDA op (test
infrastructure); on the GPU the CUDA op is compared with that CPU run."""
import copy

import numpy as np
import pytest
import torch
import torch.nn.functional as F

import oracle
from tests.test_parallel_gloo import _oracle_fn


def _tiny_sam():
    from transformers import SamConfig, SamModel
    from transformers.models.sam.configuration_sam import SamVisionConfig
    torch.manual_seed(0)
    vc = SamVisionConfig(hidden_size=32, output_channels=256, num_hidden_layers=2, num_attention_heads=2,
                         mlp_dim=64, global_attn_indexes=[1], window_size=7)
    model = SamModel(SamConfig(vision_config=vc))
    for name, p in model.named_parameters():  # prepare_model, training_utils.py:277-279
        if name.startswith("vision_encoder") or name.startswith("prompt_encoder"):
            p.requires_grad_(False)
    return model


def _batch(B=2, N=3, size=64):
    g = torch.Generator().manual_seed(1)
    inputs = {
        "pixel_values": torch.randn((B, 3, 1024, 1024), generator=g),
        "input_boxes": torch.tensor([[[100.0, 100, 400, 400], [200, 300, 800, 900], [10, 10, 1000, 500]][:N]] * B),
        "reshaped_input_sizes": torch.tensor([[1024, 1024]] * B),
        "original_sizes": torch.tensor([[size, size]] * B),
    }
    gt = (torch.rand((B, N, size, size), generator=g) < 0.3).float()
    return inputs, gt


def _seg_loss(masks, gt):  # stand-in for monai DiceCELoss(sigmoid=True), which is not installed
    return F.binary_cross_entropy_with_logits(masks, gt)


def _reference_step(model, inputs, gt_masks, optimizer, topo):
    """training_utils.py:55-68, literally (prompt = bboxes)."""
    optimizer.zero_grad()
    outputs = model(pixel_values=inputs["pixel_values"], input_boxes=inputs["input_boxes"], multimask_output=False)
    masks = F.interpolate(outputs.pred_masks.squeeze(2), (1024, 1024), mode="bilinear", align_corners=False)
    masks = masks[..., : inputs["reshaped_input_sizes"][0, 0], : inputs["reshaped_input_sizes"][0, 1]]
    masks = F.interpolate(masks, (inputs["original_sizes"][0, 0], inputs["original_sizes"][0, 1]), mode="bilinear", align_corners=False)
    train_loss = _seg_loss(masks, gt_masks)
    train_loss += topo(torch.sigmoid(masks.float()), gt_masks.float(), 0.1, feat_d=1, interp=50)
    train_loss.backward()
    optimizer.step()
    return train_loss.item()


def _oracle_topo(pred, true, lamda, feat_d, interp):
    p = F.interpolate(pred, size=(interp, interp), mode="bilinear", align_corners=True)
    t = F.interpolate(true, size=(interp, interp), mode="bilinear", align_corners=True)
    return _oracle_fn(p, t, lamda, feat_d, 2, False, p.shape[0])


def _model_inputs(inputs):
    return {k: inputs[k] for k in ("pixel_values", "input_boxes", "reshaped_input_sizes", "original_sizes")}


class _SamWithSizes(torch.nn.Module):
    """training_step calls model(**inputs): drop the two size entries the HF model does not take."""

    def __init__(self, sam):
        super().__init__()
        self.sam = sam

    def forward(self, pixel_values, input_boxes, reshaped_input_sizes=None, original_sizes=None, multimask_output=False):
        return self.sam(pixel_values=pixel_values, input_boxes=input_boxes, multimask_output=multimask_output)


@pytest.mark.timeout(300)
def test_training_step_matches_reference_loop_body_on_cpu():
    from dilabhelmholtzoct_b200.parallel import training_step
    inputs, gt = _batch()
    ref = _tiny_sam()
    ours = copy.deepcopy(ref)
    opt_ref = torch.optim.Adam(ref.mask_decoder.parameters(), lr=1e-3)
    opt_ours = torch.optim.Adam(ours.mask_decoder.parameters(), lr=1e-3)
    want = _reference_step(ref, inputs, gt, opt_ref, _oracle_topo)
    got = training_step(_SamWithSizes(ours), _model_inputs(inputs), gt, opt_ours, _seg_loss,
                        decoder_params=ours.mask_decoder.parameters(), loss_fn=_oracle_fn)
    assert abs(float(got) - want) <= 1e-5 * abs(want)
    for (n1, p1), (n2, p2) in zip(ref.named_parameters(), ours.named_parameters()):
        assert torch.allclose(p1, p2, rtol=1e-5, atol=1e-7), n1
    frozen = [p for n, p in ours.named_parameters() if n.startswith("vision_encoder")]
    assert all(p.grad is None for p in frozen)


@pytest.mark.gpu
@pytest.mark.timeout(300)
def test_training_step_on_gpu_uses_the_cuda_op_and_matches_the_oracle_on_the_same_maps():
    """The decoder runs in different arithmetic on CPU and GPU and the loss is discontinuous in its
    input, so the CUDA op is checked on exactly the maps the GPU step fed it."""
    from dilabhelmholtzoct_b200.parallel import training_step
    from dilabhelmholtzoct_b200.topological_loss import _TopoLossFn
    inputs, gt = _batch()
    gpu = _tiny_sam().cuda()
    before = [p.detach().clone() for p in gpu.mask_decoder.parameters()]
    opt = torch.optim.Adam(gpu.mask_decoder.parameters(), lr=1e-3)
    seen = {}

    def recording_op(pred, truth, lamda, feat_d, q, loss_r, gb):
        pred.retain_grad()
        out = _TopoLossFn.apply(pred, truth, lamda, feat_d, q, loss_r, gb)
        seen.update(pred=pred, truth=truth, loss=out, args=(lamda, feat_d, q, loss_r))
        return out

    ginputs = {k: v.cuda() for k, v in _model_inputs(inputs).items()}
    total = training_step(_SamWithSizes(gpu), ginputs, gt.cuda(), opt, _seg_loss,
                          decoder_params=gpu.mask_decoder.parameters(), loss_fn=recording_op)
    assert torch.isfinite(total)
    lamda, feat_d, q, loss_r = seen["args"]
    want, wgrad, _ = oracle.topo_loss(seen["pred"].detach().cpu().numpy(), seen["truth"].cpu().numpy(), lamda,
                                      feat_d=feat_d, loss_q=q, loss_r=loss_r)
    assert abs(float(seen["loss"]) - want) <= 1e-5 * abs(want)
    g = seen["pred"].grad.cpu().numpy()
    assert np.abs(g - wgrad).max() <= 1e-5 * np.abs(wgrad).max()
    after = list(gpu.mask_decoder.parameters())
    assert any(not torch.equal(a, b) for a, b in zip(after, before))  # the decoder was updated
    assert all(p.grad is None for n, p in gpu.named_parameters() if n.startswith("vision_encoder"))


def _torch_ops_step(model, inputs, gt, optimizer):
    """training_utils.py:55-68 on the GPU with PyTorch ops around the CUDA topological loss."""
    import dilabhelmholtzoct_b200 as tlb
    from oracle.dice_ce_oracle import dice_ce
    optimizer.zero_grad()
    outputs = model(pixel_values=inputs["pixel_values"], input_boxes=inputs["input_boxes"], multimask_output=False)
    masks = F.interpolate(outputs.pred_masks.squeeze(2), (1024, 1024), mode="bilinear", align_corners=False)
    masks = masks[..., : inputs["reshaped_input_sizes"][0, 0], : inputs["reshaped_input_sizes"][0, 1]]
    masks = F.interpolate(masks, (int(inputs["original_sizes"][0, 0]), int(inputs["original_sizes"][0, 1])), mode="bilinear", align_corners=False)
    loss = dice_ce(masks, gt) + tlb.topo_loss(torch.sigmoid(masks.float()), gt.float(), 0.1, feat_d=1, interp=50)
    loss.backward()
    optimizer.step()
    return loss.detach()


@pytest.mark.gpu
@pytest.mark.timeout(300)
def test_fused_training_step_equals_the_pytorch_ops_step_on_gpu():
    """parallel.training_step with this package's kernels for every line around the model (postprocess_masks,
    dice_ce_loss, fused sigmoid + down-sample + topological loss) against the same step written with PyTorch ops."""
    from dilabhelmholtzoct_b200.parallel import training_step
    inputs, gt = _batch(size=96)
    a = _tiny_sam().cuda()
    b = copy.deepcopy(a)
    before = [p.detach().clone() for p in a.mask_decoder.parameters()]
    ginputs = {k: v.cuda() for k, v in _model_inputs(inputs).items()}
    stats = {}
    got = training_step(_SamWithSizes(a), ginputs, gt.cuda(), torch.optim.SGD(a.mask_decoder.parameters(), lr=0.1),
                        decoder_params=a.mask_decoder.parameters(), stats=stats)
    want = _torch_ops_step(b, ginputs, gt.cuda(), torch.optim.SGD(b.mask_decoder.parameters(), lr=0.1))
    assert abs(float(got) - float(want)) <= 1e-4 * abs(float(want)), (float(got), float(want))
    num = den = 0.0
    for p0, pa, pb in zip(before, a.mask_decoder.parameters(), b.mask_decoder.parameters()):
        num += float(((pa - p0) - (pb - p0)).double().pow(2).sum()); den += float((pb - p0).double().pow(2).sum())
    assert den > 0 and (num / den) ** 0.5 <= 2e-3, (num, den)   # the SGD updates (= gradients) agree
    assert stats["grad_allreduce_bytes"] == 0                    # single process: nothing to reduce
