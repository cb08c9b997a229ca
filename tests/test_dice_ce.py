"""Row F2: the fused DiceCE kernel against the oracle restatement of monai.losses.DiceCELoss(sigmoid=True)
(oracle/dice_ce_oracle.py; monai is absent here, so this row's parity is UNPINNED -- the restatement itself is
pinned against an explicit formula and, through PyTorch, against the reference's arithmetic)."""
import numpy as np
import pytest
import torch

from oracle.dice_ce_oracle import dice_ce, dice_ce_explicit


def _case(B, C, H, W, seed):
    g = torch.Generator().manual_seed(seed)
    x = 3.0 * torch.randn((B, C, H, W), generator=g)
    t = (torch.rand((B, C, H, W), generator=g) < 0.3).float()
    t[:, -1] = 0.0  # padded (all-zero) component masks are part of the reference's batches (training_utils.py:453)
    return x, t


def test_oracle_forms_agree_and_match_hand_values():
    x, t = _case(2, 5, 9, 11, 0)
    a, b = dice_ce(x.double(), t.double()), dice_ce_explicit(x.double(), t.double())
    assert abs(float(a) - float(b)) <= 1e-12 * abs(float(a))
    # one pixel, two channels, logits (0, 0), target (1, 0): dice = mean(1 - (1+1e-5)/(1.5+1e-5), 1 - 1e-5/(0.5+1e-5)), ce = log 2
    x1 = torch.zeros((1, 2, 1, 1), dtype=torch.float64); t1 = torch.tensor([1.0, 0.0], dtype=torch.float64).view(1, 2, 1, 1)
    want = 0.5 * ((1 - (1 + 1e-5) / (1.5 + 1e-5)) + (1 - 1e-5 / (0.5 + 1e-5))) + np.log(2.0)
    assert abs(float(dice_ce(x1, t1)) - want) <= 1e-12


@pytest.mark.gpu
@pytest.mark.parametrize("shape", [(2, 3, 17, 23), (3, 14, 64, 64), (2, 14, 496, 512), (1, 1, 50, 50), (2, 40, 31, 33)])
def test_kernel_matches_oracle(shape):
    import dilabhelmholtzoct_b200 as tlb
    x, t = _case(*shape, seed=sum(shape))
    xg = x.cuda().requires_grad_(True)
    loss = tlb.dice_ce_loss(xg, t.cuda())
    (3.0 * loss).backward()
    xr = x.double().requires_grad_(True)
    want = dice_ce(xr, t.double())
    (3.0 * want).backward()
    assert abs(float(loss) - float(want)) <= 1e-5 * abs(float(want))
    g, w = xg.grad.cpu().double(), xr.grad
    assert float((g - w).abs().max()) <= 1e-5 * float(w.abs().max())
    # float64 / uint8 targets as the reference passes them (gt_masks are float64, training_utils.py:413)
    assert abs(float(tlb.dice_ce_loss(x.cuda(), t.double().cuda())) - float(want)) <= 1e-5 * abs(float(want))


@pytest.mark.gpu
def test_errors():
    import dilabhelmholtzoct_b200 as tlb
    x = torch.randn(2, 3, 8, 8)
    with pytest.raises(ValueError, match="CUDA"):
        tlb.dice_ce_loss(x, x)
    with pytest.raises(ValueError):
        tlb.dice_ce_loss(x.cuda(), x[:, :2].cuda())
    with pytest.raises(ValueError):
        tlb.dice_ce_loss(torch.randn(1, 65, 4, 4).cuda(), torch.randn(1, 65, 4, 4).cuda())  # more than 64 channels
