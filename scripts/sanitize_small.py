"""Small-case driver for compute-sanitizer (memcheck / racecheck): every kernel once, both dims,
shared-memory and global persistence kernels, ties, matching with truth points, backward."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import dilabhelmholtzoct_b200 as tlb
import oracle
rng = np.random.default_rng(0)
for size in (12, 40):
    maps = np.concatenate([rng.random((3, size, size)), np.round(rng.random((3, size, size)) * 3) / 3]).astype(np.float32)
    for dim in (0, 1):
        got = tlb.persistence_pairs(torch.tensor(maps, device="cuda"), dim)
        for k in range(len(maps)):
            assert np.array_equal(got[k].cpu().numpy(), oracle.cubical_pairs(maps[k], dim)), (size, dim, k)
big = rng.random((1, 270, 270)).astype(np.float32)  # global-memory kernel
for dim in (0, 1):
    got = tlb.persistence_pairs(torch.tensor(big, device="cuda"), dim)
    assert np.array_equal(got[0].cpu().numpy(), oracle.cubical_pairs(big[0], dim))
pred = torch.tensor(rng.random((2, 3, 24, 24)).astype(np.float32))
truth = torch.tensor((np.round(rng.random((2, 3, 24, 24)) * 2) / 2).astype(np.float32))
for dim in (0, 1):
    p = pred.cuda().requires_grad_(True)
    loss = tlb.topo_loss(p, truth.cuda(), 0.1, feat_d=dim, loss_r=True)
    loss.backward()
    want, wgrad, _ = oracle.topo_loss(pred.numpy(), truth.numpy(), 0.1, feat_d=dim, loss_r=True)
    assert abs(float(loss) - want) <= 1e-5 * abs(want)
    assert np.abs(p.grad.cpu().numpy() - wgrad).max() <= 1e-5 * np.abs(wgrad).max()
torch.cuda.synchronize()
print("sanitize driver ok")
