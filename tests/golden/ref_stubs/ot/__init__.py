"""Stand-in for POT's ``ot.emd2`` (see ../README.md): the exact transport LP  min <G, M>  s.t.  G 1 = a, G^T 1 = b, G >= 0,
solved with scipy's HiGHS instead of POT's network simplex; value in float64 cast back to M's dtype, gradient with respect
to M = the optimal plan (POT: ``nx.set_gradients(cost, (a, b, M), (u - mean(u), v - mean(v), G))``)."""
import numpy as np
import torch
from scipy.optimize import linprog

__version__ = "0.9-standin"


class _Emd2(torch.autograd.Function):
    @staticmethod
    def forward(ctx, a, b, M):
        Mn = M.detach().cpu().numpy().astype(np.float64)
        an, bn = a.detach().cpu().numpy().astype(np.float64), b.detach().cpu().numpy().astype(np.float64)
        R, C = Mn.shape
        A_eq = np.zeros((R + C, R * C))
        for i in range(R):
            A_eq[i, i * C:(i + 1) * C] = 1.0
        for j in range(C):
            A_eq[R + j, j::C] = 1.0
        res = linprog(Mn.ravel(), A_eq=A_eq, b_eq=np.concatenate([an, bn]), bounds=(0, None), method="highs-ds")
        if res.status != 0:
            raise RuntimeError(f"transport LP failed: {res.message}")
        G = res.x.reshape(R, C)
        ctx.G = torch.as_tensor(G, dtype=M.dtype)
        return torch.as_tensor(float((G * Mn).sum()), dtype=M.dtype)

    @staticmethod
    def backward(ctx, g):
        return None, None, ctx.G * g


def emd2(a, b, M, **kwargs):
    return _Emd2.apply(a, b, M)
