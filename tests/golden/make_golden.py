"""Generates tests/golden/pairs_small.json with the LITERAL oracle (oracle/oracle_literal.py: boundary-
matrix reduction of the gudhi cell complex) and losses/gradients with the fast oracle.

The real reference cannot be imported in the build container (torch_topological, gudhi and POT are not
installed and not pinned) -- parity is unpinned by the reference itself; these vectors pin
the fast oracle and the CUDA path to the literal restatement.   Run:  python tests/golden/make_golden.py
"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import oracle  # noqa: E402
from oracle.oracle_literal import cubical_pairs_literal  # noqa: E402


def main():
    rng = np.random.default_rng(20261018)
    cases = []
    for t in range(36):
        n = int(rng.integers(2, 13))
        mode = t % 4
        if mode == 0:
            f = rng.random((n, n))
        elif mode == 1:
            f = rng.integers(0, 4, (n, n)) / 4.0
        elif mode == 2:
            f = rng.integers(0, 2, (n, n)).astype(np.float64)
        else:
            f = np.round(rng.normal(size=(n, n)), 1)
        f = f.astype(np.float32)
        h0, h1, ess = cubical_pairs_literal(f)
        cases.append({"image": f.tolist(), "h0": h0 + [list(ess)], "h1": h1})
    losses = []
    for t in range(6):
        B, Cc, n = 2, 3, int(rng.integers(6, 15))
        pred = rng.random((B, Cc, n, n)).astype(np.float32)
        truth = (rng.random((B, Cc, n, n)) < 0.5).astype(np.float32) if t % 2 else np.round(rng.random((B, Cc, n, n)) * 4).astype(np.float32) / 4
        for feat_d in (0, 1):
            loss, grad, _ = oracle.topo_loss(pred, truth, 0.1, feat_d=feat_d, loss_q=2)
            losses.append({"pred": pred.tolist(), "truth": truth.tolist(), "feat_d": feat_d, "lamda": 0.1, "q": 2,
                           "loss": loss, "grad": grad.tolist()})
    out = os.path.join(os.path.dirname(os.path.abspath(__file__)), "pairs_small.json")
    with open(out, "w") as f:
        json.dump({"pairs": cases, "losses": losses}, f)
    print(out, os.path.getsize(out))


if __name__ == "__main__":
    main()
