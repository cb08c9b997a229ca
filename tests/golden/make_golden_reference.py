#!/usr/bin/env python
"""Pin the oracle against the REAL reference the moment its dependencies are importable.

The reference's arithmetic for this path lives in torch_topological -> gudhi / POT
(/root/reference/octsam/models/topological_loss.py:4-9).  None of them is installed in the build
container, so parity is UNPINNED today.  On any box that has them

    pip install torch-topological gudhi POT
    python tests/golden/make_golden_reference.py [/path/to/DILabHelmholtzOCT]

imports the unmodified ``octsam/models/topological_loss.py``, runs it (and the CubicalComplex layer it
uses) on the known-answer images, a seeded tie-heavy set and a few small loss cases, and writes
``tests/golden/reference_vectors.json``.  ``tests/test_reference_golden.py`` then checks the oracle (CPU)
and the CUDA path (GPU) against that file; without the file those tests are skipped with the reason
"reference deps absent".  Nothing here runs on the GPU box or in the product path.
"""
from __future__ import annotations

import importlib.util
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
OUT = os.path.join(HERE, "reference_vectors.json")


def images():
    """Known-answer images + seeded random / tie-heavy / two-valued maps (all small, square and not)."""
    from tests.kats import KATS, TIE_KATS
    out = [(name, np.array(v[0], dtype=np.float32)) for name, v in sorted(KATS.items())]
    out += [(name, np.array(v["image"], dtype=np.float32)) for name, v in sorted(TIE_KATS.items())]
    rng = np.random.default_rng(20261018)
    for levels in (2, 3, 4, 8, 32, 1024, 0):
        for size in (5, 8, 12, 24):
            f = rng.random((size, size)) if levels == 0 else rng.integers(0, levels, (size, size)) / levels
            out.append((f"rand_l{levels}_s{size}", f.astype(np.float32)))
    for shape in ((4, 9), (9, 4), (6, 16)):  # non-square: read by gudhi as W rows of H pixels (shape passed un-reversed)
        out.append((f"rect_{shape[0]}x{shape[1]}", rng.random(shape).astype(np.float32)))
    return out


def loss_cases():
    rng = np.random.default_rng(7)
    cases = []
    for (B, C, S), feat_d, q, lam in (((2, 3, 12), 1, 2, 0.1), ((2, 3, 12), 0, 2, 0.1), ((3, 2, 16), 1, 1, 0.25),
                                      ((1, 4, 10), 1, 2, 0.1), ((2, 1, 10), 1, 2, 0.1)):
        pred = rng.random((B, C, S, S)).astype(np.float32)
        truth = (rng.random((B, C, S, S)) < 0.35).astype(np.float32)
        cases.append(dict(pred=pred, truth=truth, feat_d=feat_d, q=q, lamda=lam, interp=0))
    pred = rng.random((2, 2, 40, 40)).astype(np.float32)
    truth = (rng.random((2, 2, 40, 40)) < 0.4).astype(np.float32)
    cases.append(dict(pred=pred, truth=truth, feat_d=1, q=2, lamda=0.1, interp=16))
    # non-square maps without interp: CubicalComplex passes the shape un-reversed to gudhi
    for (B, C, H, W), feat_d in (((2, 2, 8, 14), 1), ((2, 2, 14, 8), 0)):
        pred = rng.random((B, C, H, W)).astype(np.float32)
        truth = (rng.random((B, C, H, W)) < 0.4).astype(np.float32)
        cases.append(dict(pred=pred, truth=truth, feat_d=feat_d, q=2, lamda=0.1, interp=0))
    return cases


def main():
    ref_root = sys.argv[1] if len(sys.argv) > 1 else "/root/reference"
    path = os.path.join(ref_root, "octsam", "models", "topological_loss.py")
    try:
        import torch
        from torch_topological.nn import CubicalComplex
        spec = importlib.util.spec_from_file_location("reference_topological_loss", path)
        ref = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(ref)
    except Exception as e:  # ModuleNotFoundError in the build container
        print(f"reference deps absent ({type(e).__name__}: {e}); nothing written", file=sys.stderr)
        return 2
    import gudhi
    import ot
    import torch_topological
    doc = {"generator": "tests/golden/make_golden_reference.py",
           "versions": {"torch": torch.__version__, "gudhi": gudhi.__version__, "POT": ot.__version__,
                        "torch_topological": getattr(torch_topological, "__version__", "?")},
           "pairs": [], "losses": []}
    cc = CubicalComplex(dim=2, superlevel=False)  # topological_loss.py:55-58
    for name, f in images():
        if min(f.shape) < 2:
            continue
        info = cc(torch.tensor(f))  # [PersistenceInformation(dim 0), PersistenceInformation(dim 1)]
        W = f.shape[1]
        rec = {"name": name, "image": f.tolist()}
        for pi in info:
            p = np.asarray(pi.pairing).reshape(-1, 4)
            rec[f"h{pi.dimension}"] = [[int(a * W + b), int(c * W + d)] for a, b, c, d in p]
        doc["pairs"].append(rec)
    for case in loss_cases():
        p = torch.tensor(case["pred"], requires_grad=True)
        loss = ref.topo_loss(p, torch.tensor(case["truth"]), case["lamda"], interp=case["interp"],
                             feat_d=case["feat_d"], loss_q=case["q"])
        loss.backward()
        doc["losses"].append({"pred": case["pred"].tolist(), "truth": case["truth"].tolist(), "feat_d": case["feat_d"],
                              "q": case["q"], "lamda": case["lamda"], "interp": case["interp"],
                              "loss": float(loss), "grad": p.grad.numpy().tolist()})
    with open(OUT, "w") as fh:
        json.dump(doc, fh)
    print(f"wrote {OUT}: {len(doc['pairs'])} images, {len(doc['losses'])} loss cases, versions {doc['versions']}")
    return 0


if __name__ == "__main__":
    sys.exit(main())
