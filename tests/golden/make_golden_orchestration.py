#!/usr/bin/env python
"""Golden vectors from the UNMODIFIED reference file, run in this container.

``/root/reference/octsam/models/topological_loss.py`` cannot be imported as is: torch_topological / gudhi / POT are absent
(SURVEY.md 8c).  ``tests/golden/ref_stubs`` holds stand-ins for exactly the layers that file imports (README.md there):
gudhi's pairs come from this repo's oracle, ``ot.emd2`` is the same LP solved by scipy's HiGHS, everything else is a
restatement of torch-topological's public classes in plain PyTorch.  With them on ``sys.path`` the reference's own
``topo_loss`` -- its ``.squeeze()`` nesting, ``batch_iter`` filtering, WassersteinDistance per image, the mean, ``lamda``,
the ``loss_r`` regulariser, ``interp`` -- runs here under torch autograd, and this script records its loss and gradient:

    python tests/golden/make_golden_orchestration.py [/path/to/DILabHelmholtzOCT]   ->  tests/golden/orchestration_vectors.json

What the vectors pin: rows A1, A2, A4, A5, A7, A8 and A9 of SURVEY.md 8(a) (orchestration, cost matrix, backward) against
the reference's real code path.  What they do NOT pin: the persistence pairs themselves (row A3a: they come from the
oracle) -- that still needs gudhi, see make_golden_reference.py.  /root/reference is only read here, never at test time.
"""
from __future__ import annotations

import importlib.util
import json
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
OUT = os.path.join(HERE, "orchestration_vectors.json")


def cases():
    rng = np.random.default_rng(20261018)
    out = []

    def add(shape, feat_d, q, lamda=0.1, interp=0, loss_r=False, truth="binary", note=""):
        pred = rng.random(shape).astype(np.float32)
        if truth == "binary":
            t = (rng.random(shape) < 0.4).astype(np.float32)
        elif truth == "blobs":  # a few rectangles with holes: non-empty H1 ground-truth diagrams
            t = np.zeros(shape, np.float32)
            for idx in np.ndindex(*shape[:2]):
                h, w = shape[2:]
                t[idx][1:h - 1, 1:w - 1] = 1.0
                t[idx][h // 2, w // 2] = 0.0
                t[idx][2, 2] = 0.0
        else:                  # fractional values, as after the reference's own down-sampling
            t = (np.round(rng.random(shape) * 4) / 4).astype(np.float32)
        out.append(dict(pred=pred, truth=t, feat_d=feat_d, q=q, lamda=lamda, interp=interp, loss_r=loss_r, note=note))

    for feat_d in (0, 1):
        for q in (1, 2):
            add((2, 3, 10, 10), feat_d, q)
    add((2, 3, 12, 12), 1, 2, truth="blobs", note="ground truth with holes: matched pairs")
    add((2, 2, 10, 10), 1, 2, truth="levels", note="non-binary ground truth")
    add((2, 2, 10, 10), 0, 2, truth="levels")
    add((1, 4, 10, 10), 1, 2, note="B == 1: .squeeze() makes every channel its own image")
    add((1, 3, 10, 10), 0, 1, note="B == 1, H0")
    add((3, 1, 10, 10), 1, 2, note="C == 1: .squeeze() drops the channel axis")
    add((2, 3, 10, 10), 1, 2, loss_r=True, note="total-persistence regulariser")
    add((2, 2, 10, 10), 0, 2, loss_r=True, lamda=0.25)
    add((1, 3, 10, 10), 1, 2, loss_r=True, note="regulariser with B == 1")
    add((2, 2, 24, 24), 1, 2, interp=9, note="interp: both inputs down-sampled (bilinear, align_corners=True)")
    add((2, 2, 20, 28), 1, 2, interp=8, note="interp from a non-square source")
    add((2, 2, 6, 11), 1, 2, note="non-square without interp: dimensions=x.shape goes to gudhi un-reversed")
    add((2, 2, 11, 6), 0, 2, note="non-square, H0")
    add((2, 2, 10, 10), 1, 3, note="q = 3")
    return out


def main():
    ref_root = sys.argv[1] if len(sys.argv) > 1 else "/root/reference"
    path = os.path.join(ref_root, "octsam", "models", "topological_loss.py")
    sys.path.insert(0, os.path.join(HERE, "ref_stubs"))
    sys.path.insert(0, ROOT)
    spec = importlib.util.spec_from_file_location("reference_topological_loss", path)
    ref = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(ref)
    import hashlib
    doc = {"generator": "tests/golden/make_golden_orchestration.py",
           "reference_file": "octsam/models/topological_loss.py",
           "reference_sha256": hashlib.sha256(open(path, "rb").read()).hexdigest(),
           "stand_ins": "tests/golden/ref_stubs (pairs: this repo's oracle; emd2: scipy HiGHS; the rest: torch autograd)",
           "torch": torch.__version__, "cases": [], "errors": []}
    for c in cases():
        p = torch.tensor(c["pred"], requires_grad=True)
        loss = ref.topo_loss(p, torch.tensor(c["truth"]), c["lamda"], interp=c["interp"], feat_d=c["feat_d"],
                             loss_q=c["q"], loss_r=c["loss_r"])
        loss.backward()
        doc["cases"].append({"pred": c["pred"].tolist(), "truth": c["truth"].tolist(), "feat_d": c["feat_d"], "q": c["q"],
                             "lamda": c["lamda"], "interp": c["interp"], "loss_r": c["loss_r"], "note": c["note"],
                             "loss": float(loss), "grad": p.grad.numpy().tolist()})
    # behaviour at the edges of the signature (topological_loss.py:30-31, :68)
    x = torch.rand((2, 2, 6, 6))
    doc["lamda_zero_returns"] = repr(ref.topo_loss(x, x, 0.0))
    for name, args, kw in (("default feat_d=2", (x, x, 0.1), {}),
                           ("feat_d=3", (x, x, 0.1), {"feat_d": 3}),
                           ("B == C == 1", (x[:1, :1], x[:1, :1], 0.1), {"feat_d": 1})):
        try:
            ref.topo_loss(*args, **kw)
            doc["errors"].append({"case": name, "raises": None})
        except Exception as e:  # noqa: BLE001 -- the type of the crash IS the recorded behaviour
            doc["errors"].append({"case": name, "raises": type(e).__name__})
    with open(OUT, "w") as fh:
        json.dump(doc, fh)
    print(f"wrote {OUT}: {len(doc['cases'])} cases; lamda == 0 -> {doc['lamda_zero_returns']}; errors {doc['errors']}")
    return 0


if __name__ == "__main__":
    sys.exit(main())
