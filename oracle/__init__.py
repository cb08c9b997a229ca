"""CPU oracle for the topological-loss hot path -- TEST INFRASTRUCTURE ONLY.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import this package; the product package
(``dilabhelmholtzoct_b200``) never does and has no CPU fallback.

PARITY UNPINNED: the reference's arithmetic for this path lives in
``torch_topological`` / ``gudhi`` / ``POT``, which are not vendored, pinned or installed and the
reference has no tests.  The oracle restates their published algorithms; ``oracle_literal``
(boundary-matrix reduction of the literal cell complex) pins ``topo_oracle.c``.

What IS pinned against the reference's own code: everything downstream of the persistence pairs.
``tests/golden/make_golden_orchestration.py`` imports the unmodified
``/root/reference/octsam/models/topological_loss.py`` (stand-ins for its absent imports in
``tests/golden/ref_stubs``: the pairs come from this oracle, ``ot.emd2`` from scipy's LP solver, the rest is
torch autograd) and records loss and gradient; ``tests/test_zz_orchestration_golden.py`` checks
``topo_loss`` below against those vectors.  The pairs themselves (gudhi) remain unpinned.
"""
from __future__ import annotations

import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


def build(force: bool = False) -> str:
    so = os.path.join(_HERE, "libtopo_oracle.so")
    src = os.path.join(_HERE, "topo_oracle.c")
    if force or not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
        subprocess.check_call(["make", "-s", "-C", _HERE, "-B", "libtopo_oracle.so"])
    return so


def lib() -> ctypes.CDLL:
    global _LIB
    if _LIB is None:
        L = ctypes.CDLL(build())
        fp = ctypes.POINTER(ctypes.c_float)
        ip = ctypes.POINTER(ctypes.c_int32)
        L.to_cubical_pairs.argtypes = [fp, ctypes.c_int, ctypes.c_int, ctypes.c_int, ip, ctypes.c_int]
        L.to_cubical_pairs.restype = ctypes.c_int
        L.to_wasserstein.argtypes = [fp, ctypes.c_int, fp, ctypes.c_int, ctypes.c_double, ip]
        L.to_wasserstein.restype = ctypes.c_double
        L.to_topo_loss.argtypes = [fp, fp] + [ctypes.c_int] * 5 + [ctypes.c_double, ctypes.c_double,
                                                                    ctypes.c_int, ctypes.c_int, fp, fp, ip]
        L.to_topo_loss.restype = ctypes.c_int
        L.to_max_threads.restype = ctypes.c_int
        _LIB = L
    return _LIB


def _fptr(a):
    return a.ctypes.data_as(ctypes.POINTER(ctypes.c_float))


def _iptr(a):
    return a.ctypes.data_as(ctypes.POINTER(ctypes.c_int32))


def cubical_pairs(f, dim: int) -> np.ndarray:
    """(creator, destroyer) flat pixel indices of one map, gudhi emission order; for dim 0 the
    essential class (paired with argmax) comes last.  Follows torch_topological
    CubicalComplex._forward as called at /root/reference/octsam/models/topological_loss.py:62."""
    f = np.ascontiguousarray(f, dtype=np.float32)
    H, W = f.shape
    cap = H * W + 1
    out = np.empty((cap, 2), dtype=np.int32)
    n = lib().to_cubical_pairs(_fptr(f), H, W, dim, _iptr(out), cap)
    if n < 0:
        raise RuntimeError("to_cubical_pairs failed")
    return out[:n].copy()


def wasserstein(D1, D2, q: float = 2.0):
    """(cost, match) of torch_topological WassersteinDistance for one channel
    (/root/reference/octsam/models/topological_loss.py:78-82), before the 1/q root."""
    D1 = np.ascontiguousarray(D1, dtype=np.float32).reshape(-1, 2)
    D2 = np.ascontiguousarray(D2, dtype=np.float32).reshape(-1, 2)
    m = np.empty(max(len(D1), 1), dtype=np.int32)
    c = lib().to_wasserstein(_fptr(D1), len(D1), _fptr(D2), len(D2), float(q), _iptr(m))
    return c, m[:len(D1)].copy()


def topo_loss(pred, truth, lamda, feat_d=1, loss_q=2, loss_r=False, nthreads=1, want_grad=True,
              reference_shape_order=True):
    """Forward + backward of /root/reference/octsam/models/topological_loss.py:11-96 (interp=0)
    on [B,C,H,W] fp32 arrays.  Returns (loss, grad_pred or None, pair_counts[B*C,2]).
    ``reference_shape_order=False`` reads non-square maps as the images they look like (H rows of W pixels)
    instead of the way the reference does (see below); it changes nothing for H == W."""
    pred = np.ascontiguousarray(pred, dtype=np.float32)
    truth = np.ascontiguousarray(truth, dtype=np.float32)
    B, C, H, W = pred.shape
    if H != W and reference_shape_order:
        # torch_topological's CubicalComplex hands gudhi ``dimensions=x.shape`` un-reversed although gudhi's first
        # dimension is the fastest-varying one [UPSTREAM-RECALL; SURVEY.md 8a row A3a]: for H != W the reference
        # computes the persistence of the same flat buffer read as W rows of H pixels.  Flat indices (and the
        # gradient layout) are unchanged.  `cubical_pairs` keeps the plain reading (H rows of W pixels).
        H, W = W, H
    loss = np.zeros(1, dtype=np.float32)
    grad = np.empty_like(pred) if want_grad else None
    cnt = np.zeros((B * C, 2), dtype=np.int32)
    rc = lib().to_topo_loss(_fptr(pred), _fptr(truth), B, C, H, W, int(feat_d), float(loss_q), float(lamda),
                            int(bool(loss_r)), int(nthreads), _fptr(loss),
                            _fptr(grad) if want_grad else None, _iptr(cnt))
    if rc != 0:
        raise RuntimeError("to_topo_loss failed")
    return float(loss[0]), grad, cnt


def max_threads() -> int:
    return int(lib().to_max_threads())
