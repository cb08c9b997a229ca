"""Event counters of the triplet-merge phase (debug build with -DTL_STATS) and phase cycle shares."""
import sys, os, ctypes
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from dilabhelmholtzoct_b200.synthetic import make_batch

def load(name):
    L = ctypes.CDLL(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "dilabhelmholtzoct_b200", name))
    vp = ctypes.c_void_p
    L.tl_pairs_workspace_bytes.argtypes = [ctypes.c_int] * 4 + [ctypes.POINTER(ctypes.c_size_t)]
    L.tl_set_option.argtypes = [ctypes.c_int, ctypes.c_int]
    L.tl_persistence_pairs.argtypes = [vp] + [ctypes.c_int] * 4 + [vp, ctypes.c_size_t, vp, ctypes.c_int, vp, vp]
    L.tl_max_pairs.argtypes = [ctypes.c_int] * 3
    L.tl_debug_profile.argtypes = [vp, vp]
    return L

def run(L, maps, dim, stats):
    n, H, W = maps.shape
    nb = ctypes.c_size_t(0)
    L.tl_pairs_workspace_bytes(n, H, W, dim, ctypes.byref(nb))
    ws = torch.empty(nb.value, dtype=torch.uint8, device="cuda")
    cap = L.tl_max_pairs(H, W, dim)
    pairs = torch.empty((n, cap, 2), dtype=torch.int32, device="cuda")
    counts = torch.empty(n, dtype=torch.int32, device="cuda")
    out = (ctypes.c_ulonglong * 8)()
    if stats:
        L.tl_debug_stats.argtypes = [ctypes.c_void_p, ctypes.c_int]
        L.tl_debug_stats(out, 1)
    st = torch.cuda.current_stream().cuda_stream
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
    ev[0].record()
    rc = L.tl_persistence_pairs(maps.data_ptr(), n, H, W, dim, ws.data_ptr(), ws.numel(), pairs.data_ptr(), cap, counts.data_ptr(), st)
    ev[1].record(); torch.cuda.synchronize()
    assert rc == 0
    ms = ev[0].elapsed_time(ev[1])
    prof = (ctypes.c_ulonglong * 8)()
    L.tl_debug_profile(ws.data_ptr(), prof)
    res = {"ms": round(ms, 3), "pairs/map": float(counts.float().mean())}
    tot = sum(prof) or 1
    res["phase%[init,L0,flatten,census,merge,emit,compact,boruvka]"] = [round(100.0 * v / tot, 1) for v in prof][:8]
    res["cycles/map"] = int(tot / n)
    if stats:
        L.tl_debug_stats(out, 1)
        names = ["rep_hops", "merges", "merge_iters", "cas_attempts", "displaced", "cas_fail", "basins", "maps"]
        d = dict(zip(names, list(out)))
        m = max(1, d["maps"])
        res["per_map"] = {k: round(v / m, 1) for k, v in d.items()}
    return res

if __name__ == '__main__':
    pred, truth = make_batch(16, 256, 256, seed=1234, device="cuda")
    P = pred.reshape(-1, 256, 256).contiguous(); T = truth.reshape(-1, 256, 256).contiguous()
    X = torch.rand((224, 256, 256), device="cuda")
    S = torch.nn.functional.avg_pool2d(torch.rand((224, 1, 768, 768), device="cuda"), 3).reshape(224, 256, 256).contiguous()
    for name in os.environ.get("TL_PROBE_LIBS", "libtopoloss_stats.so,libtopoloss.so").split(","):
        L = load(name)
        L.tl_set_option(1, 1)  # TL_OPT_PROFILE
        stats = "stats" in name
        for tag, m in (("pred", P), ("truth", T), ("iid", X), ("smooth3", S)):
            run(L, m, 1, stats)
            print(name, tag, "dim1", run(L, m, 1, stats), flush=True)
        if os.path.exists(os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "dilabhelmholtzoct_b200", name)):
            print(name, "pred dim0", run(L, P[:64], 0, stats), flush=True)
            print(name, "truth dim0", run(L, T[:64], 0, stats), flush=True)
