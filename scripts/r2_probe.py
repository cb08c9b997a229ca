"""Round-2 probe: phase cycles of the persistence kernel on the headline maps (prediction, iid noise, ground truth with and
without the two-valued path, H0); the library is chosen with TL_PROBE_LIB (A/B of build variants: scripts/gpu_ab.sh)."""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from dilabhelmholtzoct_b200.synthetic import make_batch
from scripts.stats_probe import load, run

if __name__ == "__main__":
    pred, truth = make_batch(16, 256, 256, seed=1234, device="cuda")
    P = pred.reshape(-1, 256, 256).contiguous(); T = truth.reshape(-1, 256, 256).contiguous()
    X = torch.rand((224, 256, 256), device="cuda")
    L = load(os.environ.get("TL_PROBE_LIB", "libtopoloss.so"))
    L.tl_set_option(1, 1)  # TL_OPT_PROFILE
    for tag, m in (("pred", P), ("iid", X)):
        run(L, m, 1, False)
        print(tag, run(L, m, 1, False), flush=True)
    for nb in (1, 0):
        L.tl_set_option(2, nb)  # TL_OPT_NO_BINARY_PATH
        run(L, T, 1, False)
        print("no_binary=%d" % nb, "truth", run(L, T, 1, False), flush=True)
    print("pred dim0", run(L, P[:64], 0, False), flush=True)
