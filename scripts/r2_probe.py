"""Round-2 probe: phase cycles of the persistence kernel on the headline maps for a few runtime variants
(TL_BV_ROUNDS, TL_NO_BINARY are read per call), then stage times of the whole step."""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from dilabhelmholtzoct_b200.synthetic import make_batch
from scripts.stats_probe import load, run

if __name__ == "__main__":
    os.environ["TL_PROFILE"] = "1"
    pred, truth = make_batch(16, 256, 256, seed=1234, device="cuda")
    P = pred.reshape(-1, 256, 256).contiguous(); T = truth.reshape(-1, 256, 256).contiguous()
    X = torch.rand((224, 256, 256), device="cuda")
    L = load(os.environ.get("TL_PROBE_LIB", "libtopoloss.so"))
    for var in sys.argv[1:] or ["TL_X=0"]:
        k, v = var.split("=")
        os.environ[k] = v
        for tag, m in (("pred", P), ("iid", X)):
            run(L, m, 1, False)
            print(var, tag, run(L, m, 1, False), flush=True)
        os.environ.pop(k)
    for var in ("TL_NO_BINARY=1", "TL_NO_BINARY=0"):
        k, v = var.split("=")
        os.environ[k] = v
        run(L, T, 1, False)
        print(var, "truth", run(L, T, 1, False), flush=True)
    print("pred dim0", run(L, P[:64], 0, False), flush=True)
