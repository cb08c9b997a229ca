"""Per-SM timeline of the persistence launch's tail (matching + gradient jobs) on the C2 batch: when each SM ran out of
persistence jobs, how long it then spent in matching / gradient jobs, when it left the kernel.  TL_OPT_PROFILE build path."""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from dilabhelmholtzoct_b200 import _lib
from dilabhelmholtzoct_b200.synthetic import make_batch
from dilabhelmholtzoct_b200.topological_loss import _buffers

if __name__ == "__main__":
    L = _lib.lib()
    B, C, S = 64, 14, 256
    pred, truth = make_batch(B, S, S, seed=1234 + 2000, device="cuda")
    state, scratch = _buffers(B, C, S, S, 1, pred.device)
    loss, grad = torch.zeros((), device="cuda"), torch.empty_like(pred)
    st = torch.cuda.current_stream().cuda_stream
    args = (pred.data_ptr(), truth.data_ptr(), B, C, S, S, 1, 2.0, 0.1, 0, 0, state.data_ptr(), state.numel(),
            scratch.data_ptr(), scratch.numel(), loss.data_ptr())
    for fused in (1, 0):
        L.tl_set_option(_lib.OPT_NO_FUSED_GRAD, 1 - fused)
        L.tl_set_option(_lib.OPT_PROFILE, 0)
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
        for _ in range(3):
            L.tl_forward_backward(*args, grad.data_ptr(), st)
        ev[0].record()
        for _ in range(10):
            L.tl_forward_backward(*args, grad.data_ptr(), st)
        ev[1].record(); torch.cuda.synchronize()
        print("fused_grad=%d: %.3f ms per forward_backward" % (fused, ev[0].elapsed_time(ev[1]) / 10))
        L.tl_set_option(_lib.OPT_PROFILE, 1)
        L.tl_forward_backward(*args, grad.data_ptr(), st)
        out = (ctypes.c_ulonglong * (160 * 11))()
        n = L.tl_debug_tail_profile(scratch.data_ptr(), S, S, 1, out, 160)
        a = np.array(list(out), dtype=np.float64).reshape(160, 11)[:n]
        t_end = a[:, 1].max()
        tail_start = (a[:, 0] - t_end) / 1e3  # us before the end of the launch
        order = np.argsort(tail_start)
        print("  SMs %d; tail starts (us before the launch ends): min %.0f  p10 %.0f  median %.0f  p90 %.0f  max %.0f" % (
            n, tail_start.min(), np.percentile(tail_start, 10), np.median(tail_start), np.percentile(tail_start, 90), tail_start.max()))
        print("  per SM: matching jobs %.1f (%.1f us each), gradient jobs %.1f (%.1f us each)" % (
            a[:, 4].mean(), a[:, 2].sum() / max(1, a[:, 4].sum()) / 1e3, a[:, 5].mean(), a[:, 3].sum() / max(1, a[:, 5].sum()) / 1e3))
        busy = (a[:, 2] + a[:, 3]) / 1e3
        span = (a[:, 1] - a[:, 0]) / 1e3
        print("  tail span per SM (us): mean %.0f max %.0f; busy in jobs mean %.0f; waiting mean %.0f" % (span.mean(), span.max(), busy.mean(), (span - busy).mean()))
        ng = max(1.0, a[:, 5].sum())
        print("  gradient job cycles (thread 0): zero %.0f  record wait %.0f  math + atomics %.0f  barrier after scatter %.0f  stream-out %.0f" % (
            a[:, 6].sum() / ng, a[:, 9].sum() / ng, a[:, 7].sum() / ng, a[:, 10].sum() / ng, a[:, 8].sum() / ng))
        print("  exit (us before the end): min %.0f median %.0f" % (((a[:, 1] - t_end) / 1e3).min(), np.median((a[:, 1] - t_end) / 1e3)))
    L.tl_set_option(_lib.OPT_PROFILE, 0)
