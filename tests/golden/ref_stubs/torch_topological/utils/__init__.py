"""torch_topological.utils, restated [UPSTREAM-RECALL]: total_persistence."""
import torch


def total_persistence(D, p=2, **kwargs):
    """Sum of |death - birth|^p over the finite points of a diagram."""
    persistence = torch.diff(D)
    persistence = persistence[torch.isfinite(persistence)]
    return persistence.abs().pow(p).sum()
