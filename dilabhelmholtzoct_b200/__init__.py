"""B200-native topological loss: the one hot path of philippendres/DILabHelmholtzOCT's
``training.py --top`` (``octsam/models/topological_loss.py``) rebuilt as sm_100a CUDA kernels
behind the reference's ``topo_loss`` signature."""
from .topological_loss import persistence_pairs, topo_loss, topo_loss_from_host, wasserstein_cost  # noqa: F401

__all__ = ["topo_loss", "topo_loss_from_host", "persistence_pairs", "wasserstein_cost"]
