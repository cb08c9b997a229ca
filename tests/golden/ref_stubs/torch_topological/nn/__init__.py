"""torch_topological.nn, restated [UPSTREAM-RECALL]: CubicalComplex (gudhi replaced by the repo's oracle) and
WassersteinDistance (verbatim structure; ot.emd2 comes from the stand-in `ot` next to this package)."""
import numpy as np
import ot
import torch

import oracle
from oracle.oracle_literal import gudhi_bitmap_as_image

from .data import PersistenceInformation, batch_iter  # noqa: F401


class CubicalComplex(torch.nn.Module):
    def __init__(self, superlevel=False, dim=None):
        super().__init__()
        self.superlevel = superlevel
        self.dim = dim

    def forward(self, x):
        if self.dim is not None:
            shape = x.shape[:-self.dim]
            dims = len(shape)
        else:
            dims = len(x.shape) - 2
        if dims == 0:
            return self._forward(x)
        elif dims == 1:
            return [self._forward(x_) for x_ in x]
        elif dims == 2:
            return [[self._forward(x__) for x__ in x_] for x_ in x]
        raise RuntimeError("unsupported number of leading dimensions")

    def _forward(self, x):
        if self.superlevel:
            x = -x
        # gudhi.CubicalComplex(dimensions=x.shape, top_dimensional_cells=x.flatten()): the shape goes in UN-REVERSED
        image = gudhi_bitmap_as_image(x.detach().cpu().numpy().ravel(), tuple(x.shape))
        return [self._extract_generators_and_diagrams(x, image, dim) for dim in range(0, len(x.shape))]

    def _extract_generators_and_diagrams(self, x, image, dim):
        # cofaces_of_persistence_pairs(): regular pairs in gudhi's order, then the essential class paired with argmax(x);
        # the oracle returns exactly that list (flat indices of the top-dimensional cells = of x.ravel())
        pairs = torch.as_tensor(oracle.cubical_pairs(image, dim).astype(np.int64), dtype=torch.long).reshape(-1, 2)
        if dim == 0 and len(pairs):
            assert int(pairs[-1, 1]) == int(torch.argmax(x))  # max_index = torch.argmax(x)
        return self._create_tensors_from_pairs(x, pairs, dim)

    def _create_tensors_from_pairs(self, x, pairs, dim):
        xs = x.shape
        creators = torch.as_tensor(np.column_stack(np.unravel_index(pairs[:, 0], xs)), dtype=torch.long)
        destroyers = torch.as_tensor(np.column_stack(np.unravel_index(pairs[:, 1], xs)), dtype=torch.long)
        gens = torch.as_tensor(torch.hstack((creators, destroyers)))
        persistence_diagram = torch.stack((x.ravel()[pairs[:, 0]], x.ravel()[pairs[:, 1]]), 1)
        return PersistenceInformation(pairing=gens, diagram=persistence_diagram, dimension=dim)


class WassersteinDistance(torch.nn.Module):
    def __init__(self, p=torch.inf, q=1):
        super().__init__()
        self.p = p
        self.q = q

    def _project_to_diagonal(self, diagram):
        x = diagram[:, 0]
        y = diagram[:, 1]
        return 0.5 * torch.stack(((x + y), (x + y)), 1)

    def _distance_to_diagonal(self, diagram):
        return torch.linalg.vector_norm(diagram - self._project_to_diagonal(diagram), self.p, dim=1)

    def _make_distance_matrix(self, D1, D2):
        dist_D11 = self._distance_to_diagonal(D1)
        dist_D22 = self._distance_to_diagonal(D2)
        dist = torch.cdist(D1, D2, p=torch.inf)
        upper_blocks = torch.hstack((dist, dist_D11[:, None]))
        lower_blocks = torch.cat((dist_D22, torch.tensor(0, device=dist.device).unsqueeze(0)))
        M = torch.vstack((upper_blocks, lower_blocks))
        M = M.pow(self.q)
        return M

    def forward(self, X, Y):
        total_cost = 0.0
        for pers_info in zip(X, Y):
            D1 = pers_info[0].diagram
            D2 = pers_info[1].diagram
            n = len(D1)
            m = len(D2)
            dist = self._make_distance_matrix(D1, D2)
            a = torch.ones(n + 1, device=dist.device)
            b = torch.ones(m + 1, device=dist.device)
            a[-1] = m
            b[-1] = n
            total_cost += ot.emd2(a, b, dist)
        return total_cost.pow(1.0 / self.q)
