"""torch_topological.nn.data, restated [UPSTREAM-RECALL]: PersistenceInformation, nesting_level, batch_iter."""
import itertools
from collections import namedtuple


class PersistenceInformation(namedtuple("PersistenceInformation", ["pairing", "diagram", "dimension"], defaults=[None])):
    """Persistence information data structure: generators (pairing), diagram, dimension."""
    __slots__ = ()


def nesting_level(x):
    """Maximum number of times one can recurse into `x` while still obtaining lists (a PersistenceInformation is a leaf)."""
    if not isinstance(x, list):
        return 0
    if len(x) == 0:
        return 1
    return max(nesting_level(y) for y in x) + 1


def batch_iter(x, dim=None):
    """Iterate over a batch of sparse inputs: per batch element, the PersistenceInformation objects (of dimension `dim`)."""
    level = nesting_level(x)
    if level <= 2:           # nothing to chain: every entry of x already is a list of PersistenceInformation
        def handler(y):
            return y
    else:                    # batch x channels: chain the channels of one batch element
        handler = itertools.chain.from_iterable
    if dim is not None:
        for first in x:
            yield [*filter(lambda y: y.dimension == dim, handler(first))]
    else:
        for first in x:
            yield handler(first)
