"""Known-answer tests for the cubical persistence path.

The reference has no tests or golden vectors; these were derived by hand (tie-free cases) or from
the literal cell-complex restatement (tie-dependent creator indices) and are the pins of the oracle.
pairs = (creator pixel, destroyer pixel), flat index r*W+c, sublevel filtration, T-construction.
"""
KATS = {
    # name: (image rows, H0 regular pairs, H1 pairs, H0 essential (creator, argmax))
    "ring3x3": ([[1, 2, 3], [8, 9, 4], [7, 6, 5]], [], [(3, 4)], (0, 4)),
    "diag8conn": ([[1, 9], [9, 2]], [], [], (0, 1)),
    "two_diag_holes": ([[0, 0, 0, 0], [0, 5, 0, 0], [0, 0, 6, 0], [0, 0, 0, 0]], [], [(9, 5), (14, 10)], (0, 10)),
    "two_minima": ([[1, 5, 2], [6, 7, 8]], [(2, 1)], [], (0, 5)),
    "hole_touching_border": ([[0, 0, 0], [0, 9, 0], [0, 9, 0]], [], [], (0, 4)),
    "nested": ([[1, 1, 1, 1, 1], [1, 5, 5, 5, 1], [1, 5, 9, 5, 1], [1, 5, 5, 5, 1], [1, 1, 1, 1, 1]],
               [], [(23, 12)], (0, 12)),
    "saddle": ([[1, 2, 3, 4, 5], [16, 30, 17, 31, 6], [15, 20, 18, 21, 7], [14, 13, 19, 9, 8], [10, 11, 12, 22, 23]],
               [(20, 22)], [(5, 8), (12, 6), (12, 17)], (0, 8)),
}

# Images that tell apart the two candidate H0 tie rules (oracle/oracle_literal.py, "KNOWN RISK"): the canonical
# (value, dim, position) pairing this repo implements, and gudhi's union-find short cut for edges as recalled
# [UPSTREAM-RECALL].  Two components with EQUAL minima merge through a higher edge: the diagram is the same,
# the creator pixel is not.  Unverifiable here; tests/golden/make_golden_reference.py records the truth.
TIE_KATS = {
    "two_equal_minima": {
        "image": [[1, 5, 1], [5, 5, 5], [5, 5, 5]],
        "canonical": {"h0": [(2, 1)], "ess": (0, 1)},          # the later-positioned minimum (pixel 2) dies
        "gudhi_union_find": {"h0": [(0, 1)], "ess": (2, 1)},   # the class of the edge's first vertex (pixel 0) dies
        "h1": [],
    },
    "equal_minima_on_the_diagonal": {
        "image": [[1, 5, 5], [5, 5, 5], [5, 5, 1]],
        "canonical": {"h0": [(8, 4)], "ess": (0, 1)},
        "gudhi_union_find": {"h0": [(0, 4)], "ess": (8, 1)},
        "h1": [],
    },
}
