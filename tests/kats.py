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
