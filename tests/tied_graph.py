"""A hand-made string graph on which asg_arc_del_trans (asg.c:148-193) depends on the ORDER of two arcs with the same
(source, length) sort key -- the shape behind DESIGN.md "Tie order" (ADVICE round 1).

    v -> w1 (100)   v -> w2 (100)   w1 -> w2 (10)   w2 -> x (500)   v -> x (600)        fuzz 50

Walking v's slab in the order (w1, w2): w1 marks w2 as reduced, w2 is then skipped (mark != 1, asg.c:168) and v -> x survives the
pass over v.  In the order (w2, w1): w2 is explored first, reaches x within the bound and v -> x is reduced as well.  The
complement x' -> v' is reduced either way (through x' -> w2' -> v'), so the one-sided survivor falls to asg_arc_del_asymm
(asg.c:125-145): the graph after the call is the same, the "transitively reduced" counter is not.  The reference's in-place radix
sort leaves such ties in an input-dependent order; ours are stable.  Given the SAME slab order every implementation must agree.
"""
import numpy as np

from miniasm_b200.capi import ARC_DT

READ_LEN = 20000
FUZZ = 50
V, W1, W2, X = 0, 2, 4, 6


def tied_graph(w2_first):
    """(arcs sorted by (source, length) with the v -> w1 / v -> w2 tie in the requested order, seq, idx)."""
    rows = []

    def add(u, v, l):
        rows.append((u << 32 | l, v, READ_LEN - l))
        rows.append(((v ^ 1) << 32 | l, u ^ 1, READ_LEN - l))
    first, second = (W2, W1) if w2_first else (W1, W2)
    add(V, first, 100), add(V, second, 100)
    add(W1, W2, 10), add(W2, X, 500), add(V, X, 600)
    arcs = np.array(rows, dtype=ARC_DT)
    arcs = arcs[np.argsort(arcs["ul"], kind="stable")]          # stable: the tie keeps the order it was added in
    n_seq = 4
    seq = np.full(n_seq, READ_LEN, dtype=np.uint32)
    idx = np.zeros(2 * n_seq, dtype=np.uint64)
    src = (arcs["ul"] >> np.uint64(32)).astype(np.int64)
    for u in range(2 * n_seq):                                  # asg_arc_index, asg.c:68-80: start << 32 | count
        at = np.flatnonzero(src == u)
        if len(at):
            idx[u] = np.uint64(int(at[0]) << 32 | len(at))
    return arcs, seq, idx
