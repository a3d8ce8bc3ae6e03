"""How the parity tests compare GFA text.

Default: exact bytes.  `tie_order_free=True` is for inputs on which the reference's own output order is an accident of
its unstable in-place radix sort (ksort.h:134-183): unitig arcs with equal (unitig end, length) keys print their L lines
in whatever order the cycle-leader permutation left them.  Then the comparison is BASELINE.json's "same S/L lines,
modulo order" -- the sorted line multisets must be equal -- plus: every S / a / x line and the relative order of
everything that is not an L line must still match byte for byte, and L lines may only move within their block."""


def assert_same_gfa(got, want, tie_order_free=False):
    if got == want:
        return
    assert tie_order_free, _first_diff(got, want)
    g, w = got.split(b"\n"), want.split(b"\n")
    assert sorted(g) == sorted(w), "GFA differs beyond line order: " + _first_diff(got, want)
    assert [x for x in g if not x.startswith(b"L\t")] == [x for x in w if not x.startswith(b"L\t")], "non-L lines moved"
    # L lines are printed sorted by (from-unitig, orientation, arc length): ties may permute only lines with the same from-end
    key = lambda x: x.split(b"\t")[1:3]
    assert [key(x) for x in g if x.startswith(b"L\t")] == [key(x) for x in w if x.startswith(b"L\t")], "L lines moved across unitig ends"


def _first_diff(a, b):
    la, lb = a.split(b"\n"), b.split(b"\n")
    for i, (x, y) in enumerate(zip(la, lb)):
        if x != y:
            return f"line {i + 1}: ours {x[:200]!r} vs reference {y[:200]!r} ({len(la)} vs {len(lb)} lines)"
    return f"{len(la)} vs {len(lb)} lines, common prefix equal"
