"""The Bourke case table is the topology contract
(/root/reference/src/prim3d/Utility/marching_cubes.h:21-277).  Both packed encodings in
this repo must expand to the reference's exact 4096 bytes."""
import hashlib
import os
import re

import numpy as np
import pytest

from oracle import mc

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
TABLE_SHA256 = "19bf7699e214903d72c94c296546f2e31337d637a1e4b118c3108a0f428e809b"


def _expand_product_table():
    text = open(os.path.join(ROOT, "primitive3d_b200/csrc/mc_case_table.h")).read()
    words = [int(w, 16) for w in re.findall(r"0x([0-9a-f]{16})ull", text)]
    assert len(words) == 256
    out = np.full((256, 16), -1, np.int8)
    for c, w in enumerate(words):
        n = 3 * (w >> 60)
        for i in range(15):
            nib = (w >> (4 * i)) & 0xF
            if i < n:
                out[c, i] = nib
            else:
                assert nib == 0xF
    return out


def test_oracle_table_digest():
    assert hashlib.sha256(mc.triangle_table().tobytes()).hexdigest() == TABLE_SHA256


def test_product_table_digest():
    assert hashlib.sha256(_expand_product_table().tobytes()).hexdigest() == TABLE_SHA256


def test_table_structure():
    t = mc.triangle_table()
    ntri = (t >= 0).sum(1) // 3
    # histogram of triangle counts, SURVEY.md section 8 row a6
    assert np.bincount(ntri, minlength=6).tolist() == [2, 16, 50, 80, 76, 32]
    # every referenced edge is a sign-changing edge of that case
    corners = [(0, 1), (1, 2), (3, 2), (0, 3), (4, 5), (5, 6), (7, 6), (4, 7), (0, 4), (1, 5), (2, 6), (3, 7)]
    for c in range(256):
        for e in t[c][t[c] >= 0]:
            a, b = corners[e]
            assert ((c >> a) & 1) != ((c >> b) & 1)


def test_triangle_count_is_edges_minus_two_per_loop():
    """What k_tile's counting relies on (primitive3d_b200/csrc/mc_kernels.cu, P3D_COUNT_EULER): every case has
    E - 2 L triangles (E crossed edges, L >= 1 loops), and L > 1 only where a face has all four edges crossed or
    two opposite corners are both isolated -- the cells the kernel looks up one by one."""
    t = mc.triangle_table()
    ends = [(0, 1), (1, 2), (2, 3), (3, 0), (4, 5), (5, 6), (6, 7), (7, 4), (0, 4), (1, 5), (2, 6), (3, 7)]
    faces = [(0, 1, 2, 3), (4, 5, 6, 7), (0, 1, 5, 4), (1, 2, 6, 5), (2, 3, 7, 6), (3, 0, 4, 7)]
    opposite = [(0, 6), (1, 7), (2, 4), (3, 5)]
    neighbours = {v: [b if a == v else a for a, b in ends if v in (a, b)] for v in range(8)}
    bit = lambda c, k: (c >> k) & 1
    flagged = 0
    for c in range(256):
        ntri = int((t[c] >= 0).sum()) // 3
        edges = sum(bit(c, a) != bit(c, b) for a, b in ends)
        if edges == 0:
            assert ntri == 0
            continue
        assert (edges - ntri) % 2 == 0 and (edges - ntri) // 2 >= 1
        ambiguous = any(bit(c, f[0]) != bit(c, f[1]) and bit(c, f[1]) != bit(c, f[2]) and bit(c, f[2]) != bit(c, f[3])
                        for f in faces)
        isolated = lambda v: all(bit(c, n) != bit(c, v) for n in neighbours[v])
        diagonal = any(isolated(a) and isolated(b) for a, b in opposite)
        if ntri != edges - 2:
            assert ambiguous or diagonal, f"case {c} has {(edges - ntri) // 2} loops but would not be looked up"
        flagged += ambiguous or diagonal
    assert flagged < 256


@pytest.mark.skipif(not os.path.exists("/root/reference/src/prim3d/Utility/marching_cubes.h"),
                    reason="reference tree not mounted")
def test_table_matches_reference_header():
    text = open("/root/reference/src/prim3d/Utility/marching_cubes.h").read()
    body = text[text.index("triangle_table[256][16]"):]
    rows = [[int(v) for v in r.split(",") if v.strip()] for r in re.findall(r"\{([^{}]*)\}", body)]
    rows = np.array([r for r in rows if len(r) == 16], np.int8)
    assert np.array_equal(rows, mc.triangle_table())
