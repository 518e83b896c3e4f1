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


@pytest.mark.skipif(not os.path.exists("/root/reference/src/prim3d/Utility/marching_cubes.h"),
                    reason="reference tree not mounted")
def test_table_matches_reference_header():
    text = open("/root/reference/src/prim3d/Utility/marching_cubes.h").read()
    body = text[text.index("triangle_table[256][16]"):]
    rows = [[int(v) for v in r.split(",") if v.strip()] for r in re.findall(r"\{([^{}]*)\}", body)]
    rows = np.array([r for r in rows if len(r) == 16], np.int8)
    assert np.array_equal(rows, mc.triangle_table())
