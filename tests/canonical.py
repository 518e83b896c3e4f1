"""Order-free canonical forms for comparing meshes (test helper).

The reference's CUDA marching cubes hands out vertex and face slots with atomicAdd
(/root/reference/src/prim3d/Utility/marching_cubes.cu:104,117,130,199), so two runs of
the reference itself give differently permuted outputs.  "Same result" therefore means
(SURVEY.md section 8c):
  1. V and F equal;
  2. the sorted multiset of vertex rows is bit-identical;
  3. the multiset of triangles, each taken as its 9 corner coordinates IN EMITTED
     CORNER ORDER (winding kept, no rotation), is bit-identical.
NaN coordinates (possible when the grid holds NaN) compare equal to each other.
"""
import numpy as np


def _bits(a):
    a = np.ascontiguousarray(a, dtype=np.float32)
    b = a.view(np.uint32).copy()
    b[np.isnan(a)] = 0x7FC00000  # one NaN
    b[b == 0x80000000] = 0       # -0 == +0 never arises (coordinates >= 0) but be safe
    return b


def _sort_rows(rows):
    if rows.shape[0] == 0:
        return rows
    order = np.lexsort(rows.T[::-1])
    return rows[order]


def vertex_multiset(verts):
    return _sort_rows(_bits(verts).reshape(-1, 3))


def triangle_soup(verts, faces):
    """[F, 9] uint32 keys: corner coordinates in emitted corner order."""
    f = np.asarray(faces).astype(np.int64).reshape(-1, 3)
    return _bits(verts).reshape(-1, 3)[f].reshape(-1, 9)


def triangle_multiset(verts, faces):
    return _sort_rows(triangle_soup(verts, faces))


def assert_same_mesh(verts_a, faces_a, verts_b, faces_b, ordered_faces=False):
    assert verts_a.shape == verts_b.shape, (verts_a.shape, verts_b.shape)
    assert faces_a.shape == faces_b.shape, (faces_a.shape, faces_b.shape)
    if faces_a.size:
        assert faces_a.min() >= 0 and faces_a.max() < verts_a.shape[0]
    assert np.array_equal(vertex_multiset(verts_a), vertex_multiset(verts_b)), "vertex multisets differ"
    if ordered_faces:
        # both sides emit faces in voxel-major cell order: compare triangle by triangle
        assert np.array_equal(triangle_soup(verts_a, faces_a), triangle_soup(verts_b, faces_b)), \
            "triangles differ (ordered comparison)"
    else:
        assert np.array_equal(triangle_multiset(verts_a, faces_a), triangle_multiset(verts_b, faces_b)), \
            "triangle multisets differ"
