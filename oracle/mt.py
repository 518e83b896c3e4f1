"""numpy restatement of the reference's marching tetrahedra -- TEST INFRASTRUCTURE.

Follows /root/reference/prim3d/utility/marching_tetrahedras.py:89-235 step by step
(the reference is ~25 torch ops; numpy has the same primitives):

  :50-65   orientation test: sign of det([1|p0; 1|p1; 1|p2; 1|p3]); det < 0 -> flip
  :148     tets[flip, :2] = tets[flip][:, [1, 0]]          (IN PLACE, caller's array)
  :151-154 occ = sdf > 0; valid tets have 0 < sum(occ) < 4
  :157-160 6 edges per valid tet (base_tet_edges :33-43), each sorted (lo, hi),
           torch.unique(dim=0) -> lexicographic order of (lo, hi)
  :163-173 crossing edges (exactly one endpoint occupied) numbered 0..V-1 in that order
  :177-189 w = (-s1, s0) / (s0 + (-s1)); vert = p0*w0 + p1*w1 (separately rounded fp32)
  :193-223 table index = sum(occ_i << i); faces = [all 1-triangle tets..., all 2-triangle
           tets...] in valid-tet order, int64
  :225-234 tet_idx likewise

The one deliberate difference: the reference takes the sign from torch.det (a batched
fp32 LU whose rounding is backend dependent: LAPACK on CPU, cuSOLVER/MAGMA on GPU).
The restatement evaluates the same determinant as the triple product
(p1-p0) . ((p2-p0) x (p3-p0)) in float64, which has the same sign whenever the fp32
LU result is not rounding noise.  `orientation_margin()` returns |det| so tests can
exclude numerically degenerate tets from bit-exact comparisons (none exist in the
reference's fixture: min |det| = 2.2e-9 and all 12045 signs agree with torch.det,
checked by tests/golden/make_golden.py in the container that has the reference).

Parity status: pinned -- tests/golden/*.npz hold outputs of the real reference module
(imported from /root/reference by tests/golden/make_golden.py) for the docstring
known-answer case (:119-136), the shipped fixture and seeded random/Kuhn cases.
"""
import numpy as np

# marching_tetrahedras.py:7-46
TRIANGLE_TABLE = np.array([
    [-1, -1, -1, -1, -1, -1], [1, 0, 2, -1, -1, -1], [4, 0, 3, -1, -1, -1], [1, 4, 2, 1, 3, 4],
    [3, 1, 5, -1, -1, -1], [2, 3, 0, 2, 5, 3], [1, 4, 0, 1, 5, 4], [4, 2, 5, -1, -1, -1],
    [4, 5, 2, -1, -1, -1], [4, 1, 0, 4, 5, 1], [3, 2, 0, 3, 5, 2], [1, 3, 5, -1, -1, -1],
    [4, 1, 2, 4, 3, 1], [3, 0, 4, -1, -1, -1], [2, 0, 1, -1, -1, -1], [-1, -1, -1, -1, -1, -1]],
    dtype=np.int64)
NUM_TRIANGLES = np.array([0, 1, 1, 2, 1, 2, 2, 1, 1, 2, 2, 1, 2, 1, 1, 0], dtype=np.int64)
BASE_TET_EDGES = np.array([0, 1, 0, 2, 0, 3, 1, 2, 1, 3, 2, 3], dtype=np.int64)


def orientation_det(vertices, tets):
    p = vertices.astype(np.float64)[tets]  # [T,4,3]
    a, b, c = p[:, 1] - p[:, 0], p[:, 2] - p[:, 0], p[:, 3] - p[:, 0]
    return np.einsum("ij,ij->i", a, np.cross(b, c))


def orientation_margin(vertices, tets):
    return np.abs(orientation_det(vertices, tets))


def marching_tetrahedras(vertices, tets, sdf, return_tet_idx=False):
    """vertices f32 [P,3], tets i64 [T,4] (MUTATED in place), sdf f32 [P]
    -> verts f32 [V,3], faces i64 [F,3] (, tet_idx i64 [F])."""
    vertices = np.asarray(vertices, dtype=np.float32)
    sdf = np.asarray(sdf, dtype=np.float32)
    assert tets.dtype == np.int64

    flip = orientation_det(vertices, tets) < 0
    tets[flip, :2] = tets[flip][:, [1, 0]]

    occ_n = sdf > 0
    occ_fx4 = occ_n[tets.reshape(-1)].reshape(-1, 4)
    occ_sum = occ_fx4.sum(-1)
    valid = (occ_sum > 0) & (occ_sum < 4)

    all_edges = tets[valid][:, BASE_TET_EDGES].reshape(-1, 2)
    all_edges = np.sort(all_edges, axis=1)
    if all_edges.shape[0]:
        unique_edges, inverse = np.unique(all_edges, axis=0, return_inverse=True)
        inverse = inverse.reshape(-1)
    else:
        unique_edges, inverse = np.zeros((0, 2), np.int64), np.zeros((0,), np.int64)

    mask_edges = occ_n[unique_edges].sum(-1) == 1
    mapping = np.full(unique_edges.shape[0], -1, dtype=np.int64)
    mapping[mask_edges] = np.arange(int(mask_edges.sum()), dtype=np.int64)
    edge_idx_map = mapping[inverse].reshape(-1, 6)

    interp_v = unique_edges[mask_edges]  # [V,2]
    p = vertices[interp_v]               # [V,2,3]
    s = sdf[interp_v].copy()             # [V,2]
    s[:, 1] *= np.float32(-1)
    denom = (s[:, 0] + s[:, 1]).astype(np.float32)[:, None]
    w = (s[:, ::-1] / denom).astype(np.float32)  # (-s1, s0) / denom
    verts = ((p[:, 0] * w[:, 0, None]).astype(np.float32) +
             (p[:, 1] * w[:, 1, None]).astype(np.float32)).astype(np.float32)

    table_idx = (occ_fx4[valid] * (1 << np.arange(4, dtype=np.int64))[None, :]).sum(-1)
    ntri = NUM_TRIANGLES[table_idx]
    one, two = ntri == 1, ntri == 2
    f1 = np.take_along_axis(edge_idx_map[one], TRIANGLE_TABLE[table_idx[one]][:, :3], axis=1)
    f2 = np.take_along_axis(edge_idx_map[two], TRIANGLE_TABLE[table_idx[two]][:, :6], axis=1).reshape(-1, 3)
    faces = np.concatenate([f1, f2], axis=0).astype(np.int64).reshape(-1, 3)
    if return_tet_idx:
        tid = np.arange(tets.shape[0], dtype=np.int64)[valid]
        tet_idx = np.concatenate([tid[one], np.repeat(tid[two], 2)], axis=0)
        return verts, faces, tet_idx
    return verts, faces
