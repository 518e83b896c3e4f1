"""`prim3d.marching_tetrahedras` -- same signature and outputs as the reference's
prim3d/utility/marching_tetrahedras.py:89-235, computed by hand-written sm_100a kernels
(primitive3d_b200/csrc/mt_kernels.cu) instead of ~25 torch ops.

Contract kept from the reference:
  * `tets` is MUTATED IN PLACE: columns 0 and 1 of negatively oriented tets are swapped (:148);
  * occupancy is `sdf > 0`; vertex i is the i-th crossing edge in lexicographic (min, max) order
    (the order `torch.unique(dim=0)` defines, :157-173); positions are p0*w0 + p1*w1 with
    w = (-s1, s0) / (s0 - s1) (:177-189);
  * faces are int64, all one-triangle tets first, then all two-triangle tets (:205-223);
  * outputs live on the inputs' device; `verts` is differentiable w.r.t. `vertices` and `sdf`
    (the reference leaves :175-189 outside no_grad).

There is one compute path, the CUDA one.  CPU tensors (examples/sphere_tetrahedra.py:15-16 calls
with CPU tensors first) are staged through the GPU and the results copied back, including the
in-place flip of `tets`; without a CUDA device the call fails.
"""
import torch

import prim3d.libPrim3D as _C


class _MarchingTets(torch.autograd.Function):
    @staticmethod
    def forward(ctx, points, sdf, tets):
        verts, faces, tet_idx, edges = _C.marching_tetrahedras(points, tets, sdf)
        ctx.save_for_backward(points, sdf, edges)
        ctx.mark_non_differentiable(faces, tet_idx, edges)
        return verts, faces, tet_idx, edges

    @staticmethod
    def backward(ctx, grad_verts, _gf, _gt, _ge):
        points, sdf, edges = ctx.saved_tensors
        if grad_verts is None:  # verts took no part in the loss
            return None, None, None
        grad_points, grad_sdf = _C.marching_tetrahedras_backward(points, sdf, edges, grad_verts.to(torch.float32).contiguous())
        return grad_points, grad_sdf, None


def marching_tetrahedras(vertices, tets, sdf, return_tet_idx=False):
    """vertices float [P,3], tets int64 [T,4] (mutated), sdf float [P]
    -> (verts [V,3] in the dtype of `vertices`, faces int64 [F,3][, tet_idx int64 [F]]).

    The kernels compute in float32.  Like the reference (dtype-generic torch ops), float64 / float16 / bfloat16
    inputs get verts and gradients back in their own dtype, through differentiable casts either side of the
    float32 kernels; anything that is not a floating-point tensor is a TypeError."""
    if not torch.cuda.is_available():
        raise RuntimeError("prim3d.marching_tetrahedras needs a CUDA device")
    if not vertices.is_floating_point() or not sdf.is_floating_point():
        raise TypeError("vertices and sdf must be floating-point tensors")
    home = vertices.device
    out_dtype = vertices.dtype
    staged = not vertices.is_cuda
    points_d = vertices.cuda() if staged else vertices
    sdf_d = sdf.to(points_d.device)
    tets_d = tets.to(points_d.device)
    if tets_d.dtype != torch.int64:
        raise TypeError("tets must be int64")
    if not tets_d.is_contiguous():
        tets_d = tets_d.contiguous()
    points_d = points_d.to(torch.float32).contiguous()
    sdf_d = sdf_d.to(torch.float32).contiguous()

    if torch.is_grad_enabled() and (points_d.requires_grad or sdf_d.requires_grad):
        verts, faces, tet_idx, _ = _MarchingTets.apply(points_d, sdf_d, tets_d)
    else:  # nothing to differentiate: skip the autograd bookkeeping (about 15 us of a 0.25 ms call)
        verts, faces, tet_idx, _ = _C.marching_tetrahedras(points_d, tets_d, sdf_d)

    if tets_d.data_ptr() != tets.data_ptr():
        with torch.no_grad():
            tets.copy_(tets_d)  # the caller's tensor sees the orientation fix, like the reference
    if verts.dtype != out_dtype:
        verts = verts.to(out_dtype)
    if staged:
        verts, faces, tet_idx = verts.to(home), faces.to(home), tet_idx.to(home)
    if return_tet_idx:
        return verts, faces, tet_idx
    return verts, faces
