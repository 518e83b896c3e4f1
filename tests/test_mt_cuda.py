"""GPU parity tests of the marching-tetrahedra path against (a) outputs of the REAL reference
stored in tests/golden (made by tests/golden/make_golden.py), (b) the numpy oracle on larger
seeded inputs, (c) the real reference module run on the same GPU when it has been staged in
oracle/_ref.  Bar: bit-exact verts (fp32), identical int64 faces / tet_idx, identical in-place
orientation fix of `tets`."""
import importlib.util
import os

import numpy as np
import pytest
import torch

from oracle import inputs, mt

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
GOLDEN = ["kat", "fixture", "random0", "random1", "kuhn8", "kuhn8_noise"]


def run_capi(pts, tets, sdf, **kw):
    """p3d_mt_extract (one call) unless staged=True; every comparison against the oracle / goldens below runs on
    the one-call path, and test_one_call_and_staged_paths_agree pins the staged calls to it."""
    from primitive3d_b200 import capi
    t = torch.from_numpy(tets.copy()).cuda()
    v, f, ti, e = capi.marching_tetrahedra(torch.from_numpy(pts).cuda(), t, torch.from_numpy(sdf).cuda(), **kw)
    torch.cuda.synchronize()
    return v.cpu().numpy(), f.cpu().numpy(), ti.cpu().numpy(), e.cpu().numpy(), t.cpu().numpy()


def _same(a, b):
    for x, y in zip(a, b):
        assert x.shape == y.shape and np.array_equal(x.view(np.uint32) if x.dtype == np.float32 else x, y.view(np.uint32) if y.dtype == np.float32 else y)


@pytest.mark.parametrize("name", GOLDEN + ["kuhn48", "noisy24", "noisy40"])
def test_one_call_and_staged_paths_agree(name):
    from primitive3d_b200 import capi
    if name in GOLDEN:
        g = np.load(os.path.join(HERE, "golden", f"mt_{name}.npz"))
        pts, tets, sdf = g["points"], g["tets"], g["sdf"]
    else:
        pts, tets, sdf = inputs.kuhn_tet_grid(int(name[-2:]))
        if name.startswith("noisy"):
            sdf = np.random.default_rng(11).uniform(-1, 1, len(pts)).astype(np.float32)
    a = run_capi(pts, tets, sdf, staged=True)
    b = run_capi(pts, tets, sdf)
    _same(a, b)
    if name != "noisy40":   # 1.8 M crossing edges in 0.4 M tets: more than the guessed bucket layout takes
        assert capi.marching_tetrahedra.last_state == 0
    # capacities that are too small in every combination: the second, exact call completes
    n1n2 = max(len(a[1]), 4)
    for caps in ((1, 64, len(a[0]) + 8, len(a[1]) + 8), (n1n2, len(a[3]) * 8 + 64, 1, len(a[1]) + 8), (n1n2, len(a[3]) * 8 + 64, len(a[0]) + 8, 1),
                 (n1n2, len(a[3]) * 8 + 64, len(a[0]), len(a[1]))):
        _same(a, run_capi(pts, tets, sdf, capacities=caps))


def test_crowded_buckets_fall_back_to_the_staged_sort():
    """Every crossing edge starts at point 0 (a fan of tets around one vertex): one bucket takes all the keys, the
    one-call path reports state 2 and the wrapper completes through the staged calls."""
    from primitive3d_b200 import capi
    rng = np.random.default_rng(3)
    n = 6000
    pts = rng.normal(size=(n, 3)).astype(np.float32)
    pts[0] = 0
    others = rng.integers(1, n, size=(4000, 3))
    others = others[(others[:, 0] != others[:, 1]) & (others[:, 1] != others[:, 2]) & (others[:, 0] != others[:, 2])]
    tets = np.concatenate([np.zeros((len(others), 1), np.int64), others], 1)
    sdf = np.full(n, -1.0, np.float32)
    sdf[0] = 1.0
    v, f, ti, e, tets_after = run_capi(pts, tets, sdf, capacities=(len(tets), 256, 3 * len(tets), 3 * len(tets)))
    assert capi.marching_tetrahedra.last_state == 2
    o_tets = tets.copy()
    ov, of, oti = mt.marching_tetrahedras(pts, o_tets, sdf, True)
    assert np.array_equal(tets_after, o_tets) and np.array_equal(v.view(np.uint32), ov.view(np.uint32))
    assert np.array_equal(f, of) and np.array_equal(ti, oti)


def test_near_degenerate_tets_take_the_float64_orientation_test():
    """Tets squashed to within float32 rounding of a plane: the float32 filter cannot decide, the float64 triple
    product does, and the one-call and staged paths flip the same tets (also those of exactly zero volume)."""
    rng = np.random.default_rng(9)
    n = 4096
    pts = rng.uniform(-1, 1, size=(n, 3)).astype(np.float32)
    pts[:, 2] = (pts[:, 0] * 0.5 + pts[:, 1] * 0.25) + rng.uniform(-1, 1, n).astype(np.float32) * np.float32(3e-7)
    pts[: n // 4, 2] = pts[: n // 4, 0] * np.float32(0.5)     # exactly representable plane for a part of them
    tets = rng.integers(0, n, size=(20000, 4)).astype(np.int64)
    sdf = rng.uniform(-1, 1, n).astype(np.float32)
    a = run_capi(pts, tets, sdf, staged=True)
    b = run_capi(pts, tets, sdf)
    assert 0 < (a[4] != tets).any(1).sum() < len(tets)
    _same(a, b)


@pytest.mark.parametrize("name", GOLDEN)
def test_capi_matches_reference_golden(name):
    g = np.load(os.path.join(HERE, "golden", f"mt_{name}.npz"))
    v, f, ti, e, tets_after = run_capi(g["points"], g["tets"], g["sdf"])
    assert np.array_equal(tets_after, g["tets_after"])
    assert v.shape == g["verts"].shape and f.shape == g["faces"].shape
    assert np.array_equal(v.view(np.uint32), g["verts"].view(np.uint32))
    assert np.array_equal(f, g["faces"]) and np.array_equal(ti, g["tet_idx"])
    assert (e[:, 0] < e[:, 1]).all()


@pytest.mark.parametrize("n", [32, 64, 128])  # 128 = BASELINE configs[3]
def test_kuhn_grid_matches_oracle(n):
    pts, tets, sdf = inputs.kuhn_tet_grid(n)
    v, f, ti, e, tets_after = run_capi(pts, tets, sdf)
    o_tets = tets.copy()
    ov, of, oti = mt.marching_tetrahedras(pts, o_tets, sdf, True)
    if n == 32:
        assert (len(ov), len(of)) == (3314, 6624)
    if n == 128:
        assert (len(ov), len(of)) == (56786, 113568)
    assert np.array_equal(tets_after, o_tets)
    assert np.array_equal(v.view(np.uint32), ov.view(np.uint32))
    assert np.array_equal(f, of) and np.array_equal(ti, oti)


def test_noisy_sdf_many_valid_tets():
    pts, tets, _ = inputs.kuhn_tet_grid(24)
    sdf = np.random.default_rng(5).uniform(-1, 1, len(pts)).astype(np.float32)
    sdf[::17] = 0.0
    v, f, ti, e, tets_after = run_capi(pts, tets, sdf)
    o_tets = tets.copy()
    ov, of, oti = mt.marching_tetrahedras(pts, o_tets, sdf, True)
    assert np.array_equal(v.view(np.uint32), ov.view(np.uint32))
    assert np.array_equal(f, of) and np.array_equal(ti, oti) and np.array_equal(tets_after, o_tets)


def test_noisy_sdf_large_sort():
    """~3 M crossing-edge keys: more than 512 radix-sort tiles, i.e. the pass with the separate scan kernel (the
    smaller cases above take the pass with the scan fused into the scatter kernel)."""
    pts, tets, _ = inputs.kuhn_tet_grid(56)
    sdf = np.random.default_rng(6).uniform(-1, 1, len(pts)).astype(np.float32)
    v, f, ti, e, tets_after = run_capi(pts, tets, sdf)
    o_tets = tets.copy()
    ov, of, oti = mt.marching_tetrahedras(pts, o_tets, sdf, True)
    per_tet = np.unique(oti, return_counts=True)[1]   # 1 or 2 triangles per valid tet: 3 or 4 crossing edges
    assert 3 * int((per_tet == 1).sum()) + 4 * int((per_tet == 2).sum()) > 512 * 4096
    assert np.array_equal(tets_after, o_tets)
    assert np.array_equal(v.view(np.uint32), ov.view(np.uint32))
    assert np.array_equal(f, of) and np.array_equal(ti, oti)


def test_python_entry_point_contract():
    """prim3d.marching_tetrahedras: CUDA and CPU tensors, in-place mutation, return_tet_idx,
    empty result (reference marching_tetrahedras.py:89-94,148,225-235)."""
    import prim3d
    g = np.load(os.path.join(HERE, "golden", "mt_fixture.npz"))
    for dev in ("cuda", "cpu"):
        pts = torch.from_numpy(g["points"]).to(dev)
        sdf = torch.from_numpy(g["sdf"]).to(dev)
        tets = torch.from_numpy(g["tets"].copy()).to(dev)
        v, f, ti = prim3d.marching_tetrahedras(pts, tets, sdf, return_tet_idx=True)
        assert v.device.type == dev and f.device.type == dev and f.dtype == torch.int64
        assert np.array_equal(tets.cpu().numpy(), g["tets_after"])           # caller's tensor mutated
        assert np.array_equal(v.cpu().numpy().view(np.uint32), g["verts"].view(np.uint32))
        assert np.array_equal(f.cpu().numpy(), g["faces"]) and np.array_equal(ti.cpu().numpy(), g["tet_idx"])
        out = prim3d.marching_tetrahedras(pts, tets, sdf)
        assert len(out) == 2
    pts = torch.tensor([[0, 0, 0], [1, 0, 0], [0, 1, 0], [0, 0, 1]], dtype=torch.float32).cuda()
    v, f, ti = prim3d.marching_tetrahedras(pts, torch.tensor([[0, 1, 2, 3]]).cuda(), torch.ones(4).cuda(), True)
    assert v.shape == (0, 3) and f.shape == (0, 3) and ti.shape == (0,)


def test_tet_indices_outside_the_point_set_are_an_error():
    """The reference's torch indexing asserts on such a tet; the classify pass reports it instead of reading out
    of bounds."""
    from primitive3d_b200 import capi
    g = np.load(os.path.join(HERE, "golden", "mt_fixture.npz"))
    pts, sdf = torch.from_numpy(g["points"]).cuda(), torch.from_numpy(g["sdf"]).cuda()
    for bad in (len(g["points"]), -1, 1 << 40):
        tets = torch.from_numpy(g["tets"].copy()).cuda()
        tets[len(tets) // 2, 2] = bad
        with pytest.raises(capi.P3DError) as e:
            capi.marching_tetrahedra(pts, tets, sdf)
        assert e.value.status == capi.P3D_ERR_INVALID and "outside" in str(e.value)
    v = capi.marching_tetrahedra(pts, torch.from_numpy(g["tets"].copy()).cuda(), sdf)[0]  # and the library still works
    assert np.array_equal(v.cpu().numpy().view(np.uint32), g["verts"].view(np.uint32))


def test_wrapper_keeps_the_input_dtype_and_tolerates_unused_verts():
    """The reference is dtype-generic torch code (marching_tetrahedras.py:175-189): float64 in, float64 verts and
    gradients out.  Our kernels compute in float32; values agree to float32 rounding."""
    import prim3d
    g = np.load(os.path.join(HERE, "golden", "mt_kuhn8_noise.npz"))
    pts = torch.from_numpy(g["points"]).double().cuda().requires_grad_(True)
    sdf = torch.from_numpy(g["sdf"]).double().cuda().requires_grad_(True)
    v, f = prim3d.marching_tetrahedras(pts, torch.from_numpy(g["tets"].copy()).cuda(), sdf)
    assert v.dtype == torch.float64 and f.dtype == torch.int64
    assert np.array_equal(v.detach().float().cpu().numpy().view(np.uint32), g["verts"].view(np.uint32))
    v.sum().backward()
    assert pts.grad.dtype == torch.float64 and sdf.grad.dtype == torch.float64 and pts.grad.abs().sum() > 0
    with pytest.raises(TypeError):
        prim3d.marching_tetrahedras(pts.detach().long(), torch.from_numpy(g["tets"].copy()).cuda(), sdf.detach())
    # a loss that uses another output only: backward sees grad_verts = None for verts
    from prim3d.utility.marching_tetrahedras import _MarchingTets
    p32 = torch.from_numpy(g["points"]).cuda().requires_grad_(True)
    s32 = torch.from_numpy(g["sdf"]).cuda().requires_grad_(True)
    assert _MarchingTets.backward(type("C", (), {"saved_tensors": (p32, s32, None)})(), None, None, None, None) == (None, None, None)


def test_gradients_match_torch_autograd_of_the_reference_formula():
    """The reference's verts are differentiable w.r.t. vertices and sdf
    (marching_tetrahedras.py:175-189); compare our backward with autograd on that formula."""
    import prim3d
    g = np.load(os.path.join(HERE, "golden", "mt_kuhn8_noise.npz"))
    pts = torch.from_numpy(g["points"]).cuda().requires_grad_(True)
    sdf = torch.from_numpy(g["sdf"]).cuda().requires_grad_(True)
    tets = torch.from_numpy(g["tets"].copy()).cuda()
    v, f = prim3d.marching_tetrahedras(pts, tets, sdf)
    w = torch.randn_like(v)
    (v * w).sum().backward()
    gp, gs = pts.grad.clone(), sdf.grad.clone()
    # reference formula with autograd, on the same unique crossing edges
    p2 = torch.from_numpy(g["points"]).cuda().requires_grad_(True)
    s2 = torch.from_numpy(g["sdf"]).cuda().requires_grad_(True)
    from primitive3d_b200 import capi
    e = capi.marching_tetrahedra(p2.detach(), tets.clone(), s2.detach())[3]
    ev, es = p2[e], s2[e].clone()
    es = torch.stack([es[:, 0], -es[:, 1]], 1)
    wts = torch.flip(es, [1]) / es.sum(1, keepdim=True)
    v2 = (ev * wts[..., None]).sum(1)
    assert torch.equal(v2.detach(), v.detach())
    (v2 * w).sum().backward()
    assert torch.allclose(gp, p2.grad, rtol=1e-4, atol=1e-5)
    assert torch.allclose(gs, s2.grad, rtol=1e-3, atol=1e-3 * float(s2.grad.abs().max()))


@pytest.mark.parametrize("n", [48, 128])  # 128 = BASELINE configs[3]
def test_against_reference_module_on_gpu(n):
    """The reference's own torch implementation run on the same GPU (staged copy in oracle/_ref)."""
    path = os.path.join(ROOT, "oracle", "_ref", "ref_marching_tetrahedras.py")
    # a missing staged reference is a FAILURE of the GPU suite, not a skip
    assert os.path.exists(path), "oracle/_ref/ref_marching_tetrahedras.py is not staged (python oracle/build_ref.py)"
    spec = importlib.util.spec_from_file_location("ref_mt", path)
    ref = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(ref)
    pts, tets, sdf = inputs.kuhn_tet_grid(n)
    margin = mt.orientation_margin(pts, tets)
    assert margin.min() > 1e-6   # no numerically degenerate tets, torch.det's sign is reliable
    P, S = torch.from_numpy(pts).cuda(), torch.from_numpy(sdf).cuda()
    t_ref = torch.from_numpy(tets.copy()).cuda()
    rv, rf, rti = ref.marching_tetrahedras(P, t_ref, S, True)
    v, f, ti, e, tets_after = run_capi(pts, tets, sdf)
    assert np.array_equal(tets_after, t_ref.cpu().numpy())
    assert np.array_equal(v.view(np.uint32), rv.cpu().numpy().view(np.uint32))
    assert np.array_equal(f, rf.cpu().numpy()) and np.array_equal(ti, rti.cpu().numpy())
