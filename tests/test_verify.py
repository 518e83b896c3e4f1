"""primitive3d_b200.verify: checksums that ignore the vertex numbering but not the geometry or the face order
(CPU tests; the multi-GPU use is in bench.py and tools/check_sharded_nccl.py)."""
import json
import os
import socket
import subprocess
import sys

import numpy as np
import torch

from oracle import inputs, mc
from primitive3d_b200.verify import mesh_checksums

HERE = os.path.dirname(os.path.abspath(__file__))


def _mesh():
    g = inputs.waves((14, 12, 40))
    v, f = mc.marching_cubes(g, 0.0)
    return v, f.astype(np.int32)


def test_checksums_ignore_numbering_but_not_geometry():
    v, f = _mesh()
    base = mesh_checksums(torch.from_numpy(v), torch.from_numpy(f), chunk=777)
    perm = np.random.default_rng(0).permutation(len(v))
    inv = np.empty_like(perm)
    inv[perm] = np.arange(len(v))
    same = mesh_checksums(torch.from_numpy(v[perm]), torch.from_numpy(inv[f].astype(np.int32)))
    assert same == base                                   # renumbered vertices: same mesh
    f2 = f.copy()
    f2[[3, 4]] = f2[[4, 3]]
    assert mesh_checksums(torch.from_numpy(v), torch.from_numpy(f2))[1] != base[1]   # two faces swapped
    f3 = f.copy()
    f3[5] = f3[5][[1, 2, 0]]
    assert mesh_checksums(torch.from_numpy(v), torch.from_numpy(f3))[1] != base[1]   # a face rotated
    v2 = v.copy()
    v2[7, 1] = np.nextafter(v2[7, 1], np.float32(9))
    moved = mesh_checksums(torch.from_numpy(v2), torch.from_numpy(f))
    assert moved[0] != base[0] and moved[1] != base[1]    # one coordinate moved by an ulp


def test_two_virtual_shards_over_gloo(tmp_path):
    """The oracle mesh cut at plane X into two shards numbered shard by shard (as the multi-GPU driver does): the
    checksums taken shard-wise over gloo equal those of the whole mesh."""
    v, f = _mesh()
    base = mesh_checksums(torch.from_numpy(v), torch.from_numpy(f))
    X = 6
    # owner plane of a vertex: floor(x), except that a vertex exactly on a plane belongs to that plane
    lo = np.nonzero(v[:, 0] < X)[0]
    hi = np.nonzero(v[:, 0] >= X)[0]
    newid = np.empty(len(v), np.int64)
    newid[lo] = np.arange(len(lo))
    newid[hi] = len(lo) + np.arange(len(hi))
    tri_x = np.floor(v[f].min(1)[:, 0])                   # the cell a face belongs to (faces are voxel-major)
    cut = int(np.searchsorted(tri_x, X))
    assert (tri_x[:cut] < X).all() and (tri_x[cut:] >= X).all()
    g = newid[f].astype(np.int32)
    path = str(tmp_path / "shards.npz")
    np.savez(path, v0=v[lo], v1=v[hi], f0=g[:cut], f1=g[cut:], voff=[0, len(lo)], foff=[0, cut], plane=[0.0, float(X)])
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    worker = os.path.join(HERE, "_gloo_checksum_worker.py")
    procs = [subprocess.Popen([sys.executable, worker, str(r), "2", str(port), path], stdout=subprocess.PIPE,
                              stderr=subprocess.PIPE, text=True) for r in range(2)]
    for p in procs:
        out, err = p.communicate(timeout=180)
        assert p.returncode == 0, err[-2000:]
        assert tuple(json.loads(out.strip().splitlines()[-1])["sums"]) == base
