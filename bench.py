#!/usr/bin/env python3
"""bench.py -- throughput of the marching-cubes hot path on B200 (driver contract).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

One "step" = one full extraction (tile pass: classify + count + look-back + vertices -> {V,F}
readback -> face allocation -> face pass: chunk scan + faces) of a synthetic gyroid SDF
that is already resident in HBM.

  N = 1   gyroid 1024^3 fp32, BASELINE.json configs[2] (the HBM-roofline case)
  N > 1   gyroid 2048^3 fp32 sharded into dim-0 slabs with one halo plane, configs[4]; the total
          work is the same at every N ("scaling": "strong"); each rank builds its slab on its
          own GPU; outputs stay sharded (primitive3d_b200/sharded.py)

metric = Gvoxel/s = Rx*Ry*Rz / t, t from CUDA events on the launching stream, max over ranks.
The line also carries the per-kernel roofline numbers (kernels timed one by one with CUDA events
in a separate loop of the same run), the whole-path achieved GB/s, a CPU baseline timed on this
box's host cores on a bounded sample, the end-to-end number with HOST buffers (N = 1: the C-ABI call
p3d_mc_extract_host, pinned grid in, pinned mesh out, slabs pipelined so that upload, extraction and
download overlap; the reference's call shape, prim3d.marching_cubes + the caller's copy back, is
timed beside it), and the SM clocks sampled while the GPU was busy.

--impl reference times the reference's CPU path for the same workload: there is no GPU in that
arm; it runs the OpenMP port of the reference algorithm (oracle/mc_oracle.c) on all host threads
on a bounded sample (PyMCubes itself, the package the reference's cpu=True mode wraps, is not
installable here; its single-threaded stand-in is the `cpu_baseline` of the main arm).
"""
import argparse
import ctypes
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

KNOWN = {1024: (40621056, 81103132), 2048: (162441216, 324595996), 512: (10111488, 20157724),
         256: (2500608, 4972828), 128: (635904, 1261852)}


def ncu_traffic(kernel, n):
    """DRAM bytes per launch of `kernel` from the committed `ncu --set full` capture of this rank's slab (n = its dim-0
    planes, halo included; profiles/ncu_traffic.json, written from the .ncu-rep files); None if there is none."""
    try:
        for rec in json.load(open(os.path.join(ROOT, "profiles", "ncu_traffic.json")))[kernel]:
            if rec.get("planes") == n:
                return rec["dram_bytes_read"] + rec["dram_bytes_write"]
    except Exception:
        pass
    return None


def measured_peak_gbs():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        try:
            return float(json.load(open(path))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks/throttle reasons while the GPU is busy (B200_PROFILING.md recipe)."""
    Q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
        "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index):
        self.rows, self.proc, self.index = [], None, index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "50"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append((time.time(), [c.strip() for c in line.split(",")]))

    def stop(self, t0=None, t1=None):
        if self.proc is None:
            return None
        time.sleep(0.12)
        self.proc.terminate()
        rows = [r for t, r in self.rows if len(r) >= 7 and (t0 is None or t0 - 0.05 <= t <= t1 + 0.05)]
        if not rows:
            rows = [r for _, r in self.rows if len(r) >= 7]
        if not rows:
            return None
        sm = [float(r[0]) for r in rows if r[0].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({n for r in rows for n, v in zip(names, r[3:7]) if v.lower().startswith("active")})
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": float(rows[0][1]),
                "power_w_max": max(float(r[2]) for r in rows), "samples": len(rows), "reasons": reasons}


def gyroid_cuda(n, x0, x1, device, periods=8):
    """Gyroid(N, P) planes [x0, x1) built on the device from the two fp32 tables; same separately
    rounded fp32 ops as workloads.gyroid (bit-identical, tests/test_mc_cuda.py)."""
    import torch
    from primitive3d_b200 import workloads as inputs  # table generator only (numpy sin/cos)
    s, c = (torch.from_numpy(t).to(device) for t in inputs.gyroid_tables(n, periods))
    g = s[x0:x1, None, None] * c[None, :, None]
    g = g + s[None, :, None] * c[None, None, :]
    g = g + s[None, None, :] * c[x0:x1, None, None]
    return g.contiguous()


def cpu_baseline_pymcubes(n, budget_planes):
    """The reference's CPU path is mcubes.marching_cubes (prim3d/utility/marching_cubes.py:66-81),
    single-threaded.  Timed on a bounded sample: the first `budget_planes` planes of the workload."""
    import numpy as np
    sys.path.insert(0, os.path.join(ROOT, "oracle", "pymcubes_compat"))
    import mcubes
    from oracle import inputs
    g = inputs.gyroid(n, x0=0, x1=budget_planes)
    t = time.perf_counter()
    v, f = mcubes.marching_cubes(g, 0.0)
    best = time.perf_counter() - t
    return {"value": g.size / best / 1e9, "unit": "Gvoxel/s", "cores": 1, "kind": "port",
            "sample": f"planes [0,{budget_planes}) of gyroid {n}^3 ({g.size / 1e6:.0f} Mvoxel), one run, "
                      f"PyMCubes-compatible restatement (oracle/pymcubes_compat.c), float32->float64 conversion included",
            "host_cpus": os.cpu_count(), "seconds": best, "V": int(v.shape[0]), "F": int(f.shape[0])}


def run_reference_arm(args):
    """CPU arm: OpenMP port of the reference algorithm on all host threads, bounded sample."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import numpy as np
    from oracle import inputs, mc
    n = 1024 if args.gpus == 1 else 2048
    planes = args.sample_planes or (256 if n == 1024 else 64)
    g = inputs.gyroid(n, x0=0, x1=planes)
    threads = os.cpu_count() or 1
    times = []
    for i in range(args.warmup + args.steps):
        t = time.perf_counter()
        v, f = mc.marching_cubes(g, 0.0, [0, 0, 0], [float(n)] * 3, threads=threads)
        dt = time.perf_counter() - t
        if i >= args.warmup:
            times.append(dt)
    ms = 1e3 * sum(times) / len(times)
    value = g.size / (ms * 1e-3) / 1e9
    sample = f"planes [0,{planes}) of gyroid {n}^3 ({g.size / 1e6:.0f} Mvoxel) per step"
    fraction = planes / n   # of the grid named in `config`: the line is a RATE measured on that part
    line = {"impl": "reference", "metric": "marching_cubes_throughput", "value": value, "unit": "Gvoxel/s",
            "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms,
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": workload_config(n, args.gpus), "sample_fraction": fraction,
            "cpu_baseline": {"value": value, "unit": "Gvoxel/s", "cores": threads, "kind": "port", "sample": sample,
                             "sample_fraction": fraction,
                             "what": "oracle/mc_oracle.c: OpenMP restatement of marching_cubes.cu (the reference has "
                                     "no CPU implementation of its own; its cpu=True mode calls PyMCubes)"},
            "e2e": {"value": value, "unit": "Gvoxel/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line))


def workload_config(n, gpus, counts=None):
    V, F = counts if counts else KNOWN.get(n, (0, 0))
    return {"workload": f"gyroid SDF {n}^3 fp32 marching cubes, thresh 0 (SURVEY.md Appendix B, P=8)",
            "grid": [n, n, n], "V": V, "F": F, "algorithmic_bytes": 4 * n ** 3 + 12 * V + 12 * F,
            "sharding": "single GPU" if gpus == 1 else f"dim-0 slabs x{gpus}, 1 halo plane, outputs sharded",
            "l2": "input (4.3 GB / 34 GB) far larger than the 126 MB L2: no flush needed between steps"}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--size", type=int, default=0, help="override the cubic grid size (default 1024 / 2048)")
    ap.add_argument("--sample-planes", type=int, default=0)
    ap.add_argument("--no-extras", action="store_true", help="skip cpu_baseline / e2e / reference-CUDA legs")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)

    if args.impl == "reference":
        run_reference_arm(args)
        return

    import numpy as np
    import torch
    import torch.distributed as dist
    from primitive3d_b200 import capi, sharded

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (there is no CPU fallback)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    n = args.size or (1024 if args.gpus == 1 else 2048)

    sampler = ClockSampler(local) if rank == 0 else None
    if sampler:
        sampler.start()

    x0, x1h = sharded.slab_with_halo(n, world, rank)
    slab = gyroid_cuda(n, x0, x1h, dev)
    torch.cuda.synchronize()

    # our kernels per step: k_tile, k_round_sums, k_faces_rows (+ k_export_p2p / k_export_exchange, k_wait_p2p,
    # k_apply_exchange on several GPUs)
    launches_per_step = 3 if world == 1 else 5
    exchange = "nccl all-gather"

    if world == 1:
        # one GPU: the reference-facing module itself, prim3d.libPrim3D.marching_cubes (pybind -> p3d_mc_extract)
        import prim3d
        box_lo, box_hi = [0.0, 0.0, 0.0], [float(n)] * 3

        def step():
            v, f = prim3d._C.marching_cubes(slab, 0.0, box_lo, box_hi)
            return sharded.SlabMesh(v, f, 0, 0, v.shape[0], f.shape[0])
    else:
        # several GPUs: ONE C call per step, p3d_mc_sharded_extract over a raw NCCL communicator (tile pass, exchange
        # payload, ncclAllGather, face pass, one host wait); P3D_BENCH_DRIVER=python times the torch.distributed driver
        if os.environ.get("P3D_BENCH_DRIVER", "c") == "python":
            def step():
                return sharded.marching_cubes_slab(slab, 0.0, x0, n)
        elif os.environ.get("P3D_BENCH_EXCHANGE", "nccl") != "p2p":
            comm = sharded.nccl_comm_init()

            def step():
                return sharded.marching_cubes_slab_c(slab, 0.0, x0, n, comm, rank, world)
        else:
            # P3D_BENCH_EXCHANGE=p2p: the same one C call with the shard-boundary exchange over peer memory (NVLink stores
            # into the neighbours' mailboxes + flags, p3d_mc_sharded_extract_p2p): no collective call inside the timed
            # region.  Measured 1.62 against 1.73 ms per step on 8 GPUs (profiles/r2h_bench_gyroid2048_n8_p2p.json); not
            # the default because the teardown of the IPC mappings across 8 exiting processes could not be re-checked
            # within the round's GPU budget.
            exchange = "p2p"
            launches_per_step = 6
            peer = sharded.PeerExchange(slab.shape[1], slab.shape[2])

            def step():
                return sharded.marching_cubes_slab_p2p(slab, 0.0, x0, n, peer)

    for _ in range(args.warmup):
        out = step()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t_wall0 = time.time()
    e0.record()
    for _ in range(args.steps):
        out = step()
    e1.record()
    torch.cuda.synchronize()
    t_wall1 = time.time()
    elapsed = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.barrier()
        dist.all_reduce(elapsed, op=dist.ReduceOp.MAX)
    ms = float(elapsed.item()) / args.steps
    V_tot, F_tot = out.num_vertices_total, out.num_faces_total
    if n in KNOWN:
        assert (V_tot, F_tot) == KNOWN[n], f"counts {(V_tot, F_tot)} differ from the known answer {KNOWN[n]}"

    # ---- end to end with HOST buffers: pinned host slab -> device -> extraction -> host mesh ----
    # N = 1 goes through the reference-facing API prim3d.marching_cubes; N > 1 through the sharded
    # driver (the reference has no multi-GPU API).  Every rank takes part (collectives inside).
    e2e = None
    if not args.no_extras:
        import prim3d
        host = torch.empty(slab.shape, dtype=torch.float32).pin_memory()
        host.copy_(slab)
        hv = torch.empty((out.vertices.shape[0], 3), dtype=torch.float32).pin_memory()
        hf = torch.empty((out.faces.shape[0], 3), dtype=torch.int32).pin_memory()
        del out
        def timed_e2e(call):
            ts = []
            for i in range(2 + 3):
                torch.cuda.synchronize()
                if world > 1:
                    dist.barrier()
                t = time.perf_counter()
                call()
                torch.cuda.synchronize()
                dt = torch.tensor([time.perf_counter() - t], dtype=torch.float64, device=dev)
                if world > 1:
                    dist.all_reduce(dt, op=dist.ReduceOp.MAX)
                if i >= 2:
                    ts.append(float(dt.item()))
            return statistics.mean(ts)

        def via_reference_api():       # the reference's call shape: .cuda() inside the wrapper, caller copies back
            v, f = prim3d.marching_cubes(host, 0.0)
            hv.copy_(v, non_blocking=True)
            hf.copy_(f, non_blocking=True)

        def via_sharded_driver():      # the shard uploads in plane ranges while the ranges on the device are extracted
            sharded.marching_cubes_slab_host(host, 0.0, x0, n, out_vertices=hv, out_faces=hf)

        def via_host_abi():            # C ABI with host buffers: slabs pipelined, upload / extraction / download overlap
            capi.marching_cubes_host(host, 0.0, vertices_out=hv, faces_out=hf)

        io = torch.tensor([host.numel() * 4, hv.numel() * 4 + hf.numel() * 4], dtype=torch.int64, device=dev)
        if world > 1:
            dist.all_reduce(io)
        if world == 1:
            sec = timed_e2e(via_host_abi)
            sec_ref_api = timed_e2e(via_reference_api)
            api = "p3d_mc_extract_host (C ABI, pinned host grid in, pinned host mesh out, slab-pipelined)"
        else:
            sec, sec_ref_api = timed_e2e(via_sharded_driver), None
            api = ("primitive3d_b200.sharded.marching_cubes_slab_host: pinned host shard in, pinned host mesh out, upload in "
                   "plane ranges overlapped with their tile passes, one all-gather, face passes, D2H")
        e2e = {"value": n ** 3 / sec / 1e9, "unit": "Gvoxel/s", "h2d_bytes_per_step": int(io[0]),
               "d2h_bytes_per_step": int(io[1]), "ms_per_step": sec * 1e3, "api": api}
        if sec_ref_api is not None:
            e2e["via_prim3d_marching_cubes"] = {"value": n ** 3 / sec_ref_api / 1e9, "ms_per_step": sec_ref_api * 1e3,
                                                "api": "prim3d.marching_cubes(pinned host tensor) + the caller's D2H of vertices and "
                                                       "faces (the reference's call shape, marching_cubes.py:86-95: the mesh comes back "
                                                       "on the device, so the 1.46 GB download cannot overlap the 4.3 GB upload; floor = "
                                                       "both transfers back to back)"}
        del host, hv, hf

    # ---- several GPUs: the same workload on ONE GPU in the same job (rank 0): the base of the strong-scaling factor,
    # and numbering-independent checksums of the sharded mesh against those of the single-GPU mesh ----
    verification = strong = None
    if world > 1 and not args.no_extras:
        from primitive3d_b200 import verify
        out = step()
        sums = verify.mesh_checksums(out.vertices, out.faces, out.v_offset, out.f_offset, float(x0))
        del out
        if rank == 0:
            torch.cuda.empty_cache()
            big = torch.empty((n, n, n), dtype=torch.float32, device=dev)
            for xs in range(0, n, 256):
                big[xs:xs + 256] = gyroid_cuda(n, xs, min(xs + 256, n), dev)
            desc1 = capi.McDesc.make(big.shape, 0.0)
            for _ in range(2):
                o1 = capi.mc_extract(desc1, big, V_tot + V_tot // 16, F_tot + F_tot // 16)
            torch.cuda.synchronize()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            for _ in range(5):
                o1 = capi.mc_extract(desc1, big, V_tot + V_tot // 16, F_tot + F_tot // 16)
            b.record()
            torch.cuda.synchronize()
            t1 = a.elapsed_time(b) / 5
            del big
            one = verify.mesh_checksums(o1[0], o1[1], single=True)
            strong = {"t1_ms": t1, "tN_ms": ms, "speedup": t1 / ms, "efficiency": t1 / ms / world,
                      "note": "t1 = the same grid on one GPU (rank 0), same job, p3d_mc_extract"}
            verification = {"vertex_checksum": f"{sums[0]:016x}", "triangle_checksum": f"{sums[1]:016x}",
                            "single_gpu_vertex_checksum": f"{one[0]:016x}", "single_gpu_triangle_checksum": f"{one[1]:016x}",
                            "equals_single_gpu": tuple(sums) == tuple(one) and (o1[2], o1[3]) == (V_tot, F_tot),
                            "what": "primitive3d_b200/verify.py: vertex multiset hash and order-sensitive triangle hash "
                                    "(corner coordinates, global face order), independent of the vertex numbering"}
            assert verification["equals_single_gpu"], verification
            del o1
            torch.cuda.empty_cache()
        dist.barrier()

    if rank != 0:
        if world > 1:
            dist.barrier()
            dist.destroy_process_group()
        return

    peak, peak_src = measured_peak_gbs()
    b_alg = 4 * n ** 3 + 12 * V_tot + 12 * F_tot
    value = n ** 3 / (ms * 1e-3) / 1e9
    line = {"metric": "marching_cubes_throughput", "value": value, "unit": "Gvoxel/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": workload_config(n, world, (V_tot, F_tot)), "gpu_launches": launches_per_step * args.steps,
            "path": {"achieved_gbs": b_alg / (ms * 1e-3) / 1e9, "frac_of_peak": b_alg / (ms * 1e-3) / 1e9 / (peak * world),
                     "algorithmic_bytes": b_alg, "peak_gbs_per_gpu": peak, "peak_source": peak_src}}

    if world > 1:
        line["config"]["exchange"] = ("peer memory: NVLink stores into the neighbours' mailboxes + flags (p3d_mc_sharded_extract_p2p), "
                                      "no collective call in the step" if exchange == "p2p" else
                                      "one ncclAllGather of every rank's first-plane table + counts (p3d_mc_sharded_extract)")

    # ---- per-kernel timing on rank 0's slab (CUDA events on the launching stream) ----
    desc = capi.McDesc.make(slab.shape, 0.0, [0, 0, 0], [float(n)] * 3, owned_x=sharded.slab_range(n, world, 0)[1],
                            x_origin=0, global_rx=n)
    V, F, ws, verts = capi.mc_count(desc, slab, vertex_capacity=sharded.capacity_for(slab.shape))
    assert verts.shape[0] >= V
    faces = torch.empty((F, 3), dtype=torch.int32, device=dev)
    L = capi.lib()
    stream = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)

    def time_kernel(fn, reps=10, before=None):
        ts = []
        for _ in range(reps + 2):
            if before:
                before()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            fn()
            b.record()
            torch.cuda.synchronize()
            ts.append(a.elapsed_time(b))
        return statistics.mean(ts[2:])

    stage = lambda k: capi.check(L.p3d_mc_debug_stage(ctypes.byref(desc), slab.data_ptr(), ws.data_ptr(), k,
                                                      verts.data_ptr(), verts.shape[0], stream))
    emit_faces = lambda: capi.check(L.p3d_mc_faces(ctypes.byref(desc), ws.data_ptr(), faces.data_ptr(), 0, stream))
    nvox = slab.numel()
    k_ms = {"tile_pass": time_kernel(lambda: stage(1), before=lambda: stage(0))}
    stage(0), stage(1)   # a consistent workspace for the face pass
    k_ms["faces"] = time_kernel(emit_faces)
    k_bytes = {"tile_pass": 4 * nvox + 12 * V, "faces": 12 * F}
    kernels = {k: {"ms": k_ms[k], "algorithmic_bytes": k_bytes[k], "achieved_gbs": k_bytes[k] / (k_ms[k] * 1e-3) / 1e9,
                   "share_of_step": k_ms[k] / sum(k_ms.values())} for k in k_ms}
    dom = max(k_ms, key=k_ms.get)
    line["roofline"] = {"bound": "hbm", "kernel": {"tile_pass": "k_tile", "faces": "k_faces_rows"}[dom],
                        "achieved": kernels[dom]["achieved_gbs"], "peak": peak, "unit": "GB/s",
                        "frac": kernels[dom]["achieved_gbs"] / peak,
                        "traffic": ncu_traffic({"tile_pass": "k_tile", "faces": "k_faces_rows"}[dom], slab.shape[0]),
                        "path_frac": line["path"]["frac_of_peak"], "peak_source": peak_src,
                        "algorithmic_bytes_per_launch": k_bytes[dom], "launch_ms": k_ms[dom]}
    line["kernels"] = kernels
    line["host_overhead_ms_per_step"] = ms - sum(k_ms.values()) if world == 1 else None
    del verts, faces, ws

    line["e2e"] = e2e
    if strong:
        line["strong_scaling"], line["verification"] = strong, verification
    if not args.no_extras and world == 1:
        import prim3d
        # ---- the reference's own CUDA kernels on this GPU (largest size its int32 indexing allows here) ----
        ref_so = os.path.join(ROOT, "oracle", "_ref", "libPrim3D_ref.so")
        if os.path.exists(ref_so):
            try:
                import importlib.util
                spec = importlib.util.spec_from_file_location("libPrim3D_ref", ref_so)
                ref = importlib.util.module_from_spec(spec)
                spec.loader.exec_module(ref)
                g512 = gyroid_cuda(512, 0, 512, dev)
                t_ref = time_kernel(lambda: ref.marching_cubes(g512, 0.0, [0, 0, 0], [512.0] * 3), reps=5)
                t_our = time_kernel(lambda: prim3d._C.marching_cubes(g512, 0.0, [0, 0, 0], [512.0] * 3), reps=5)
                line["reference_cuda_gyroid512"] = {"reference_ms": t_ref, "ours_ms": t_our, "speedup": t_ref / t_our,
                                                    "note": "reference marching_cubes.cu compiled unmodified for sm_100a"}
                del g512
                # the reference's own example sizes (BASELINE configs[0], [1]): latency-bound, not roofline cases
                from primitive3d_b200 import workloads as oin   # input generators only
                small = {"sphere128": torch.from_numpy(oin.sphere_int64(128).astype(np.float32)).to(dev),
                         "bunny66": torch.from_numpy(np.load(os.path.join(ROOT, "tests", "golden", "mc_bunny66.npz"))["grid"]).to(dev)}
                small["bunny256"] = torch.from_numpy(oin.upsample_trilinear(small["bunny66"].cpu().numpy(), 256)).to(dev)
                line["reference_cuda_examples"] = {}
                for name, gs in small.items():
                    box = [float(v) for v in gs.shape]
                    t_r = time_kernel(lambda: ref.marching_cubes(gs, 0.0, [0, 0, 0], box), reps=10)
                    t_o = time_kernel(lambda: prim3d._C.marching_cubes(gs, 0.0, [0, 0, 0], box), reps=10)
                    line["reference_cuda_examples"][name] = {"reference_ms": t_r, "ours_ms": t_o, "speedup": t_r / t_o}
                del small
            except Exception as exc:  # the comparison is informative only
                line["reference_cuda_gyroid512"] = {"error": str(exc)[:200]}

        # ---- marching tetrahedra, BASELINE configs[3] (Kuhn tet grid 128^3), beside the reference's torch ops on this GPU ----
        try:
            from primitive3d_b200 import workloads as oin   # input generator only
            pts, tets, sdf = oin.kuhn_tet_grid(128)
            P_, T_, S_ = torch.from_numpy(pts).to(dev), torch.from_numpy(tets).to(dev), torch.from_numpy(sdf).to(dev)
            tv, tf = prim3d.marching_tetrahedras(P_, T_.clone(), S_)
            t_mt = time_kernel(lambda: prim3d.marching_tetrahedras(P_, T_.clone(), S_), reps=5)
            t_clone = time_kernel(lambda: T_.clone(), reps=5)
            nT, nP, nV, nF = T_.shape[0], P_.shape[0], tv.shape[0], tf.shape[0]
            b_mt = 32 * nT + 16 * nP + 12 * nV + 24 * nF
            leg = {"ms": t_mt - t_clone, "T": nT, "P": nP, "V": nV, "F": nF, "algorithmic_bytes": b_mt,
                   "achieved_gbs": b_mt / ((t_mt - t_clone) * 1e-3) / 1e9,
                   "note": "prim3d.marching_tetrahedras on CUDA tensors (tets cloned per call: the call flips them in place; "
                           "clone time subtracted)"}
            ref_py = os.path.join(ROOT, "oracle", "_ref", "ref_marching_tetrahedras.py")
            if os.path.exists(ref_py):
                import importlib.util
                spec = importlib.util.spec_from_file_location("ref_marching_tetrahedras", ref_py)
                rmt = importlib.util.module_from_spec(spec)
                spec.loader.exec_module(rmt)
                t_rmt = time_kernel(lambda: rmt.marching_tetrahedras(P_, T_.clone(), S_), reps=3)
                leg["reference_torch_ms"] = t_rmt - t_clone
                leg["speedup"] = (t_rmt - t_clone) / (t_mt - t_clone)
            line["tets_kuhn128"] = leg
            del P_, T_, S_, tv, tf
        except Exception as exc:
            line["tets_kuhn128"] = {"error": str(exc)[:200]}

        # ---- the multi-GPU workload (gyroid 2048^3) on this single GPU: the base of the strong-scaling claim ----
        try:
            torch.cuda.empty_cache()
            n2 = 2048
            big = torch.empty((n2, n2, n2), dtype=torch.float32, device=dev)
            for xs in range(0, n2, 256):
                big[xs:xs + 256] = gyroid_cuda(n2, xs, xs + 256, dev)
            for _ in range(2):
                o2 = sharded.marching_cubes_slab(big, 0.0, 0, n2)
            torch.cuda.synchronize()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            for _ in range(5):
                o2 = sharded.marching_cubes_slab(big, 0.0, 0, n2)
            b.record()
            torch.cuda.synchronize()
            ms2 = a.elapsed_time(b) / 5
            assert (o2.num_vertices_total, o2.num_faces_total) == KNOWN[n2]
            line["single_gpu_gyroid2048"] = {"ms_per_step": ms2, "value": n2 ** 3 / (ms2 * 1e-3) / 1e9, "unit": "Gvoxel/s",
                                             "note": "configs[4] on one GPU: divide the N-GPU ms_per_step into this for "
                                                     "the strong-scaling factor"}
            del big, o2
            torch.cuda.empty_cache()
        except Exception as exc:
            line["single_gpu_gyroid2048"] = {"error": str(exc)[:200]}

        # ---- CPU baseline on a bounded sample ----
        line["cpu_baseline"] = cpu_baseline_pymcubes(n, args.sample_planes or 256)
    else:
        line["cpu_baseline"] = None

    if sampler:
        line["clocks"] = sampler.stop()
        if line["clocks"]:
            line["clocks"]["window"] = "whole measurement phase of this process (timed region is too short to sample alone)"
    print(json.dumps(line))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
