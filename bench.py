#!/usr/bin/env python
"""bench.py -- headline benchmark of the model3d ray-tracing hot path on B200.

Workload (BASELINE.json configs[1], "C2"): raw ray batch, 2^24 random rays (origin ~N(0,I),
direction uniform on S^2, as BenchmarkMeshFirstRayCollisions, model3d/collisions_test.go:
382-394) against the BVH of a 1,003,520-triangle icosphere (NewMeshIcosphere(0,1,224)),
first-hit only.  Metric: Mrays/s.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]

One process per GPU (torchrun for N>1); the ray batch shards by rank with no data-path
collective (weak scaling: every rank traces its own 2^24 rays).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

ICO_N = 224
N_RAYS = 1 << 24
SEED = 20260
NODE_BYTES, TRI_BYTES, RAY_IO_BYTES = 80, 48, 64


def make_rays(n, seed):
    rng = np.random.default_rng(seed)
    org = rng.standard_normal((n, 3), dtype=np.float32)
    d = rng.standard_normal((n, 3), dtype=np.float32)
    d /= np.linalg.norm(d, axis=1, keepdims=True)
    return org, d


def make_mesh():
    # NewMeshIcosphere(0, 1, 224) restated in numpy on the product side (model3d_b200/meshes.py)
    from model3d_b200 import meshes
    return meshes.NewMeshIcosphere((0, 0, 0), 1.0, ICO_N).astype(np.float32).reshape(-1, 9)


class ClockSampler:
    """Samples nvidia-smi clocks / throttle reasons while the timed region runs."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index = index
        self.lines = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                 "--format=csv,noheader,nounits", "-lms", "100"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                mx.append(float(f[2]))
            except ValueError:
                continue
            for nm, v in zip(names, f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": float(np.median(sm)) if sm else None,
                "sm_max_mhz": float(max(mx)) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons)}


def cpu_reference_rate(tris, n_sample, threads, steps=1, warmup=0):
    """Times the oracle port (restated Go reference, float64, unculled binary BVH) on the
    host cores.  Returns (Mrays/s, seconds per step, sample size)."""
    from oracle import pyoracle as O
    col = O.Collider(tris)
    org, d = make_rays(n_sample, SEED + 99)
    for _ in range(warmup):
        col.first_hits(org[: n_sample // 8], d[: n_sample // 8], threads=threads)
    t0 = time.perf_counter()
    for _ in range(steps):
        col.first_hits(org, d, threads=threads)
    dt = (time.perf_counter() - t0) / steps
    return n_sample / dt / 1e6, dt, n_sample


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle import pyoracle as O
    threads = O.hardware_threads()
    tris = make_mesh()
    n_sample = 1 << 21
    rate, dt, ns = cpu_reference_rate(tris, n_sample, threads, steps=max(1, args.steps), warmup=min(1, args.warmup))
    sample = "%d of the 2^24 rays per step, %d steps" % (ns, max(1, args.steps))
    line = {
        "impl": "reference", "metric": "first_hit_Mrays_per_s", "value": rate, "unit": "Mrays/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
        "data": "synthetic",
        "config": {"workload": "C2 raw ray batch: 2^24 random rays vs 1,003,520-triangle icosphere BVH, first hit",
                   "mesh": "NewMeshIcosphere(0,1,224)", "rays_per_gpu": N_RAYS, "ray_mix": "A: origin~N(0,I), dir uniform S^2"},
        "cpu_baseline": {"value": rate, "unit": "Mrays/s", "cores": threads, "kind": "port", "sample": sample,
                         "note": "C++ float64 restatement of the Go reference (no Go toolchain in the image)"},
        "e2e": {"value": rate, "unit": "Mrays/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


C3_WORKLOAD = "C3 cornell_box (examples/renderings/cornell_box), RecursiveRayTracer MaxDepth 5, Cutoff 1e-4, Antialias 1, PhongFocusPoint 0.3"


C5_WORKLOAD = "C5 cornell_box with the ceiling light as MeshAreaLight, BidirPathTracer MaxDepth 10, MinDepth 3, RouletteDelta 0.2, PowerHeuristic 2, Antialias 1, Cutoff 1e-4"
C5_KW = dict(max_depth=10, min_depth=3, roulette_delta=0.2, power_heuristic=2.0, antialias=1.0, cutoff=1e-4)


C4_WORKLOAD = "C4 showcase (examples/renderings/showcase, 326,136 triangles in 7 meshes + 2 spheres + rect + cylinder, vase omitted: asset missing), RecursiveRayTracer MaxDepth 10, Cutoff 1e-4, Antialias 1, SphereFocusPoint 0.3, fixed spp"
WORKLOAD_NAME = {"c3": C3_WORKLOAD, "c4": C4_WORKLOAD, "c5": C5_WORKLOAD}


def cornell_tracer(spp, workload="c3", seed=1234, ctx=None):
    from model3d_b200 import examples
    kw = {"ctx": ctx} if ctx is not None else {}
    if workload == "c4":
        spec = examples.showcase(hd=True)
        psc = examples.build_product(spec, **kw)
        tr = examples.product_tracer(spec, psc, spec["max_depth"], spp, cutoff=spec["cutoff"],
                                     antialias=spec["antialias"], seed=seed)
        return spec, psc, tr
    spec = examples.cornell_box()
    psc = examples.build_product(spec, **kw)
    if workload == "c5":
        tr = examples.product_bidir(spec, psc, num_samples=spp, seed=seed, **C5_KW)
    else:
        tr = examples.product_tracer(spec, psc, 5, spp, cutoff=1e-4, antialias=1.0, seed=seed)
    return spec, psc, tr


def path_measure(wl, spp, size, steps, warmup, rank, world, local_rank, reduce_mode="fused", lib_devices=0):
    """Times one path-tracing workload: `steps` frames of W x H x spp samples, samples sharded by
    index over the GPUs (strong scaling), CUDA events on the launching stream, max over ranks.

    How the per-pixel sums of the shards meet (the only exchange step of the path):
      one process per GPU (torchrun), reduce_mode "fused": every rank's path_flush kernel adds
          straight into rank 0's accumulator (CUDA IPC mapping over NVLink, red.add; model3d_b200/
          distributed.SharedAccumulator); a 4-byte all_reduce per frame is the only NCCL call;
      reduce_mode "nccl": local accumulators + one NCCL reduce per frame (round 1, kept for A/B);
      one process, lib_devices > 1: m3d_ctx_create_multi -- the library shards and reduces inside
          m3d_render_path / m3d_render_bidir (one host thread per device, peer-mapped red.add).
    Returns a dict (identical on every rank)."""
    import torch
    import torch.distributed as dist
    from model3d_b200 import _native as N, distributed as D
    dev = torch.device("cuda", local_rank)
    ctx = N.MultiContext(list(range(lib_devices))) if lib_devices > 1 else N.default_context(local_rank)
    W = H = size
    spec, psc, tr = cornell_tracer(spp, wl, ctx=ctx)
    if wl == "c4":
        W, H = spec["size"]  # "HD" 960x640 (showcase/main.go:61-69)
    part, my_spp = D.sample_shard(spp, rank, world)
    stream = torch.cuda.current_stream().cuda_stream
    fused = world > 1 and reduce_mode == "fused"
    sa = None
    fused_note = ""
    if fused:
        # every rank must take the same path: agree on whether the IPC mapping came up everywhere
        ok = torch.ones(1, dtype=torch.int32, device=dev)
        try:
            sa = D.SharedAccumulator(W * H * 12, rank, world, D._CudaIpcBackend(ctx), nbuf=2)
        except Exception as e:
            ok.zero_()
            fused_note = "fused flush unavailable (%s); " % str(e)[:120]
        dist.all_reduce(ok, op=dist.ReduceOp.MIN)
        if int(ok.item()) == 0:
            if sa is not None:
                sa.close()
            sa, fused = None, False
            fused_note = fused_note or "fused flush unavailable on another rank; "
    if fused:
        accs = [torch.as_tensor(D.DevicePointer(sa.ptr(k), (H, W, 3)), device=dev) if rank == 0 else None
                for k in range(2)]
    else:
        accs = [torch.zeros((H, W, 3), dtype=torch.float32, device=dev)]
    mean = torch.empty((H, W, 3), dtype=torch.float32, device=dev) if rank == 0 else None
    info = {"rays": 0, "launches": 0}
    marks = []

    def step(k, timed):
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(3)] if timed else None
        if timed:
            ev[0].record()
        if fused:
            st = tr.RenderSumsDevice(W, H, psc, sa.ptr(k), partition=part + (N.PART_ATOMIC,),
                                     sample_count=my_spp, stream=stream)
        else:
            accs[0].zero_()
            st = tr.RenderSumsDevice(W, H, psc, accs[0].data_ptr(), partition=part, sample_count=my_spp,
                                     stream=stream)
        if timed:
            ev[1].record()
        if fused:
            sa.barrier(dev)  # every rank's flush has landed in rank 0's memory
            if rank == 0:
                torch.mul(accs[k % 2], 1.0 / spp, out=mean)  # colorSum / numSamples (ray_renderer.go:150)
                accs[k % 2].zero_()
        else:
            if world > 1:
                dist.reduce(accs[0], dst=0, op=dist.ReduceOp.SUM)
            if rank == 0:
                torch.mul(accs[0], 1.0 / spp, out=mean)
        if timed:
            ev[2].record()
            marks.append(ev)
        info["rays"], info["launches"] = st["rays"], st["launches"]
        if timed:
            info["kernel_ms"] = info.get("kernel_ms", 0.0) + st["kernel_ms"] / steps

    for k in range(warmup):
        step(k, False)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    torch.cuda.synchronize()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    for k in range(steps):
        step(warmup + k, True)
    ev1.record()
    torch.cuda.synchronize()
    clocks = sampler.stop() if rank == 0 else None
    render_ms = sum(e[0].elapsed_time(e[1]) for e in marks) / steps
    sync_ms = sum(e[1].elapsed_time(e[2]) for e in marks) / steps
    t = torch.tensor([ev0.elapsed_time(ev1) / steps, float(info["rays"]), render_ms, sync_ms,
                      info.get("kernel_ms", 0.0)], dtype=torch.float64, device=dev)
    per_rank = [[float(x) for x in t.tolist()]]
    if world > 1:
        dist.barrier()
        gathered = [torch.zeros_like(t) for _ in range(world)]
        dist.all_gather(gathered, t)
        per_rank = [[float(x) for x in g.tolist()] for g in gathered]
    ms = max(r[0] for r in per_rank)
    total_rays = sum(r[1] for r in per_rank)
    checksum = float(mean.sum().item()) if rank == 0 else 0.0
    if sa is not None:
        torch.cuda.synchronize()
        dist.barrier()
        del accs
        sa.close()
    n_dev = max(world, lib_devices, 1)
    return {"W": W, "H": H, "spp": spp, "ms_per_step": ms, "total_rays": total_rays, "clocks": clocks,
            "launches": int(info["launches"]), "n_devices": n_dev,
            "render_ms_per_rank": [round(r[2], 3) for r in per_rank],
            "reduce_wait_ms_per_rank": [round(r[3], 3) for r in per_rank],
            "kernel_ms_per_rank": [round(r[4], 3) for r in per_rank],
            "image_mean": checksum / (W * H * 3),
            "reduce": fused_note + ("none (one GPU)" if n_dev == 1 else
                       "library: m3d_ctx_create_multi, peer-mapped red.add inside path_flush" if lib_devices > 1 else
                       "fused: path_flush red.add into rank 0's accumulator over a CUDA IPC / NVLink mapping, "
                       "one 4-byte all_reduce per frame as the barrier" if fused else
                       "nccl: one reduce of the W*H*3 float32 sums per frame"),
            "tracer": tr, "scene": psc, "spec": spec}


def run_path(args):
    """BASELINE configs[2..4]: cornell_box / showcase with the RecursiveRayTracer, cornell_box with
    the BidirPathTracer, --spp samples per pixel per step.  N GPUs: samples sharded by index
    (strong scaling); see path_measure for how the shards are reduced."""
    import torch
    import torch.distributed as dist

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    dev = torch.device("cuda", local_rank)
    lib_devices = args.gpus if (world == 1 and args.gpus > 1) else 0  # one process driving several GPUs
    tstream = torch.cuda.Stream(device=dev)
    torch.cuda.synchronize()
    torch.cuda.set_stream(tstream)
    m = path_measure(args.workload, args.spp, args.size, args.steps, max(3, args.warmup), rank, world, local_rank,
                     reduce_mode=args.reduce, lib_devices=lib_devices)
    W, H, spp, ms_step, total_rays = m["W"], m["H"], m["spp"], m["ms_per_step"], m["total_rays"]
    tr, psc = m["tracer"], m["scene"]
    n_dev = m["n_devices"]
    samples = W * H * spp
    value = samples / (ms_step * 1e-3) / 1e6
    if rank == 0:
        # end to end: the host-buffer call (sums copied back to host memory every step)
        e2e = None
        if world == 1 and not args.no_e2e:
            tr.RenderSums(W, H, psc)
            t0 = time.perf_counter()
            k = max(1, args.steps // 2)
            for _ in range(k):
                tr.RenderSums(W, H, psc)
            dt = (time.perf_counter() - t0) / k
            e2e = {"value": samples / dt / 1e6, "unit": "Msamples/s", "ms_per_step": dt * 1e3,
                   "h2d_bytes_per_step": 0, "d2h_bytes_per_step": W * H * 12,
                   "api": "m3d_render_%s (host image buffers)" % ("bidir" if args.workload == "c5" else "path")}
        line = {
            "metric": "path_traced_Msamples_per_s", "value": value, "unit": "Msamples/s", "n_gpus": n_dev,
            "steps": args.steps, "warmup": max(3, args.warmup), "ms_per_step": ms_step, "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": WORKLOAD_NAME[args.workload], "width": W, "height": H,
                       "spp": spp,
                       "rays_per_sample": total_rays / samples, "Mrays_per_s": total_rays / (ms_step * 1e-3) / 1e6,
                       "sharding": "sample index; " + m["reduce"],
                       "render_ms_per_rank": m["render_ms_per_rank"],
                       "reduce_wait_ms_per_rank": m["reduce_wait_ms_per_rank"],
                       "kernel_ms_per_rank": m["kernel_ms_per_rank"], "image_mean": m["image_mean"],
                       "l2": "path state streams through HBM (> L2 per batch); scene BVH is L2/L1 resident"},
            "clocks": m["clocks"], "gpu_launches": m["launches"] * args.steps * (n_dev if lib_devices else 1),
        }
        if e2e:
            line["e2e"] = e2e
        # Path state that crosses HBM per traced ray: queue entry 40 B written and
        # read, raw hit 16 + 16, hit record 48 + 48, throughput 16 + 16 = 256 B.  The path tracers
        # are bound by issue rate and dependent-load latency, not by this stream; the fraction says so.
        # N GPUs: the bytes of all GPUs against N times one GPU's peak.
        peak, peak_src = hbm_peak()
        achieved = total_rays * 256.0 / (ms_step * 1e-3) / 1e9
        line["roofline"] = {"bound": "hbm", "achieved": achieved, "peak": peak * n_dev, "unit": "GB/s",
                            "frac": achieved / (peak * n_dev), "traffic": None,
                            "peak_source": peak_src + (" x %d GPUs" % n_dev if n_dev > 1 else ""),
                            "bytes_per_ray": 256.0, "kernel": "trace_first_hit_kernel + path_resolve / path_sample"
                            if args.workload != "c5" else "bidir_connect_kernel + trace_first_hit_kernel"}
        if world == 1 and not args.no_cpu_baseline:
            rate, _, threads, sample = path_cpu_rate(args.workload, 1)
            line["cpu_baseline"] = {"value": rate, "unit": "Msamples/s", "cores": threads, "kind": "port",
                                    "sample": sample,
                                    "note": "C++ float64 restatement of the Go renderer on all host threads"}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


C1_WORKLOAD = ("C1 model3d.Sphere -> MarchingCubesSearch(0.01, 8) mesh (376,832 triangles) rendered by "
               "render3d.RayCaster 512x512 in the SaveRendering setup, one frame per step")


def run_c1(args):
    """BASELINE configs[0]: one RayCaster frame per step (raygen + first-hit traversal + float64
    finish + Phong shading).  N GPUs: row bands, one frame each per step, gathered by summing the
    zero-initialised band images (NCCL reduce), inside the timed region."""
    import torch
    import torch.distributed as dist
    from model3d_b200 import _native as N, distributed as D, examples, render3d as R

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    dev = torch.device("cuda", local_rank)
    N.default_context(local_rank)
    spec = examples.c1_scene()
    psc = examples.build_product(spec)
    W = H = 512
    cam = spec["camera"]
    lt = spec["lights"][0]
    rc = R.RayCaster(Camera=R.NewCameraAt(cam["src"], cam["dst"], cam["fov"]),
                     Lights=[R.PointLight(lt["origin"], lt["color"])])
    band = (rank * H // world, (rank + 1) * H // world)
    acc = torch.zeros((H, W, 3), dtype=torch.float32, device=dev)
    ts = torch.cuda.Stream(device=dev)
    torch.cuda.synchronize()
    torch.cuda.set_stream(ts)
    launches = [0]

    def step():
        acc.zero_()
        st = rc.RenderDevice(W, H, psc, acc.data_ptr(), partition=band, stream=ts.cuda_stream)
        launches[0] = st["launches"]
        D.reduce_sums(acc, dst=0)

    for _ in range(max(3, args.warmup)):
        step()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    for _ in range(args.steps):
        step()
    ev1.record()
    torch.cuda.synchronize()
    t = torch.tensor([ev0.elapsed_time(ev1)], dtype=torch.float64, device=dev)
    clocks = sampler.stop() if rank == 0 else None
    if world > 1:
        dist.barrier()
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_step = float(t.item()) / args.steps
    if rank == 0:
        e2e = None
        if world == 1 and not args.no_e2e:
            img = R.Image(W, H)
            rc.Render(img, psc)
            t0 = time.perf_counter()
            k = max(3, args.steps // 4)
            for _ in range(k):
                rc.Render(img, psc)
            dt = (time.perf_counter() - t0) / k
            e2e = {"value": W * H / dt / 1e6, "unit": "Mrays/s", "ms_per_step": dt * 1e3,
                   "h2d_bytes_per_step": W * H * 12, "d2h_bytes_per_step": W * H * 12,
                   "api": "m3d_render_raycast (host image buffer in and out)"}
        line = {"metric": "first_hit_Mrays_per_s", "value": W * H / (ms_step * 1e-3) / 1e6, "unit": "Mrays/s",
                "n_gpus": world, "steps": args.steps, "warmup": max(3, args.warmup), "ms_per_step": ms_step,
                "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32",
                "data": "synthetic",
                "config": {"workload": C1_WORKLOAD, "width": W, "height": H, "triangles": 376832,
                           "frames_per_s": 1e3 / ms_step, "sharding": "row bands, NCCL reduce of the band images",
                           "l2": "the 262,144-ray frame and the 21 MB BVH fit in L2; a launch-latency-bound case"},
                "clocks": clocks, "gpu_launches": int(launches[0]) * args.steps}
        if e2e:
            line["e2e"] = e2e
        # algorithmic bytes per primary ray, counted on the same camera rays (untimed pass)
        dirs = R.CasterRays(rc.Camera, W, H).astype(np.float32)
        orgs = np.tile(np.asarray(cam["src"], np.float32), (W * H, 1))
        cst = psc.Cast(orgs, dirs, counters=True)["stats"]
        v_nodes, t_tris = cst["nodes_visited"] / (W * H), cst["tris_tested"] / (W * H)
        b_ray = RAY_IO_BYTES + v_nodes * NODE_BYTES + t_tris * TRI_BYTES + 12  # + the pixel written
        peak, peak_src = hbm_peak()
        achieved = W * H * b_ray / (ms_step * 1e-3) / 1e9
        line["roofline"] = {"bound": "hbm", "achieved": achieved, "peak": peak * world, "unit": "GB/s",
                            "frac": achieved / (peak * world), "traffic": None, "peak_source": peak_src,
                            "bytes_per_ray": b_ray, "nodes_per_ray": v_nodes, "tris_per_ray": t_tris,
                            "kernel": "raygen_camera + trace_first_hit_kernel + finish_scene_hits + shade_raycast",
                            "note": "a 262,144-ray frame is four dependent launches of ~50 us: launch latency, "
                                    "not bandwidth, bounds it (the same kernels reach 0.61 on the 2^24-ray batch)"}
        if world == 1 and not args.no_cpu_baseline:
            ref = c1_cpu_rate(3)
            line["cpu_baseline"] = {"value": ref[0], "unit": "Mrays/s", "cores": ref[2], "kind": "port",
                                    "sample": "the full 512x512 frame, 3 frames",
                                    "note": "C++ float64 restatement of the Go RayCaster on all host threads"}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def c1_cpu_rate(frames):
    """The oracle's RayCaster (float64 restatement of raycast.go:15-39) on the C1 scene, all host
    threads: (Mrays/s, seconds per frame, threads)."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import scenes
    from oracle import pyoracle as O
    threads = O.hardware_threads()
    spec = scenes.c1_scene()
    osc = scenes.build_oracle(spec)
    cam = spec["camera"]
    ocam = O.camera_at(cam["src"], cam["dst"], cam["fov"])
    lt = spec["lights"][0]
    ol = O.PointLight()
    ol.origin[:], ol.color[:], ol.quad_dropoff = lt["origin"], lt["color"], 0
    W = H = 512
    k = max(1, frames)
    t0 = time.perf_counter()
    for _ in range(k):
        osc.render_raycast(ocam, [ol], W, H, threads=threads)
    dt = (time.perf_counter() - t0) / k
    return W * H / dt / 1e6, dt, threads


def run_reference_c1(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    W = H = 512
    rate, dt, threads = c1_cpu_rate(min(max(1, args.steps), 10))
    line = {"impl": "reference", "metric": "first_hit_Mrays_per_s", "value": rate, "unit": "Mrays/s",
            "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt * 1e3,
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": C1_WORKLOAD},
            "cpu_baseline": {"value": rate, "unit": "Mrays/s", "cores": threads, "kind": "port",
                             "sample": "the full 512x512 frame per step"},
            "e2e": {"value": rate, "unit": "Mrays/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


def hbm_peak():
    """(GB/s, source): the driver-measured copy bandwidth of this pool's B200s, else the profiling
    recipe's fallback."""
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    return (float(peaks.get("hbm_gbs", 6650.0)),
            "measured (MEASURED_PEAKS.json hbm_gbs)" if "hbm_gbs" in peaks else "fallback 6.65 TB/s")


def path_cpu_rate(workload, steps):
    """The oracle's float64 restatement of the reference renderer for this workload on all host
    threads, on a bounded sample (a small frame at a few spp): (Msamples/s, s per step, threads, sample)."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import scenes
    from oracle import pyoracle as O
    threads = O.hardware_threads()
    spec = scenes.showcase(hd=True) if workload == "c4" else scenes.cornell_box()
    osc = scenes.build_oracle(spec)
    cam = spec["camera"]
    ocam = O.camera_at(cam["src"], cam["dst"], cam["fov"])
    W = H = 256
    spp = 8
    if workload == "c4":
        W, H, spp = 240, 160, 4
    if workload == "c5":
        W = H = 128
        bp, lights = scenes.oracle_bidir_params(spec, num_samples=spp, seed=3, **C5_KW)
    elif workload == "c4":
        pp = scenes.oracle_path_params(spec, osc, 10, spp, cutoff=1e-4, antialias=1.0, seed=3)
    else:
        pp = scenes.oracle_path_params(spec, osc, 5, spp, cutoff=1e-4, antialias=1.0, seed=3)
    t0 = time.perf_counter()
    k = max(1, steps)
    for _ in range(k):
        if workload == "c5":
            osc.render_bidir(ocam, lights, bp, W, H, threads=threads)
        else:
            osc.render_path(ocam, [], pp, W, H, threads=threads)
    dt = (time.perf_counter() - t0) / k
    return W * H * spp / dt / 1e6, dt, threads, "%dx%d at %d spp per step" % (W, H, spp)


def run_reference_path(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    rate, dt, threads, sample = path_cpu_rate(args.workload, args.steps)
    line = {"impl": "reference", "metric": "path_traced_Msamples_per_s", "value": rate, "unit": "Msamples/s",
            "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt * 1e3,
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": WORKLOAD_NAME[args.workload]},
            "cpu_baseline": {"value": rate, "unit": "Msamples/s", "cores": threads, "kind": "port",
                             "sample": sample},
            "e2e": {"value": rate, "unit": "Msamples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


# BASELINE configs[2..4] at their own sample counts (the driver-run lines must be
# measured at 256 / 256 / 1024 spp, >= 3 warm-up and >= 5 timed frames each)
SECONDARY = (("c3", 256), ("c4", 256), ("c5", 1024))


def secondary_path_metrics(rank, world, local_rank, reduce_mode="fused", steps=5, warmup=3):
    """Path-tracing measurements appended to the headline line (BASELINE metric: 'Mrays/s first-hit
    AND path-traced samples/s'): cornell_box 1024x1024 with the RecursiveRayTracer (C3, 256 spp) and
    the BidirPathTracer (C5, 1024 spp), the showcase scene 960x640 (C4, 256 spp); samples sharded
    over the ranks, every rank's flush kernel reducing into rank 0's accumulator inside the timed
    region.  CUDA-event timed, max over ranks; per-rank render / reduce-wait split reported."""
    out = {}
    for wl, spp in SECONDARY:
        m = path_measure(wl, spp, 1024, steps, warmup, rank, world, local_rank, reduce_mode=reduce_mode)
        W, H, ms = m["W"], m["H"], m["ms_per_step"]
        out[wl] = {"workload": WORKLOAD_NAME[wl], "width": W, "height": H, "spp": spp, "steps": steps,
                   "warmup": warmup,
                   "Msamples_per_s": W * H * spp / (ms * 1e-3) / 1e6,
                   "Mrays_per_s": m["total_rays"] / (ms * 1e-3) / 1e6,
                   "ms_per_frame": ms, "render_ms_per_rank": m["render_ms_per_rank"],
                   "reduce_wait_ms_per_rank": m["reduce_wait_ms_per_rank"],
                   "kernel_ms_per_rank": m["kernel_ms_per_rank"], "clocks": m["clocks"],
                   "scaling": "strong (sample shards); " + m["reduce"]}
        del m
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=None,
                    help="timed steps (default: 100 first-hit batches / RayCaster frames, 10 C3 / C4 frames, 5 C5 frames)")
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200")
    ap.add_argument("--rays", type=int, default=N_RAYS)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-secondary", action="store_true",
                    help="skip the short path-tracing measurements appended to the headline line")
    ap.add_argument("--workload", default="c2", choices=["c1", "c2", "c3", "c4", "c5"],
                    help="c2: raw first-hit ray batch (headline); c1: marching-cubes sphere, RayCaster 512x512; "
                         "c3: cornell_box RecursiveRayTracer 1024x1024; c4: showcase HD; "
                         "c5: cornell_box BidirPathTracer 1024x1024")
    ap.add_argument("--reduce", default="fused", choices=["fused", "nccl"],
                    help="N > 1 path tracing: fused = every rank's flush kernel adds into rank 0's accumulator "
                         "over NVLink (default); nccl = local accumulators + one NCCL reduce per frame")
    ap.add_argument("--spp", type=int, default=0,
                    help="samples per pixel per step (default: the BASELINE value, c3/c4 256, c5 1024)")
    ap.add_argument("--size", type=int, default=1024, help="frame width == height (c3)")
    args = ap.parse_args()
    if not args.spp:
        args.spp = 1024 if args.workload == "c5" else 256
    if args.steps is None:  # a C5 frame at 1024 spp takes seconds: keep the default run within minutes
        args.steps = {"c3": 10, "c4": 10, "c5": 5}.get(args.workload, 100)
    if args.impl == "reference":
        if args.workload in ("c3", "c4", "c5"):
            run_reference_path(args)
        elif args.workload == "c1":
            run_reference_c1(args)
        else:
            run_reference(args)
        return
    if args.workload in ("c3", "c4", "c5"):
        run_path(args)
        return
    if args.workload == "c1":
        run_c1(args)
        return

    import torch
    import torch.distributed as dist

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    dev = torch.device("cuda", local_rank)

    from model3d_b200 import MeshCollider
    from model3d_b200 import _native as N

    tris = make_mesh()
    ctx = N.Context(local_rank)
    col = MeshCollider(tris, ctx=ctx)
    info = col.Info()

    n = args.rays
    org, d = make_rays(n, SEED + rank)
    # device-resident SoA inputs (float4): (ox,oy,oz,tmin) (dx,dy,dz,tmax)
    o4 = torch.zeros((n, 4), dtype=torch.float32)
    d4 = torch.full((n, 4), float("inf"), dtype=torch.float32)
    o4[:, :3] = torch.from_numpy(org)
    d4[:, :3] = torch.from_numpy(d)
    o4, d4 = o4.to(dev), d4.to(dev)
    h0 = torch.empty((n, 4), dtype=torch.float32, device=dev)
    h1 = torch.empty((n, 4), dtype=torch.float32, device=dev)
    # a dedicated (non-default) torch stream: the library launches on the stream it is
    # given, and torch.cuda.Event only sees torch's current stream
    tstream = torch.cuda.Stream(device=dev)
    torch.cuda.synchronize()
    torch.cuda.set_stream(tstream)
    stream = tstream.cuda_stream
    assert stream != 0

    def step():
        col.FirstRayCollisionsDevice(o4.data_ptr(), d4.data_ptr(), n, h0.data_ptr(), h1.data_ptr(), stream=stream)

    # counters pass (separate kernel instantiation, untimed): nodes / triangles per ray
    st = col.FirstRayCollisionsDevice(o4.data_ptr(), d4.data_ptr(), n, h0.data_ptr(), h1.data_ptr(),
                                      stream=stream, counters=True)
    torch.cuda.synchronize()
    nodes_per_ray = st["nodes_visited"] / n
    tris_per_ray = st["tris_tested"] / n
    hit_frac = float((h0[:, 3].contiguous().view(torch.int32) >= 0).float().mean().item())

    for _ in range(max(3, args.warmup)):
        step()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    torch.cuda.synchronize()
    ev0 = torch.cuda.Event(enable_timing=True)
    ev1 = torch.cuda.Event(enable_timing=True)
    ev0.record()
    for _ in range(args.steps):
        step()
    ev1.record()
    torch.cuda.synchronize()
    ms_total = ev0.elapsed_time(ev1)
    clocks = sampler.stop() if rank == 0 else None
    t = torch.tensor([ms_total], dtype=torch.float64, device=dev)
    if world > 1:
        dist.barrier()
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_total = float(t.item())
    ms_step = ms_total / args.steps
    value = world * n / (ms_step * 1e-3) / 1e6

    # end to end through the host-buffer C-ABI call, H2D + kernels + D2H inside the timed region.
    # Headline e2e: caller buffers from m3d_host_alloc (pinned; what the cgo binding hands out) and
    # all of t / prim / normal returned.  Beside it: the same call on pageable numpy arrays (what a
    # caller gets who ignores m3d_host_alloc) and with only t + prim returned (8 B/ray back).
    e2e = None
    if not args.no_e2e:
        import ctypes as C
        f32p, i32p = C.POINTER(C.c_float), C.POINTER(C.c_int32)
        org_p, d_p = N.host_empty((n, 3), np.float32), N.host_empty((n, 3), np.float32)
        org_p[:], d_p[:] = org, d
        t_p, prim_p = N.host_empty(n, np.float32), N.host_empty(n, np.int32)
        nrm_p = N.host_empty((n, 3), np.float32)

        def call(o, dd, t, pr, nr):
            N.check(N.lib().m3d_mesh_first_ray_collisions(
                col.h, o.ctypes.data_as(f32p), dd.ctypes.data_as(f32p), C.c_int64(n),
                t.ctypes.data_as(f32p), pr.ctypes.data_as(i32p),
                nr.ctypes.data_as(f32p) if nr is not None else None, None, C.c_uint32(0), None))

        def timed(fn, reps):
            for _ in range(2):
                fn()
            if world > 1:
                dist.barrier()
            t0 = time.perf_counter()
            for _ in range(reps):
                fn()
            ctx.synchronize()
            tt = torch.tensor([(time.perf_counter() - t0) / reps], dtype=torch.float64, device=dev)
            if world > 1:
                dist.all_reduce(tt, op=dist.ReduceOp.MAX)
            return float(tt.item())

        e2e_steps = max(2, min(args.steps // 2, 10))
        dt = timed(lambda: call(org_p, d_p, t_p, prim_p, nrm_p), e2e_steps)
        e2e = {"value": world * n / dt / 1e6, "unit": "Mrays/s", "ms_per_step": dt * 1e3,
               "h2d_bytes_per_step": n * 24, "d2h_bytes_per_step": n * 20,
               "api": "m3d_mesh_first_ray_collisions (host buffers from m3d_host_alloc: pinned)"}
        dt_min = timed(lambda: call(org_p, d_p, t_p, prim_p, None), e2e_steps)
        e2e["t_prim_only"] = {"value": world * n / dt_min / 1e6, "ms_per_step": dt_min * 1e3,
                              "d2h_bytes_per_step": n * 8}
        t_g, prim_g, nrm_g = np.empty(n, np.float32), np.empty(n, np.int32), np.empty((n, 3), np.float32)
        dt_pg = timed(lambda: call(org, d, t_g, prim_g, nrm_g), 2)
        e2e["pageable"] = {"value": world * n / dt_pg / 1e6, "ms_per_step": dt_pg * 1e3,
                           "note": "same call on ordinary (pageable) numpy arrays: the driver stages the copies"}
        assert np.array_equal(prim_g, prim_p)
        del t_g, prim_g, nrm_g
        # what the host link allows: the same bytes copied both ways at once with no compute at all
        # (two streams, pinned memory), all ranks at the same time -- every GPU of the box shares the
        # host's memory system, so at N > 1 this bound, not the GPUs, is what e2e scales with
        d_in = torch.empty(n * 24, dtype=torch.uint8, device=dev)
        d_out = torch.empty(n * 20, dtype=torch.uint8, device=dev)
        h_in = torch.from_numpy(org_p.view(np.uint8).reshape(-1)), torch.from_numpy(d_p.view(np.uint8).reshape(-1))
        h_out_np = N.host_empty(n * 20, np.uint8)
        h_out = torch.from_numpy(h_out_np)
        s_in, s_out = torch.cuda.Stream(device=dev), torch.cuda.Stream(device=dev)

        def copies():
            with torch.cuda.stream(s_in):
                d_in[:n * 12].copy_(h_in[0], non_blocking=True)
                d_in[n * 12:].copy_(h_in[1], non_blocking=True)
            with torch.cuda.stream(s_out):
                h_out.copy_(d_out, non_blocking=True)
            s_in.synchronize()
            s_out.synchronize()
        cb = timed(copies, 5)
        e2e["copy_bound_ms"] = cb * 1e3
        e2e["frac_of_copy_bound"] = cb / dt
        e2e["copy_bound_note"] = "%d rank(s) copying %d MB in + %d MB out each, concurrently, no compute" % (
            world, n * 24 >> 20, n * 20 >> 20)
        del d_in, d_out, h_out, h_out_np, h_in

    # mix B: coherent primary rays of a 4096x4096 pinhole camera at (0,-3,0) looking at
    # the origin, fov pi/3.6 -- same mesh, same kernels, reported beside the incoherent headline
    mix_b = None
    if not args.no_secondary and n == N_RAYS:
        from model3d_b200 import render3d as R
        cam = R.NewCameraAt((0.0, -3.0, 0.0), (0.0, 0.0, 0.0), np.pi / 3.6)
        db = torch.from_numpy(R.CasterRays(cam, 4096, 4096).astype(np.float32))
        o4.zero_()
        o4[:, 1] = -3.0
        d4[:, :3] = db.to(dev)
        del db
        stb = col.FirstRayCollisionsDevice(o4.data_ptr(), d4.data_ptr(), n, h0.data_ptr(), h1.data_ptr(),
                                           stream=stream, counters=True)
        for _ in range(3):
            step()
        torch.cuda.synchronize()
        eb0, eb1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        kb = max(5, min(args.steps, 20))
        eb0.record()
        for _ in range(kb):
            step()
        eb1.record()
        torch.cuda.synchronize()
        tb = torch.tensor([eb0.elapsed_time(eb1) / kb], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(tb, op=dist.ReduceOp.MAX)
        hit_b = float((h0[:, 3].contiguous().view(torch.int32) >= 0).float().mean().item())
        mix_b = {"ray_mix": "B: 4096x4096 pinhole camera at (0,-3,0) looking at the origin, fov pi/3.6",
                 "Mrays_per_s": world * n / (float(tb.item()) * 1e-3) / 1e6, "ms_per_step": float(tb.item()),
                 "hit_fraction": hit_b, "nodes_per_ray": stb["nodes_visited"] / n, "tris_per_ray": stb["tris_tested"] / n}
        mix_b["bytes_per_ray"] = RAY_IO_BYTES + mix_b["nodes_per_ray"] * NODE_BYTES + mix_b["tris_per_ray"] * TRI_BYTES
        mix_b["achieved_GBps"] = n * mix_b["bytes_per_ray"] / (mix_b["ms_per_step"] * 1e-3) / 1e9

    secondary = None
    if not args.no_secondary:
        try:
            secondary = secondary_path_metrics(rank, world, local_rank, reduce_mode=args.reduce)
        except Exception as e:  # the headline must still be reported
            secondary = {"error": str(e)[:200]}
    if rank == 0:
        peak, peak_src = hbm_peak()
        bytes_per_ray = RAY_IO_BYTES + nodes_per_ray * NODE_BYTES + tris_per_ray * TRI_BYTES
        achieved = n * bytes_per_ray / (ms_step * 1e-3) / 1e9
        traffic = None
        try:
            traffic = json.load(open(os.path.join(ROOT, "profiles", "trace_dram_traffic.json"))).get("bytes_per_launch")
        except Exception:
            pass
        line = {
            "metric": "first_hit_Mrays_per_s", "value": value, "unit": "Mrays/s", "n_gpus": world,
            "steps": args.steps, "warmup": max(3, args.warmup), "ms_per_step": ms_step,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic",
            "config": {
                "workload": "C2 raw ray batch: 2^24 random rays vs 1,003,520-triangle icosphere BVH, first hit",
                "mesh": "NewMeshIcosphere(0,1,224)", "triangles": int(info["num_triangles"]),
                "bvh_nodes": int(info["num_nodes"]), "bvh_bytes": int(info["device_bytes"]),
                "rays_per_gpu": n, "ray_mix": "A: origin~N(0,I), dir uniform S^2", "hit_fraction": hit_frac,
                "l2": "ray/hit streams (1 GiB per step) exceed L2; the 56 MB BVH is deliberately L2-resident",
                "final_hit_refine": "float64", "nodes_per_ray": nodes_per_ray, "tris_per_ray": tris_per_ray,
            },
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s",
                         "frac": achieved / peak, "traffic": traffic, "peak_source": peak_src,
                         "bytes_per_ray": bytes_per_ray, "kernel": "trace_first_hit_kernel",
                         "compulsory_stream_GBps": n * RAY_IO_BYTES / (ms_step * 1e-3) / 1e9,
                         "frac_hbm_compulsory": n * RAY_IO_BYTES / (ms_step * 1e-3) / 1e9 / peak,
                         "note": "achieved / frac follow SURVEY 8d (all algorithmic bytes over the HBM copy peak); "
                                 "the BVH is L2 resident, so the node + triangle fetches are also set against "
                                 "what L2 delivers on this GPU, measured live (m3d_measure_l2_bandwidth)"},
            "clocks": clocks,
            "gpu_launches": 2 * args.steps,
        }
        try:
            # L2 -> SM bandwidth over a working set of the BVH's size: coalesced stream, and the
            # traversal's own access pattern (32 different 80-byte records per warp instruction)
            import ctypes as C
            bvh_mb = max(8, int(info["device_bytes"]) >> 20)
            l2 = {}
            for mode, key in ((0, "stream"), (1, "gather")):
                g = C.c_double()
                N.check(N.lib().m3d_measure_l2_bandwidth(ctx.h, C.c_int32(mode), C.c_int64(bvh_mb << 20),
                                                         C.c_int32(3), C.byref(g)))
                l2[key] = g.value
            fetch = n * (nodes_per_ray * NODE_BYTES + tris_per_ray * TRI_BYTES) / (ms_step * 1e-3) / 1e9
            line["roofline"].update({
                "node_tri_fetch_GBps": fetch, "l2_stream_peak_GBps": l2["stream"], "l2_gather_peak_GBps": l2["gather"],
                "frac_l2_stream": fetch / l2["stream"], "frac_l2_gather": fetch / l2["gather"],
                "l2_working_set_MB": bvh_mb})
        except Exception as e:
            line["roofline"]["l2_error"] = str(e)[:200]
        if e2e:
            line["e2e"] = e2e
        if mix_b:
            line["coherent_rays"] = mix_b
        if secondary:
            line["path_tracing"] = secondary
        if not args.no_cpu_baseline and world == 1:
            from oracle import pyoracle as O
            threads = O.hardware_threads()
            rate, dt, ns = cpu_reference_rate(tris, 1 << 21, threads)
            reps = 1
            if dt < 4.0:
                # aim at ~10 s of CPU work: the full 2^24-ray batch, repeated as needed
                reps = int(max(1, min(8, round(10.0 / max(dt * 8, 1e-3)))))
                rate, dt, ns = cpu_reference_rate(tris, 1 << 24, threads, steps=reps)
            line["cpu_baseline"] = {"value": rate, "unit": "Mrays/s", "cores": threads, "kind": "port",
                                    "sample": "%d of the 2^24 rays x %d passes, %.1f s per pass" % (ns, reps, dt),
                                    "note": "C++ float64 restatement of the Go reference on all host threads "
                                            "(no Go toolchain in the image)"}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
