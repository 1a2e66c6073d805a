#!/usr/bin/env python
"""bench.py -- skinned vertices/s of the vertex-animation hot path on B200, beside the CPU path.

Contract (see the task statement): `python bench.py --gpus N --steps K --warmup W` prints ONE JSON
line on rank 0.  A *step* is one animate pass over the whole crowd (every joint matrix changed, so
every blend is recomputed and every vertex re-skinned).

Workload (BASELINE.json configs[2], the configuration the metric "... at 1/2/4/8 B200" is quoted on):
a crowd of 10 000 instances of a 20 000-vertex, 128-joint character (4096 4-weight blends),
position + normal animated (stride 24).  The crowd is sharded by instance over the N ranks with no
data-path collective ("scaling": "strong": the total is fixed at 10 000 instances).

  value     whole-job vertices/s, palettes already resident in HBM, CUDA events, max over ranks
  e2e       the same through the C-ABI with HOST buffers: palettes H2D from pinned memory and the
            animated vertex arrays D2H every step (what GeomVertexData::animate_vertices hands back)
  roofline  the dominant kernel (skin): algorithmic bytes / summed per-launch CUDA-event time
  cpu_baseline / --impl reference: the reference's own CPU animate_vertices (oracle/_ref, built from
            /root/reference) on the box's host cores, one worker process per core.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import tempfile
import threading
import time

import numpy as np

# stdout carries exactly one JSON line: file descriptor 1 is pointed at stderr for the whole run (NCCL prints its
# "NCCL version ..." banner there from native code) and the line is written to the saved descriptor at the end
_JSON_FD = os.dup(1)
os.dup2(2, 1)


def emit(line: dict) -> None:
    os.write(_JSON_FD, (json.dumps(line) + "\n").encode())

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

TOTAL_INSTANCES = 10_000
METRIC = "skinned vertices/sec"
UNIT = "vertices/s"


def workload_mesh():
    from panda3d_b200 import synth
    return synth.config_c3_mesh()


def config_dict(args, n_gpus):
    return {
        "workload": "C3 crowd: 10000 instances x (20000 vertices, 128 joints, 4096 4-weight blends), "
                    "pos+normal float32 stride 24, u16 transform_blend, rows=all",
        "instances_total": args.instances, "vertices_per_step": args.instances * 20_000,
        "sharding": f"instance i -> rank i % {n_gpus}, no data-path collective",
        "l2": "per-rank output (>= 600 MB) exceeds the 126 MB L2; no flush needed",
    }


# ------------------------------------------------------------------------------------ clocks
class ClockSampler(threading.Thread):
    """Samples SM clock / throttle reasons through NVML while the timed regions run."""

    def __init__(self, index: int):
        super().__init__(daemon=True)
        self.index = index
        self.samples = []
        self.reasons = set()
        self.max_mhz = None
        self.stop_flag = threading.Event()
        self.ok = False
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = int(pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM))
            self.ok = True
        except Exception:
            self.nv = None

    def run(self):
        if not self.ok:
            return
        nv = self.nv
        names = {
            getattr(nv, "nvmlClocksEventReasonHwSlowdown", 0x8): "hw_slowdown",
            getattr(nv, "nvmlClocksEventReasonHwThermalSlowdown", 0x40): "hw_thermal_slowdown",
            getattr(nv, "nvmlClocksEventReasonSwThermalSlowdown", 0x20): "sw_thermal_slowdown",
            getattr(nv, "nvmlClocksEventReasonSwPowerCap", 0x4): "sw_power_cap",
            getattr(nv, "nvmlClocksEventReasonHwPowerBrakeSlowdown", 0x80): "hw_power_brake",
        }
        while not self.stop_flag.is_set():
            try:
                util = nv.nvmlDeviceGetUtilizationRates(self.h).gpu
                mhz = int(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                try:
                    mask = int(nv.nvmlDeviceGetCurrentClocksEventReasons(self.h))
                except Exception:
                    mask = int(nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h))
                self.samples.append((mhz, util))
                for bit, name in names.items():
                    if mask & bit:
                        self.reasons.add(name)
            except Exception:
                pass
            self.stop_flag.wait(0.02)

    def summary(self):
        if not self.ok or not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons), "samples": 0}
        loaded = [m for m, u in self.samples if u > 0] or [m for m, _ in self.samples]
        return {"sm_mhz": statistics.median(loaded), "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons),
                "samples": len(self.samples)}


# ------------------------------------------------------------------------------------ host binding
def bind_to_gpu_cpus(torch, local_rank: int):
    """Binds this rank to the CPUs next to its GPU (sysfs local_cpulist) BEFORE any page-locked buffer is allocated, so
    the buffers the PCIe copies target are first touched on the GPU's own NUMA node."""
    info = {"bound": False}
    try:
        props = torch.cuda.get_device_properties(local_rank)
        bus = "%04x:%02x:%02x.0" % (getattr(props, "pci_domain_id", 0), getattr(props, "pci_bus_id", 0), getattr(props, "pci_device_id", 0))
        base = f"/sys/bus/pci/devices/{bus}"
        info["pci"] = bus
        info["numa_node"] = open(base + "/numa_node").read().strip()
        cpulist = open(base + "/local_cpulist").read().strip()
        info["local_cpulist"] = cpulist
        cpus = set()
        for part in cpulist.split(","):
            if "-" in part:
                a, b = part.split("-")
                cpus |= set(range(int(a), int(b) + 1))
            elif part.strip():
                cpus.add(int(part))
        allowed = os.sched_getaffinity(0)
        if cpus & allowed and (cpus & allowed) != allowed:
            os.sched_setaffinity(0, cpus & allowed)
            info["bound"] = True
        info["cpus"] = len(os.sched_getaffinity(0))
    except Exception as exc:   # a box without that sysfs entry: run unbound, say so
        info["error"] = repr(exc)
    return info


# ------------------------------------------------------------------------------------ CPU arms
def cpu_reference_run(mesh_synth, steps: int, warmup: int, characters_per_worker: int, workers: int):
    """Times the reference's own animate_vertices (oracle/_ref/bin/ref_harness bench) with one worker
    process per core -- in-process threads scale negatively in the reference.  Falls
    back to the oracle port (oracle/skin_oracle.c) when oracle/_ref is absent."""
    from oracle import oracle
    from panda3d_b200 import synth
    from panda3d_b200.casefile import save_case
    flat = mesh_synth.flat
    frames = 4
    pal = synth.make_palettes(flat.num_transforms, frames, 4242)
    rows = flat.num_rows
    if oracle.have_ref():
        with tempfile.TemporaryDirectory() as tmp:
            recipe = os.path.join(tmp, "bench.recipe")
            save_case(recipe, mesh_synth.recipe(pal))
            env = dict(os.environ)
            t0 = time.time()
            procs = [subprocess.Popen([oracle.REF_HARNESS, "bench", recipe, str(steps), str(warmup), str(characters_per_worker)],
                                      stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, env=env, text=True) for _ in range(workers)]
            outs = [p.communicate()[0] for p in procs]
            wall = time.time() - t0
        rates, mean_s = [], []
        for o in outs:
            line = [ln for ln in o.splitlines() if ln.startswith("{")][-1]
            j = json.loads(line)
            rates.append(j["verts_per_s_mean"])
            mean_s.append(j["mean_s"])
        return {"value": float(sum(rates)), "kind": "reference", "cores": workers, "ms_per_step": 1e3 * max(mean_s),
                "wall_s": wall, "vertices_per_step": rows * characters_per_worker * workers}
    # the oracle port, one process per core
    import multiprocessing as mp
    ctx = mp.get_context("fork")
    q = ctx.Queue()

    def worker():
        om = oracle.OracleMesh(flat)
        ts = []
        for it in range(warmup + steps):
            t = time.perf_counter()
            for c in range(characters_per_worker):
                om.animate(pal[(it + c) % frames])
            if it >= warmup:
                ts.append(time.perf_counter() - t)
        q.put(sum(ts) / len(ts))

    t0 = time.time()
    ps = [ctx.Process(target=worker) for _ in range(workers)]
    [p.start() for p in ps]
    means = [q.get() for _ in ps]
    [p.join() for p in ps]
    wall = time.time() - t0
    return {"value": float(sum(rows * characters_per_worker / m for m in means)), "kind": "port", "cores": workers,
            "ms_per_step": 1e3 * max(means), "wall_s": wall, "vertices_per_step": rows * characters_per_worker * workers}


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    sm = workload_mesh()
    workers = os.cpu_count() or 1
    chars = max(1, args.cpu_characters)
    r = cpu_reference_run(sm, args.steps, args.warmup, chars, workers)
    sample = f"{chars} character(s) per worker x {workers} worker processes per step (of {args.instances} instances)"
    line = {
        "impl": "reference", "metric": METRIC, "value": r["value"], "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": r["ms_per_step"], "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": config_dict(args, args.gpus),
        "cpu_baseline": {"value": r["value"], "unit": UNIT, "cores": r["cores"], "kind": r["kind"], "sample": sample},
        "e2e": {"value": r["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit(line)


# ------------------------------------------------------------------------------------ GPU arm
def run_gpu_arm(args):
    import torch
    from panda3d_b200 import _capi, synth

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus and world > 1:
        args.gpus = world
    dist = None
    torch.cuda.set_device(local_rank)
    binding = bind_to_gpu_cpus(torch, local_rank)
    if world > 1:
        import torch.distributed as dist_mod
        dist = dist_mod
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x: float) -> float:
        if dist is None:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def sum_over_ranks(x: float) -> float:
        if dist is None:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return float(t.item())

    sm = workload_mesh()
    flat = sm.flat
    rows, T = flat.num_rows, flat.num_transforms
    n_local = len(range(rank, args.instances, world))
    stream = torch.cuda.Stream()
    sampler = ClockSampler(local_rank)
    with torch.cuda.stream(stream):
        ctx = _capi.Context(local_rank, stream.cuda_stream)
        dmesh = _capi.Mesh(ctx, flat)
        stride = dmesh.out_strides[0]
        pitch = (rows * stride + 255) // 256 * 256
        # device memory is torch's (plumbing): palettes, outputs
        pal_host = torch.from_numpy(synth.make_palettes(T, n_local, 1000 + rank).reshape(n_local, T * 16)).pin_memory()
        pal_dev = torch.empty((n_local, T * 16), dtype=torch.float32, device="cuda")
        out_dev = torch.empty((n_local, pitch), dtype=torch.uint8, device="cuda")
        grp = _capi.Group(dmesh, n_local, palettes_dev=pal_dev.data_ptr(), outputs_dev=[out_dev.data_ptr()])
        assert grp.output_pitch(0) == pitch
        pal_dev.copy_(pal_host, non_blocking=True)
        stream.synchronize()

        sampler.start()
        # ---------------- kernel-only: palettes resident in HBM
        for _ in range(args.warmup):
            grp.animate()
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for _ in range(args.steps):
            grp.animate()
        e1.record(stream)
        barrier()
        own_ms = e0.elapsed_time(e1) / args.steps
        ms = max_over_ranks(own_ms)
        launches_per_step = grp.timings().launches

        # ---------------- per-launch timing of the kernels (roofline), same workload, same stream
        grp.set_profiling(True)
        blend_ms, skin_ms, skin_launches = [], [], 0
        for _ in range(min(args.steps, 10)):
            grp.animate()
            t = grp.timings()
            blend_ms.append(t.blend_ms)
            skin_ms.append(t.skin_ms)
            skin_launches = t.skin_launches
        grp.set_profiling(False)

        # ---------------- end to end through the C-ABI with host buffers
        e2e_steps = max(2, min(args.steps, args.e2e_steps))
        # With several GPUs on one host the device->host links are NOT equal (8 x B200 VM: 13 GB/s for GPUs 0-3, 18 GB/s
        # for GPUs 4-7 when all copy at once, tools/pcie_probe.py) and the step lasts as long as the slowest rank: the
        # end-to-end leg deals the instances in proportion to the rate every rank measures for itself, all ranks copying.
        n_e, grp_e, pal_host_e, e2e_split = n_local, grp, pal_host, None
        if dist is not None:
            probe_dev = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
            probe_host = torch.empty(256 << 20, dtype=torch.uint8).pin_memory()
            probe_host.copy_(probe_dev, non_blocking=True)
            barrier()
            e0.record(stream)
            for _ in range(2):
                probe_host.copy_(probe_dev, non_blocking=True)
            e1.record(stream)
            stream.synchronize()
            rate = 2 * probe_dev.numel() / (e0.elapsed_time(e1) * 1e-3) / 1e9
            rates = torch.zeros(world, dtype=torch.float64, device="cuda")
            rates[rank] = rate
            dist.all_reduce(rates)
            rates = rates.cpu().tolist()
            del probe_dev, probe_host
            if max(rates) > 1.1 * min(rates):
                raw = [args.instances * r / sum(rates) for r in rates]
                counts_e = [int(x) for x in raw]
                for k in sorted(range(world), key=lambda k: raw[k] - counts_e[k], reverse=True)[:args.instances - sum(counts_e)]:
                    counts_e[k] += 1
                n_e = counts_e[rank]
                reps = -(-max(n_e, 1) // n_local)
                pal_host_e = pal_host.repeat(reps, 1)[:max(n_e, 1)].contiguous().pin_memory()   # instance j plays palette j % n_local
                grp_e = _capi.Group(dmesh, max(n_e, 1))
                e2e_split = {"instances_per_rank": counts_e, "d2h_gbs_per_rank_all_copying": rates}
            barrier()
        out_host = torch.empty((max(n_e, 1), rows * stride), dtype=torch.uint8).pin_memory()
        h2d = n_e * T * 64
        d2h = n_e * rows * stride

        def e2e_step():   # p3cs_group_animate_host: upload / kernels / download pipelined in chunks of instances
            if n_e > 0:
                grp_e.animate_host_ptr(pal_host_e.data_ptr(), None, [out_host.data_ptr()])
            ctx.sync()

        e2e_step()
        barrier()
        t0 = time.perf_counter()
        e0.record(stream)
        for _ in range(e2e_steps):
            e2e_step()
        e1.record(stream)
        barrier()
        e2e_ms = max_over_ranks(max(e0.elapsed_time(e1), (time.perf_counter() - t0) * 1e3) / e2e_steps)
        check_row = out_host[n_e // 2].clone() if n_e > 0 else None   # for the sanity check below (the bare copies overwrite out_host)
        # the same bytes as bare copies (palettes up, arrays down, back to back on one stream), all ranks at once: what
        # the PCIe links and the host memory of this box give a program that does nothing else
        bare_dev = torch.empty((max(n_e, 1), rows * stride), dtype=torch.uint8, device="cuda") if grp_e is not grp else None
        pal_dev_e = torch.empty((max(n_e, 1), T * 16), dtype=torch.float32, device="cuda") if grp_e is not grp else pal_dev
        barrier()
        e0.record(stream)
        for _ in range(e2e_steps):
            pal_dev_e.copy_(pal_host_e, non_blocking=True)
            if bare_dev is not None:
                out_host.copy_(bare_dev, non_blocking=True)
            else:
                out_host.copy_(out_dev[:, :rows * stride] if pitch != rows * stride else out_dev, non_blocking=True)
        e1.record(stream)
        barrier()
        bare_ms = max_over_ranks(e0.elapsed_time(e1) / e2e_steps)
        del bare_dev
        sampler.stop_flag.set()
        sampler.join(timeout=2)

        # a cheap sanity check that the timed work was real: one instance against the oracle
        check = None
        if rank == 0 and not args.no_check and check_row is not None:
            from oracle import oracle
            inst = n_e // 2
            got = check_row.numpy().reshape(rows, stride)
            explained, unexplained, expect = oracle.unexplained_differences(flat, [got], pal_host_e[inst].numpy())
            g, e = got.view(np.float32).astype(np.float64), expect[0].view(np.float32).astype(np.float64)
            check = {"instance": inst, "max_abs_err": float(np.abs(g - e).max()), "differing_bytes": explained + unexplained,
                     "differing_bytes_outside_libm_normals": unexplained}
            # every byte must equal the oracle's (but for the last place of sinf / cosf in the rare uniform-scale normal
            # branch), or the timed work was not the reference's computation and no number is reported
            assert unexplained == 0 and check["max_abs_err"] <= 1e-5, f"bench check failed: {check}"

        # ---------------- "next" row f1: joint hierarchy on the GPU -- the host sends one frame number per
        # instance (40 KB) instead of the palettes (82 MB); pose + animate timed together on the device
        joints = None
        if not args.no_pose:
            skel = synth.make_skeleton(T, 64, seed=4321)
            dskel = _capi.Skeleton(ctx, skel)
            frames_host = (np.arange(n_local, dtype=np.int32) * 7 + 3) % 64
            for _ in range(3):
                grp.pose(dskel, frames_host)
                grp.animate()
            barrier()
            e0.record(stream)
            for k in range(args.steps):
                grp.pose(dskel, (frames_host + k) % 64)
                grp.animate()
            e1.record(stream)
            barrier()
            j_ms = max_over_ranks(e0.elapsed_time(e1) / args.steps)
            e0.record(stream)
            for k in range(args.steps):
                grp.pose(dskel, (frames_host + k) % 64)
            e1.record(stream)
            barrier()
            p_ms = max_over_ranks(e0.elapsed_time(e1) / args.steps)
            joints = {"pose_plus_animate_ms_per_step": j_ms, "pose_only_ms_per_step": p_ms, "joints_per_step": int(n_local * T),
                      "h2d_bytes_per_step": int(n_local * 4), "note": "p3cs_group_pose: channel tables -> compose_matrix -> "
                      "net transforms -> skinning matrices on the GPU (one AnimControl, no frame blending)"}
            pal_dev.copy_(pal_host, non_blocking=True)   # restore the palettes the checks below expect
            stream.synchronize()
            dskel.close()

        # ---------------- "next" row f4: animated tight bounds reduced by the same kernel
        bounds = None
        if not args.no_pose:
            grp.enable_bounds()
            for _ in range(3):
                grp.animate()
            barrier()
            e0.record(stream)
            for _ in range(args.steps):
                grp.animate()
            e1.record(stream)
            barrier()
            b_ms = max_over_ranks(e0.elapsed_time(e1) / args.steps)
            b = grp.read_bounds(n_local // 2, 1)[0]
            grp.read_output_ptr(0, n_local // 2, 1, out_host.data_ptr())
            ctx.sync()
            v = out_host[0].numpy().reshape(rows, stride)[:, :12].copy().view(np.float32)
            ok = bool(np.array_equal(b[:3], v.min(axis=0)) and np.array_equal(b[3:6], v.max(axis=0)))
            bounds = {"animate_with_bounds_ms_per_step": b_ms, "launches_per_step": int(grp.timings().launches), "exact": ok,
                      "note": "p3cs_group_enable_bounds: min/max/|v|^2max of the animated vertex column per instance, fused"}
            grp.enable_bounds(enable=False)

        # ---------------- optional: NCCL gather of the finished arrays to the rendering GPU (rank 0)
        gather = None
        if dist is not None:
            from panda3d_b200 import crowd
            view = out_dev[:n_local]
            per = -(-args.instances // world)
            recv = torch.empty((world, per, pitch), dtype=torch.uint8, device="cuda") if rank == 0 else None
            crowd.gather_outputs(view, args.instances, root=0, recv=recv, reorder=False)
            barrier()
            e0.record(stream)
            for _ in range(3):
                crowd.gather_outputs(view, args.instances, root=0, recv=recv, reorder=False)
            e1.record(stream)
            barrier()
            g_ms = max_over_ranks(e0.elapsed_time(e1) / 3)
            inbound = (args.instances - n_local) * pitch if rank == 0 else 0
            gather = {"ms_per_step": g_ms, "inbound_bytes_rank0": int(sum_over_ranks(float(inbound))),
                      "note": "dist.gather (NCCL) of every rank's [instances, pitch] uint8 rows to rank 0, not part of `value`"}

        # ---------------- optional: fused skin + gather -- every rank's kernel stores its finished tiles
        # straight into the rendering GPU's buffer over NVLink (peer-mapped output, no collective)
        fused = None
        if dist is not None and not args.no_fused:
            try:
                from panda3d_b200 import crowd
                per = -(-args.instances // world)
                target = crowd.RenderTarget(ctx, world, rank, per, pitch, root=0)
                grp2 = _capi.Group(dmesh, n_local, palettes_dev=pal_dev.data_ptr(), outputs_dev=[target.slab_ptr])
                for _ in range(3):
                    grp2.animate()
                barrier()
                e0.record(stream)
                for _ in range(args.steps):
                    grp2.animate()
                e1.record(stream)
                barrier()
                f_ms = max_over_ranks(e0.elapsed_time(e1) / args.steps)
                ok = None
                if rank == 0 and world > 1 and not args.no_check:
                    from oracle import oracle
                    n1 = len(range(1, args.instances, world))
                    pal1 = synth.make_palettes(T, n1, 1000 + 1)[n1 // 2]
                    expect, _, _ = oracle.OracleMesh(flat).animate(pal1)
                    got = target.read_row(1, n1 // 2, rows * stride).reshape(rows, stride)
                    gf = got.view(np.float32).astype(np.float64)
                    ef = expect[0].view(np.float32).astype(np.float64)
                    rel = float((np.abs(gf - ef) / np.maximum(np.abs(ef).max(axis=1, keepdims=True), 1e-30)).max())
                    ok = {"max_rel_err": rel, "differing_bytes": int((got != expect[0]).sum()), "within_1e-5": bool(rel <= 1e-5)}
                fused = {"ms_per_step": f_ms, "remote_instance_vs_oracle": ok,
                         "note": "animate with outputs peer-mapped on GPU 0: skinning and the gather to the rendering GPU in one kernel"}
                grp2.close()
                barrier()
                target.close()
            except Exception as exc:  # reported, never hidden
                fused = {"error": repr(exc)}

        # ---------------- skin + gather with a SINK-AWARE split: the rendering GPU keeps the share of the crowd that makes its
        # compute time equal to the time the other ranks' rows need on its inbound NVLink (crowd.sink_aware_split)
        gathered = None
        if dist is not None and not args.no_fused and isinstance(fused, dict) and "ms_per_step" in fused:
            try:
                from panda3d_b200 import crowd
                total_bytes = args.instances * pitch
                remote_even = (args.instances - len(range(0, args.instances, world))) * pitch
                ingress_gbs = remote_even / (fused["ms_per_step"] * 1e-3) / 1e9      # measured: the even deal is link-bound
                compute_ms_all = ms * world                                              # one GPU, whole crowd
                gather_ms_all = total_bytes / (ingress_gbs * 1e9) * 1e3
                def run_split(counts, steps):
                    n_w = counts[rank]
                    target = crowd.RenderTarget(ctx, world, rank, 0, pitch, root=0, counts=counts)
                    reps = -(-max(n_w, 1) // n_local)
                    pal_w = pal_dev.repeat(reps, 1)[:max(n_w, 1)].contiguous()           # weighted instance j plays palette j % n_local
                    grp3 = _capi.Group(dmesh, max(n_w, 1), palettes_dev=pal_w.data_ptr(), outputs_dev=[target.slab_ptr]) if n_w > 0 else None
                    for _ in range(3):
                        if grp3 is not None:
                            grp3.animate()
                    barrier()
                    e0.record(stream)
                    for _ in range(steps):
                        if grp3 is not None:
                            grp3.animate()
                    e1.record(stream)
                    barrier()
                    own = e0.elapsed_time(e1) / steps
                    return target, grp3, pal_w, max_over_ranks(own), own

                # first pass with the model's rates, second pass with the rates the first pass measured (the per-rank time
                # of the even deal carries its launch overhead, which overstates what the whole crowd costs one GPU)
                counts = crowd.sink_aware_split(args.instances, world, compute_ms_all, gather_ms_all, root=0)
                target, grp3, pal_w, g_ms, w_own_ms = run_split(counts, max(3, args.steps // 2))
                root_ms = w_own_ms if rank == 0 else 0.0
                remote_ms = w_own_ms if rank != 0 else 0.0
                root_ms, remote_ms = sum_over_ranks(root_ms), max_over_ranks(remote_ms)
                if counts[0] > 0 and counts[0] < args.instances and root_ms > 0 and remote_ms > 0:
                    compute_ms_all = root_ms * args.instances / counts[0]
                    gather_ms_all = remote_ms * args.instances / (args.instances - counts[0])
                    refined = crowd.sink_aware_split(args.instances, world, compute_ms_all, gather_ms_all, root=0)
                    if refined != counts:
                        if grp3 is not None:
                            grp3.close()
                        barrier()
                        target.close()
                        counts = refined
                        target, grp3, pal_w, g_ms, w_own_ms = run_split(counts, args.steps)
                ok = None
                if rank == 0 and not args.no_check:
                    from oracle import oracle
                    om = oracle.OracleMesh(flat)
                    ok = {}
                    for r_chk, j in ((0, counts[0] - 1), (1, counts[1] // 2), (world - 1, counts[world - 1] - 1)):
                        if counts[r_chk] == 0:
                            continue
                        n_r = len(range(r_chk, args.instances, world))
                        pal_r = synth.make_palettes(T, n_r, 1000 + r_chk)[j % n_r]
                        got = target.read_row(r_chk, j, rows * stride).reshape(rows, stride)
                        explained, unexplained, _ = oracle.unexplained_differences(flat, [got], pal_r, None, om)
                        ok[f"rank{r_chk}_instance{j}"] = {"differing_bytes": explained + unexplained, "outside_libm_normals": unexplained}
                        assert unexplained == 0, f"gathered check failed: {ok}"
                gathered = {"value": args.instances * flat.skinned_rows() / (g_ms * 1e-3), "unit": UNIT, "ms_per_step": g_ms,
                            "instances_per_rank": counts, "share_rank0": counts[0] / args.instances,
                            "nvlink_ingress_gbs_measured_even_deal": ingress_gbs,
                            "nvlink_ingress_frac_of_770": ingress_gbs / 770.0,
                            "model": {"compute_ms_all_on_one_gpu": compute_ms_all, "gather_ms_all_over_the_root_link": gather_ms_all},
                            "even_deal_ms_per_step": fused["ms_per_step"], "this_rank_ms": w_own_ms, "check": ok,
                            "note": "every rank's kernel stores its rows into the rendering GPU's buffer (peer-mapped, TMA bulk stores "
                                    "over NVLink); rank 0 keeps the share that balances its compute against its inbound link"}
                if grp3 is not None:
                    grp3.close()
                barrier()
                target.close()
            except AssertionError:
                raise
            except Exception as exc:  # reported, never hidden
                gathered = {"error": repr(exc)}

        # ---------------- the other BASELINE.json configurations, device time per animate call (rank 0, N=1):
        # parity-test cases, not bench lines -- reported so the single-mesh latencies sit next to the crowd number
        others = None
        if world == 1 and not args.no_configs:
            others = other_configs(torch, ctx, stream, _capi, synth, args.instances, hbm_peak()[0])

    total_verts = float(sum_over_ranks(float(n_local * flat.skinned_rows())))
    value = total_verts / (ms * 1e-3)
    e2e_value = total_verts / (e2e_ms * 1e-3)
    h2d_total, d2h_total = int(sum_over_ranks(float(h2d))), int(sum_over_ranks(float(d2h)))

    if rank == 0:
        peak, peak_src = hbm_peak()
        # dominant (only) kernel = strip_kernel: one launch per step.  Algorithmic bytes of an instanced
        # crowd: the source mesh and its tables are shared by all instances and
        # served from L2, so the compulsory HBM traffic per instance is its output rows (stride bytes per
        # vertex) plus its palette (T x 64 B); blend records never leave shared memory.
        # average launch duration of the dominant kernel: the timed region holds nothing else (one launch per step, back to
        # back on the launch stream), so it is that region's CUDA-event time over its launches -- rank 0's own, not the max
        # over ranks.  The per-launch event pairs of the profiling pass (each with its own launch gap) are kept beside it.
        skin_s = own_ms * 1e-3 / max(1, launches_per_step)
        skin_s_event_pairs = statistics.mean(skin_ms) * 1e-3
        alg_bytes_per_step = n_local * (rows * stride + T * 64)
        logical_bytes_per_step = n_local * rows * (2 * stride + 2)
        achieved = alg_bytes_per_step / skin_s / 1e9
        traffic = None
        pj_limiter = None   # what actually binds the kernel, from the committed ncu capture (not measured live)
        prof = os.path.join(ROOT, "profiles", "r02_run_kernel_c3.json")
        if os.path.exists(prof):
            pj = json.load(open(prof))
            if pj.get("instances") == n_local:
                traffic = pj.get("dram_bytes_per_launch")
                pj_limiter = pj.get("limiter")
        roofline = {"bound": "hbm", "kernel": "p3cs::run_kernel<kRow24>", "achieved": achieved, "peak": peak, "unit": "GB/s",
                    "frac": achieved / peak, "frac_of_nominal_8TBs": achieved / 8000.0, "traffic": traffic, "peak_source": peak_src,
                    "launches_per_step": launches_per_step, "avg_launch_ms": 1e3 * skin_s,
                    "avg_launch_ms_event_pairs": 1e3 * skin_s_event_pairs / max(1, skin_launches),
                    "algorithmic_bytes_per_launch": alg_bytes_per_step,
                    "algorithmic_bytes_per_vertex": alg_bytes_per_step / (n_local * rows),
                    "logical_gbs": logical_bytes_per_step / skin_s / 1e9,
                    "limiter": pj_limiter,
                    "note": "run kernel (one thread per run of rows, blend in registers, two palette copies); on this 24 B/vertex crowd the SM's L1TEX/LSU data pipe (80 % of its peak) and instruction issue (75 %) bind together, DRAM traffic equals the algorithmic bytes: see DESIGN.md"}
        cpu = None
        if world == 1 and not args.no_cpu:
            workers = os.cpu_count() or 1
            r = cpu_reference_run(sm, 100, 3, max(1, args.cpu_characters), workers)
            cpu = {"value": r["value"], "unit": UNIT, "cores": r["cores"], "kind": r["kind"], "wall_s": r.get("wall_s"),
                   "sample": f"{args.cpu_characters} character(s) x {workers} worker processes x 100 iterations of the same mesh "
                             f"(every joint matrix changed per iteration)"}
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic", "config": config_dict(args, world),
            "roofline": roofline, "cpu_baseline": cpu,
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d_total, "d2h_bytes_per_step": d2h_total,
                    "ms_per_step": e2e_ms, "steps": e2e_steps,
                    "pcie_roofline": {"bare_copies_ms_per_step": bare_ms, "bare_gbs_whole_job": (h2d_total + d2h_total) / bare_ms / 1e6,
                                      "e2e_over_bare": e2e_ms / bare_ms,
                                      "note": "the same H2D + D2H bytes as plain copies on every rank at once, no kernel: the box's PCIe / host-memory ceiling"},
                    "host_binding": binding, "split_by_link_rate": e2e_split},
            "gpu_launches": launches_per_step * args.steps, "clocks": sampler.summary(), "check": check,
            "gather": gather, "fused_gather": fused, "gathered": gathered, "gathered_value": (gathered or {}).get("value"), "joints_on_gpu": joints, "tight_bounds": bounds, "other_configs": others,
        }
        emit(line)
    if grp_e is not grp:
        grp_e.close()
    grp.close()
    dmesh.close()
    ctx.close()
    if dist is not None:
        dist.destroy_process_group()


def hbm_peak():
    peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(peaks_path):
        return float(json.load(open(peaks_path))["hbm_gbs"]), "MEASURED_PEAKS.json hbm_gbs (measured copy)"
    return 6650.0, "fallback 6.65 TB/s (B200_PROFILING.md)"


def other_configs(torch, ctx, stream, _capi, synth, instances, peak_gbs):
    """Device time per animate call of configs[1], [3], [4] (C2 single mesh, C4 morph mesh, C5 mixed crowd), L2
    flushed between calls (a 256 MB buffer is rewritten), CUDA events on the launch stream."""
    flush = torch.zeros(256 << 20, dtype=torch.uint8, device="cuda")
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    out = {}
    specs = {
        "c2_100k_single_mesh": [(synth.config_c2("random").flat, 1)],
        "c4_250k_morph_mesh": [(synth.config_c4().flat, 1)],
        "c5_mixed_crowd_2M": [(m.flat, n) for m, n in synth.config_c5_meshes()],
    }
    # the same crowd with the loader's full row layout (vertex normal texcoord tangent binormal, stride 56):
    # 2.3x the bytes per vertex, where the kernel IS bound by HBM
    specs["c3_crowd_stride56_pos_normal_uv_tangent_binormal"] = [
        (synth.make_mesh(20_000, 128, 4096, seed=33, tbn=True, texcoord=True, index_order="runs").flat, instances)]
    # ... and with a random blend per vertex (no runs at all: SURVEY 8d lists this order for C2): every row needs its own
    # four-matrix blend, the kernel is bound by instruction issue, not by HBM
    specs["c3_crowd_random_blend_index_per_vertex"] = [(synth.config_c3_mesh(index_order="random").flat, instances)]
    for name, spec in specs.items():
        meshes, groups, verts = [], [], 0
        for i, (fm, n) in enumerate(spec):
            dm = _capi.Mesh(ctx, fm)
            g = _capi.Group(dm, n)
            g.set_palettes(synth.make_palettes(fm.num_transforms, n, 31 + i).reshape(n, -1))
            if fm.num_sliders:
                g.set_sliders(synth.make_sliders(fm.num_sliders, n, 41 + i))
            meshes.append(dm)
            groups.append(g)
            verts += n * fm.skinned_rows()
        for _ in range(3):
            _capi.animate_groups(groups)
        stream.synchronize()
        total, reps = 0.0, 20
        for _ in range(reps):
            flush.add_(1)
            e0.record(stream)
            _capi.animate_groups(groups)
            e1.record(stream)
            stream.synchronize()
            total += e0.elapsed_time(e1)
        us = total / reps * 1e3
        alg = sum(g.n * (g.mesh.num_rows * sum(g.mesh.out_strides) + g.mesh.flat.num_transforms * 64) for g in groups)
        out[name] = {"us_per_call": us, "vertices": verts, "vertices_per_s": verts / (us * 1e-6),
                     "launches": int(sum(g.timings().launches for g in groups)),
                     "crowd_algorithmic_gbs": alg / (us * 1e-6) / 1e9, "frac_of_hbm_peak": alg / (us * 1e-6) / 1e9 / peak_gbs}
        for g in groups:
            g.close()
        for dm in meshes:
            dm.close()
    del flush
    out["single_call_latency"] = single_call_latency(torch, ctx, _capi, synth)
    return out


def single_call_latency(torch, ctx, _capi, synth):
    """Wall-clock microseconds of ONE synchronous p3cs_animate_vertices call (host palette in; output left on the device
    or written into registered host memory), beside the floor of this box: launch of an empty kernel + polling wait."""
    import ctypes as C
    from panda3d_b200.casefile import load_case
    from panda3d_b200.flat import mesh_from_case

    def best(fn, n=300, reps=5):
        b = 1e9
        for _ in range(reps):
            t0 = time.perf_counter()
            for it in range(n):
                fn(it)
            b = min(b, (time.perf_counter() - t0) / n * 1e6)
        return b

    s = torch.cuda.Stream()
    x = torch.zeros(1, device="cuda")

    def empty(it):
        x.add_(1)
        while not s.query():
            pass
    with torch.cuda.stream(s):
        floor = best(empty)
    res = {"empty_kernel_launch_plus_poll_us": floor,
           "note": "best of 5 x 300 calls, time.perf_counter around the C-ABI call; a call cannot be faster than the floor"}
    L = _capi.lib()
    cases = {"c1_panda_walk4": mesh_from_case(load_case(os.path.join(ROOT, "tests", "golden", "panda_walk4.p3cs"))),
             "c2_100k": synth.config_c2("random").flat}
    for name, flat in cases.items():
        dm = _capi.Mesh(ctx, flat)
        g = _capi.Group(dm, 1)
        pal = np.ascontiguousarray(synth.make_palettes(flat.num_transforms, 8, 3), np.float32)
        pp = [pal[i].ctypes.data for i in range(8)]
        host = np.empty((flat.num_rows, dm.out_strides[0]), np.uint8)
        ptrs = (C.c_void_p * 1)(host.ctypes.data)
        fn, gh = L.p3cs_animate_vertices, g.handle
        for it in range(20):
            fn(gh, 0, pp[it & 7], None, None)
        on_device = best(lambda it: fn(gh, 0, pp[it & 7], None, None))
        ctx.host_register(host)
        try:
            for it in range(20):
                fn(gh, 0, pp[it & 7], None, ptrs)
            to_host = best(lambda it: fn(gh, 0, pp[it & 7], None, ptrs))
        finally:
            ctx.host_unregister(host)
        out_bytes = flat.num_rows * dm.out_strides[0]
        res[name] = {"rows": flat.num_rows, "us_per_call_output_on_device": on_device, "us_per_call_registered_host_output": to_host,
                     "output_bytes": out_bytes}
        g.close()
        dm.close()
    return res


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=30)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--instances", type=int, default=TOTAL_INSTANCES)
    ap.add_argument("--e2e-steps", type=int, default=5)
    ap.add_argument("--cpu-characters", type=int, default=32, help="characters per CPU worker process per step")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-check", action="store_true")
    ap.add_argument("--no-pose", action="store_true", help="skip the joint-hierarchy-on-GPU measurement")
    ap.add_argument("--no-configs", action="store_true", help="skip the device timings of the other BASELINE configurations")
    ap.add_argument("--no-fused", action="store_true", help="skip the peer-memory fused skin+gather measurement (N>1)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)
    if args.impl == "reference":
        run_reference_arm(args)
    else:
        run_gpu_arm(args)


if __name__ == "__main__":
    main()
