#!/usr/bin/env python
"""Where does the end-to-end (host-buffer) ceiling of a multi-GPU box come from?  (94 GB/s aggregate over
eight PCIe links.)  Run under torchrun, one rank per GPU:

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 tools/pcie_probe.py

Per rank: the GPU's PCI address, its NUMA node and local CPU list (sysfs), then device->host copies of 1 GB into
page-locked memory: (a) one rank at a time, (b) all ranks at once, (c) all at once after the rank bound itself to its
GPU's local CPUs and allocated a fresh buffer there (first touch).  Rank 0 also prints the box's topology.
"""
import json
import os
import subprocess
import sys
import time

import torch
import torch.distributed as dist


def sysfs(path):
    try:
        return open(path).read().strip()
    except OSError:
        return None


def parse_cpulist(s):
    cpus = []
    for part in (s or "").split(","):
        if "-" in part:
            a, b = part.split("-")
            cpus += list(range(int(a), int(b) + 1))
        elif part.strip():
            cpus.append(int(part))
    return cpus


def d2h_rate(dev_buf, host_buf, reps=4):
    host_buf.copy_(dev_buf, non_blocking=True)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps):
        host_buf.copy_(dev_buf, non_blocking=True)
    torch.cuda.synchronize()
    return dev_buf.numel() * reps / (time.perf_counter() - t0) / 1e9


def main():
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    props = torch.cuda.get_device_properties(local)
    bus = "%04x:%02x:%02x.0" % (getattr(props, "pci_domain_id", 0), getattr(props, "pci_bus_id", 0), getattr(props, "pci_device_id", 0))
    base = f"/sys/bus/pci/devices/{bus}"
    info = {"rank": rank, "pci": bus, "numa_node": sysfs(base + "/numa_node"), "local_cpulist": sysfs(base + "/local_cpulist"),
            "affinity_at_start": len(os.sched_getaffinity(0)), "link_speed": sysfs(base + "/current_link_speed"),
            "link_width": sysfs(base + "/current_link_width")}
    n = 1 << 30
    dev_buf = torch.empty(n, dtype=torch.uint8, device="cuda")
    host_buf = torch.empty(n, dtype=torch.uint8).pin_memory()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # (a) one rank at a time
    for r in range(world):
        barrier()
        if r == rank:
            info["solo_gbs"] = d2h_rate(dev_buf, host_buf)
    # (b) all at once
    barrier()
    info["concurrent_gbs"] = d2h_rate(dev_buf, host_buf)
    barrier()
    # (c) bound to the GPU's local CPUs, fresh buffer
    cpus = parse_cpulist(info["local_cpulist"])
    try:
        if cpus:
            os.sched_setaffinity(0, set(cpus) & os.sched_getaffinity(0) or os.sched_getaffinity(0))
        info["affinity_bound"] = len(os.sched_getaffinity(0))
        host2 = torch.empty(n, dtype=torch.uint8).pin_memory()
        host2.fill_(1)
        barrier()
        info["concurrent_bound_gbs"] = d2h_rate(dev_buf, host2)
        barrier()
        info["solo_bound_gbs"] = None
        for r in range(world):
            barrier()
            if r == rank:
                info["solo_bound_gbs"] = d2h_rate(dev_buf, host2)
        # host -> device for completeness
        barrier()
        t0 = time.perf_counter()
        for _ in range(4):
            dev_buf.copy_(host2, non_blocking=True)
        torch.cuda.synchronize()
        info["concurrent_h2d_gbs"] = n * 4 / (time.perf_counter() - t0) / 1e9
    except Exception as exc:  # reported, not hidden
        info["bound_error"] = repr(exc)
    os.makedirs("gpurun_out", exist_ok=True)
    with open(f"gpurun_out/pcie_rank{rank}.json", "w") as f:
        json.dump(info, f)
    print(json.dumps(info), flush=True)
    if rank == 0:
        box = {}
        for name, cmd in (("topo", ["nvidia-smi", "topo", "-m"]), ("lscpu", ["lscpu"]), ("numactl", ["numactl", "-H"]),
                          ("meminfo_nodes", ["sh", "-c", "ls /sys/devices/system/node/ | tr '\\n' ' '"])):
            try:
                box[name] = subprocess.run(cmd, capture_output=True, text=True, timeout=20).stdout[-4000:]
            except Exception as exc:
                box[name] = repr(exc)
        with open("gpurun_out/pcie_box.json", "w") as f:
            json.dump(box, f, indent=1)
    barrier()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
