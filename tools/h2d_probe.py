#!/usr/bin/env python
"""Where does the end-to-end (host-buffer) path stop scaling over the GPUs of one box?  

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29517 tools/h2d_probe.py

Every rank owns one GPU and copies config-2 cubes (67 MB, pinned host memory -> device, one cudaMemcpyAsync per cube, exactly
what rsp_submit_cpi issues) for a fixed time.  The copy rate is measured with 1, 2, 4, ... ranks active AT THE SAME TIME (the
others wait at a barrier), once with default pinned allocations (wherever the allocating thread happens to run) and once with
the allocation made and first touched on the CPUs the GPU's PCI device is local to (sysfs local_cpulist).  Rank 0 prints one
JSON object: the topology every rank sees and the per-rank / aggregate GB/s of each case."""
import json
import os
import time

import torch
import torch.distributed as dist

CUBE_BYTES = 16 * 64 * 8192 * 8
N_CUBES = 4
SECONDS = 0.6


def pci_sysfs(dev: int):
    pr = torch.cuda.get_device_properties(dev)
    if all(hasattr(pr, k) for k in ("pci_bus_id", "pci_device_id", "pci_domain_id")) and isinstance(pr.pci_bus_id, int):
        busid = f"{pr.pci_domain_id:04x}:{pr.pci_bus_id:02x}:{pr.pci_device_id:02x}.0"
    else:
        import pynvml
        pynvml.nvmlInit()
        busid = pynvml.nvmlDeviceGetPciInfo(pynvml.nvmlDeviceGetHandleByIndex(dev)).busId
        if isinstance(busid, bytes):
            busid = busid.decode()
        busid = busid.lower()
        if len(busid.split(":")[0]) == 8:          # nvml gives an 8-digit domain, sysfs uses 4
            busid = busid[4:]
    base = f"/sys/bus/pci/devices/{busid}"
    out = {"busid": busid}
    for k in ("numa_node", "local_cpulist", "current_link_speed", "current_link_width", "max_link_speed", "max_link_width"):
        try:
            out[k] = open(f"{base}/{k}").read().strip()
        except OSError:
            out[k] = None
    return out


def parse_cpulist(s):
    cpus = set()
    for part in (s or "").split(","):
        if "-" in part:
            a, b = part.split("-")
            cpus.update(range(int(a), int(b) + 1))
        elif part.strip():
            cpus.add(int(part))
    return cpus


def pinned_pool(local_cpus):
    """N_CUBES cubes of pinned host memory; with local_cpus the pages are first touched by a thread bound to those CPUs."""
    old = os.sched_getaffinity(0)
    if local_cpus:
        try:
            os.sched_setaffinity(0, local_cpus & old or old)
        except OSError:
            pass
    t = torch.empty(N_CUBES * CUBE_BYTES, dtype=torch.uint8)
    t.fill_(1)                                  # first touch on the bound CPUs
    rc = torch.cuda.cudart().cudaHostRegister(t.data_ptr(), t.numel(), 0)
    os.sched_setaffinity(0, old)
    assert int(rc) == 0, rc
    return t


def copy_rate(host, devbuf, stream, active: bool):
    """GB/s of this rank over ~SECONDS of back-to-back cube copies (0 when the rank sits this case out)."""
    dist.barrier()
    torch.cuda.synchronize()
    if not active:
        dist.barrier()
        return 0.0
    n = 0
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with torch.cuda.stream(stream):
        e0.record()
        t0 = time.perf_counter()
        while time.perf_counter() - t0 < SECONDS:
            for i in range(N_CUBES):
                devbuf[i % 2].copy_(host[i * CUBE_BYTES:(i + 1) * CUBE_BYTES], non_blocking=True)
            n += N_CUBES
            if n % 16 == 0:
                stream.synchronize()            # bound the queue depth
        e1.record()
    stream.synchronize()
    ms = e0.elapsed_time(e1)
    dist.barrier()
    return n * CUBE_BYTES / ms / 1e6


def main():
    rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
    local = int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    for k, v in (("RANK", "0"), ("WORLD_SIZE", "1"), ("MASTER_ADDR", "127.0.0.1"), ("MASTER_PORT", "29517")):
        os.environ.setdefault(k, v)            # plain `python tools/h2d_probe.py` = one rank
    dist.init_process_group("nccl" if world > 1 else "gloo", device_id=torch.device("cuda", local) if world > 1 else None)
    topo = pci_sysfs(local)
    topo["sched_affinity"] = f"{min(os.sched_getaffinity(0))}-{max(os.sched_getaffinity(0))} ({len(os.sched_getaffinity(0))} cpus)"
    stream = torch.cuda.Stream()
    devbuf = torch.empty((2, CUBE_BYTES), dtype=torch.uint8, device="cuda")
    pools = {"default": pinned_pool(None), "numa_local": pinned_pool(parse_cpulist(topo.get("local_cpulist")))}
    results = {}
    actives = sorted({1, 2, 4, 8, world} & set(range(1, world + 1)))
    for mode, host in pools.items():
        for k in actives:
            r = copy_rate(host, devbuf, stream, rank < k)
            rates = [None] * world
            dist.all_gather_object(rates, r)
            results[f"{mode}/{k}_active"] = {"per_rank_gbs": [round(x, 1) for x in rates[:k]], "aggregate_gbs": round(sum(rates), 1),
                                             "cubes_per_s": round(sum(rates) * 1e9 / CUBE_BYTES, 0)}
    topos = [None] * world
    dist.all_gather_object(topos, topo)
    if rank == 0:
        try:
            nodes = sorted(d for d in os.listdir("/sys/devices/system/node") if d.startswith("node"))
        except OSError:
            nodes = []
        print(json.dumps({"world": world, "cube_bytes": CUBE_BYTES, "host_numa_nodes": nodes, "host_cpus": os.cpu_count(),
                          "gpus": topos, "h2d": results}, indent=1))
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
